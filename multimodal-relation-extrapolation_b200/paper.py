"""The paper-side (main.py / module/) entry points of the hot path, mirrored over the library:
  evaluate(...)          main.evaluate (main.py:217-272): TransE-L1 candidate ranking, ties//2, MRR + Hits@1/3/10
  PaperScorer            module.NegativeSampling._calc / .evaluate (module/NegativeSampling.py:142-168, 294-305)
  NegativeSampling       the sampler + structural-loss half of module.NegativeSampling (:114-140, 177-185, 196-229,
                         307-314, 321-375): neg_sample_fn / generate_eval_list / scoring_fn / regularization
  build_test_candidates  utils/gen_mode_candidates.py:15-39 (regenerates the missing {mode}_candidates.json)
  load_tasks, gen_e1rel_e2, gen_rel2candidates, evaluate_from_dir
                         module/utils.py:194-207, utils/gen_e1r_e2_all.py:14-19, utils/gen_rel2candidates.py:14-27
  zsl_rank_metrics       ZSLmodule.eval's rank/metric block (module/zsl_module.py:699-745): Hits@10/5/1 + MRR
  ZSLEvaluator           ZSLmodule.eval end to end (module/zsl_module.py:635-745): Extractor + cosine-mean + rank on the GPU
"""
import numpy as np
import torch

from . import _lib as L
from . import engine


class _TripleScoreFn(torch.autograd.Function):
    """score of n explicit (h, t, r) index triples into a row table `ent` [rows, D] and a relation-row table `rel`, differentiable
    with respect to both tables: mre_score_triples forward, mre_transe_backward / mre_bilinear_backward backward (float atomics
    into dense gradients of the tables' shapes)"""

    @staticmethod
    def forward(ctx, ent, rel, h, t, r, scorer, p_norm, norm_flag, lib_ctx):
        n = h.numel()
        out = torch.empty(n, dtype=torch.float32, device=ent.device)
        L.check(L.lib().mre_score_triples(lib_ctx._h, scorer, ent.data_ptr(), None, rel.data_ptr(), None, ent.shape[1], h.data_ptr(),
                                          t.data_ptr(), r.data_ptr(), n, p_norm, int(norm_flag), out.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))
        ctx.save_for_backward(ent, rel, h, t, r, out)
        ctx.cfg = (scorer, p_norm, norm_flag, lib_ctx)
        return out

    @staticmethod
    def backward(ctx, dscore):
        ent, rel, h, t, r, score = ctx.saved_tensors
        scorer, p_norm, norm_flag, lib_ctx = ctx.cfg
        ge, gr = torch.zeros_like(ent), torch.zeros_like(rel)
        dscore = dscore.contiguous()
        st = torch.cuda.current_stream().cuda_stream
        if scorer == L.TRANSE:
            L.check(L.lib().mre_transe_backward(lib_ctx._h, ent.data_ptr(), rel.data_ptr(), ent.shape[1], h.data_ptr(), t.data_ptr(),
                                                r.data_ptr(), h.numel(), p_norm, int(norm_flag), score.data_ptr(), dscore.data_ptr(),
                                                ge.data_ptr(), gr.data_ptr(), st))
        else:
            L.check(L.lib().mre_bilinear_backward(lib_ctx._h, scorer, ent.data_ptr(), None, rel.data_ptr(), None, ent.shape[1],
                                                  h.data_ptr(), t.data_ptr(), r.data_ptr(), h.numel(), dscore.data_ptr(),
                                                  ge.data_ptr(), None, gr.data_ptr(), None, st))
        return ge, gr, None, None, None, None, None, None, None


class PaperScorer:
    """score_model 'transe': ||(h + r) - t||_1 without normalisation (score_norm_flag False :31, p_norm 1 :47);
    'distmult': sum h*r*t.  Inputs are embedding ROWS ([n, D] tensors) as in the reference's _calc(h, t, r); the result is
    differentiable with respect to them (the reference trains the RGCN / M3AE outputs through this score)."""

    SCORERS = {"transe": L.TRANSE, "distmult": L.DISTMULT}

    def __init__(self, p_norm=1, score_norm_flag=False, device=0):
        self.p_norm, self.norm_flag = p_norm, score_norm_flag
        self.ctx = engine.Context(device)
        self.device = torch.device("cuda", device)

    def score_indexed(self, ent, rel, h, t, r, score_model="transe"):
        """scores of the triples (ent[h], rel[r], ent[t]) without gathering the rows: ent [rows, D], rel [rows', D] float32,
        h / t / r int64 index vectors of equal length, all on the scorer's device"""
        ent, rel = ent.to(self.device, torch.float32).contiguous(), rel.to(self.device, torch.float32).contiguous()
        h, t, r = (x.to(self.device, torch.int64).contiguous() for x in (h, t, r))
        if h.numel() and (int(torch.stack([h.min(), t.min(), r.min()]).min()) < 0 or int(torch.stack([h.max(), t.max()]).max()) >= ent.shape[0]
                          or int(r.max()) >= rel.shape[0]):
            raise IndexError("triple index out of range of the row tables")
        return _TripleScoreFn.apply(ent, rel, h, t, r, self.SCORERS[score_model], self.p_norm, self.norm_flag, self.ctx)

    def _calc(self, h, t, r, mode="normal", score_model="transe"):
        n = max(h.shape[0], t.shape[0], r.shape[0])
        D = h.shape[-1]
        rows = [x.reshape(-1, D).to(self.device, torch.float32) for x in (h, t, r)]
        rows = [x if x.shape[0] == n else x.repeat(n // x.shape[0], 1) for x in rows]      # the view(-1, B, D) broadcast (:148-151)
        ent = torch.cat(rows[:2])                        # [2n, D]: heads then tails
        idx = torch.arange(n, device=self.device)
        tails = idx + n                                  # kept alive until the call returns
        return self.score_indexed(ent, rows[2], idx, tails, idx, score_model)

    def evaluate(self, h, r, t, score_model="transe"):
        if score_model != "transe":
            print("invalid scoring model!")
            return None
        return self._calc(h, t, r)


class NegativeSampling(PaperScorer):
    """module.NegativeSampling without the feature producers (UnifiedModel = M3AE + RGCN is out of scope): the caller
    passes the node features `x` and relation embeddings the model produced.  Same constructor keywords; `whole_triples`
    is the reference's (h, r, t) tuple of GLOBAL train ids (main.py:55-68).  The sampler runs on the GPU
    (mre_sample_subgraph): distributionally the reference's random.sample + np.in1d loop, bit-reproducible per
    (seed, call number) instead of Python's unseeded `random`."""

    def __init__(self, args=None, whole_triples=None, model=None, loss_fn=None, regul_rate=0.5, neg_ent=1,
                 sampling_mode="normal", bern_flag=False, filter_flag=True, score_norm_flag=False, *, num_entities=None,
                 num_relations=None, seed=0, device=0, index=None):
        super().__init__(p_norm=1, score_norm_flag=score_norm_flag, device=device)
        if sampling_mode != "normal":
            raise NotImplementedError("the reference only implements sampling_mode='normal' (module/NegativeSampling.py:119)")
        self.args, self.model, self.loss_fn, self.regul_rate = args, model, loss_fn, regul_rate
        self.neg_ent, self.bern_flag, self.filter_flag, self.sampling_mode = neg_ent, bern_flag, filter_flag, sampling_mode
        self.seed, self.calls = int(seed), 0
        self.index = index
        if index is None and whole_triples is not None:
            h, r, t = (np.ascontiguousarray(np.asarray(x), dtype=np.int64) for x in whole_triples)
            E = int(num_entities) if num_entities is not None else int(max(h.max(), t.max())) + 1
            R = int(num_relations) if num_relations is not None else int(r.max()) + 1
            self.index = engine.KGIndex.from_arrays(E, R, (h, t, r))
        if self.index is not None and self.index.device is None:
            self.index.to_device(device)

    @staticmethod
    def _l2g_array(local_global_id):
        if isinstance(local_global_id, dict):
            n = max(local_global_id) + 1 if local_global_id else 0
            a = np.full(n, -1, np.int64)
            for k, v in local_global_id.items():
                a[int(k)] = int(v)
            return a
        return np.ascontiguousarray(np.asarray(local_global_id), dtype=np.int64)

    def neg_sample_fn(self, local_global_id, node_list, edge_index, edge_type, *, device_out=False):
        """-> (expand_edge_index int32 [2, n(1+neg_ent)], expand_edge_type int32 [n(1+neg_ent)]), CPU tensors as the
        reference returns them (device tensors with device_out=True, saving the round trip the reference then undoes)"""
        assert edge_index.shape[0] == 2
        if self.index is None:
            raise L.MreError("neg_sample_fn needs the train triples (whole_triples=...)")
        dev = self.device
        to = lambda a: torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a).to(dev, torch.int64).contiguous()
        eh, et, er = to(edge_index[0]), to(edge_index[1]), to(edge_type)
        nodes = to(node_list)
        l2g = to(self._l2g_array(local_global_id))
        n = eh.numel()
        out = torch.empty((3, n * (1 + self.neg_ent)), dtype=torch.int32, device=dev)
        L.check(L.lib().mre_sample_subgraph(
            self.ctx._h, self.index._h, self.seed, self.calls, 0, eh.data_ptr(), et.data_ptr(), er.data_ptr(), n,
            nodes.data_ptr(), nodes.numel(), l2g.data_ptr(), l2g.numel(), self.neg_ent, int(bool(self.bern_flag)),
            int(bool(self.filter_flag)), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
            torch.cuda.current_stream().cuda_stream))
        self.calls += 1
        if device_out:
            return out[:2], out[2]
        o = out.cpu()
        return o[:2], o[2]

    def generate_eval_list(self, local_global_id, edge_index, edge_type):
        mapped_node_list = torch.arange(int(torch.as_tensor(edge_index).max()))      # excludes the max id, as :210 does
        return self.neg_sample_fn(local_global_id, mapped_node_list, edge_index, edge_type)

    def scoring_fn(self, local_global_id, x, relations, edge_index, edge_type):
        """:102-109 -- _calc(h = x[edge_index[0]], t = x[edge_index[1]], r = relations) without materialising the gathered rows;
        differentiable with respect to x and relations"""
        ei = torch.as_tensor(edge_index)
        return self.score_indexed(x, relations, ei[0].long(), ei[1].long(), torch.arange(relations.shape[0], device=self.device))

    def _get_positive_score(self, score, num_pos_samples):
        return score[:num_pos_samples].view(-1, num_pos_samples).permute(1, 0)

    def _get_negative_score(self, score, num_pos_samples):
        return score[num_pos_samples:].view(-1, num_pos_samples).permute(1, 0)

    def regularization(self, x, relations, edge_index, edge_type):
        bh, bt = x[edge_index[0].long()], x[edge_index[1].long()]
        return (torch.mean(bh ** 2) + torch.mean(bt ** 2) + torch.mean(relations ** 2)) / 3

    def struct_loss(self, local_global_id, x, rel_emb, edge_index, edge_type):
        """forward's structural part (:196-229): sample, score, margin loss (+ regul_rate * regularization)"""
        node_list = torch.arange(int(torch.as_tensor(edge_index).max()))
        ei, etype = self.neg_sample_fn(local_global_id, node_list, edge_index, edge_type, device_out=True)
        rel_expand = rel_emb.repeat(1 + self.neg_ent, 1).to(self.device)
        score = self.scoring_fn(local_global_id, x.to(self.device), rel_expand, ei, etype)
        n = len(edge_type)
        loss = self.loss_fn(self._get_positive_score(score, n), self._get_negative_score(score, n))
        if self.regul_rate != 0:
            loss = loss + self.regul_rate * self.regularization(x.to(self.device), rel_expand, ei, etype)
        return loss


def build_test_candidates(triples, rel2candidates, e1rel_e2, entity2id=None):
    """triples: iterable of (head, rel, tail) symbols -> {rel: {"head\\trel\\ttail": [tail, cand, ...]}} with the true tail
    first, then every candidate of the relation that is a known entity (utils/gen_mode_candidates.py:31-32), neither a known
    tail of (head, rel) nor the true tail (:33-34)."""
    out = {}
    for head, rel, tail in triples:
        known = set(e1rel_e2.get(head + rel, ()))
        cands = [tail] + [c for c in rel2candidates[rel]
                          if (entity2id is None or c in entity2id) and c not in known and c != tail]
        out.setdefault(rel, {})["\t".join((head, rel, tail))] = cands
    return out


# ---- the reference's data files: loaders and the generators of the artefacts its repository does not bundle
def load_tasks(data_path, mode="test"):
    """{mode}_tasks_zsl.json -> (tasks dict, (h, r, t) int64 id arrays, entity2id, relation2id) with entity2ids_zsl.json /
    relation2ids.json, as load_appendix_data reads them (module/utils.py:194-207)"""
    import json
    import os
    e_id = json.load(open(os.path.join(data_path, "entity2ids_zsl.json")))
    r_id = json.load(open(os.path.join(data_path, "relation2ids.json")))
    tasks = json.load(open(os.path.join(data_path, f"{mode}_tasks_zsl.json")))
    h, r, t = [], [], []
    for rel in tasks:
        for head, rela, tail in tasks[rel]:
            h.append(e_id[head]); r.append(r_id[rela]); t.append(e_id[tail])
    return tasks, (np.asarray(h, np.int64), np.asarray(r, np.int64), np.asarray(t, np.int64)), e_id, r_id


def gen_e1rel_e2(triples):
    """e1rel_e2_all.json (utils/gen_e1r_e2_all.py:14-19): head + relation symbol -> list of tails, from (head, rel, tail) symbols"""
    out = {}
    for head, rel, tail in triples:
        out.setdefault(head + rel, []).append(tail)
    return out


def gen_rel2candidates(triples, entities, n=300, seed=None):
    """rel2candidates_all.json (utils/gen_rel2candidates.py:14-27): `n` entities drawn without replacement per relation that
    occurs in the triples (the reference uses the unseeded `random` module; pass `seed` for a reproducible file)"""
    import random
    rng = random.Random(seed)
    rels = []
    for _, rel, _ in triples:
        if rel not in rels:
            rels.append(rel)
    entities = list(entities)
    return {rel: rng.sample(entities, n) for rel in rels}


def evaluate_from_dir(data_path, ent_embs, rel_embs, mode="test", hits_at_k=(1, 3, 10), ranker=None, verbose=True):
    """main.evaluate end to end from a reference dataset directory: reads entity2ids_zsl.json, relation2ids.json,
    {mode}_tasks_zsl.json, rel2candidates_all.json, uses {mode}_candidates.json / e1rel_e2_all.json when present and
    regenerates them otherwise (from the tasks themselves -- the bundled ZS datasets carry no train.tsv)."""
    import json
    import os
    tasks, _, e2id, r2id = load_tasks(data_path, mode)
    cand_file = os.path.join(data_path, f"{mode}_candidates.json")
    if os.path.exists(cand_file):
        test_candidates = json.load(open(cand_file))
    else:
        rel2cand = json.load(open(os.path.join(data_path, "rel2candidates_all.json")))
        triples = [tuple(tri) for rel in tasks for tri in tasks[rel]]
        e1_file = os.path.join(data_path, "e1rel_e2_all.json")
        e1rel_e2 = json.load(open(e1_file)) if os.path.exists(e1_file) else gen_e1rel_e2(triples)
        test_candidates = build_test_candidates(triples, rel2cand, e1rel_e2, e2id)
    return evaluate(ent_embs, rel_embs, e2id, r2id, test_candidates, hits_at_k=hits_at_k, ranker=ranker, verbose=verbose)


def _plan_candidates(test_candidates, e2id, r2id):
    """per-triple candidate lists (true first) -> per-relation candidate groups S_r = union of the lists + per-query
    exclusion lists S_r minus own list, so that one fused pass with a CSR filter reproduces every per-triple rank"""
    q_h, q_t, q_r, excl, groups, counts, names = [], [], [], [], [], [], []
    for rel, items in test_candidates.items():
        lists = []
        for key, cands in items.items():
            head, rela, _ = key.split("\t")
            ids = np.fromiter((e2id[c] for c in cands), np.int64, len(cands))
            q_h.append(e2id[head]); q_r.append(r2id[rela]); q_t.append(int(ids[0]))
            lists.append(ids)
        if not lists:
            continue
        S = np.unique(np.concatenate(lists))
        groups.append(S)
        counts.append(len(lists))
        names.append(rel)
        for ids in lists:
            excl.append(np.setdiff1d(S, ids[1:], assume_unique=False))   # includes the true tail: never ranked against itself
    return (np.asarray(q_h, np.int64), np.asarray(q_t, np.int64), np.asarray(q_r, np.int64)), excl, groups, counts, names


def evaluate(ent_embs, rel_embs, e2id, r2id, test_candidates, hits_at_k=(1, 3, 10), ranker=None, verbose=True, return_ranks=False):
    """main.evaluate: ranks the true tail of every test triple among its candidate list with the TransE-L1 scorer,
    rank = #(n < p) + #(n == p) // 2 + 1.  Returns (mrr, hits@k...) over all triples (the reference prints them and returns
    nothing, main.py:263-272); prints the reference's lines.  return_ranks=True appends the int64 ranks in file order."""
    ranker = ranker or engine.Ranker(device=ent_embs.device.index or 0 if ent_embs.is_cuda else 0)
    dev = ranker.device
    ent = ent_embs.detach().to(dev, torch.float32).contiguous()
    rel = rel_embs.detach().to(dev, torch.float32).contiguous()
    (q_h, q_t, q_r), excl, groups, counts, names = _plan_candidates(test_candidates, e2id, r2id)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    fptr = np.concatenate([[0], np.cumsum([len(x) for x in excl])]).astype(np.int64)
    fidx = np.concatenate(excl) if excl else np.zeros(0, np.int64)
    cg = engine.CandidateGroups.from_lists(counts, groups, dev)
    c = ranker.rank("transe", (ent, rel), to(q_h), to(q_t), to(q_r), 1, p_norm=1, normalize=False, groups=cg,
                    filt_csr=(to(fptr), to(fidx))).cpu().numpy()
    ranks = c[2].astype(np.int64) + c[3] // 2 + 1
    if verbose:
        lo = 0
        for name, n in zip(names, counts):
            rk = ranks[lo:lo + n].astype(np.float64)
            lo += n
            print("Relation: %s| Number %d | mrr: %.4f | hit1: %.4f | hit3: %.4f | hit10: %.4f " % (
                name, n, (1.0 / rk).mean(), (rk <= 1).mean(), (rk <= 3).mean(), (rk <= 10).mean()))
    rk = ranks.astype(np.float64)
    mrr = float((1.0 / rk).mean())
    hits = [float((rk <= k).mean()) for k in hits_at_k]
    if verbose:
        print(f"[Final Scores] MRR: {mrr} \t" + " \t".join(f"Hits@{k}: {h}" for k, h in zip(hits_at_k, hits)))
    if return_ranks:
        return (mrr, *hits, ranks)
    return (mrr, *hits)


def zsl_rank_metrics(score_lists):
    """ZSLmodule.eval's metric block: each entry is a similarity vector (higher is better) whose index 0 is the true
    candidate; rank = 1 + position of index 0 in the descending argsort (numpy's order for exact ties, as the
    reference).  Returns (hits10, hits5, mrr) -- the tuple ZSLmodule.eval returns -- plus hits1."""
    ranks = np.asarray([list(np.argsort(s))[::-1].index(0) + 1 for s in score_lists], np.float64)
    return float((ranks <= 10).mean()), float((ranks <= 5).mean()), float((1.0 / ranks).mean()), float((ranks <= 1).mean())


class ZSLEvaluator:
    """ZSLmodule.eval (module/zsl_module.py:635-745) over the library: the Extractor's per-entity halves are computed once
    (mre_zsl_entity_features), then every (head, candidate) pair of every test triple is scored and ranked in one sweep
    (mre_zsl_rank) instead of one Extractor forward per triple.

    extractor_state: the reference Extractor's state_dict (tensors or numpy), names symbol_emb.weight, gcn_w.*, fc1.*, fc2.*,
    reshape_layer.*, support_encoder.proj1.*, support_encoder.proj2.*, support_encoder.layer_norm.*;
    connections [num_ents, max_neighbor, 2] and e1_degrees as ZSLmodule.build_connection leaves them (:239-268);
    ent_symbol[e] = symbol2id of the entity whose ent2id is e."""

    STATE = (("symbol_emb", "symbol_emb.weight"), ("gcn_w", "gcn_w.weight"), ("gcn_b", "gcn_w.bias"), ("fc1_w", "fc1.weight"),
             ("fc1_b", "fc1.bias"), ("fc2_w", "fc2.weight"), ("fc2_b", "fc2.bias"), ("reshape_w", "reshape_layer.weight"),
             ("reshape_b", "reshape_layer.bias"), ("proj1_w", "support_encoder.proj1.weight"), ("proj1_b", "support_encoder.proj1.bias"),
             ("proj2_w", "support_encoder.proj2.weight"), ("proj2_b", "support_encoder.proj2.bias"),
             ("ln_g", "support_encoder.layer_norm.weight"), ("ln_b", "support_encoder.layer_norm.bias"))

    def __init__(self, extractor_state, connections, e1_degrees, ent_symbol, device=0):
        self.ctx = engine.Context(device)
        self.device = torch.device("cuda", device)
        to = lambda a, dt: torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a).to(self.device, dt).contiguous()
        self._w = {f: to(extractor_state[k], torch.float32) for f, k in self.STATE}
        self.D = int(self._w["symbol_emb"].shape[1])
        self.model = L.ZslModel(D=self.D, ln_eps=1e-5, **{f: t.data_ptr() for f, t in self._w.items()})
        conn = np.asarray(connections)
        self.conn = to(conn[:, :, 1] if conn.ndim == 3 else conn, torch.int64)              # neighbour symbols only (:51)
        n_ent = self.conn.shape[0]
        deg = e1_degrees
        if isinstance(deg, dict):
            deg = [deg.get(i, 0) for i in range(n_ent)]
        self.deg = to(deg, torch.float32)
        self.ent_symbol = to(ent_symbol, torch.int64)
        self.A = torch.empty((n_ent, self.D), dtype=torch.float32, device=self.device)
        self.B = torch.empty_like(self.A)
        L.check(L.lib().mre_zsl_entity_features(self.ctx._h, self.model, self.ent_symbol.data_ptr(), self.conn.data_ptr(),
                                                self.deg.data_ptr(), n_ent, int(self.conn.shape[1]), self.A.data_ptr(),
                                                self.B.data_ptr(), torch.cuda.current_stream().cuda_stream))

    def rank(self, q_head, q_rel, cand_lists, rel_vecs, want_scores=False):
        """q_head [T] entity ids, q_rel [T] rows of rel_vecs [n_rel, n_vec, D], cand_lists: T arrays of entity ids (true first)
        -> (counts int32 [4, T] on the device, scores float32 [P] or None)"""
        to = lambda a, dt: torch.as_tensor(np.asarray(a) if not isinstance(a, torch.Tensor) else a).to(self.device, dt).contiguous()
        T = len(cand_lists)
        ptr = np.concatenate([[0], np.cumsum([len(c) for c in cand_lists])]).astype(np.int64)
        P = int(ptr[-1])
        flat = np.concatenate([np.asarray(c, np.int64) for c in cand_lists]) if P else np.zeros(0, np.int64)
        d_ptr, d_idx, d_head, d_rel = to(ptr, torch.int64), to(flat, torch.int64), to(q_head, torch.int64), to(q_rel, torch.int64)
        rv = to(rel_vecs, torch.float32)
        assert rv.dim() == 3 and rv.shape[2] == self.D
        counts = torch.zeros((4, max(T, 1)), dtype=torch.int32, device=self.device)[:, :T].contiguous()
        scores = torch.empty(max(P, 1), dtype=torch.float32, device=self.device) if want_scores else None
        L.check(L.lib().mre_zsl_rank(self.ctx._h, self.model, self.A.data_ptr(), self.B.data_ptr(), self.A.shape[0], d_head.data_ptr(), d_rel.data_ptr(),
                                     d_ptr.data_ptr(), d_idx.data_ptr(), T, P, rv.data_ptr(), rv.shape[0], rv.shape[1],
                                     scores.data_ptr() if want_scores else None, counts.data_ptr(),
                                     torch.cuda.current_stream().cuda_stream))
        return counts, (scores[:P] if want_scores else None)

    def eval(self, test_candidates, relation_vecs, ent2id, mode="test", verbose=True, ties="pessimistic"):
        """test_candidates: {relation: {"head\trel\ttail": [true tail, candidates...]}} ({mode}_candidates.json, :647-649);
        relation_vecs: {relation: [test_sample, D] array} = generate_model.generate(...) per relation (:657-660).
        Returns (hits10, hits5, mrr) as ZSLmodule.eval does (:745) and prints its lines."""
        rels = list(test_candidates.keys())
        rv = np.stack([np.asarray(relation_vecs[r], np.float32) for r in rels]) if rels else np.zeros((0, 1, self.D), np.float32)
        q_head, q_rel, lists, per_rel = [], [], [], []
        for ri, rel in enumerate(rels):
            n0 = len(lists)
            for key, cands in test_candidates[rel].items():
                head = key.split("\t")[0]
                q_head.append(ent2id[head]); q_rel.append(ri)
                lists.append(np.fromiter((ent2id[c] for c in cands), np.int64, len(cands)))
            per_rel.append((rel, n0, len(lists)))
        counts, _ = self.rank(np.asarray(q_head, np.int64), np.asarray(q_rel, np.int64), lists, rv)
        c = counts.cpu().numpy()
        ranks = (c[0] + 1 + (c[1] if ties == "pessimistic" else 0)).astype(np.float64)
        if verbose:
            print("##EVALUATING ON %s DATA" % mode.upper())
            for rel, a, b in per_rel:
                r = ranks[a:b]
                if len(r):
                    print("{} Hits10:{:.3f}, Hits5:{:.3f}, Hits1:{:.3f} MRR:{:.3f}".format(
                        mode + rel, (r <= 10).mean(), (r <= 5).mean(), (r <= 1).mean(), (1.0 / r).mean()))
        if len(ranks) == 0:
            return float("nan"), float("nan"), float("nan")
        h10, h5, h1, mrr = (ranks <= 10).mean(), (ranks <= 5).mean(), (ranks <= 1).mean(), (1.0 / ranks).mean()
        if verbose:
            print("############   " + mode + "    #############")
            print("HITS10: {:.3f}".format(h10)); print("HITS5: {:.3f}".format(h5)); print("HITS1: {:.3f}".format(h1))
            print("MAP: {:.3f}".format(mrr)); print("###################################")
        return float(h10), float(h5), float(mrr)
