"""The paper-side (main.py / module/) entry points of the hot path, mirrored over the library:
  evaluate(...)          main.evaluate (main.py:217-272): TransE-L1 candidate ranking, ties//2, MRR + Hits@1/3/10
  PaperScorer            module.NegativeSampling._calc / .evaluate (module/NegativeSampling.py:142-168, 294-305)
  build_test_candidates  utils/gen_mode_candidates.py:15-39 (regenerates the missing {mode}_candidates.json)
  zsl_rank_metrics       ZSLmodule.eval's rank/metric block (module/zsl_module.py:699-745): Hits@10/5/1 + MRR
"""
import numpy as np
import torch

from . import _lib as L
from . import engine


class PaperScorer:
    """score_model 'transe': ||(h + r) - t||_1 without normalisation (score_norm_flag False :31, p_norm 1 :47);
    'distmult': sum h*r*t.  Inputs are embedding ROWS ([n, D] tensors) as in the reference's _calc(h, t, r)."""

    def __init__(self, p_norm=1, score_norm_flag=False, device=0):
        self.p_norm, self.norm_flag = p_norm, score_norm_flag
        self.ctx = engine.Context(device)
        self.device = torch.device("cuda", device)

    def _calc(self, h, t, r, mode="normal", score_model="transe"):
        n = max(h.shape[0], t.shape[0], r.shape[0])
        D = h.shape[-1]
        rows = [x.reshape(-1, D).to(self.device, torch.float32) for x in (h, t, r)]
        rows = [x if x.shape[0] == n else x.repeat(n // x.shape[0], 1) for x in rows]
        ent = torch.cat(rows[:2]).contiguous()          # [2n, D]: heads then tails
        rel = rows[2].contiguous()
        idx = torch.arange(n, device=self.device)
        out = torch.empty(n, dtype=torch.float32, device=self.device)
        scorer = {"transe": L.TRANSE, "distmult": L.DISTMULT}[score_model]
        L.check(L.lib().mre_score_triples(self.ctx._h, scorer, ent.data_ptr(), None, rel.data_ptr(), None, D, idx.data_ptr(),
                                          (idx + n).data_ptr(), idx.data_ptr(), n, self.p_norm, int(self.norm_flag), out.data_ptr(),
                                          torch.cuda.current_stream().cuda_stream))
        return out

    def evaluate(self, h, r, t, score_model="transe"):
        if score_model != "transe":
            print("invalid scoring model!")
            return None
        return self._calc(h, t, r)


def build_test_candidates(triples, rel2candidates, e1rel_e2):
    """triples: iterable of (head, rel, tail) symbols -> {rel: {"head\\trel\\ttail": [tail, cand, ...]}} with the true tail
    first, then every candidate of the relation that is neither a known tail of (head, rel) nor the true tail."""
    out = {}
    for head, rel, tail in triples:
        known = set(e1rel_e2.get(head + rel, ()))
        cands = [tail] + [c for c in rel2candidates[rel] if c not in known and c != tail]
        out.setdefault(rel, {})["\t".join((head, rel, tail))] = cands
    return out


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


def evaluate(ent_embs, rel_embs, e2id, r2id, test_candidates, hits_at_k=(1, 3, 10), ranker=None, verbose=True):
    """main.evaluate: ranks the true tail of every test triple among its candidate list with the TransE-L1 scorer,
    rank = #(n < p) + #(n == p) // 2 + 1.  Returns (mrr, hits@k...) over all triples; prints the reference's lines."""
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
    return (mrr, *hits)


def zsl_rank_metrics(score_lists):
    """ZSLmodule.eval's metric block: each entry is a similarity vector (higher is better) whose index 0 is the true
    candidate; rank = 1 + position of index 0 in the descending argsort (numpy's order for exact ties, as the
    reference).  Returns (hits10, hits5, mrr) -- the tuple ZSLmodule.eval returns -- plus hits1."""
    ranks = np.asarray([list(np.argsort(s))[::-1].index(0) + 1 for s in score_lists], np.float64)
    return float((ranks <= 10).mean()), float((ranks <= 5).mean()), float((1.0 / ranks).mean()), float((ranks <= 1).mean())
