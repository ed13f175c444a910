"""CPU restatement of the paper-side (main.py / module/ / utils/) scorer, ranker and sampler -- TEST INFRASTRUCTURE ONLY.

Pinned to the reference's own class: tests/golden/make_golden_paper.py imports /root/reference/module/NegativeSampling.py
(with the two modules it cannot import here, module.model and module.vqgan, stubbed), runs its _calc / evaluate /
neg_sample_fn and main.py's evaluate() on seeded inputs, and asserts that THIS file reproduces every output bit for bit
(scores: same torch CPU expressions; sampler: same control flow on the same Mersenne-Twister stream under random.seed(k);
main.evaluate: same ranks, same printed summary).  The outputs are committed as tests/golden/golden_paper.npz, and
tests/test_paper_golden.py re-checks this file against them on any box.  Pure-python loops: small cases only.
"""
import numpy as np


def known_tails(h, r, t):
    """e1rel_e2: (head, relation) -> list of tails, utils/gen_e1r_e2_all.py:14-19 (ids instead of strings)."""
    out = {}
    for hh, rr, tt in zip(h, r, t):
        out.setdefault((int(hh), int(rr)), []).append(int(tt))
    return out


def build_candidates(h, r, t, rel2candidates, e1rel_e2):
    """Per-triple candidate lists, utils/gen_mode_candidates.py:15-39: the true tail first, then every
    entity of rel2candidates[rel] that is neither a known tail of (h, rel) nor the true tail."""
    out = []
    for hh, rr, tt in zip(h, r, t):
        known = set(e1rel_e2.get((int(hh), int(rr)), ()))
        cands = [int(tt)] + [int(e) for e in rel2candidates[int(rr)] if int(e) not in known and int(e) != int(tt)]
        out.append(np.asarray(cands, np.int64))
    return out


def paper_transe_scores(ent, rel, h, r, cands):
    """model.evaluate(h, r, t): module/NegativeSampling.py:294-305 -- ||(h + r) - t||_1, no normalisation
    (score_norm_flag False :31, p_norm 1 :47); float32, numpy's pairwise summation order."""
    u = (ent[h].astype(np.float32) + rel[r].astype(np.float32))[None, :] - ent[cands].astype(np.float32)
    return np.abs(u).sum(-1, dtype=np.float32)


def torch_calc(h, t, r, mode="normal", score_model="transe", score_norm_flag=False, p_norm=1):
    """NegativeSampling._calc (module/NegativeSampling.py:142-168) on torch CPU tensors: the reference's expression, hence
    its bits (F.normalize rows when score_norm_flag :144-147, the view for 1-vs-all modes :148-151, h + (r - t) for
    head_batch else (h + r) - t :152-155, torch.norm(p) :156; distmult h * (r * t) / (h * r) * t, torch.sum :158-168)."""
    import torch
    import torch.nn.functional as F
    h, t, r = (torch.as_tensor(x) for x in (h, t, r))
    if score_model == "transe" and score_norm_flag:
        h, r, t = F.normalize(h, 2, -1), F.normalize(r, 2, -1), F.normalize(t, 2, -1)
    if mode != "normal":
        B, D = r.shape[0], r.shape[-1]
        h, t, r = h.view(-1, B, D), t.view(-1, B, D), r.view(-1, B, D)
    if score_model == "transe":
        u = h + (r - t) if mode == "head_batch" else (h + r) - t
        return torch.norm(u, p_norm, -1).flatten()
    u = h * (r * t) if mode == "head_batch" else (h * r) * t
    return torch.sum(u, -1).flatten()


def torch_evaluate(h, r, t, score_norm_flag=False, p_norm=1):
    """NegativeSampling.evaluate (module/NegativeSampling.py:294-305): ||(h + r) - t||_p on rows"""
    return torch_calc(h, t, r, "normal", "transe", score_norm_flag, p_norm)


def _mrr_f32(ranks):
    """sum([1.0 / rank ...]) / len(ranks) as main.py:252,263 evaluate it: `rank` is a 0-dim torch LONG tensor there, so
    1.0 / rank is a float32 tensor and Python's sum() accumulates sequentially in float32"""
    acc = np.float32(0.0)
    for x in ranks:
        acc = np.float32(acc + np.float32(1.0) / np.float32(x))
    return float(np.float32(acc / np.float32(len(ranks))))


def evaluate_candidates(ent, rel, e2id, r2id, test_candidates, hits_at_k=(1, 3, 10)):
    """main.evaluate (main.py:217-272) on a {mode}_candidates.json dict: per test triple the head / relation rows repeated
    against the candidates' rows (:236-244), model.evaluate (:245), rank = #(n < p) + #(n == p) // 2 + 1 (:246-250).
    Returns (ranks in file order, per-query score vectors, per-relation (name, n, mrr, hit1, hit3, hit10), (mrr, hits...))."""
    import torch
    ent, rel = torch.as_tensor(ent), torch.as_tensor(rel)
    ranks, scores, per_rel = [], [], []
    for query, items in test_candidates.items():
        tmp = []
        for key, cands in items.items():
            head, rela, _ = key.split("\t")
            hrow = ent[e2id[head]].repeat(len(cands), 1)
            rrow = rel[r2id[rela]].repeat(len(cands), 1)
            trow = torch.stack([ent[e2id[c]] for c in cands])
            s = torch_evaluate(hrow, rrow, trow).numpy()
            scores.append(s)
            tmp.append(rank_ties_half(s))
        ranks.extend(tmp)
        per_rel.append((query, len(items), _mrr_f32(tmp)) + tuple(
            sum(1.0 if x <= k else 0.0 for x in tmp) / len(tmp) for k in hits_at_k))
    mrr = _mrr_f32(ranks)
    hits = [sum(1.0 if x <= k else 0.0 for x in ranks) / len(ranks) for k in hits_at_k]
    return ranks, scores, per_rel, (mrr, *hits)


def rank_ties_half(scores):
    """main.py:245-250: p = scores[0]; rank = #(n < p) + #(n == p) // 2 + 1."""
    p, n = scores[0], scores[1:]
    return int((n < p).sum()) + int((n == p).sum()) // 2 + 1


def rank_argsort_desc(scores):
    """module/zsl_module.py:705-706: 1 + position of candidate 0 in the descending argsort."""
    order = list(np.argsort(scores))[::-1]
    return order.index(0) + 1


def summarize(ranks, ks):
    """main.py:263-266 / zsl_module.py:739-745: MRR and Hits@k as plain means over all test triples."""
    ranks = np.asarray(ranks, np.float64)
    return float((1.0 / ranks).mean()), [float((ranks <= k).mean()) for k in ks]


class ReferenceSubgraphSampler:
    """Restatement of module/NegativeSampling.py's sampler with the reference's own control flow and Python `random`
    (neg_sample_fn :114-140, __normal_batch :321-349, __corrupt_head/__corrupt_tail :351-375, __count_htr :59-93), ids as
    ints.  With rng = random.Random(k) it reproduces the reference class under random.seed(k) bit for bit
    (tests/golden/make_golden_paper.py asserts it; golden_paper.npz holds the reference's vectors).  The product's sampler
    draws from Philox instead of Python's Mersenne Twister, so IT is compared with this one statistically (head/tail split,
    uniformity over the admissible nodes, layout, no leaks) and bit for bit with its own Philox CPU replay."""

    def __init__(self, whole_triples, neg_ent=1, filter_flag=True, rng=None):
        import random
        self.random = rng or random.Random(0)
        self.neg_ent, self.filter_flag = neg_ent, filter_flag
        self.h_of_tr, self.t_of_hr = {}, {}
        h, r, t = whole_triples
        for hh, tt, rr in zip(h, t, r):
            self.t_of_hr.setdefault((int(hh), int(rr)), []).append(int(tt))
            self.h_of_tr.setdefault((int(tt), int(rr)), []).append(int(hh))
        self.h_of_tr = {k: np.array(list(set(v))) for k, v in self.h_of_tr.items()}
        self.t_of_hr = {k: np.array(list(set(v))) for k, v in self.t_of_hr.items()}

    def _corrupt(self, table, key, local_global_id, node_list, num_max):
        try:
            tmp = np.asarray(self.random.sample(list(node_list), k=num_max), np.int64)
        except ValueError:
            tmp = np.asarray(self.random.sample(list(node_list), k=len(node_list)), np.int64)
        if not self.filter_flag:
            return tmp
        compare = np.asarray([local_global_id[int(x)] for x in tmp], np.int64)
        mask = np.isin(compare, table.get(key, np.zeros(0, np.int64)), invert=True)
        return tmp[mask]

    def _normal_batch(self, local_global_id, node_list, h, t, r, neg_size):
        nh = sum(1 for _ in range(neg_size) if self.random.random() < 0.5)
        nt = neg_size - nh
        out = []
        for table, key, want in ((self.h_of_tr, (local_global_id[t], r), nh), (self.t_of_hr, (local_global_id[h], r), nt)):
            got, cur = [], 0
            while cur < want:
                tmp = self._corrupt(table, key, local_global_id, node_list, (want - cur) * 2)
                got.append(tmp)
                cur += len(tmp)
            out.append(np.concatenate(got)[:want] if got else np.zeros(0, np.int64))
        return out

    def neg_sample_fn(self, local_global_id, node_list, edge_index, edge_type):
        bh, bt, br = (np.asarray(x, np.int64) for x in (edge_index[0], edge_index[1], edge_type))
        hs, ts, rs = (np.repeat(x.reshape(-1, 1), 1 + self.neg_ent, axis=-1) for x in (bh, bt, br))
        for i, (h, t, r) in enumerate(zip(bh.tolist(), bt.tolist(), br.tolist())):
            last = 1
            neg_head, neg_tail = self._normal_batch(local_global_id, node_list, h, t, r, self.neg_ent)
            hs[i][last:last + len(neg_head)] = neg_head
            last += len(neg_head)
            ts[i][last:last + len(neg_tail)] = neg_tail
        return np.stack([hs.T.flatten(), ts.T.flatten()]).astype(np.int32), rs.T.flatten().astype(np.int32)
