"""numpy restatement of the paper-side (main.py / module/ / utils/) ranking code -- TEST INFRASTRUCTURE ONLY.

The paper half of the reference cannot be imported here or anywhere offline (module.vqgan is absent,
torch_geometric / ml_collections / skimage are not installed, the data blobs are missing), and
the reference holds no test or golden vector for it: PARITY UNPINNED for this file beyond the checks
tests/ run between it and oracle/openke_torch.py (whose tensor expressions ARE pinned to the reference's
OpenKE modules, and main.py's scorer is the same expression).  Pure-python loops: small cases only.
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
