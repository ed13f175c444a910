"""BASELINE.json's full sizes (2 M entities x 256):
  * counts of sampled queries BIT-EXACT against the C oracle's sequential-FP32 scores of the same tables (oracle/kge_oracle.c:
    orc_transe_scores / orc_distmult_scores; ~0.7 s per query on one host core): raw = #(s_j < s_true), ties, and
    raw - filtered = #known entities scoring below s_true; Model.predict's materialised float32[E] equals the oracle's vector;
  * additivity over a partition of the candidate set: counts(all entities) = sum of the counts over disjoint candidate groups;
  * a planted duplicate of the true entity is an exact tie, never a strict win;
  * shard additivity of the integer metric sums (what the N-GPU all-reduce relies on) on the full FB15K237 protocol.
"""
import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import kge_oracle as ko

pytestmark = pytest.mark.gpu

E2M, D2M, R2M, Q2M = 2_000_000, 256, 1000, 384


@pytest.fixture(scope="module")
def big(mre):
    g = torch.Generator(device="cuda").manual_seed(192)
    ent = torch.randn(E2M, D2M, device="cuda", generator=g) / D2M ** 0.5
    rel = torch.randn(R2M, D2M, device="cuda", generator=g) / D2M ** 0.5
    rng = np.random.default_rng(5)
    q_h, q_t, q_r = rng.integers(0, E2M, Q2M), rng.integers(0, E2M, Q2M), rng.integers(0, R2M, Q2M)
    ent[E2M - 1] = ent[int(q_t[0])]                      # a duplicate of query 0's true tail: an exact tie
    lists = [np.unique(np.concatenate([[t], rng.integers(0, E2M, 1 + rng.geometric(1 / 2.5))])) for t in q_t.tolist()]
    fptr = np.concatenate([[0], np.cumsum([len(x) for x in lists])]).astype(np.int64)
    fidx = np.concatenate(lists)
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    return dict(ent=ent, rel=rel, q=(d(q_h), d(q_t), d(q_r)), q_host=(q_h, q_t, q_r), lists=lists, csr=(d(fptr), d(fidx)),
                rk=mre.engine.Ranker(device=0), eng=mre.engine, ent_host=ent.cpu().numpy(), rel_host=rel.cpu().numpy())


@pytest.mark.parametrize("scorer", ["transe", "distmult"])
def test_full_size_counts_match_own_scores_and_partition(big, scorer):
    rk, eng = big["rk"], big["eng"]
    tabs = (big["ent"], big["rel"])
    q_h, q_t, q_r = big["q"]
    c = rk.rank(scorer, tabs, q_h, q_t, q_r, 1, filt_csr=big["csr"]).cpu().numpy()
    assert np.all(c[2] <= c[0]) and np.all(c[3] < c[1])
    # (1) the C oracle's sequential-FP32 scores of the same tables for a sample of queries
    score_fn = ko.transe_scores if scorer == "transe" else ko.distmult_scores
    for q in (0, 1, Q2M // 2, Q2M - 1):
        h, t, r = (int(a[q]) for a in big["q_host"])
        s = score_fn(big["ent_host"], big["rel_host"], 1, 1, h, t, r) if scorer == "transe" else score_fn(big["ent_host"], big["rel_host"], 1, h, t, r)
        st = s[t]
        lt, eq = int((s < st).sum()), int((s == st).sum())    # raw_eq counts every candidate at the true score, the true one included
        assert (c[0][q], c[1][q]) == (lt, eq), (scorer, q)
        known = big["lists"][q]
        known = known[known != t]
        k_lt, k_eq = int((s[known] < st).sum()), int((s[known] == st).sum())
        assert (c[2][q], c[3][q]) == (lt - k_lt, eq - 1 - k_eq), (scorer, q)     # filtered: minus the known ones, minus itself
        if q in (0, Q2M - 1):                                                    # Model.predict's vector is the oracle's, bit for bit
            assert np.array_equal(rk.predict(scorer, tabs, q_h, q_t, q_r, 1, query=q).cpu().numpy(), s)
    # (2) the planted duplicate of query 0's true tail ties exactly
    assert c[3][0] >= 1
    # (3) additivity over a partition of the candidates into three ragged groups (every query ranks against each part)
    cuts = [0, 700_001, 1_333_333, E2M]
    total = np.zeros_like(c)
    qptr = np.array([0, Q2M], np.int64)
    for a, b in zip(cuts[:-1], cuts[1:]):
        cand = torch.arange(a, b, device="cuda", dtype=torch.int64)
        groups = eng.CandidateGroups(qptr, np.array([0, b - a], np.int64), cand)
        total += rk.rank(scorer, tabs, q_h, q_t, q_r, 1, groups=groups, filt_csr=big["csr"]).cpu().numpy()
    # the true entity lies in exactly one part; ties with itself are never counted, so the parts add up exactly
    assert np.array_equal(total, c)


def test_full_protocol_shard_additivity(mre, fb15k237):
    """FB15K237, all 2 x 20 466 queries (the OpenKE protocol): the integer metric sums of 3 contiguous shards add up to the
    sums of the whole run -- the property the multi-GPU all-reduce relies on -- and MRR from the rank histogram equals the
    float64 reciprocal-rank sum"""
    eng = mre.engine
    E, R, D = fb15k237.E, fb15k237.R, 200
    ent, rel = (torch.from_numpy(t).cuda() for t in gu.xavier_tables(gu.SEED, [(E, D), (R, D)]))
    ix = eng.KGIndex.from_arrays(E, R, fb15k237.train, fb15k237.valid, fb15k237.test).to_device(0)
    th, tt, tr = ix.test_triples()
    d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    q_h, q_t, q_r = d(np.repeat(th, 2)), d(np.repeat(tt, 2)), d(np.repeat(tr, 2))
    side = d(np.tile(np.array([0, 1], np.uint8), len(th)))
    rk = eng.Ranker(device=0)
    whole = rk.metrics(rk.rank("transe", (ent, rel), q_h, q_t, q_r, side, index=ix, normalize=True), side, "strict", hist_len=E + 2)
    sums = torch.zeros_like(whole["sums"]); rr = torch.zeros_like(whole["rr"]); hist = torch.zeros_like(whole["hist"])
    Q = q_h.numel()
    for rank in range(3):
        lo, hi = mre.dist.DistContext.shard_of(Q, rank, 3)
        sl = slice(lo, hi)
        part = rk.metrics(rk.rank("transe", (ent, rel), q_h[sl].contiguous(), q_t[sl].contiguous(), q_r[sl].contiguous(),
                                  side[sl].contiguous(), index=ix, normalize=True), side[sl].contiguous(), "strict", hist_len=E + 2)
        sums += part["sums"]; rr += part["rr"]; hist += part["hist"]
    assert torch.equal(sums, whole["sums"]) and torch.equal(hist, whole["hist"])
    assert torch.allclose(rr, whole["rr"], rtol=1e-12, atol=0)
    m = mre.dist.metrics_from_hist(whole["hist"])
    assert m["n"] == Q and np.isclose(m["mrr"], float(whole["rr"].sum().item()) / Q, rtol=1e-12)
