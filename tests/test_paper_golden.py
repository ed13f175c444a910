"""The paper half (main.py / module/NegativeSampling.py) against tests/golden/golden_paper.npz -- outputs of the REFERENCE's own
class and of main.evaluate compiled from the reference's source (tests/golden/make_golden_paper.py):

  not gpu: oracle/paper_oracle.py reproduces every golden vector BIT FOR BIT (scores, sampler under random.seed(k), ranks and
           the printed summary of main.evaluate) -- the pin of the paper-side oracle;
  gpu:     the CUDA path (paper.PaperScorer / paper.NegativeSampling / paper.evaluate, all through the C ABI) against the same
           vectors: scores within 1e-6 of the score scale, gradients within 5e-5 of the gradient scale, main.evaluate's ranks
           EXACT (every golden query has an empty 1e-5 tie band; exact ties are planted), MRR/Hits within 1e-4, the sampler's
           layout / positives / filter / head-tail split against the reference's vectors.
"""
import random

import numpy as np
import pytest
import torch

import golden_util as gu
from oracle import paper_oracle as po


@pytest.fixture(scope="module")
def g():
    return gu.load("golden_paper.npz")


# ------------------------------------------------------------------------------------------------ oracle pin (CPU)
@pytest.mark.parametrize("case", gu.PAPER_CALC_CASES, ids=[c[0] for c in gu.PAPER_CALC_CASES])
def test_oracle_calc_bit_exact_vs_reference(g, case):
    name, score_model, mode, norm, B, k = case
    h, t, r = (torch.from_numpy(x) for x in gu.paper_calc_inputs(gu.SEED, name))
    mine = po.torch_calc(h, t, r, mode, score_model, norm).numpy()
    assert mine.shape == (B * k,) and np.array_equal(mine, g["calc_" + name])
    if score_model == "transe" and mode == "normal":
        assert np.array_equal(po.torch_evaluate(h, r, t, norm).numpy(), g["evaluate_" + name])


def test_oracle_sampler_bit_exact_vs_reference_under_seed(g):
    whole, E, R, cases = gu.paper_subgraph_cases(gu.SEED)
    for ci, c in enumerate(cases):
        for filt in (True, False):
            smp = po.ReferenceSubgraphSampler(whole, neg_ent=c["neg_ent"], filter_flag=filt, rng=random.Random(c["py_seed"]))
            l2g = {i: int(x) for i, x in enumerate(c["l2g"])}
            ei = np.stack([c["edge_h"], c["edge_t"]])
            nodes = np.arange(int(ei.max()))
            mi, mt = smp.neg_sample_fn(l2g, nodes, ei, c["edge_r"])
            tag = f"samp{ci}_{'f' if filt else 'nf'}"
            assert np.array_equal(mi, g[tag + "_ei"]) and np.array_equal(mt, g[tag + "_et"])


def test_oracle_main_evaluate_bit_exact_vs_reference(g):
    ents, rels, e2id, r2id, ent, rel, cand = gu.paper_eval_setup(gu.SEED)
    ranks, scores, per_rel, final = po.evaluate_candidates(ent, rel, e2id, r2id, cand)
    assert np.array_equal(np.asarray(ranks), g["eval_ranks"])
    assert np.array_equal(np.concatenate(scores), g["eval_scores"])
    assert np.array_equal(np.asarray(final), g["eval_final"])
    lines = ["Relation: %s| Number %d | mrr: %.4f | hit1: %.4f | hit3: %.4f | hit10: %.4f " % p for p in per_rel]
    assert lines == [str(x) for x in g["eval_lines"]]
    assert int(g["eval_band"].max()) == 0 and int((g["eval_ties"] > 0).sum()) >= 10      # what the GPU test relies on


# ------------------------------------------------------------------------------------------------ CUDA path (GPU)
def _scale(case, h, t, r):
    """sum of |terms| per row: what a float32 summation-order difference is relative to"""
    name, score_model, mode, norm, B, k = case
    n = max(len(h), len(t))
    rep = lambda x: x if len(x) == n else np.tile(x, (n // len(x), 1))
    h, t, r = rep(h), rep(t), rep(r)
    if norm:
        nz = lambda x: x / np.maximum(np.linalg.norm(x, axis=1, keepdims=True), 1e-12)
        h, t, r = nz(h), nz(t), nz(r)
    u = (h + r - t) if score_model == "transe" else h * r * t
    return np.abs(u).sum(1)


@pytest.mark.gpu
@pytest.mark.parametrize("case", gu.PAPER_CALC_CASES, ids=[c[0] for c in gu.PAPER_CALC_CASES])
def test_cuda_calc_vs_reference(mre, g, case):
    name, score_model, mode, norm, B, k = case
    h, t, r = gu.paper_calc_inputs(gu.SEED, name)
    sc = mre.paper.PaperScorer(p_norm=1, score_norm_flag=norm, device=0)
    got = sc._calc(torch.from_numpy(h).cuda(), torch.from_numpy(t).cuda(), torch.from_numpy(r).cuda(), mode=mode,
                   score_model=score_model).cpu().numpy()
    want = g["calc_" + name]
    assert got.shape == want.shape
    assert np.all(np.abs(got - want) <= 1e-6 * _scale(case, h, t, r) + 1e-7), float(np.abs(got - want).max())
    if score_model == "transe" and mode == "normal":
        ev = sc.evaluate(torch.from_numpy(h).cuda(), torch.from_numpy(r).cuda(), torch.from_numpy(t).cuda()).cpu().numpy()
        assert np.array_equal(ev, got)
        assert sc.evaluate(torch.from_numpy(h).cuda(), torch.from_numpy(r).cuda(), torch.from_numpy(t).cuda(), score_model="distmult") is None


@pytest.mark.gpu
@pytest.mark.parametrize("case", [c for c in gu.PAPER_CALC_CASES if c[2] == "normal"], ids=lambda c: c[0])
def test_cuda_calc_gradients_vs_autograd_on_reference_expression(mre, case):
    """_calc is differentiable in the reference (it trains the RGCN / M3AE outputs): d(sum w_i score_i)/d rows against torch
    autograd on the pinned reference expression"""
    name, score_model, mode, norm, B, k = case
    h, t, r = gu.paper_calc_inputs(gu.SEED, name)
    w = np.random.default_rng(5).standard_normal(B).astype(np.float32)
    ref = [torch.from_numpy(x).requires_grad_() for x in (h, t, r)]
    (po.torch_calc(ref[0], ref[1], ref[2], mode, score_model, norm) * torch.from_numpy(w)).sum().backward()
    sc = mre.paper.PaperScorer(p_norm=1, score_norm_flag=norm, device=0)
    mine = [torch.from_numpy(x).cuda().requires_grad_() for x in (h, t, r)]
    (sc._calc(mine[0], mine[1], mine[2], mode=mode, score_model=score_model) * torch.from_numpy(w).cuda()).sum().backward()
    for a, b in zip(mine, ref):
        scale = float(b.grad.abs().max())
        assert float((a.grad.cpu() - b.grad).abs().max()) <= 5e-5 * scale


@pytest.mark.gpu
def test_cuda_main_evaluate_vs_reference(mre, g, capsys):
    ents, rels, e2id, r2id, ent, rel, cand = gu.paper_eval_setup(gu.SEED)
    out = mre.paper.evaluate(torch.from_numpy(ent).cuda(), torch.from_numpy(rel).cuda(), e2id, r2id, cand, return_ranks=True)
    mrr, h1, h3, h10, ranks = out
    assert np.array_equal(np.asarray(ranks), g["eval_ranks"])               # every tie band is empty; exact ties are exact
    ref = g["eval_final"]
    assert abs(mrr - ref[0]) <= 1e-6 and (h1, h3, h10) == tuple(ref[1:])     # the reference sums 1/rank in float32
    printed = [ln for ln in capsys.readouterr().out.splitlines() if ln.startswith("Relation: ")]
    assert len(printed) == len(g["eval_lines"])
    for mine, theirs, row in zip(printed, g["eval_lines"], g["eval_per_rel"]):
        assert mine.split("| mrr")[0] == str(theirs).split("| mrr")[0]       # relation name and triple count
        nums = [float(x.split(":")[1]) for x in mine.split("|")[2:]]
        assert np.allclose(nums, row, atol=1.01e-4)                          # %.4f of float64 vs float32 sums


@pytest.mark.gpu
def test_cuda_sampler_vs_reference_vectors(mre, g):
    """the GPU sampler draws from Philox, the reference from Python's Mersenne Twister: against the reference's vectors the
    deterministic part must be EQUAL (layout, positive block, relation tiling, one corrupted side per negative, heads before
    tails, ids inside the node list, no train triple when filtering) and the random part statistically alike (head/tail split)"""
    whole, E, R, cases = gu.paper_subgraph_cases(gu.SEED)
    th, tr, tt = (np.asarray(x, np.int64) for x in whole)
    known = set(zip(th.tolist(), tr.tolist(), tt.tolist()))
    tot_ref = tot_mine = tot = 0
    for ci, c in enumerate(cases):
        ref_ei, ref_et = g[f"samp{ci}_f_ei"], g[f"samp{ci}_f_et"]
        ns = mre.paper.NegativeSampling(whole_triples=(th, tr, tt), neg_ent=c["neg_ent"], filter_flag=True, num_entities=E, num_relations=R,
                                        seed=gu.SEED + ci, device=0)
        l2g = {i: int(x) for i, x in enumerate(c["l2g"])}
        ei_t = torch.from_numpy(np.stack([c["edge_h"], c["edge_t"]]))
        nodes = torch.arange(int(ei_t.max()))
        ei, et = ns.neg_sample_fn(l2g, nodes, ei_t, torch.from_numpy(c["edge_r"]))
        ei, et = ei.numpy(), et.numpy()
        n, neg = len(c["edge_h"]), c["neg_ent"]
        assert ei.shape == ref_ei.shape and ei.dtype == ref_ei.dtype and et.dtype == ref_et.dtype
        assert np.array_equal(et, ref_et)                                                    # relation tiling
        assert np.array_equal(ei[:, :n], ref_ei[:, :n])                                      # the positive block
        for arr in (ei, ref_ei):
            hs, ts = arr[0].reshape(1 + neg, n), arr[1].reshape(1 + neg, n)
            ch, ct = hs[1:] != hs[0], ts[1:] != ts[0]
            assert not np.any(ch & ct)
            for b in range(n):
                if ch[:, b].any() and ct[:, b].any():
                    assert np.max(np.nonzero(ch[:, b])[0]) < np.min(np.nonzero(ct[:, b])[0])
            assert arr.max() <= max(int(nodes.max()), int(ei_t.max()))
            for o in range(n, arr.shape[1]):                                                 # zero leaks, in both
                b = o % n
                if arr[0, o] != arr[0, b] or arr[1, o] != arr[1, b]:
                    assert (l2g[int(arr[0, o])], int(et[o]), l2g[int(arr[1, o])]) not in known
        tot_ref += int((ref_ei[0].reshape(1 + neg, n)[1:] != ref_ei[0, :n]).sum())
        tot_mine += int((ei[0].reshape(1 + neg, n)[1:] != ei[0, :n]).sum())
        tot += n * neg
    # head corruptions ~ Binomial(tot, 1/2) minus the (rare) draws that return the original id: both within 4 sigma of tot/2
    sigma = 0.5 * np.sqrt(tot)
    assert abs(tot_ref - tot / 2) < 4 * sigma + 0.03 * tot and abs(tot_mine - tot / 2) < 4 * sigma + 0.03 * tot


@pytest.mark.gpu
def test_struct_loss_backpropagates_to_features(mre):
    """NegativeSampling.struct_loss must carry gradient to x and rel_emb through the margin term (regul_rate = 0
    leaves no other path): compare with autograd on the reference expressions over the SAME sampled negatives"""
    whole, E, R, cases = gu.paper_subgraph_cases(gu.SEED)
    th, tr, tt = (np.asarray(x, np.int64) for x in whole)
    c = cases[1]
    loss_fn = mre.openke.module.loss.MarginLoss(margin=3.0)                                  # main.py:70
    ns = mre.paper.NegativeSampling(whole_triples=(th, tr, tt), loss_fn=loss_fn, regul_rate=0.0, neg_ent=c["neg_ent"], num_entities=E,
                                    num_relations=R, seed=3, device=0)
    rng = np.random.default_rng(1)
    n_local, n = len(c["l2g"]), len(c["edge_h"])
    x0 = (rng.standard_normal((n_local, 200)) * 0.3).astype(np.float32)
    r0 = (rng.standard_normal((n, 200)) * 0.3).astype(np.float32)
    x, rel = torch.from_numpy(x0).cuda().requires_grad_(), torch.from_numpy(r0).cuda().requires_grad_()
    l2g = {i: int(v) for i, v in enumerate(c["l2g"])}
    ei_t = torch.from_numpy(np.stack([c["edge_h"], c["edge_t"]]))
    et_t = torch.from_numpy(c["edge_r"])
    loss = ns.struct_loss(l2g, x, rel, ei_t, et_t)
    loss.backward()
    assert x.grad is not None and float(x.grad.abs().max()) > 0 and float(rel.grad.abs().max()) > 0
    # the same draw again (calls counter rewound), scored with the pinned torch expressions on the CPU
    ns.calls -= 1
    ei, et = ns.neg_sample_fn(l2g, torch.arange(int(ei_t.max())), ei_t, et_t)
    xr, rr = torch.from_numpy(x0).requires_grad_(), torch.from_numpy(r0).requires_grad_()
    rel_expand = rr.repeat(1 + c["neg_ent"], 1)
    score = po.torch_calc(xr[ei[0].long()], xr[ei[1].long()], rel_expand)
    p, nn_ = score[:n].view(-1, n).permute(1, 0), score[n:].view(-1, n).permute(1, 0)
    want = torch.max(p - nn_, torch.tensor(-3.0)).mean() + 3.0                               # module/loss.py:20-24
    want.backward()
    assert np.isclose(loss.item(), want.item(), rtol=1e-5)
    for a, b in ((x.grad, xr.grad), (rel.grad, rr.grad)):
        assert float((a.cpu() - b).abs().max()) <= 5e-5 * float(b.abs().max())
