"""The paper's subgraph negative sampler (module/NegativeSampling.py:114-140, 321-375 -> mre_sample_subgraph).

The reference draws with Python's unseeded `random`, so parity is (a) bit-exact between the kernel and the CPU replay of
the same Philox stream (oracle/kge_oracle.c:orc_sample_subgraph_philox), (b) distributional between that replay and
oracle/paper_oracle.py:ReferenceSubgraphSampler -- the reference's own control flow (random.sample + np.in1d retry loop):
layout, head/tail split ~ Binomial(neg, 1/2) with heads first, uniform over the admissible nodes, zero train leaks.
"""
import random

import numpy as np
import pytest
import torch

import helpers
from oracle import paper_oracle as po


def make_subgraph(ds, seed, n_local, n_edges):
    """a subgraph in LOCAL ids: nodes = a random subset of entities; edges = train triples inside it (+ random ones)"""
    rng = np.random.default_rng(seed)
    l2g = rng.choice(ds.E, n_local, replace=False).astype(np.int64)
    g2l = {int(g): i for i, g in enumerate(l2g)}
    th, tt, tr = ds.oracle.train_triples()
    inside = [i for i in range(len(th)) if int(th[i]) in g2l and int(tt[i]) in g2l]
    pick = rng.choice(inside, min(n_edges // 2, len(inside)), replace=False) if inside else np.zeros(0, np.int64)
    eh = [g2l[int(th[i])] for i in pick] + rng.integers(0, n_local, n_edges - len(pick)).tolist()
    et = [g2l[int(tt[i])] for i in pick] + rng.integers(0, n_local, n_edges - len(pick)).tolist()
    er = [int(tr[i]) for i in pick] + rng.integers(0, ds.R, n_edges - len(pick)).tolist()
    return l2g, np.asarray(eh, np.int64), np.asarray(et, np.int64), np.asarray(er, np.int64)


def leaks(ds, l2g, ei, etype, n_edges):
    """negatives (slots >= 1) whose corrupted side completes a train triple"""
    bad = 0
    for o in range(n_edges, ei.shape[1]):
        b = o % n_edges
        h, t, r = int(ei[0, o]), int(ei[1, o]), int(etype[o])
        changed = (h != int(ei[0, b])) or (t != int(ei[1, b]))
        if changed and ds.oracle.count_train_leaks(np.array([l2g[h]]), np.array([l2g[t]]), np.array([r]), 0, 1):
            bad += 1
    return bad


@pytest.fixture(scope="module")
def dense():
    """a small, dense graph: known-answer sets cover a large share of the node list, so the filter matters"""
    return helpers.synthetic_graph(11, 60, 3, 1500, 0, 0)


def test_replay_layout_and_no_leaks(dense):
    l2g, eh, et, er = make_subgraph(dense, 1, 40, 200)
    nodes = np.arange(39, dtype=np.int64)          # arange(max(edge_index)): the reference leaves the max id out (:210)
    neg = 10
    ei, etype = dense.oracle.sample_subgraph_philox(192, 0, eh, et, er, nodes, l2g, neg)
    n = len(eh)
    assert ei.shape == (2, n * (1 + neg)) and ei.dtype == np.int32 and etype.dtype == np.int32
    assert np.array_equal(ei[0, :n], eh) and np.array_equal(ei[1, :n], et)
    assert np.array_equal(etype, np.tile(er, 1 + neg))
    hs, ts = ei[0].reshape(1 + neg, n), ei[1].reshape(1 + neg, n)
    # every negative changes at most one side; heads first, then tails (slots 1..nh, nh+1..neg)
    for b in range(n):
        ch = hs[1:, b] != eh[b]
        ct = ts[1:, b] != et[b]
        assert not np.any(ch & ct)
        if ch.any() and ct.any():
            assert np.max(np.nonzero(ch)[0]) < np.min(np.nonzero(ct)[0])
    assert hs[1:][hs[1:] != eh[None, :]].max() <= nodes.max() and ts[1:][ts[1:] != et[None, :]].max() <= nodes.max()
    assert leaks(dense, l2g, ei, etype, n) == 0
    # without the filter the same stream does emit known answers on this dense graph
    ei_nf, _ = dense.oracle.sample_subgraph_philox(192, 0, eh, et, er, nodes, l2g, neg, filt=0)
    assert leaks(dense, l2g, ei_nf, etype, n) > 0
    # a different step is a different draw; the same step is the same draw
    again, _ = dense.oracle.sample_subgraph_philox(192, 0, eh, et, er, nodes, l2g, neg)
    other, _ = dense.oracle.sample_subgraph_philox(192, 1, eh, et, er, nodes, l2g, neg)
    assert np.array_equal(again, ei) and not np.array_equal(other, ei)


def test_replay_distribution_matches_reference_sampler(dense):
    """the same edge sampled many times: head/tail split and the per-node frequencies agree with the reference's loop"""
    l2g, eh, et, er = make_subgraph(dense, 2, 30, 40)
    nodes = np.arange(29, dtype=np.int64)
    neg, reps = 10, 300
    th, tt, tr = dense.oracle.train_triples()
    ref = po.ReferenceSubgraphSampler((th, tr, tt), neg_ent=neg, rng=random.Random(5))
    l2g_dict = {i: int(g) for i, g in enumerate(l2g)}
    n = len(eh)
    mine_h = np.zeros((n, 30)); mine_t = np.zeros((n, 30)); ref_h = np.zeros((n, 30)); ref_t = np.zeros((n, 30))
    nh_mine = nh_ref = 0
    for rep in range(reps):
        ei, _ = dense.oracle.sample_subgraph_philox(7, rep, eh, et, er, nodes, l2g, neg)
        ri, _ = ref.neg_sample_fn(l2g_dict, nodes.tolist(), np.stack([eh, et]), er)
        for src, Hc, Tc in ((ei, mine_h, mine_t), (ri, ref_h, ref_t)):
            hs, ts = src[0].reshape(1 + neg, n), src[1].reshape(1 + neg, n)
            for b in range(n):
                for k in range(1, 1 + neg):
                    if hs[k, b] != eh[b]: Hc[b, hs[k, b]] += 1
                    elif ts[k, b] != et[b]: Tc[b, ts[k, b]] += 1
    # split: both ~ neg/2 head corruptions per edge (a draw that re-picks the original id is invisible; rare and equal)
    tot = reps * n * neg
    assert abs(mine_h.sum() / tot - 0.5) < 0.02 and abs(ref_h.sum() / tot - 0.5) < 0.02
    # support: a node is drawn by one sampler iff the other may draw it (the admissible set is the same)
    for b in range(0, n, 5):
        for mine, refc in ((mine_h[b], ref_h[b]), (mine_t[b], ref_t[b])):
            adm_m, adm_r = mine > 0, refc > 0
            # never-drawn nodes agree except for low-probability misses
            assert (adm_m ^ adm_r).sum() <= 2
            p_m, p_r = mine / max(mine.sum(), 1), refc / max(refc.sum(), 1)
            assert np.abs(p_m - p_r).max() < 0.035     # ~1/25 per node, 1500 draws per side: sigma ~ 0.005


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.gpu
@pytest.mark.parametrize("n_local,n_edges,neg,bern,filt,seed,step", [
    (40, 200, 10, 0, 1, 192, 0), (60, 1000, 1, 0, 1, 5, 3), (25, 64, 25, 1, 1, 1 << 40, (1 << 33) + 1),
    (40, 300, 4, 0, 0, 9, 2), (3, 10, 5, 0, 1, 1, 1), (40, 50, 0, 0, 1, 2, 2)])
def test_kernel_bit_exact_vs_cpu_replay(mre, dense, n_local, n_edges, neg, bern, filt, seed, step):
    import ctypes as C
    eng, L = mre.engine, mre._lib
    ix = eng.KGIndex.from_arrays(dense.E, dense.R, dense.train).to_device(0)
    ctx = eng.Context(0)
    l2g, eh, et, er = make_subgraph(dense, n_local, n_local, n_edges)
    nodes = np.arange(n_local - 1, dtype=np.int64)
    want_ei, want_type = dense.oracle.sample_subgraph_philox(seed, step, eh, et, er, nodes, l2g, neg, bern=bern, filt=filt, stream=2)
    out = torch.empty((3, n_edges * (1 + neg)), dtype=torch.int32, device="cuda")
    d = [dev(x) for x in (eh, et, er, nodes, l2g)]
    L.check(L.lib().mre_sample_subgraph(ctx._h, ix._h, seed, step, 2, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), n_edges,
                                        d[3].data_ptr(), len(nodes), d[4].data_ptr(), len(l2g), neg, bern, filt,
                                        out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), None))
    got = out.cpu().numpy()
    assert np.array_equal(got[:2], want_ei) and np.array_equal(got[2], want_type)


@pytest.mark.gpu
def test_negative_sampling_class_dropin(mre, dense):
    """paper.NegativeSampling: the reference's call sequence (neg_sample_fn -> scoring_fn -> margin loss) on device"""
    from oracle import kge_oracle as ko
    paper = mre.paper
    th, tt, tr = dense.oracle.train_triples()
    from importlib import import_module
    MarginLoss = import_module("multimodal-relation-extrapolation_b200.openke.module.loss").MarginLoss
    ns = paper.NegativeSampling(whole_triples=(th, tr, tt), loss_fn=MarginLoss(margin=3.0).cuda(), neg_ent=10, seed=3,
                                num_entities=dense.E, num_relations=dense.R)
    l2g, eh, et, er = make_subgraph(dense, 4, 40, 120)
    l2g_dict = {i: int(g) for i, g in enumerate(l2g)}
    edge_index = torch.from_numpy(np.stack([eh, et]))
    edge_type = torch.from_numpy(er)
    nodes = torch.arange(int(edge_index.max()))
    ei, etype = ns.neg_sample_fn(l2g_dict, nodes, edge_index, edge_type)
    assert ei.dtype == torch.int32 and not ei.is_cuda and ei.shape == (2, 120 * 11) and etype.shape == (120 * 11,)
    want_ei, want_type = dense.oracle.sample_subgraph_philox(3, 0, eh, et, er, nodes.numpy(), l2g, 10)
    assert np.array_equal(ei.numpy(), want_ei) and np.array_equal(etype.numpy(), want_type)
    assert leaks(dense, l2g, ei.numpy(), etype.numpy(), 120) == 0
    # score + loss against torch on the reference's expressions (module/NegativeSampling.py:142-157, module/loss.py:20-24)
    g = torch.Generator().manual_seed(0)
    x = torch.randn(40, 200, generator=g)
    rel_emb = torch.randn(120, 200, generator=g)
    loss = ns.struct_loss(l2g_dict, x, rel_emb, edge_index, edge_type)      # second call: step 1
    ei2, _ = dense.oracle.sample_subgraph_philox(3, 1, eh, et, er, nodes.numpy(), l2g, 10)
    ei2 = torch.from_numpy(ei2).long()
    rel_expand = rel_emb.repeat(11, 1)
    score = torch.norm((x[ei2[0]] + rel_expand) - x[ei2[1]], 1, -1)
    p, nsc = score[:120].view(-1, 120).permute(1, 0), score[120:].view(-1, 120).permute(1, 0)
    want = torch.max(p - nsc, torch.tensor(-3.0)).mean() + 3.0
    want = want + 0.5 * (torch.mean(x[ei2[0]] ** 2) + torch.mean(x[ei2[1]] ** 2) + torch.mean(rel_expand ** 2)) / 3
    assert torch.allclose(loss.cpu().flatten(), want.flatten(), rtol=1e-5, atol=1e-5)
