"""CPU: the C-ABI library loads without a GPU, exports every symbol include/mre_b200.h declares, its host-side index
builder (Reader.h's restatement) equals the oracle's, and compute entry points fail LOUDLY without a device."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import torch

import helpers


def test_library_exports_every_declared_symbol(mre):
    L = mre._lib
    want = L.exported_symbols_in_header()
    assert len(want) >= 25
    out = subprocess.run(["nm", "-D", "--defined-only", L.LIB_PATH], capture_output=True, text=True, check=True).stdout
    have = {line.split()[-1] for line in out.splitlines() if " T " in line}
    assert [s for s in want if s not in have] == []
    lib = L.lib()
    for s in want:
        assert hasattr(lib, s)
    assert lib.mre_abi_version() == 1


def test_no_triton_no_oracle_in_product():
    """the product package never imports the oracle or a compatibility layer"""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "multimodal-relation-extrapolation_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import triton" not in text and "torch.compile" not in text
                if f.endswith(".py"):
                    assert "from oracle" not in text and "import oracle" not in text


def test_index_matches_oracle(mre, fb15k237):
    eng = mre.engine
    ix = eng.KGIndex.from_arrays(fb15k237.E, fb15k237.R, fb15k237.train, fb15k237.valid, fb15k237.test)
    o = fb15k237.oracle
    assert (ix.ent_tot, ix.rel_tot) == (fb15k237.E, fb15k237.R)
    assert (ix.train_tot, ix.valid_tot, ix.test_tot, ix.triple_tot) == (o.train_total, o.valid_total, o.test_total, o.triple_total)
    for mine, ref in zip(ix.test_triples(), o.test_triples()):
        assert np.array_equal(mine, ref)
    for mine, ref in zip(ix.train_triples(), o.train_triples()):
        assert np.array_equal(mine, ref)
    for mine, ref in zip(ix.means(), o.means()):
        assert np.array_equal(mine, ref)       # float32 tph / hpt, bit for bit
    rng = np.random.default_rng(0)
    th, tt, tr = o.test_triples()
    for i in rng.integers(0, len(th), 50):
        assert ix.find(th[i], tt[i], tr[i]) and o.find(th[i], tt[i], tr[i])
    for _ in range(200):
        h, t, r = int(rng.integers(0, fb15k237.E)), int(rng.integers(0, fb15k237.E)), int(rng.integers(0, fb15k237.R))
        assert ix.find(h, t, r) == o.find(h, t, r)


def test_index_from_dir_equals_from_arrays(mre, tmp_path):
    from oracle import ref_driver as rd
    eng = mre.engine
    ds = helpers.synthetic_graph(1, 50, 4, 600, 40, 30)
    d = rd.write_benchmark_dir(str(tmp_path / "kg"), ds.E, ds.R, ds.train, ds.valid, ds.test)
    a = eng.KGIndex.from_dir(d)
    b = eng.KGIndex.from_arrays(ds.E, ds.R, ds.train, ds.valid, ds.test)
    assert (a.train_tot, a.valid_tot, a.test_tot, a.triple_tot) == (b.train_tot, b.valid_tot, b.test_tot, b.triple_tot)
    for x, y in zip(a.test_triples() + a.train_triples() + a.means(), b.test_triples() + b.train_triples() + b.means()):
        assert np.array_equal(x, y)
    with pytest.raises(mre.MreError, match="cannot open"):
        eng.KGIndex.from_dir(str(tmp_path / "missing"))


def test_bad_arguments_are_reported(mre):
    eng = mre.engine
    z = np.zeros(1, np.int64)
    with pytest.raises(mre.MreError, match="out of range"):
        eng.KGIndex.from_arrays(5, 2, (np.array([7]), z, z))
    with pytest.raises(mre.MreError, match="positive"):
        eng.KGIndex.from_arrays(0, 2, (z, z, z))


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_compute_fails_loudly_without_a_gpu(mre):
    """no CPU fallback: creating a workspace (the first thing every compute call needs) raises"""
    eng = mre.engine
    with pytest.raises(mre.MreError, match="CUDA|device"):
        eng.Context(0)
    ix = eng.KGIndex.from_arrays(5, 2, (np.array([1]), np.array([2]), np.array([0])))
    with pytest.raises(mre.MreError):
        ix.to_device(0)
    # the paper-side entry points too: the ZSL evaluator and the candidate-list evaluate have no host path to fall back to
    import golden_util as gu
    n_symbols, conn, deg, heads, rels, cands, rel_vecs = gu.synthetic_zsl_setup(D=16, max_nb=4)
    with pytest.raises(mre.MreError, match="CUDA|device"):
        mre.paper.ZSLEvaluator(gu.seeded_extractor_weights(1, n_symbols, 16), conn, deg, np.arange(300), device=0)
