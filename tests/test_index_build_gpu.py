"""The GPU index build (mre_index_create_device, csrc/index_build.cu: LSD radix sort + flag / scan / compact + relation counters)
against the host build (mre_index_create + mre_index_to_device), which tests/test_oracle_cpu.py and tests/test_abi_cpu.py pin
to the compiled reference's importTrainFiles / importTestFiles (Reader.h:53-257): every host getter and every device-resident
column must hold the same bits."""
import numpy as np
import pytest

COLUMNS = range(11)


def assert_same_index(a, b):
    for name in ("ent_tot", "rel_tot", "train_tot", "valid_tot", "test_tot", "triple_tot"):
        assert getattr(a, name) == getattr(b, name), name
    for fn in ("train_triples", "valid_triples", "test_triples"):
        for x, y in zip(getattr(a, fn)(), getattr(b, fn)()):
            assert np.array_equal(x, y), fn
    for x, y in zip(a.means(), b.means()):
        assert np.array_equal(x, y, equal_nan=True)
    for c in COLUMNS:
        x, y = a.device_column(c), b.device_column(c)
        assert x.shape == y.shape and np.array_equal(x, y, equal_nan=True), c


@pytest.mark.gpu
def test_device_build_matches_host_build_on_fb15k237(mre, fb15k237):
    eng = mre.engine
    args = (fb15k237.E, fb15k237.R, fb15k237.train, fb15k237.valid, fb15k237.test)
    host = eng.KGIndex.from_arrays(*args).to_device(0)
    dev = eng.KGIndex.from_arrays_device(*args, device=0)
    assert dev.build_ms > 0
    assert_same_index(host, dev)
    h, t, r = fb15k237.test
    for i in (0, 17, len(h) - 1):
        assert dev.find(int(h[i]), int(t[i]), int(r[i])) == host.find(int(h[i]), int(t[i]), int(r[i]))
    # the device tables serve a rank job and the sampler exactly as the uploaded ones do
    import torch
    rng = np.random.default_rng(0)
    ent = torch.from_numpy(rng.standard_normal((fb15k237.E, 64)).astype(np.float32)).cuda()
    rel = torch.from_numpy(rng.standard_normal((fb15k237.R, 64)).astype(np.float32)).cuda()
    q = slice(0, 512)
    dv = lambda a: torch.from_numpy(np.ascontiguousarray(a[q])).cuda()
    rk = eng.Ranker(device=0)
    side = torch.from_numpy((np.arange(512) % 2).astype(np.uint8)).cuda()
    c_host = rk.rank("transe", (ent, rel), dv(h), dv(t), dv(r), side, index=host).cpu().numpy()
    c_dev = rk.rank("transe", (ent, rel), dv(h), dv(t), dv(r), side, index=dev).cpu().numpy()
    assert np.array_equal(c_host, c_dev)
    sa, sb = eng.Sampler(host, rk.ctx, seed=5), eng.Sampler(dev, rk.ctx, seed=5)
    for x, y in zip(sa.sample(0, 1024, 4), sb.sample(0, 1024, 4)):
        assert torch.equal(x, y)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["wide_ids", "heavy_duplicates", "train_only", "tiny", "one_relation", "empty"])
def test_device_build_matches_host_build_on_synthetic_graphs(mre, case):
    eng = mre.engine
    rng = np.random.default_rng(11)

    def draw(n, E, R):
        return rng.integers(0, E, n), rng.integers(0, E, n), rng.integers(0, R, n)
    if case == "wide_ids":              # three 8-bit digit passes per column; tiles that do not divide the list
        E, R = 3_000_017, 70_001
        train, valid, test = draw(300_001, E, R), draw(4_099, E, R), draw(5_003, E, R)
        train = tuple(np.concatenate([a, b[:100]]) for a, b in zip(train, test))        # triples shared between splits
    elif case == "heavy_duplicates":    # few distinct triples: long runs for the de-duplication and the distinct-pair counters
        E, R = 50, 3
        train, valid, test = draw(40_000, E, R), draw(3_000, E, R), draw(3_000, E, R)
    elif case == "train_only":
        E, R = 1_000, 11
        train, valid, test = draw(10_000, E, R), None, None
    elif case == "tiny":
        E, R = 5, 2
        train = (np.array([4, 0, 4]), np.array([1, 2, 1]), np.array([1, 0, 1]))
        valid, test = (np.array([3]), np.array([3]), np.array([0])), (np.array([0]), np.array([2]), np.array([0]))
    elif case == "empty":                # no triple at all: every table empty, tph / hpt = 0 / 0
        E, R = 4, 3
        train = valid = test = None
        train = (np.zeros(0, np.int64),) * 3
    else:                                # R = 1: a relation column of one bit
        E, R = 70_000, 1
        train, valid, test = draw(100_000, E, R), draw(1_000, E, R), draw(1_000, E, R)
    host = eng.KGIndex.from_arrays(E, R, train, valid, test).to_device(0)
    dev = eng.KGIndex.from_arrays_device(E, R, train, valid, test, device=0)
    assert_same_index(host, dev)


@pytest.mark.gpu
def test_device_build_rejects_out_of_range_ids(mre):
    eng = mre.engine
    train = (np.array([0, 1, 7]), np.array([1, 2, 0]), np.array([0, 0, 0]))
    with pytest.raises(mre._lib.MreError, match="train triple 2"):
        eng.KGIndex.from_arrays_device(5, 2, train, device=0)


@pytest.mark.gpu
def test_from_dir_on_device_matches_host(mre, tmp_path):
    """KGIndex.from_dir(path, device=0) (mre_index_create_from_dir_device: host parse, GPU build, type_constrain.txt loaded when
    present) against KGIndex.from_dir(path).to_device(0); a TestDataLoader built on it serves the Tester"""
    import helpers
    from oracle import ref_driver as rd
    eng = mre.engine
    ds = helpers.synthetic_graph(21, 500, 9, 8000, 300, 300)
    d = rd.write_benchmark_dir(str(tmp_path / "kg"), ds.E, ds.R, ds.train, ds.valid, ds.test)
    rng = np.random.default_rng(2)
    with open(d + "/type_constrain.txt", "w") as f:            # importTypeFiles' format: per relation a head line, then a tail line
        f.write(f"{ds.R}\n")
        for r in range(ds.R):
            for _ in range(2):
                ids = np.unique(rng.integers(0, ds.E, 40))
                f.write(f"{r}\t{len(ids)}\t" + "\t".join(map(str, ids.tolist())) + "\n")
    host = eng.KGIndex.from_dir(d).to_device(0)
    dev = eng.KGIndex.from_dir(d, device=0)
    assert_same_index(host, dev)
    assert dev.has_type_constrain
    for side in (0, 1):
        for x, y in zip(host.type_constrain(side), dev.type_constrain(side)):
            assert np.array_equal(x, y)
    loader = mre.openke.data.TestDataLoader(d, "link", index=dev)
    assert loader.get_triple_tot() == host.test_tot
