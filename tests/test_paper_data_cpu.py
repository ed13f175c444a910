"""Host-side loaders / generators of the paper's data artefacts: module/utils.py:194-207,
utils/gen_e1r_e2_all.py:14-19, utils/gen_rel2candidates.py:14-27, utils/gen_mode_candidates.py:15-39."""
import json
import os

import numpy as np
import pytest


def write_dataset(d):
    ents = [f"/m/e{i}" for i in range(40)] + ["/m/unknown_to_zsl"]
    e2id = {e: i for i, e in enumerate(ents[:40])}
    r2id = {"/r/a": 0, "/r/b": 1}
    tasks = {"/r/a": [["/m/e1", "/r/a", "/m/e2"], ["/m/e1", "/r/a", "/m/e3"], ["/m/e4", "/r/a", "/m/e2"]],
             "/r/b": [["/m/e5", "/r/b", "/m/e6"]]}
    rel2cand = {"/r/a": ["/m/e2", "/m/e3", "/m/e7", "/m/e8", "/m/unknown_to_zsl"], "/r/b": ["/m/e6", "/m/e9"]}
    for name, obj in (("entity2ids_zsl.json", e2id), ("relation2ids.json", r2id), ("test_tasks_zsl.json", tasks),
                      ("rel2candidates_all.json", rel2cand)):
        json.dump(obj, open(os.path.join(d, name), "w"))
    return e2id, r2id, tasks, rel2cand


def test_loaders_and_generators(mre, tmp_path):
    paper = mre.paper
    e2id, r2id, tasks, rel2cand = write_dataset(str(tmp_path))
    t2, (h, r, t), e2, r2 = paper.load_tasks(str(tmp_path), "test")
    assert t2 == tasks and e2 == e2id and r2 == r2id
    assert h.tolist() == [1, 1, 4, 5] and r.tolist() == [0, 0, 0, 1] and t.tolist() == [2, 3, 2, 6]
    triples = [tuple(x) for rel in tasks for x in tasks[rel]]
    e1 = paper.gen_e1rel_e2(triples)
    assert e1["/m/e1/r/a"] == ["/m/e2", "/m/e3"] and e1["/m/e4/r/a"] == ["/m/e2"]
    cands = paper.build_test_candidates(triples, rel2cand, e1, e2id)
    # true tail first; known tails of (head, rel), the true tail and entities missing from entity2id are dropped
    assert cands["/r/a"]["/m/e1\t/r/a\t/m/e2"] == ["/m/e2", "/m/e7", "/m/e8"]
    assert cands["/r/a"]["/m/e4\t/r/a\t/m/e2"] == ["/m/e2", "/m/e3", "/m/e7", "/m/e8"]
    assert cands["/r/b"]["/m/e5\t/r/b\t/m/e6"] == ["/m/e6", "/m/e9"]
    r2c = paper.gen_rel2candidates(triples, list(e2id), n=5, seed=3)
    assert set(r2c) == {"/r/a", "/r/b"} and all(len(set(v)) == 5 and set(v) <= set(e2id) for v in r2c.values())
    assert r2c == paper.gen_rel2candidates(triples, list(e2id), n=5, seed=3)


@pytest.mark.gpu
def test_evaluate_from_dir(mre, tmp_path):
    import torch
    from oracle import paper_oracle as po
    paper = mre.paper
    e2id, r2id, tasks, rel2cand = write_dataset(str(tmp_path))
    g = torch.Generator().manual_seed(0)
    ent, rel = torch.randn(40, 16, generator=g), torch.randn(2, 16, generator=g)
    out = paper.evaluate_from_dir(str(tmp_path), ent.cuda(), rel.cuda(), verbose=False)
    triples = [tuple(x) for rel_ in tasks for x in tasks[rel_]]
    cands = paper.build_test_candidates(triples, rel2cand, paper.gen_e1rel_e2(triples), e2id)
    ranks = []
    for rel_ in cands:
        for key, lst in cands[rel_].items():
            hd = key.split("\t")[0]
            ids = np.asarray([e2id[c] for c in lst])
            ranks.append(po.rank_ties_half(po.paper_transe_scores(ent.numpy(), rel.numpy(), e2id[hd], r2id[rel_], ids)))
    mrr, hits = po.summarize(ranks, (1, 3, 10))
    assert np.isclose(out[0], mrr) and np.allclose(out[1:], hits)
