"""Generates tests/golden/golden_zsl.npz by running the REFERENCE's own Extractor (module/zsl_module.py:16-106, imported
from /root/reference with the three modules it cannot import here stubbed) + the reference's scoring expression
(sklearn cosine_similarity(...).mean(axis=1), zsl_module.py:699-706) on a seeded synthetic graph, and asserts that
oracle/zsl_oracle.py restates it bit for bit.  Build-container only; the tests read just the .npz."""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.dont_write_bytecode = True
from oracle import zsl_oracle as zo  # noqa: E402

REF = "/root/reference"
SEED, N_ENT, N_REL, D, MAX_NB, T = 192, 300, 3, 200, 50, 14


def synthetic_setup():
    return zo.synthetic_zsl_setup(SEED, N_ENT, N_REL, D, MAX_NB, T)


def main():
    sys.path.insert(0, REF)
    for name in ("module.model", "module.spectral_norm", "module.utils"):
        sys.modules[name] = types.ModuleType(name)
    sys.modules["module.model"].MaskedMultimodalAutoencoder = object
    sys.modules["module.spectral_norm"].spectral_norm = lambda x: x
    from module.zsl_module import Extractor          # the reference class
    from sklearn.metrics.pairwise import cosine_similarity

    n_symbols, conn, deg, heads, rels, cands, rel_vecs = synthetic_setup()
    w = zo.seeded_extractor_weights(SEED, n_symbols, D)
    ext = Extractor(D, n_symbols, embed=w["symbol_emb.weight"])
    sd = {k: torch.from_numpy(v) for k, v in w.items()}
    missing = ext.load_state_dict(sd, strict=True)
    ext.eval()
    out = {"scores": [], "ranks": [], "ptr": [0]}
    probe_vecs = []
    with torch.no_grad():
        for t in range(T):
            h, c = int(heads[t]), cands[t]
            query = torch.from_numpy(np.stack([np.full(len(c), h), c], 1))                       # symbol ids == entity ids here
            meta = (torch.from_numpy(conn[np.full(len(c), h)]), torch.from_numpy(deg[np.full(len(c), h)]),
                    torch.from_numpy(conn[c]), torch.from_numpy(deg[c]))
            vecs, _ = ext(query, query, meta, meta)                                              # zsl_module.py:690-693
            vecs = vecs.numpy()
            mine = zo.extractor_query_vectors(w, query.numpy(), conn[np.full(len(c), h)][:, :, 1], deg[np.full(len(c), h)],
                                              conn[c][:, :, 1], deg[c])
            assert np.array_equal(vecs, mine), f"restatement differs from the reference Extractor on triple {t}"
            scores = cosine_similarity(vecs, rel_vecs[rels[t]]).mean(axis=1)                     # zsl_module.py:699-701
            assert scores.dtype == np.float32
            mine_s = zo.cosine_mean_scores(vecs, rel_vecs[rels[t]])
            assert np.allclose(scores, mine_s, rtol=0, atol=2e-7)
            sort = list(np.argsort(scores))[::-1]
            rank = sort.index(0) + 1                                                             # zsl_module.py:705-706
            lo, hi = zo.rank_interval(scores)
            assert lo <= rank <= hi
            out["scores"].append(scores); out["ranks"].append(rank); out["ptr"].append(out["ptr"][-1] + len(c))
            probe_vecs.append(vecs[:2])
    ranks = np.asarray(out["ranks"], np.float64)
    np.savez_compressed(os.path.join(HERE, "golden_zsl.npz"), seed=SEED, n_ent=N_ENT, n_rel=N_REL, D=D, max_nb=MAX_NB,
                        scores=np.concatenate(out["scores"]), ranks=np.asarray(out["ranks"], np.int64), ptr=np.asarray(out["ptr"], np.int64),
                        probe_vecs=np.concatenate(probe_vecs),
                        metrics=np.asarray([(ranks <= 10).mean(), (ranks <= 5).mean(), (1.0 / ranks).mean()]))   # (hits10, hits5, mrr), :745
    print("golden_zsl.npz written; ranks", out["ranks"])


if __name__ == "__main__":
    main()
