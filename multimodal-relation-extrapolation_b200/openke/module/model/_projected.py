"""Shared by TransH and TransD: the link-prediction ranking of a translation model whose entities are projected per relation.

The reference runs `_transfer` on all E rows for every 1-vs-all query (TransH.py:66-96, TransD.py:92-131).  The projected table
depends on the relation only, so the relations of a run are projected once each (mre_relation_project, csrc/project.cu) into
one stacked table [n_rel * E, D]; every relation's queries then rank against their own slice as a candidate group of the
TransE kernel, with the known-true lists re-based to the stacked ids."""
import numpy as np
import torch

from .... import _lib as L
from .... import engine

CHUNK_BYTES = 2 << 30          # stacked-table budget per rank call (the gathered copy inside the library doubles it)


def known_lists(index, q_h, q_t, q_r, side):
    """per query: sorted distinct known answers over all splits of the index -- tails of (h, r) for a tail query, heads of (t, r)
    for a head query (what _find filters, Corrupt.h:166-177) -> CSR (ptr, idx)"""
    trip = [index.train_triples(), index.valid_triples(), index.test_triples()]
    h, t, r = (np.concatenate([s[k] for s in trip]) for k in range(3))
    R = index.rel_tot
    n_of = np.zeros(len(q_h), np.int64)
    lo_of = np.zeros(len(q_h), np.int64)
    vals = {}
    for s, (fixed, ans, qf) in enumerate(((t, h, q_t), (h, t, q_h))):       # side 0: heads of (t, r); side 1: tails of (h, r)
        key = fixed.astype(np.int64) * R + r
        order = np.lexsort((ans, key))
        key, val = key[order], ans[order].astype(np.int64)
        keep = np.ones(len(key), bool)
        keep[1:] = (key[1:] != key[:-1]) | (val[1:] != val[:-1])
        key, vals[s] = key[keep], val[keep]
        sel = side == s
        qk = qf[sel].astype(np.int64) * R + q_r[sel]
        lo, hi = np.searchsorted(key, qk, "left"), np.searchsorted(key, qk, "right")
        n_of[sel], lo_of[sel] = hi - lo, lo
    ptr = np.concatenate([[0], np.cumsum(n_of)]).astype(np.int64)
    idx = np.zeros(int(ptr[-1]), np.int64)
    for s in (0, 1):
        sel = np.nonzero(side == s)[0]
        n = n_of[sel]
        if n.sum() == 0:
            continue
        within = np.arange(int(n.sum())) - np.repeat(np.cumsum(n) - n, n)
        idx[np.repeat(ptr[sel], n) + within] = vals[s][np.repeat(lo_of[sel], n) + within]
    return ptr, idx


class RelationProjected:
    """mixin: rank_queries() for models that define project_kind, projection_tables() and norm_flag / p_norm"""

    def rank_queries(self, q_h, q_t, q_r, side, index):
        """host int64 query arrays + uint8 side array -> device int32 counts [4, Q] (raw_lt, raw_eq, filt_lt, filt_eq)"""
        dev = self.device()
        ctx, rk = self.ctx(), self.ranker()
        ent, ent_aux, rel, rel_aux = (x.detach().contiguous() if x is not None else None for x in self.projection_tables())
        E, D = ent.shape
        if rel.shape[1] != D:
            raise L.MreError("TransD with dim_e != dim_r is not supported by the projected ranking path")
        Q = len(q_h)
        counts = torch.zeros((4, Q), dtype=torch.int32, device=dev)
        order = np.argsort(q_r, kind="stable")
        fptr, fidx = known_lists(index, q_h, q_t, q_r, side)
        rel_hat = torch.nn.functional.normalize(rel, 2, -1) if self.norm_flag else rel      # r enters _calc unprojected
        rels_sorted, first = np.unique(q_r[order], return_index=True)
        bounds = np.concatenate([first, [Q]])
        per = max(1, int(CHUNK_BYTES // (E * D * 4)))
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
        st = torch.cuda.current_stream().cuda_stream
        for c0 in range(0, len(rels_sorted), per):
            rels = rels_sorted[c0:c0 + per]
            n_rel = len(rels)
            sel = order[bounds[c0]:bounds[min(c0 + per, len(rels_sorted))]]           # this chunk's queries, grouped by relation
            stacked = torch.empty((n_rel * E, D), dtype=torch.float32, device=dev)
            rels_d = to(rels.astype(np.int64))
            L.check(L.lib().mre_relation_project(ctx._h, self.project_kind, ent.data_ptr(), ent_aux.data_ptr() if ent_aux is not None else None,
                                                 rel_aux.data_ptr(), rels_d.data_ptr(), n_rel, E, D, int(self.norm_flag), stacked.data_ptr(), st))
            slot = np.searchsorted(rels, q_r[sel])
            off = slot.astype(np.int64) * E
            qptr = np.concatenate([[0], np.cumsum(np.bincount(slot, minlength=n_rel))]).astype(np.int64)
            cptr = (np.arange(n_rel + 1, dtype=np.int64) * E)
            groups = engine.CandidateGroups(qptr, cptr, torch.arange(n_rel * E, dtype=torch.int64, device=dev))
            n = fptr[sel + 1] - fptr[sel]
            sub_ptr = np.concatenate([[0], np.cumsum(n)]).astype(np.int64)
            within = np.arange(int(sub_ptr[-1])) - np.repeat(sub_ptr[:-1], n)
            sub_idx = fidx[np.repeat(fptr[sel], n) + within] + np.repeat(off, n)
            c = rk.rank("transe", (stacked, rel_hat), to(q_h[sel] + off), to(q_t[sel] + off), to(q_r[sel]), to(side[sel]), p_norm=self.p_norm,
                        normalize=False, groups=groups, filt_csr=(to(sub_ptr), to(sub_idx)))
            counts[:, to(sel)] = c
        return counts
