"""TrainDataLoader -- the loader the reference lost to its .gitignore (openke/data/), re-created
against the contract its callers and the C side pin down:
  OpenKE/examples/train_transe_FB15K237.py:9-20  constructor keywords
  OpenKE/openke/base/Base.cpp:161-174            sampling(h, t, r, y, B, negRate, negRelRate, mode, filter, p, val_loss)
  OpenKE/openke/module/model/TransE.py:51-54     the "head_batch"/"tail_batch" shapes of cross sampling
Batches come from the Philox kernel (mre_sample) instead of Base.so's pthread sampler.  `device_batches=True` keeps the
batch on the GPU (torch tensors) so the fused Trainer never round-trips through the host; the default hands out numpy
arrays exactly like the reference.
"""
import numpy as np

from ... import engine


class TrainDataSampler(object):
    def __init__(self, nbatches, datasampler):
        self.nbatches = nbatches
        self.datasampler = datasampler
        self.batch = 0

    def __iter__(self):
        return self

    def __next__(self):
        self.batch += 1
        if self.batch > self.nbatches:
            raise StopIteration()
        return self.datasampler()

    def __len__(self):
        return self.nbatches


class TrainDataLoader(object):
    def __init__(self, in_path="./", tri_file=None, ent_file=None, rel_file=None, batch_size=None, nbatches=None, threads=8,
                 sampling_mode="normal", bern_flag=False, filter_flag=True, neg_ent=1, neg_rel=0, seed=0, stream_id=0,
                 device=0, device_batches=False, index=None):
        if tri_file is not None or ent_file is not None or rel_file is not None:
            raise NotImplementedError("explicit tri_file/ent_file/rel_file paths: pass in_path (benchmark directory)")
        if neg_rel:
            raise NotImplementedError("relation corruption (neg_rel > 0, Corrupt.h:86-163) is outside the hot path")
        self.in_path = in_path
        self.work_threads = threads      # accepted for compatibility; the sampler is one GPU thread per slot
        self.nbatches = nbatches
        self.batch_size = batch_size
        self.bern = int(bool(bern_flag))
        self.filter = int(bool(filter_flag))  # ignored by the reference too (Base.cpp:116,119)
        self.negative_ent = neg_ent
        self.negative_rel = neg_rel
        self.sampling_mode = sampling_mode
        self.cross_sampling_flag = 0
        self.device_batches = device_batches
        self.index = index if index is not None else engine.KGIndex.from_dir(in_path)
        self.ctx = engine.Context(device)
        if self.index.device is None:
            self.index.to_device(device)
        self.sampler = engine.Sampler(self.index, ctx=self.ctx, seed=seed, stream_id=stream_id)
        self.rel_tot, self.ent_tot, self.tripleTotal = self.index.rel_tot, self.index.ent_tot, self.index.train_tot
        if self.batch_size is None:
            self.batch_size = self.tripleTotal // self.nbatches
        if self.nbatches is None:
            self.nbatches = self.tripleTotal // self.batch_size
        self.step = 0
        self.batch_seq_size = 0
        self._size_buffers()

    def _size_buffers(self):
        """host batch arrays of B (1 + neg_ent + neg_rel) rows; re-made whenever a setter changed the sizes (the library
        writes exactly that many elements into them)"""
        n = self.batch_size * (1 + self.negative_ent + self.negative_rel)
        if n != self.batch_seq_size and not self.device_batches:
            self.batch_h = np.zeros(n, dtype=np.int64)
            self.batch_t = np.zeros(n, dtype=np.int64)
            self.batch_r = np.zeros(n, dtype=np.int64)
            self.batch_y = np.zeros(n, dtype=np.float32)
        self.batch_seq_size = n

    def _draw_prefetched(self):
        """device batches in "normal" mode: batch k was drawn on a side stream while the consumer worked on batch k - 1 (its
        Philox stream depends on (seed, step) only), into one of two buffer sets; the consumer's stream waits for the draw, the
        next draw waits until the consumer's work on the buffer it overwrites has been enqueued and finished"""
        import torch
        n = self.batch_seq_size
        dev = self.sampler.device
        st = getattr(self, "_pf", None)
        if st is None or st["n"] != n or st["key"] != (self.batch_size, self.negative_ent, self.bern):
            st = {"n": n, "key": (self.batch_size, self.negative_ent, self.bern), "side": torch.cuda.Stream(device=dev), "next": None,
                  "bufs": [tuple([torch.empty(n, dtype=torch.int64, device=dev) for _ in range(3)] + [torch.empty(n, dtype=torch.float32, device=dev)])
                           for _ in range(2)], "free": [None, None]}
            self._pf = st

        def launch():
            k = self.step & 1
            with torch.cuda.stream(st["side"]):
                if st["free"][k] is not None:
                    st["side"].wait_event(st["free"][k])
                batch = self.sampler.sample(self.step, self.batch_size, self.negative_ent, mode=0, bern=self.bern, out=st["bufs"][k])
                ev = torch.cuda.Event()
                ev.record(st["side"])
            self.step += 1
            return batch, ev, k

        if st["next"] is None:
            st["next"] = launch()
        batch, ev, k = st["next"]
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ev)
        # the PREVIOUS batch's buffer (the other set) is free once everything enqueued so far on the consumer's stream is done
        st["free"][1 - k] = torch.cuda.Event()
        st["free"][1 - k].record(cur)
        st["next"] = launch()
        return batch

    def _draw(self, mode):
        self._size_buffers()
        if self.device_batches and mode == 0:
            return self._draw_prefetched()
        step = self.step
        self.step += 1
        if self.device_batches:
            return self.sampler.sample(step, self.batch_size, self.negative_ent, mode=mode, bern=self.bern)
        return self.sampler.sample_host(step, self.batch_size, self.negative_ent, mode=mode, bern=self.bern,
                                        out=(self.batch_h, self.batch_t, self.batch_r, self.batch_y))

    def sampling(self):
        h, t, r, y = self._draw(0)
        return {"batch_h": h, "batch_t": t, "batch_r": r, "batch_y": y, "mode": "normal"}

    def sampling_head(self):
        h, t, r, y = self._draw(-1)
        B = self.batch_size
        return {"batch_h": h, "batch_t": t[:B], "batch_r": r[:B], "batch_y": y, "mode": "head_batch"}

    def sampling_tail(self):
        h, t, r, y = self._draw(1)
        B = self.batch_size
        return {"batch_h": h[:B], "batch_t": t, "batch_r": r[:B], "batch_y": y, "mode": "tail_batch"}

    def cross_sampling(self):
        self.cross_sampling_flag = 1 - self.cross_sampling_flag
        if self.cross_sampling_flag == 0:
            return self.sampling_head()
        return self.sampling_tail()

    # ---- setters / getters of the reference loader
    def set_work_threads(self, work_threads):
        self.work_threads = work_threads

    def set_in_path(self, in_path):
        self.in_path = in_path

    def set_nbatches(self, nbatches):
        self.nbatches = nbatches

    def set_batch_size(self, batch_size):
        self.batch_size = batch_size
        self.nbatches = self.tripleTotal // self.batch_size

    def set_ent_neg_rate(self, rate):
        self.negative_ent = rate

    def set_rel_neg_rate(self, rate):
        if rate:
            raise NotImplementedError("relation corruption is outside the hot path")

    def set_bern_flag(self, bern):
        self.bern = int(bool(bern))

    def set_filter_flag(self, filter):
        self.filter = int(bool(filter))

    def get_batch_size(self):
        return self.batch_size

    def get_ent_tot(self):
        return self.ent_tot

    def get_rel_tot(self):
        return self.rel_tot

    def get_triple_tot(self):
        return self.tripleTotal

    def __iter__(self):
        if self.sampling_mode == "normal":
            return TrainDataSampler(self.nbatches, self.sampling)
        return TrainDataSampler(self.nbatches, self.cross_sampling)

    def __len__(self):
        return self.nbatches
