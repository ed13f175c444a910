"""Trainer (OpenKE/openke/config/Trainer.py:18-99): same constructor, same epoch/batch loop, same optimiser choices.
With a fusable strategy (any library scorer + library loss, no regulariser) and opt_method "sgd" one step is: Philox sample on
the device -> mre_ns_train_step (forward + loss + backward) -> mre_sgd_update; nothing crosses PCIe.  Data-parallel
(`dist=DistContext()`): the embedding tables and their gradients are moved into NVLink peer memory once (PeerGroup) and the
update becomes mre_dp_sgd_step -- gradient reduce-scatter + SGD + weight all-gather in ONE kernel per rank, no collective
library on the step; if peer memory cannot be set up the step falls back to NCCL all-reduces of the gradient tables.  Any
other combination runs the strategy's autograd forward and the torch optimiser on the same kernels' scores."""
import os

import torch
import torch.optim as optim

from ... import engine


class Trainer(object):
    def __init__(self, model=None, data_loader=None, train_times=1000, alpha=0.5, use_gpu=True, opt_method="sgd", save_steps=None,
                 checkpoint_dir=None, dist=None):
        self.work_threads = 8
        self.train_times = train_times
        self.opt_method = opt_method
        self.optimizer = None
        self.lr_decay = 0
        self.weight_decay = 0
        self.alpha = alpha
        self.model = model
        self.data_loader = data_loader
        self.use_gpu = use_gpu
        self.save_steps = save_steps
        self.checkpoint_dir = checkpoint_dir
        self.dist = dist
        self.losses = []
        self.peer = None              # PeerGroup holding the tables + gradients when data-parallel and fused
        self._peer_tried = False

    def _adopt_peer_memory(self):
        """data-parallel + fused: re-home the embedding tables and their gradient tables, back to back, in one peer-memory region
        (parameters keep their identity: only .data / .grad are re-pointed), so that the update is mre_dp_sgd_step"""
        if self._peer_tried or self.dist is None or self.dist.world == 1:
            return
        self._peer_tried = True
        import torch.distributed as td
        from ... import dist as mdist
        tabs = self.model.model.tables()
        sizes = [t.numel() for t in tabs]
        n = (sum(sizes) + 3) & ~3
        ok = torch.ones(1, device=tabs[0].device)
        pg = None
        try:
            assert all(s % 4 == 0 for s in sizes)          # every table starts 16-byte aligned inside the flat buffer
            pg = mdist.PeerGroup(self.model.model.ctx(), n)
        except Exception as e:  # noqa: BLE001
            print(f"peer memory unavailable ({e}); data-parallel steps use NCCL all-reduces")
            ok.zero_()
        td.all_reduce(ok, op=td.ReduceOp.MIN)
        if ok.item() == 0:
            return
        off = 0
        for t, s in zip(tabs, sizes):
            w, g = pg.weights[off:off + s].view_as(t), pg.grads[off:off + s].view_as(t)
            w.copy_(t.data)
            if t.grad is not None:
                g.copy_(t.grad)
            t.data, t.grad = w, g
            off += s
        td.barrier()
        torch.cuda.synchronize()
        self.peer = pg

    def _fused(self):
        return (self.use_gpu and hasattr(self.model, "can_fuse") and self.model.can_fuse()
                and self.opt_method.lower() == "sgd" and self.weight_decay == 0 and self.lr_decay == 0)

    def to_var(self, x, use_gpu):
        if isinstance(x, torch.Tensor):
            return x.cuda() if use_gpu else x
        t = torch.from_numpy(x)
        return t.cuda() if use_gpu else t

    def train_one_step(self, data):                                             # Trainer.py:43-54
        batch = {"batch_h": self.to_var(data["batch_h"], self.use_gpu), "batch_t": self.to_var(data["batch_t"], self.use_gpu),
                 "batch_r": self.to_var(data["batch_r"], self.use_gpu), "batch_y": self.to_var(data["batch_y"], self.use_gpu),
                 "mode": data["mode"]}
        if self._fused():
            self._adopt_peer_memory()
            loss = self.model.fused_step(batch)
            if self.peer is not None:
                self.peer.sgd_step(self.alpha / self.dist.world)      # mean over ranks of the per-rank mean losses
                return loss
            tabs = self.model.model.tables()
            if self.dist is not None:
                self.dist.all_reduce_grads([t.grad for t in tabs])
            for t in tabs:
                engine.sgd_update(self.model.model.ctx(), t.data, t.grad, self.alpha)
            return loss
        self.optimizer.zero_grad()
        loss = self.model(batch)
        loss.backward()
        if self.dist is not None:
            self.dist.all_reduce_grads([p.grad for p in self.model.parameters() if p.grad is not None])
        self.optimizer.step()
        return loss

    def _build_optimizer(self):                                                 # Trainer.py:60-86
        if self.optimizer is not None or self._fused():
            return
        params = [p for p in self.model.parameters() if p.requires_grad]
        m = self.opt_method.lower()
        if m == "adagrad":
            self.optimizer = optim.Adagrad(params, lr=self.alpha, lr_decay=self.lr_decay, weight_decay=self.weight_decay)
        elif m == "adadelta":
            self.optimizer = optim.Adadelta(params, lr=self.alpha, weight_decay=self.weight_decay)
        elif m == "adam":
            self.optimizer = optim.Adam(params, lr=self.alpha, weight_decay=self.weight_decay)
        else:
            self.optimizer = optim.SGD(params, lr=self.alpha, weight_decay=self.weight_decay)

    def run(self):                                                              # Trainer.py:56-99
        if self.use_gpu:
            self.model.cuda()
        self._build_optimizer()
        print("Finish initializing...")
        for epoch in range(self.train_times):
            res = None
            for data in self.data_loader:
                loss = self.train_one_step(data)
                res = loss.detach() if res is None else res + loss.detach()
            self.losses.append(float(res))       # one device->host read per epoch (the reference reads .item() per batch)
            if self.save_steps and self.checkpoint_dir and (epoch + 1) % self.save_steps == 0:
                print("Epoch %d has finished, saving..." % (epoch))
                self.model.model.save_checkpoint(os.path.join(self.checkpoint_dir + "-" + str(epoch) + ".ckpt"))

    # ---- setters of the reference Trainer
    def set_model(self, model):
        self.model = model

    def to_cuda(self):
        self.model.cuda()

    def set_use_gpu(self, use_gpu):
        self.use_gpu = use_gpu

    def set_alpha(self, alpha):
        self.alpha = alpha

    def set_lr_decay(self, lr_decay):
        self.lr_decay = lr_decay

    def set_weight_decay(self, weight_decay):
        self.weight_decay = weight_decay

    def set_opt_method(self, opt_method):
        self.opt_method = opt_method

    def set_train_times(self, train_times):
        self.train_times = train_times

    def set_save_steps(self, save_steps, checkpoint_dir=None):
        self.save_steps = save_steps
        if not self.checkpoint_dir:
            self.set_checkpoint_dir(checkpoint_dir)

    def set_checkpoint_dir(self, checkpoint_dir):
        self.checkpoint_dir = checkpoint_dir
