"""Multi-process check of the data-parallel Trainer (run under torchrun on a multi-GPU box):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/check_trainer_multi.py
every rank trains TransE on a synthetic graph through openke.config.Trainer(dist=DistContext()): the fused step must have adopted
peer memory, the replicas must hold bit-identical tables after every epoch, and the loss must fall."""
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

import mre_b200
from oracle import ref_driver as rd

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dctx = mre_b200.dist.DistContext()
ok = mre_b200.openke
rng = np.random.default_rng(0)
E, R = 500, 9
split = lambda n: (rng.integers(0, E, n), rng.integers(0, E, n), rng.integers(0, R, n))
d = tempfile.mkdtemp(prefix=f"mre_multi_{rank}_")
rd.write_benchmark_dir(d, E, R, split(8000), split(200), split(200))
ld = ok.data.TrainDataLoader(in_path=d + "/", nbatches=8, bern_flag=1, neg_ent=5, seed=192, stream_id=rank, device=local, device_batches=True)
torch.manual_seed(0)
model = ok.module.model.TransE(E, R, dim=64, p_norm=1, norm_flag=True)
strat = ok.module.strategy.NegativeSampling(model=model, loss=ok.module.loss.MarginLoss(margin=5.0), batch_size=ld.get_batch_size())
tr = ok.config.Trainer(model=strat, data_loader=ld, train_times=6, alpha=0.5, use_gpu=True, opt_method="sgd", dist=dctx)
tr.run()
assert tr.peer is not None, "the trainer did not adopt peer memory"
tr.peer.check()
w = model.ent_embeddings.weight.data
ws = [torch.zeros_like(w) for _ in range(world)]
dist.all_gather(ws, w.clone())
same = all(torch.equal(x, ws[0]) for x in ws)
if rank == 0:
    print(f"world {world}: losses {['%.3f' % x for x in tr.losses]}, replicas identical: {same}")
    assert same and tr.losses[-1] < tr.losses[0]
dist.barrier()
tr.peer.close()
dist.destroy_process_group()
