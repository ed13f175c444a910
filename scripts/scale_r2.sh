#!/bin/bash
# Round-2 multi-GPU pass (run on an 8-GPU box through `gpurun --gpus 8`): peer-memory checks, the data-parallel trainer, the
# training bench with the fused exchange and with NCCL, and the default (strong-scaled) bench line at N = 8 and N = 4.
set -u
O=gpurun_out/r2s
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 200 $TR --nproc-per-node 8 --master-port 29521 scripts/check_peer_multi.py > $O/check_peer_n8.log 2>&1; echo "check_peer rc $?"
timeout 200 $TR --nproc-per-node 8 --master-port 29522 scripts/check_trainer_multi.py > $O/check_trainer_n8.log 2>&1; echo "check_trainer rc $?"
for ar in fused nccl; do
  o=""; [ $ar = nccl ] && o="--opt train_allreduce=nccl"
  timeout 300 $TR --nproc-per-node 8 --master-port 29523 bench.py --gpus 8 --workload train --steps 100 --warmup 10 --no-extra $o > $O/train_n8_$ar.json 2> $O/train_n8_$ar.err; echo "train $ar rc $?"
done
timeout 300 $TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 --workload train --steps 100 --warmup 10 --no-extra > $O/train_n4_fused.json 2> $O/train_n4_fused.err; echo "train n4 rc $?"
timeout 600 $TR --nproc-per-node 8 --master-port 29525 bench.py --gpus 8 --steps 6 --warmup 3 > $O/default_n8.json 2> $O/default_n8.err; echo "default n8 rc $?"
timeout 600 $TR --nproc-per-node 4 --master-port 29526 bench.py --gpus 4 --steps 4 --warmup 3 --no-extra > $O/default_n4.json 2> $O/default_n4.err; echo "default n4 rc $?"
tail -3 $O/check_peer_n8.log $O/check_trainer_n8.log
