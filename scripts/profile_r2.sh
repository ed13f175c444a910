#!/bin/bash
# Round-2 profiling pass (run on the GPU box through gpurun): launch lists + one `ncu --set full` capture per hot kernel.
# Every ncu run follows a plain run of the same command that exited 0 (the bench lines of gpurun_out/r2*/).
set -u
O=gpurun_out/r2p
mkdir -p $O
NCU="ncu --clock-control none"
# launch lists (gpu__time_duration per launch; cold-cache, serialised: the SHARE of the step is what counts)
$NCU --metrics gpu__time_duration.sum -c 600 --csv --log-file $O/launches_default.csv python bench.py --steps 2 --warmup 3 --queries 16384 > $O/launches_default.log 2>&1
$NCU --metrics gpu__time_duration.sum -c 200 --csv --log-file $O/launches_train.csv python bench.py --workload train --steps 3 --warmup 3 --no-extra > $O/launches_train.log 2>&1
# full captures
$NCU --set full --import-source on -k regex:transe_rank_kernel -s 3 -c 1 -o $O/transe_rank_2m python bench.py --workload synthetic2m --queries 16384 --steps 2 --warmup 3 --no-extra > $O/ncu_transe_2m.log 2>&1
$NCU --set full --import-source on -k regex:transe_rank_kernel -s 3 -c 1 -o $O/transe_rank_db15k python bench.py --workload db15k_zs --steps 2 --warmup 3 --no-extra > $O/ncu_transe_db.log 2>&1
$NCU --set full --import-source on -k regex:transe_rank_kernel -s 3 -c 1 -o $O/transe_rank_fbzs python bench.py --workload fb15k237_zs --steps 2 --warmup 3 --no-extra > $O/ncu_transe_fbzs.log 2>&1
$NCU --set full --import-source on -k regex:bilinear_rank_kernel -s 3 -c 1 -o $O/bilinear_distmult python bench.py --workload distmult --steps 2 --warmup 3 --no-extra > $O/ncu_bil_dm.log 2>&1
$NCU --set full --import-source on -k regex:bilinear_rank_kernel -s 3 -c 1 -o $O/bilinear_complex python bench.py --workload complex --steps 2 --warmup 3 --no-extra > $O/ncu_bil_cx.log 2>&1
$NCU --set full --import-source on -k regex:zsl_tc_kernel -s 2 -c 1 -o $O/zsl_tc python bench.py --workload zsl --steps 2 --warmup 3 --no-extra > $O/ncu_zsl.log 2>&1
$NCU --set full --import-source on -k "regex:transe_fwd_kernel|transe_bwd_kernel|ns_loss_kernel|sgd_kernel|sample_kernel|dp_sgd_kernel" -s 12 -c 5 -o $O/train_kernels python bench.py --workload train --steps 3 --warmup 3 --no-extra > $O/ncu_train.log 2>&1
$NCU --set full --import-source on -k "regex:known_score|known_compare|transe_query_kernel|bil_query_kernel|bil_table_kernel|metrics_kernel" -s 12 -c 6 -o $O/prepass_complex python bench.py --workload complex --steps 3 --warmup 3 --no-extra > $O/ncu_prepass.log 2>&1
# the single-product (FP16) mode of the bilinear path: the measurement behind DESIGN 3.2's negative result
for w in distmult complex; do for pr in 1 3; do python bench.py --workload $w --no-extra --steps 50 --opt bil_products=$pr > $O/bench_${w}_p$pr.json 2> $O/bench_${w}_p$pr.err; done; done
ls -la $O
