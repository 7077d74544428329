#!/bin/bash
# Round 2, multi-GPU call:  gpurun --gpus N -- 'bash scripts/r2_multi.sh N'
# Weak-scaling bench line at N GPUs (eager launches + one NCCL all-reduce) and the BASELINE configs[4] sweep (DE vs D
# LNCDE, depth-3 rows).  (The all-reduce captured inside the CUDA graph - LNCDE_BENCH_GRAPH_NCCL=1 - hung at N = 2 in
# round 2 as it had at N = 8 in round 1, so it is no longer part of this script; the eager step is within 0.1 % of it.)
set -u
N=${1:-2}
O=gpurun_out
mkdir -p $O
P=$((29500 + N))
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port"
$TR $P bench.py --gpus $N --steps 20 --warmup 3 > $O/r2_bench_n$N.json 2> $O/r2_bench_n$N.err
timeout 900 $TR $((P + 20)) scripts/lncde_sweep.py --quick --out $O/r2_lncde_sweep_n$N.json > $O/r2_lncde_sweep_n$N.log 2>&1
tail -c 600 $O/r2_bench_n$N.json; echo; tail -2 $O/r2_lncde_sweep_n$N.log | cut -c1-600
