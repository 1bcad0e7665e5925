#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 scripts/step_timeline_mg.py > gpurun_out/i${N}_timeline_exchange.txt 2> gpurun_out/i${N}_err1.txt; echo rc=$?
HPS_BENCH_HEAD_AT=tail timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29545 scripts/step_timeline_mg.py > gpurun_out/i${N}_timeline_tail.txt 2> gpurun_out/i${N}_err2.txt; echo rc=$?
grep span gpurun_out/i${N}_timeline_*.txt
