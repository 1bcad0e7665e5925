#!/bin/bash
# multi-GPU probes: usage scripts/gpu_round2_mg.sh NGPUS
N=${1:-2}
mkdir -p gpurun_out
for mode in ${MODES:-nccl_eager own iteration}; do
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 scripts/multi_gpu_probe.py $mode > gpurun_out/mg${N}_$mode.txt 2>&1
  echo "$mode rc=$?" >> gpurun_out/mg${N}_status.txt
done
if [ -n "${BENCH:-}" ]; then
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 ${BENCH_FLAGS:-} > gpurun_out/mg${N}_bench.json 2> gpurun_out/mg${N}_bench.err
  echo "bench rc=$?" >> gpurun_out/mg${N}_status.txt
  HPS_BENCH_EXCHANGE=nccl timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus $N --steps 20 --warmup 5 --no-extras > gpurun_out/mg${N}_bench_nccl.json 2> gpurun_out/mg${N}_bench_nccl.err
  echo "bench(nccl) rc=$?" >> gpurun_out/mg${N}_status.txt
fi
cat gpurun_out/mg${N}_status.txt
grep -h "^\[" gpurun_out/mg${N}_*.txt
