#!/bin/bash
# Run on the GPU box via gpurun.  Launch list (per-launch device time) + one full capture of the top kernel.
# Usage: scripts/gpu_profile.sh <tag> <kernel-regex (demangled)>
set -u
TAG=${1:-r01}
KREGEX=${2:-'mlp_bwd_kernel<\(int\)1>'}
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3"
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_${TAG}.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -s 600 -c 500 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${KREGEX}" -s 3 -c 1 \
    -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/ | tail -12
