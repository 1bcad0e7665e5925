#!/bin/bash
# Run on the GPU box via gpurun.  Launch list (per-launch device time) + one full capture of a chosen kernel.
# Usage: scripts/gpu_profile.sh <tag> <kernel-regex (demangled)> [skip] [count]
set -u
TAG=${1:-r02}
KREGEX=${2:-'mlp_dgrad_tc_kernel<\(int\)1>'}
SKIP=${3:-700}
COUNT=${4:-400}
FULLCOUNT=${5:-1}
mkdir -p gpurun_out
CMD="python bench.py --steps 3 --warmup 3 --no-extras"
timeout 280 $CMD > gpurun_out/plain_${TAG}.log 2>&1 || { echo "plain run failed"; tail -20 gpurun_out/plain_${TAG}.log; exit 1; }
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -s $SKIP -c $COUNT --csv \
    --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu_launch_${TAG}.log 2>&1
echo "launch list rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:${KREGEX}" -s 3 -c $FULLCOUNT \
    -f -o gpurun_out/prof_${TAG} $CMD > gpurun_out/ncu_full_${TAG}.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/ | grep ${TAG}
