#!/bin/bash
# N GPUs: the split exchange (feature gradients behind the scatter) against the single one; peer test; timeline
N=${1:-2}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_render.py -m gpu -x -q -k "peer or live or fused" > gpurun_out/j${N}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/j${N}_pytest.log
run() {  # name, env...
    name=$1; shift
    env "$@" timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 100 --warmup 10 --no-extras > gpurun_out/j${N}_$name.json 2> gpurun_out/j${N}_$name.err
    echo "$name rc=$?"
    python - "gpurun_out/j${N}_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "%.4f ms" % d["ms_per_step"], "%.2f M rays/s" % (d["value"] / 1e6), "sustained %.2f M" % (d["sustained"]["value"] / 1e6), "e2e %.2f M" % (d["e2e"]["value"] / 1e6), d.get("exchange"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
run split X=1
run single HPS_BENCH_SPLIT_EXCHANGE=0
run split2 X=1
run split_headex HPS_BENCH_HEAD_AT=exchange
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 scripts/step_timeline_mg.py > gpurun_out/j${N}_timeline.txt 2> gpurun_out/j${N}_err1.txt; echo rc=$?
grep span gpurun_out/j${N}_timeline.txt
