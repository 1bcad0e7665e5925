#!/bin/bash
# N GPUs: bench with the next head beside the backward tail (default) and beside the exchange
N=${1:-8}
mkdir -p gpurun_out
run() {  # name, env...
    name=$1; shift
    env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 100 --warmup 10 --no-extras > gpurun_out/hp${N}_$name.json 2> gpurun_out/hp${N}_$name.err
    echo "$name rc=$?"
    python - "gpurun_out/hp${N}_$name.json" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "%.4f ms" % d["ms_per_step"], "%.2f M rays/s" % (d["value"] / 1e6), "sustained %.2f M" % (d["sustained"]["value"] / 1e6), "e2e %.2f M" % (d["e2e"]["value"] / 1e6), d.get("exchange"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
run tail X=1
run exchange HPS_BENCH_HEAD_AT=exchange
