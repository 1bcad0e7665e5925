#!/bin/bash
mkdir -p gpurun_out
HPS_TIMELINE_REPS=3 timeout 200 python scripts/step_timeline.py > gpurun_out/k_timeline.txt 2> gpurun_out/k_err1.txt
run() {  # name, env...
    name=$1; shift
    env "$@" timeout 300 python bench.py --no-extras --steps 200 --warmup 10 > gpurun_out/k_$name.json 2> gpurun_out/k_$name.err
    python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/k_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "%.4f ms" % d["ms_per_step"], "sustained %.3f M" % (d["sustained"]["value"] / 1e6), "e2e %.3f M" % (d["e2e"]["value"] / 1e6), d["parity"].get("ok"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
run a X=1
run b X=1
grep span gpurun_out/k_timeline.txt
