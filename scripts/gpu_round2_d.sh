#!/bin/bash
# scheduling of the backward tail after the critical-path trims: timelines (3 replays each) + bench per variant
mkdir -p gpurun_out
HPS_EARLY_ZERO=0 HPS_TIMELINE_REPS=3 HPS_GRAPH_DOT=gpurun_out/d_graph_noearly.dot timeout 200 python scripts/step_timeline.py > gpurun_out/d_timeline_noearly.txt 2>/dev/null
HPS_EARLY_ZERO=1 HPS_TIMELINE_REPS=3 HPS_GRAPH_DOT=gpurun_out/d_graph_early.dot timeout 200 python scripts/step_timeline.py > gpurun_out/d_timeline_early.txt 2>/dev/null
run() {  # name, env...
    name=$1; shift
    env "$@" timeout 300 python bench.py --no-extras --steps 100 --warmup 10 > gpurun_out/d_$name.json 2> gpurun_out/d_$name.err
    python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/d_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "%.4f ms" % d["ms_per_step"], "sustained %.3f M" % (d["sustained"]["value"] / 1e6), "e2e %.3f M" % (d["e2e"]["value"] / 1e6), d["parity"].get("ok"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
run noearly HPS_EARLY_ZERO=0
run noearly_scatter HPS_EARLY_ZERO=0 HPS_SCATTER_STREAM=1
run noearly_after HPS_EARLY_ZERO=0 HPS_TAIL_AFTER_SCATTER=1
run noearly_tailprio HPS_EARLY_ZERO=0 HPS_TAIL_PRIORITY=-1
run early_scatter HPS_EARLY_ZERO=1 HPS_SCATTER_STREAM=1
run early_after HPS_EARLY_ZERO=1 HPS_TAIL_AFTER_SCATTER=1
run early_tailprio HPS_EARLY_ZERO=1 HPS_TAIL_PRIORITY=-1
run noearly_again HPS_EARLY_ZERO=0
