#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/f_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/f_status.txt
HPS_TIMELINE_REPS=2 timeout 200 python scripts/step_timeline.py > gpurun_out/f_timeline.txt 2> gpurun_out/f_err1.txt
run() {  # name, env...
    name=$1; shift
    env "$@" timeout 300 python bench.py --no-extras --steps 100 --warmup 10 > gpurun_out/f_$name.json 2> gpurun_out/f_$name.err
    python - "$name" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/f_%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "%.4f ms" % d["ms_per_step"], "sustained %.3f M" % (d["sustained"]["value"] / 1e6), "e2e %.3f M" % (d["e2e"]["value"] / 1e6), d["parity"].get("ok"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
}
run default X=1
run nozero HPS_ZERO_IN_COMPOSITE=0
run tailprio HPS_TAIL_PRIORITY=-1
run nozero_tailprio HPS_ZERO_IN_COMPOSITE=0 HPS_TAIL_PRIORITY=-1
run default2 X=1
run tailprio2 HPS_TAIL_PRIORITY=-1
cat gpurun_out/f_status.txt; tail -3 gpurun_out/f_pytest.log; grep span gpurun_out/f_timeline.txt
