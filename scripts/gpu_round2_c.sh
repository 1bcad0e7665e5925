#!/bin/bash
# round-2 critical-path trims (one-launch row write-back, exposure backward early, gradient pool zeroed by the forward,
# shared-memory carve-out on the pack / gather / compositing kernels): tests, timeline, bench with and without
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/c_status.txt
timeout 200 python scripts/step_timeline.py > gpurun_out/c_timeline.txt 2>/dev/null; echo "timeline rc=$?" >> gpurun_out/c_status.txt
timeout 300 python bench.py --no-extras --steps 100 --warmup 10 > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo "bench rc=$?" >> gpurun_out/c_status.txt
HPS_EARLY_ZERO=0 timeout 300 python bench.py --no-extras --steps 100 --warmup 10 > gpurun_out/c_bench_noearly.json 2> gpurun_out/c_bench_noearly.err; echo "bench(no early zero) rc=$?" >> gpurun_out/c_status.txt
timeout 300 python bench.py --no-extras --steps 100 --warmup 10 > gpurun_out/c_bench2.json 2> gpurun_out/c_bench2.err; echo "bench2 rc=$?" >> gpurun_out/c_status.txt
cat gpurun_out/c_status.txt; tail -3 gpurun_out/c_pytest.log
python - <<'PY'
import json
for f in ("c_bench", "c_bench_noearly", "c_bench2"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("sustained", {}).get("value"), d["parity"].get("ok"))
    except Exception as e:
        print(f, "unreadable", e)
PY
