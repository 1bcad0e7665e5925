#!/bin/bash
# round-2 end state on one GPU: whole GPU suite, smoke, the driver's bench command, the CPU arm, timelines
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/z_pytest.log 2>&1; echo "pytest rc=$?" > gpurun_out/z_status.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.txt 2>&1; echo "smoke rc=$?" >> gpurun_out/z_status.txt
timeout 600 python bench.py > gpurun_out/z_bench.json 2> gpurun_out/z_bench.err; echo "bench rc=$?" >> gpurun_out/z_status.txt
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/z_bench_ref.json 2> gpurun_out/z_bench_ref.err; echo "bench ref rc=$?" >> gpurun_out/z_status.txt
HPS_TIMELINE_REPS=3 timeout 200 python scripts/step_timeline.py > gpurun_out/z_timeline.txt 2>/dev/null
cat gpurun_out/z_status.txt
