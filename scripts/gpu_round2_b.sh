#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_scale.py tests/test_gpu_kernels.py -m gpu -q -s > gpurun_out/b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/b_status.txt
timeout 300 python scripts/step_timeline.py > gpurun_out/b_timeline_overlap.txt 2> gpurun_out/b_timeline.err; echo "timeline rc=$?" >> gpurun_out/b_status.txt
HPS_BENCH_OVERLAP=0 timeout 300 python scripts/step_timeline.py > gpurun_out/b_timeline_plain.txt 2>> gpurun_out/b_timeline.err; echo "timeline plain rc=$?" >> gpurun_out/b_status.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/b_bench.json 2> gpurun_out/b_bench.err; echo "bench rc=$?" >> gpurun_out/b_status.txt
cat gpurun_out/b_status.txt
