#!/bin/bash
# round 2, GPU call A: full GPU test suite, search-kernel A/B, bench with extras, bench A/Bs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=12 -x --deselect tests/test_gpu_scale.py > gpurun_out/a_pytest_old.log 2>&1; echo "pytest(old) rc=$?" >> gpurun_out/a_status.txt
timeout 900 python -m pytest tests/test_gpu_scale.py -m gpu -q -s > gpurun_out/a_pytest_scale.log 2>&1; echo "pytest(scale) rc=$?" >> gpurun_out/a_status.txt
timeout 300 python scripts/knn_group_bench.py > gpurun_out/a_knn.txt 2>&1; echo "knn rc=$?" >> gpurun_out/a_status.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err; echo "bench rc=$?" >> gpurun_out/a_status.txt
HPS_BENCH_OVERLAP=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/a_bench_nooverlap.json 2> gpurun_out/a_bench_nooverlap.err; echo "bench(no overlap) rc=$?" >> gpurun_out/a_status.txt
HPS_KNN_IMPL=warp timeout 300 python bench.py --steps 20 --warmup 5 --no-extras > gpurun_out/a_bench_knnwarp.json 2> gpurun_out/a_bench_knnwarp.err; echo "bench(knn warp) rc=$?" >> gpurun_out/a_status.txt
cat gpurun_out/a_status.txt
