#!/bin/bash
# A/B of neighbour-search build variants on one box (build them first, here: scripts/build_knn_variants.sh A:64:6 C:128:6)
mkdir -p gpurun_out
L=hierarchical-point-slam_b200/libhps_b200.so
cp $L /tmp/libhps_base.so
for v in base A C base; do
    if [ $v = base ]; then cp /tmp/libhps_base.so $L; else cp hierarchical-point-slam_b200/build/variants/libhps_$v.so $L; fi
    for rep in 1 2; do
        timeout 300 python bench.py --no-extras --steps 200 --warmup 10 > gpurun_out/v_${v}_$rep.json 2> gpurun_out/v_${v}_$rep.err
        python - "$v" "$rep" <<'PY'
import json, sys
try:
    d = json.loads(open("gpurun_out/v_%s_%s.json" % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
    print(sys.argv[1], sys.argv[2], "%.4f ms" % d["ms_per_step"], "sustained %.3f M" % (d["sustained"]["value"] / 1e6), "knn %.1f us" % (1e3 * d["kernels_ms_per_step"]["hps_knn_query"]), d["parity"].get("ok"), d["parity"].get("knn_bit_exact"))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
    done
    timeout 200 python scripts/knn_group_bench.py 2>/dev/null | grep "group 1" | grep " 5000 \|307200" 
done
cp /tmp/libhps_base.so $L
