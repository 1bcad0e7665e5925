#!/bin/bash
# Build neighbour-search variants of libhps_b200.so for scripts/gpu_knn_variants.sh (run HERE, before gpurun: the built
# files travel with the snapshot).  Each argument: NAME:STAGE_CELLS:CTAS_PER_SM, e.g.  A:64:6 C:128:6
set -e
cd "$(dirname "$0")/../hierarchical-point-slam_b200"
python build.py > /dev/null
mkdir -p build/variants
for v in "$@"; do
    IFS=: read -r name cells ctas <<< "$v"
    nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -fvisibility=default \
        -DHPS_KNN_STAGE_CELLS=$cells -DHPS_KNN_CTAS_PER_SM=$ctas -Xptxas -v -c csrc/grid.cu -o build/variants/grid_$name.o 2>&1 \
        | grep -A2 "knn_query_kernelILi1ELb1" | grep "Used\|spill"
    objs=$(ls build/*.o | grep -v "/grid.o")
    nvcc -gencode arch=compute_100a,code=sm_100a --shared -o build/variants/libhps_$name.so $objs build/variants/grid_$name.o -lcudart
    echo "built build/variants/libhps_$name.so (stage cells $cells, $ctas CTAs per SM)"
done
