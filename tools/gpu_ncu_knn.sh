#!/bin/bash
# full ncu capture of the two-pass kNN kernels (variant $2, default 2).  usage (under gpurun): bash tools/gpu_ncu_knn.sh <tag> [variant]
TAG=${1:-r01j}; V=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"knn2_tc_kernel|knn2_refine" --launch-skip 4 -c 2 \
    -o $OUT/knn2_v$V -f python tools/time_knn_variants.py 64 4096 20 $V > $OUT/ncu_full_v$V.log 2>&1
tail -2 $OUT/ncu_full_v$V.log
