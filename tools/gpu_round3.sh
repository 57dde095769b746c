#!/bin/bash
# GPU-box visit for the two-pass fp16 kNN: exactness tests of all variants, timings, per-kernel launch list, full captures.
# usage (under gpurun): bash tools/gpu_round3.sh <tag> [full]
TAG=${1:-r01g}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests/test_kernels_gpu.py -m gpu -x -q -k "knn_tensor_core" > $OUT/pytest_knn.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_knn.log
tail -15 $OUT/pytest_knn.log
timeout 300 python tools/time_knn_variants.py 64 4096 20 > $OUT/time_knn_c2.log 2>&1; cat $OUT/time_knn_c2.log
timeout 300 python tools/time_knn_variants.py 16 16384 32 > $OUT/time_knn_c5.log 2>&1; cat $OUT/time_knn_c5.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:knn2 -c 200 --csv --log-file $OUT/launches_knn.csv \
    python tools/time_knn_variants.py 64 4096 20 1 > $OUT/ncu_knn.log 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open("$OUT/launches_knn.csv")) if len(r) > 5 and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4][:70]].append(float(r[-1].replace(",", "")))
for k, v in agg.items():
    print(f"{k:72s} n={len(v):3d} median={sorted(v)[len(v)//2]:10.1f}")
PY
if [ "$2" = "full" ]; then
  for v in 1 2; do
    timeout 600 ncu --set full --clock-control none --import-source on -k regex:"knn2_tc_kernel|knn2_refine" --launch-skip 4 -c 2 \
        -o $OUT/knn2_v$v -f python tools/time_knn_variants.py 64 4096 20 $v > $OUT/ncu_full_v$v.log 2>&1
    tail -2 $OUT/ncu_full_v$v.log
  done
fi
