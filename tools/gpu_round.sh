#!/bin/bash
# One GPU-box visit: parity tests, bench lines, ncu launch list + full captures of the top kernels.
# usage (under gpurun): bash tools/gpu_round.sh <tag> [skip_tests]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.csv 2>&1
if [ -z "$2" ]; then
  timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
  echo "pytest rc=$?" >> $OUT/pytest_gpu.log
  tail -5 $OUT/pytest_gpu.log
fi
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_tf32.json 2> $OUT/bench_tf32.err
tail -c 600 $OUT/bench_tf32.json
timeout 600 python bench.py --steps 10 --warmup 3 --precision fp32 > $OUT/bench_fp32.json 2> $OUT/bench_fp32.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_tf32.csv \
    python tools/prof_step.py tf32 64 > $OUT/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn|edgeconv|gemm_tc' -c 14 \
    -o $OUT/prof_top python tools/prof_step.py tf32 64 > $OUT/ncu_full.log 2>&1
ls -la $OUT
