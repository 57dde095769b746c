#!/bin/bash
# GPU-box visit: parity tests + C2 / C3 bench lines.  usage (under gpurun): bash tools/gpu_round2.sh <tag>
TAG=${1:-r01c}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -5 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 > $OUT/bench_c2_tf32.json 2> $OUT/bench_c2_tf32.err
tail -c 300 $OUT/bench_c2_tf32.json; tail -3 $OUT/bench_c2_tf32.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload c3 > $OUT/bench_c3_tf32.json 2> $OUT/bench_c3_tf32.err
tail -c 300 $OUT/bench_c3_tf32.json; tail -3 $OUT/bench_c3_tf32.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload c3 --precision fp32 > $OUT/bench_c3_fp32.json 2> $OUT/bench_c3_fp32.err
tail -3 $OUT/bench_c3_fp32.err
ls -la $OUT
