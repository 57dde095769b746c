#!/bin/bash
# GPU-box visit: kNN variant timings, full parity suite, C2 / C3 bench lines.  usage (under gpurun): bash tools/gpu_round4.sh <tag>
TAG=${1:-r01m}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 300 python tools/time_knn_variants.py 64 4096 20 > $OUT/time_knn_c2.log 2>&1; cat $OUT/time_knn_c2.log
timeout 1500 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> $OUT/pytest_gpu.log
tail -6 $OUT/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench_c2_tf32.json 2> $OUT/bench_c2_tf32.err
python - <<PY
import json
d = json.loads(open("$OUT/bench_c2_tf32.json").read().strip().splitlines()[-1])
print("C2", d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"])
for k, v in d["kernel_breakdown"].items(): print("   ", k, v["ms_per_step"])
PY
tail -3 $OUT/bench_c2_tf32.err
timeout 600 python bench.py --steps 10 --warmup 3 --workload c3 > $OUT/bench_c3_tf32.json 2> $OUT/bench_c3_tf32.err
python - <<PY
import json
d = json.loads(open("$OUT/bench_c3_tf32.json").read().strip().splitlines()[-1])
print("C3", d["value"], d["ms_per_step"], d["e2e"]["value"])
for k, v in list(d["kernel_breakdown"].items())[:12]: print("   ", k, v["ms_per_step"])
PY
tail -3 $OUT/bench_c3_tf32.err
