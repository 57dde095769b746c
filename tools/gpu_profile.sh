#!/bin/bash
# ncu evidence for the C2 step: launch list (gpu__time_duration) + one full capture of the top kernels.
# usage (under gpurun): bash tools/gpu_profile.sh <tag>
TAG=${1:-r01r}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_c2_tf32.csv \
    python tools/prof_step.py tf32 64 > $OUT/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'knn2_|edgeconv_dg_tc|knn_grid_search|gemm_tf32_kernel|edge_gather_ext' -c 24 \
    -o $OUT/prof_top python tools/prof_step.py tf32 64 > $OUT/ncu_full.log 2>&1
timeout 300 ncu -i $OUT/prof_top.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct > $OUT/ncu_full_top_kernels.csv 2>/dev/null
timeout 300 python bench.py --steps 20 --warmup 3 > $OUT/bench_c2_tf32.json 2> $OUT/bench_c2_tf32.err
timeout 300 python bench.py --steps 10 --warmup 3 --precision fp32 > $OUT/bench_c2_fp32.json 2> $OUT/bench_c2_fp32.err
timeout 300 python bench.py --steps 10 --warmup 3 --workload c3 > $OUT/bench_c3_tf32.json 2> $OUT/bench_c3_tf32.err
ls -la $OUT
