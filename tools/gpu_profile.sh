#!/bin/bash
# ncu evidence for the C2 step: launch list (gpu__time_duration) + one full capture of the top kernels + the bench lines.
# usage (under gpurun): bash tools/gpu_profile.sh <tag> [precision]      (precision: f16 (default) | tf32 | fp32)
TAG=${1:-r02}
PREC=${2:-f16}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/launches_c2_$PREC.csv \
    python tools/prof_step.py $PREC 64 > $OUT/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:'knn2_|knn_grid_lockstep|knn_xyz_rescue|edgeconv_dg|edge_gather|gemm_tf32_kernel|gemm_h2|pointwise_mlp2|vlad_finish|hidden_gate' -c 44 \
    -o $OUT/prof_top python tools/prof_step.py $PREC 64 > $OUT/ncu_full.log 2>&1
timeout 300 ncu -i $OUT/prof_top.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed,launch__registers_per_thread,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct > $OUT/ncu_full_top_kernels.csv 2>/dev/null
# gpurun copies at most 64 MiB back: keep the report only if it is small (the CSV above holds what profiles/ needs)
[ $(stat -c %s $OUT/prof_top.ncu-rep 2>/dev/null || echo 0) -gt 40000000 ] && rm -f $OUT/prof_top.ncu-rep
timeout 600 python bench.py --steps 20 --warmup 3 > $OUT/bench_all_$PREC.json 2> $OUT/bench_all_$PREC.err
ls -la $OUT
