#!/bin/bash
# per-launch time of the fused EdgeConv kernel inside two C2 steps (under gpurun)
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:edgeconv_dg -c 4 --csv --log-file gpurun_out/edge.csv python tools/prof_step.py tf32 64 > /dev/null 2>&1
grep edgeconv gpurun_out/edge.csv | awk -F'","' '{print $NF}' | tr -d '"' | tr '\n' ' '; echo
