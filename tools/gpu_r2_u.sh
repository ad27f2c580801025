#!/bin/bash
# Round 2, visit U: evidence for the padding-compacted encoder - launch list of one timed step (eager launches under ncu), one
# `ncu --set full` capture of the packed three-lane attention kernel.
mkdir -p gpurun_out
ATLAS_B200_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "atlas_b200_timed/" --csv --log-file gpurun_out/u_launches_step.csv python bench.py --steps 1 --warmup 3 --profile-step > gpurun_out/u_ncu_launch.log 2>&1
wc -l gpurun_out/u_launches_step.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_lanes_kernel -s 4 -c 1 -f -o gpurun_out/prof_lanes_packed python tools/prof_ops.py packed 3 > gpurun_out/u_ncu_lanes_packed.log 2>&1
tail -2 gpurun_out/u_ncu_lanes_packed.log | cut -c1-300
ls -la gpurun_out/prof_lanes_packed.ncu-rep
