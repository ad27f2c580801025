#!/bin/bash
# Second profiling visit: launch list of one training step + full captures of the attention backward kernels.
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off \
    --csv --log-file gpurun_out/launches_train.csv python tools/perf_train.py launchlist > gpurun_out/ncu_launch_train.log 2>&1
wc -l gpurun_out/launches_train.csv
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_dq2_kernel -s 2 -c 1 -o gpurun_out/prof_attn_bwd_dq -f \
    python tools/prof_ops.py attn_bwd 3 > gpurun_out/ncu_attn_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_dkv2_kernel -s 2 -c 1 -o gpurun_out/prof_attn_bwd_dkv -f \
    python tools/prof_ops.py attn_bwd 3 >> gpurun_out/ncu_attn_bwd.log 2>&1
grep -h "attention bwd\|Report" gpurun_out/ncu_attn_bwd.log
ls -la gpurun_out | head -20
