#!/bin/bash
# One GPU visit for the round's evidence: launch lists of one bench step / one training step + full captures of the
# dominant kernels.  Numbers printed under ncu are never bench values.
mkdir -p gpurun_out
# (1) launch list of the forward step (per-kernel gpu__time_duration, serialised / cold-cache: shares, not absolutes);
#     eager launches so that every kernel of the step appears as its own row
ATLAS_B200_CUDA_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "atlas_b200_timed/" \
    --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --profile-step > gpurun_out/ncu_launch.log 2>&1
wc -l gpurun_out/launches.csv
# (2) launch list of one training step (FiD-base forward + backward, 1 query)
ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "atlas_b200_train/" \
    --csv --log-file gpurun_out/launches_train.csv python tools/perf_train.py launchlist > gpurun_out/ncu_launch_train.log 2>&1
wc -l gpurun_out/launches_train.csv
# (3) full captures: attention backward (dq kernel), MN-major weight-gradient GEMM
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_dq_kernel -s 2 -c 1 -o gpurun_out/prof_attn_bwd_dq -f \
    python tools/prof_ops.py attn_bwd 3 > gpurun_out/ncu_attn_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_bwd_dkv_kernel -s 2 -c 1 -o gpurun_out/prof_attn_bwd_dkv -f \
    python tools/prof_ops.py attn_bwd 3 >> gpurun_out/ncu_attn_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 -o gpurun_out/prof_wgrad -f \
    python tools/prof_ops.py wgrad 3 > gpurun_out/ncu_wgrad.log 2>&1
tail -2 gpurun_out/ncu_attn_bwd.log gpurun_out/ncu_wgrad.log
ls -la gpurun_out
