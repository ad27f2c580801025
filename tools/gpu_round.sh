#!/bin/bash
# One GPU visit for the round's evidence: launch list of one bench step + full captures of the dominant kernels.
mkdir -p gpurun_out
# (1) launch list (per-kernel gpu__time_duration, serialised / cold-cache: shares, not absolutes); eager launches so that
#     every kernel of the step appears as its own row
ATLAS_B200_CUDA_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "atlas_b200_timed/" \
    --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --profile-step > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-300
wc -l gpurun_out/launches.csv
# (2) full capture of the CTA-pair GEMM
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 2 -c 1 -o gpurun_out/prof_gemm2 -f \
    python tools/prof_ops.py gemm 3 > gpurun_out/ncu_gemm.log 2>&1
tail -2 gpurun_out/ncu_gemm.log
ls -la gpurun_out
