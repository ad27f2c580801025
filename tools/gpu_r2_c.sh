#!/bin/bash
# Round 2, visit C: three-lane attention kernel v2 (48-key blocks, double-buffered S): correctness + timing breakdown
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_fullsize_gpu.py -q -p no:cacheprovider -k "attention or untied or fid_base" 2>&1 | tail -8 > gpurun_out/c_tests.log
cat gpurun_out/c_tests.log
{
echo "== lanes kernel"; python tools/prof_ops.py attention 10
for d in 1 2 4 3 7 15; do echo "== lanes debug=$d"; ATLAS_B200_ATTN_DEBUG=$d python tools/prof_ops.py attention 10; done
} > gpurun_out/c_attn.log 2>&1
cat gpurun_out/c_attn.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_lanes_kernel -s 2 -c 1 -f -o gpurun_out/prof_attn_lanes2 python tools/prof_ops.py attention 3 > gpurun_out/c_ncu.log 2>&1
tail -2 gpurun_out/c_ncu.log
