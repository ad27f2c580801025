#!/bin/bash
# Round 2, visit B: where does the three-lane attention kernel spend its time?
mkdir -p gpurun_out
{
echo "== lanes kernel"; python tools/prof_ops.py attention 10
echo "== first-generation kernel"; ATLAS_B200_ATTN_LANES=0 python tools/prof_ops.py attention 10
for d in 1 2 4 8 3 7 15; do echo "== lanes debug=$d"; ATLAS_B200_ATTN_DEBUG=$d python tools/prof_ops.py attention 10; done
} > gpurun_out/b_attn.log 2>&1
cat gpurun_out/b_attn.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_lanes_kernel -s 2 -c 1 -f -o gpurun_out/prof_attn_lanes python tools/prof_ops.py attention 3 > gpurun_out/b_ncu.log 2>&1
tail -3 gpurun_out/b_ncu.log
timeout 300 python -m pytest tests/test_fullsize_gpu.py -q -k untied -p no:cacheprovider 2>&1 | tail -3
