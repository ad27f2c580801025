#!/bin/bash
# Round 2, visit E (after the container was re-created and the earlier visits' outputs were lost): full GPU suite on the
# three-way attention dispatch (default = 96-key lanes kernel), attention A/B/C timing, TMEM + softmax microbenchmarks,
# the bench line of both arms, one ncu capture of the lanes kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/e_smi.txt 2>&1
timeout 1100 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/e_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/e_suite.log
tail -15 gpurun_out/e_suite.log
{
echo "== lanes v1 (96-key)"; python tools/prof_ops.py attention 10
echo "== first generation"; ATLAS_B200_ATTN_LANES=0 python tools/prof_ops.py attention 10
echo "== lanes v2 (48-key, double-buffered S)"; ATLAS_B200_ATTN_LANES=2 python tools/prof_ops.py attention 10
for d in 1 2 4 8 15; do echo "== lanes v2 debug=$d"; ATLAS_B200_ATTN_LANES=2 ATLAS_B200_ATTN_DEBUG=$d python tools/prof_ops.py attention 10; done
echo "== attention bwd"; python tools/prof_ops.py attn_bwd 10
echo "== gemm"; python tools/prof_ops.py gemm 10
} > gpurun_out/e_attn.log 2>&1
cat gpurun_out/e_attn.log
./tools/ubench/tmem_ubench > gpurun_out/e_tmem_ubench.log 2>&1; cat gpurun_out/e_tmem_ubench.log
./tools/ubench/softmax_ubench > gpurun_out/e_softmax_ubench.log 2>&1; cat gpurun_out/e_softmax_ubench.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/e_bench.err; head -c 6000 gpurun_out/e_bench.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 --ref-budget-s 45 > gpurun_out/e_ref.json 2> gpurun_out/e_ref.err
echo "ref rc=$?"; tail -3 gpurun_out/e_ref.err; head -c 1500 gpurun_out/e_ref.json
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_lanes_kernel -s 2 -c 1 -f -o gpurun_out/prof_attn_lanes python tools/prof_ops.py attention 3 > gpurun_out/e_ncu.log 2>&1
tail -2 gpurun_out/e_ncu.log
