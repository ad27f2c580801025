#!/bin/bash
# Round 2, visit G: three-lane attention with its own P columns + early S issue (correctness, A/B against the 96-key version
# and the first generation), GEMM with L2 prefetch of the next work item's A rows (A/B), tests of the work since visit F.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_fullsize_gpu.py tests/test_dropout_gpu.py tests/test_models_gpu.py "tests/test_search_gpu.py::test_streamed_index_load" tests/test_linear_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/g_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/g_tests.log; tail -12 gpurun_out/g_tests.log
{
echo "== lanes (64-key, separate P, early S)"; timeout 120 python tools/prof_ops.py attention 10
echo "== lanes96"; ATLAS_B200_ATTN_LANES=3 timeout 120 python tools/prof_ops.py attention 10
echo "== first generation"; ATLAS_B200_ATTN_LANES=0 timeout 120 python tools/prof_ops.py attention 10
} > gpurun_out/g_attn.log 2>&1
cat gpurun_out/g_attn.log
{
echo "== pair + L2 prefetch (default)"; timeout 200 python tools/prof_ops.py gemm4 10
echo "== pair, no prefetch"; ATLAS_B200_GEMM_PREFETCH=0 timeout 200 python tools/prof_ops.py gemm4 10
} > gpurun_out/g_gemm.log 2>&1
cat gpurun_out/g_gemm.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/g_bench.err; python - <<'PY'
import json
try:
    l = json.load(open("gpurun_out/g_bench.json"))
    print("value", l["value"], "ms", l["ms_per_step"], "e2e", l["e2e"]["value"], "roofline", l["roofline"]["achieved"], l["roofline"]["frac"],
          "attn", l["roofline"]["attention_kernel"], "gemm ms", l["roofline"]["kernel_ms_per_step"])
    print("train", l["train"].get("value"), l["train"].get("ms_per_step"), "xl", str(l["train_xl"])[:200])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_lanes_kernel -s 2 -c 1 -f -o gpurun_out/prof_attn_lanes64 python tools/prof_ops.py attention 3 > gpurun_out/g_ncu.log 2>&1
tail -2 gpurun_out/g_ncu.log
