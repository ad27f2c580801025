#!/bin/bash
# Round 2, visit H: GEMM epilogue through shared memory + TMA stores (correctness, A/B), tcgen05 dK/dV kernel behind the
# warp-MMA dQ kernel through the backward / training suites (ATLAS_B200_ATTN_BWD_TC=3), bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_backward_gpu.py tests/test_models_gpu.py tests/test_fullsize_gpu.py -q -x --timeout 300 -p no:cacheprovider > gpurun_out/h_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/h_tests.log; tail -8 gpurun_out/h_tests.log
{
echo "== TMA-store epilogue (default)"; timeout 200 python tools/prof_ops.py gemm4 10
echo "== direct stores (ATLAS_B200_GEMM_TMA_STORE=0)"; ATLAS_B200_GEMM_TMA_STORE=0 timeout 200 python tools/prof_ops.py gemm4 10
} > gpurun_out/h_gemm.log 2>&1
cat gpurun_out/h_gemm.log
ATLAS_B200_ATTN_BWD_TC=3 timeout 900 python -m pytest tests/test_backward_gpu.py tests/test_train_gpu.py tests/test_dropout_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/h_tests_tc3.log 2>&1
echo "tc3 tests rc=$?" >> gpurun_out/h_tests_tc3.log; tail -6 gpurun_out/h_tests_tc3.log
{
echo "== attention bwd (warp-MMA)"; timeout 120 python tools/prof_ops.py attn_bwd 10
echo "== attention bwd TC=3"; ATLAS_B200_ATTN_BWD_TC=3 timeout 120 python tools/prof_ops.py attn_bwd 10
echo "== attention bwd TC=2"; ATLAS_B200_ATTN_BWD_TC=2 timeout 120 python tools/prof_ops.py attn_bwd 10
} > gpurun_out/h_attn_bwd.log 2>&1
cat gpurun_out/h_attn_bwd.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference > gpurun_out/h_bench.json 2> gpurun_out/h_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/h_bench.err | cut -c1-200; python - <<'PY'
import json
try:
    l = json.load(open("gpurun_out/h_bench.json"))
    print("value", l["value"], "ms", l["ms_per_step"], "e2e", l["e2e"]["value"], "roofline", l["roofline"]["achieved"], l["roofline"]["frac"],
          "attn", l["roofline"]["attention_kernel"]["ms_per_step"], "gemm ms", l["roofline"]["kernel_ms_per_step"])
    print("train", l["train"].get("value"), l["train"].get("ms_per_step"), str(l["train"].get("kernels"))[:300])
    print("mips", l["mips"]["value"], l["mips"]["ms_per_step"], l["mips"]["roofline"]["frac"])
except Exception as e:
    print("bench parse failed", e)
PY
