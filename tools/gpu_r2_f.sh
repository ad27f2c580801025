#!/bin/bash
# Round 2, visit F: 4-CTA-cluster multicast GEMM (correctness + A/B against the pair kernel and cuBLAS), dropout / optimiser
# tests, the three tests fixed after visit E, launch list of the bench step.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_linear_gpu.py tests/test_dropout_gpu.py tests/test_optim_gpu.py -q -x --timeout 300 -p no:cacheprovider > gpurun_out/f_new_tests.log 2>&1
echo "new tests rc=$?" >> gpurun_out/f_new_tests.log; tail -12 gpurun_out/f_new_tests.log
{
echo "== quad (default)"; timeout 200 python tools/prof_ops.py gemm4 10
echo "== pair (ATLAS_B200_GEMM_QUAD=0)"; ATLAS_B200_GEMM_QUAD=0 timeout 200 python tools/prof_ops.py gemm4 10
} > gpurun_out/f_gemm.log 2>&1
cat gpurun_out/f_gemm.log
timeout 600 python -m pytest tests/test_fullsize_gpu.py tests/test_round2_gpu.py tests/test_models_gpu.py tests/test_train_gpu.py -q --timeout 300 -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/f_tests2.log
cat gpurun_out/f_tests2.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f_bench.json 2> gpurun_out/f_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/f_bench.err; python - <<'PY'
import json
try:
    l = json.load(open("gpurun_out/f_bench.json"))
    print("value", l["value"], "ms", l["ms_per_step"], "e2e", l["e2e"]["value"], "roofline", l["roofline"]["achieved"], l["roofline"]["frac"],
          "attn", l["roofline"]["attention_kernel"]["ms_per_step"], "gemm ms", l["roofline"]["kernel_ms_per_step"])
    print("train", l["train"].get("value"), l["train"].get("ms_per_step"), "xl", str(l["train_xl"])[:300])
except Exception as e:
    print("bench parse failed", e)
PY
ATLAS_B200_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "atlas_b200_timed/" --csv --log-file gpurun_out/f_launches_step.csv python bench.py --steps 1 --warmup 3 --profile-step > gpurun_out/f_ncu_launch.log 2>&1
wc -l gpurun_out/f_launches_step.csv
