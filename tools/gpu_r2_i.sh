#!/bin/bash
# Round 2, visit I: full GPU suite on the current defaults (lanes kernel with deferred output, double-buffered split-KV kernel,
# tcgen05 dK/dV backward by default, TMA-store GEMM epilogue), full-size parity numbers, A/Bs, bench + launch list.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/i_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/i_suite.log; tail -8 gpurun_out/i_suite.log
timeout 600 python -m pytest tests/test_fullsize_gpu.py -q -s -p no:cacheprovider 2>&1 | grep -E "full size|Contriever-base|passed|failed" > gpurun_out/i_fullsize_parity.log
cat gpurun_out/i_fullsize_parity.log
{
echo "== lanes (deferred output)"; timeout 120 python tools/prof_ops.py attention 10
echo "== attention bwd (default: TC=3)"; timeout 120 python tools/prof_ops.py attn_bwd 10
echo "== gemm4 default"; timeout 200 python tools/prof_ops.py gemm4 10
echo "== gemm4 quad cluster"; ATLAS_B200_GEMM_QUAD=1 timeout 200 python tools/prof_ops.py gemm4 10
} > gpurun_out/i_ops.log 2>&1
cat gpurun_out/i_ops.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err
echo "bench rc=$?"; tail -2 gpurun_out/i_bench.err | cut -c1-200; python - <<'PY'
import json
try:
    l = json.load(open("gpurun_out/i_bench.json"))
    print("value", l["value"], "ms", l["ms_per_step"], "e2e", l["e2e"]["value"], "roofline", l["roofline"]["achieved"], l["roofline"]["frac"],
          "attn", l["roofline"]["attention_kernel"]["ms_per_step"], "gemm ms", l["roofline"]["kernel_ms_per_step"])
    print("train", l["train"].get("value"), l["train"].get("ms_per_step"), str(l["train"].get("kernels"))[:400])
    print("mips", l["mips"]["value"], l["mips"]["ms_per_step"], l["mips"]["roofline"]["frac"])
    print("gpu_reference", l.get("gpu_reference", {}).get("value"), l.get("gpu_reference", {}).get("ours_over_reference_e2e"))
    print("xl", str(l["train_xl"])[:300])
except Exception as e:
    print("bench parse failed", e)
PY
ATLAS_B200_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "atlas_b200_timed/" --csv --log-file gpurun_out/i_launches_step.csv python bench.py --steps 1 --warmup 3 --profile-step > gpurun_out/i_ncu_launch.log 2>&1
wc -l gpurun_out/i_launches_step.csv
