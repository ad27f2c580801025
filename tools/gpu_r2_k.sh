#!/bin/bash
# Round 2, visit K: cross-attention stream kernel + masked-key-block skipping (lanes kernel, stream kernel): full suite,
# A/B benches (skipping off, stream kernel off), chunk sizes, launch list of the step.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/k_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/k_suite.log; tail -8 gpurun_out/k_suite.log
run() {  # name, env...
  name=$1; shift
  env "$@" timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-xl > gpurun_out/k_bench_$name.json 2> gpurun_out/k_bench_$name.err
  python - <<PY
import json
try:
    l = json.load(open("gpurun_out/k_bench_$name.json"))
    print("$name: value", round(l["value"], 1), "ms", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"], 1), "attn", round(l["roofline"]["attention_kernel"]["ms_per_step"], 2), "gemm", round(l["roofline"]["kernel_ms_per_step"], 2), round(l["roofline"]["frac"], 3), "train", round(l["train"].get("value", 0)), "refresh", l["refresh"].get("value"))
    print("   e2e phases", l["e2e"].get("phases_ms_synchronised"))
except Exception as e:
    print("$name: bench parse failed", e)
PY
}
run default X=1
run noskip ATLAS_B200_ATTN_SKIP_MASKED=0
run nostream ATLAS_B200_XATTN_STREAM=0
run chunk1024 ATLAS_B200_XATTN_CHUNK=1024
run chunk256 ATLAS_B200_XATTN_CHUNK=256
ATLAS_B200_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "atlas_b200_timed/" --csv --log-file gpurun_out/k_launches_step.csv python bench.py --steps 1 --warmup 3 --profile-step > gpurun_out/k_ncu_launch.log 2>&1
wc -l gpurun_out/k_launches_step.csv
