#!/bin/bash
# Round 2, visit Q: compacted cross K|V projection (live encoder tiles only): tests + A/B bench.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/q_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/q_suite.log; tail -6 gpurun_out/q_suite.log
run() {
  name=$1; shift
  env "$@" timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-xl > gpurun_out/q_bench_$name.json 2> gpurun_out/q_bench_$name.err
  python - <<PY
import json
try:
    l = json.load(open("gpurun_out/q_bench_$name.json"))
    print("$name: value", round(l["value"], 1), "ms", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"], 1), "attn", round(l["roofline"]["attention_kernel"]["ms_per_step"], 2), "gemm", round(l["roofline"]["kernel_ms_per_step"], 2), round(l["roofline"]["frac"], 3), "flops", l["roofline"]["algorithmic_flops_per_step"], "clocks", l["clocks"].get("sm_mhz"))
except Exception as e:
    print("$name: bench parse failed", e)
PY
}
run compact X=1
run dense ATLAS_B200_XKV_COMPACT=0
run compact2 X=1
