#!/bin/bash
# Round 2, visit L: backward kernels skip masked key blocks (dQ / dK / dV), stream-kernel chunk 1024: full suite, bench, train A/B.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/l_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/l_suite.log; tail -6 gpurun_out/l_suite.log
run() {
  name=$1; shift
  env "$@" timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline $EXTRA > gpurun_out/l_bench_$name.json 2> gpurun_out/l_bench_$name.err
  python - <<PY
import json
try:
    l = json.load(open("gpurun_out/l_bench_$name.json"))
    print("$name: value", round(l["value"], 1), "ms", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"], 1), "attn", round(l["roofline"]["attention_kernel"]["ms_per_step"], 2), "gemm", round(l["roofline"]["kernel_ms_per_step"], 2), round(l["roofline"]["frac"], 3))
    print("   train", round(l["train"].get("value", 0)), l["train"].get("ms_per_step"), str(l["train"].get("kernels"))[:330])
    print("   xl", str(l.get("train_xl"))[:200])
    print("   gpu_reference", l.get("gpu_reference", {}).get("value"), l.get("gpu_reference", {}).get("ours_over_reference_e2e"), "refresh", l["refresh"].get("value"), l["refresh"].get("roofline", {}).get("frac"))
    print("   mips", l["mips"]["value"], l["mips"]["ms_per_step"], l["mips"]["roofline"]["frac"], "generate", l["generate"].get("value"))
except Exception as e:
    print("$name: bench parse failed", e)
PY
}
EXTRA="" run default X=1
EXTRA="--no-gpu-reference --no-xl" run noskip ATLAS_B200_ATTN_SKIP_MASKED=0
