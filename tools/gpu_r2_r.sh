#!/bin/bash
# Round 2, visit R: decode steps skip all-padding key tiles, train leg with the padded-passage figure, GEMM FLOP accounting of the
# device-side-M projections: full suite + the complete bench line.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/r_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/r_suite.log; tail -6 gpurun_out/r_suite.log
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r_bench.json 2> gpurun_out/r_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    l = json.load(open("gpurun_out/r_bench.json"))
    print("value", round(l["value"], 1), "ms", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"], 1), "attn", round(l["roofline"]["attention_kernel"]["ms_per_step"], 2), "gemm", round(l["roofline"]["kernel_ms_per_step"], 2), round(l["roofline"]["achieved"]), round(l["roofline"]["frac"], 3), "flops", l["roofline"]["algorithmic_flops_per_step"], "clocks", l["clocks"].get("sm_mhz"))
    print("   train", round(l["train"].get("value", 0)), l["train"].get("ms_per_step"), "padded", l["train"].get("padded_passages"))
    print("   generate", {k: l["generate"].get(k) for k in ("value", "ms_per_decode_step", "ms_per_generate")}, str(l["generate"].get("roofline"))[:200])
    print("   gpu_reference", l.get("gpu_reference", {}).get("value"), l.get("gpu_reference", {}).get("ours_over_reference_e2e"), "cpu", l.get("cpu_baseline", {}).get("value"))
    print("   mips", l["mips"]["value"], "xl", l["train_xl"].get("value"), "refresh", l["refresh"].get("value"), l["refresh"].get("roofline", {}).get("frac"))
except Exception as e:
    print("bench parse failed", e)
PY
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 --ref-budget-s 45 > gpurun_out/r_ref.json 2> gpurun_out/r_ref.err
echo "ref rc=$?"; head -c 600 gpurun_out/r_ref.json
