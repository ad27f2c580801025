#!/bin/bash
# Round 2, visit V (final state): packed-encoder tests incl. the Contriever encoder on packed rows, full suite, complete bench line.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_packed_encoder_gpu.py -q --timeout 120 -p no:cacheprovider > gpurun_out/v_packed.log 2>&1
rc=$?; echo "packed rc=$rc" >> gpurun_out/v_packed.log; tail -6 gpurun_out/v_packed.log | cut -c1-400
if [ $rc -ne 0 ]; then
  echo "FALLBACK: ATLAS_B200_BERT_PACKED=0 for the rest of this visit"; export ATLAS_B200_BERT_PACKED=0
  grep -n "FAILED\|Error\|assert" gpurun_out/v_packed.log | head -20 | cut -c1-300
fi
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider --deselect tests/test_packed_encoder_gpu.py > gpurun_out/v_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/v_suite.log; tail -6 gpurun_out/v_suite.log | cut -c1-400
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    l = json.load(open("gpurun_out/v_bench.json"))
    r = l["roofline"]
    print("value", round(l["value"], 1), "ms", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"], 1), "attn", round(r["attention_kernel"]["ms_per_step"], 2), "gemm(big)", round(r["kernel_ms_per_step"], 2), r["kernel_launches_per_step"], round(r["achieved"]), round(r["frac"], 3), "all", round(r["all_gemm_launches"]["frac"], 3), "clocks", l["clocks"].get("sm_mhz"))
    print("   padded_encoder", l.get("padded_encoder", {}).get("value"), "gpu_reference", l.get("gpu_reference", {}).get("value"), l.get("gpu_reference", {}).get("ours_over_reference_e2e"), "cpu", l.get("cpu_baseline", {}).get("value"))
    print("   e2e phases", l["e2e"].get("phases_ms_synchronised"))
    print("   train", round(l["train"].get("value", 0)), "generate", l["generate"].get("value"), "mips", l["mips"]["value"], "xl", l["train_xl"].get("value"), "refresh", l["refresh"].get("value"), l["refresh"].get("roofline", {}).get("frac"))
except Exception as e:
    print("bench parse failed", e)
PY
