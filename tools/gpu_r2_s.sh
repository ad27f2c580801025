#!/bin/bash
# Round 2, visit S: the padding-compacted FiD encoder (packed rows: segment tables, packed three-lane attention, device-side row
# counts in every encoder GEMM): new tests, the full suite, the bench line with the padded-encoder comparator.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_packed_encoder_gpu.py -q --timeout 120 -p no:cacheprovider > gpurun_out/s_packed.log 2>&1
echo "packed rc=$?" >> gpurun_out/s_packed.log; tail -15 gpurun_out/s_packed.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider --deselect tests/test_packed_encoder_gpu.py > gpurun_out/s_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/s_suite.log; tail -8 gpurun_out/s_suite.log | cut -c1-300
timeout 600 python bench.py --steps 20 --warmup 3 --no-xl --no-cpu-baseline > gpurun_out/s_bench.json 2> gpurun_out/s_bench.err
echo "bench rc=$?"; tail -3 gpurun_out/s_bench.err | cut -c1-300; python - <<'PY'
import json
try:
    l = json.load(open("gpurun_out/s_bench.json"))
    print("value", round(l["value"], 1), "ms", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"], 1), "attn", round(l["roofline"]["attention_kernel"]["ms_per_step"], 2), "gemm", round(l["roofline"]["kernel_ms_per_step"], 2), round(l["roofline"]["achieved"]), round(l["roofline"]["frac"], 3), "flops", l["roofline"]["algorithmic_flops_per_step"], "clocks", l["clocks"].get("sm_mhz"))
    print("   padded_encoder", l.get("padded_encoder"))
    print("   e2e phases", l["e2e"].get("phases_ms_synchronised"))
    print("   train", round(l["train"].get("value", 0)), "generate", {k: l["generate"].get(k) for k in ("value", "ms_per_decode_step", "ms_encoder_and_first_step")}, str(l["generate"].get("roofline"))[:300])
    print("   gpu_reference", l.get("gpu_reference", {}).get("value"), l.get("gpu_reference", {}).get("ours_over_reference_e2e"))
except Exception as e:
    print("bench parse failed", e)
PY
