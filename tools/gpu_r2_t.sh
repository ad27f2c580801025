#!/bin/bash
# Round 2, visit T: packed attention with work-balanced CTA ranges (A/B), GEMM FLOP accounting of device-side row counts, full suite,
# the complete bench line of the final code state, smoke().
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_packed_encoder_gpu.py -q --timeout 120 -p no:cacheprovider > gpurun_out/t_packed.log 2>&1
echo "packed rc=$?" >> gpurun_out/t_packed.log; tail -5 gpurun_out/t_packed.log | cut -c1-300
{
timeout 200 python tools/prof_ops.py packed 20
ATLAS_B200_PACKED_BALANCE=0 timeout 200 python tools/prof_ops.py packed 20
} > gpurun_out/t_packed_ab.log 2>&1; cat gpurun_out/t_packed_ab.log | cut -c1-400
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider --deselect tests/test_packed_encoder_gpu.py > gpurun_out/t_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/t_suite.log; tail -4 gpurun_out/t_suite.log | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/t_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/t_smoke.log | cut -c1-300
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/t_bench.json 2> gpurun_out/t_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    l = json.load(open("gpurun_out/t_bench.json"))
    print("value", round(l["value"], 1), "ms", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"], 1), "attn", round(l["roofline"]["attention_kernel"]["ms_per_step"], 2), "gemm", round(l["roofline"]["kernel_ms_per_step"], 2), round(l["roofline"]["achieved"]), round(l["roofline"]["frac"], 3), "flops", l["roofline"]["algorithmic_flops_per_step"], "clocks", l["clocks"].get("sm_mhz"))
    print("   padded_encoder", l.get("padded_encoder", {}).get("value"))
    print("   e2e phases", l["e2e"].get("phases_ms_synchronised"))
    print("   train", round(l["train"].get("value", 0)), "generate", {k: l["generate"].get(k) for k in ("value", "ms_per_decode_step", "ms_encoder_and_first_step")}, {k: l["generate"]["roofline"].get(k) for k in ("achieved", "frac", "dense_equivalent_GBps")})
    print("   gpu_reference", l.get("gpu_reference", {}).get("value"), l.get("gpu_reference", {}).get("ours_over_reference_e2e"), "cpu", l.get("cpu_baseline", {}).get("value"))
    print("   mips", l["mips"]["value"], "xl", l["train_xl"].get("value"), "refresh", l["refresh"].get("value"))
except Exception as e:
    print("bench parse failed", e)
PY
