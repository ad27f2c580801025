#!/bin/bash
# Round 2, visit N: evidence - smoke(), ncu --set full captures of the round's kernels, bench line of the final code state.
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/n_smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/n_smoke.log
timeout 900 python -m pytest tests/test_models_gpu.py tests/test_round2_gpu.py tests/test_atlas_gpu.py tests/test_fullsize_gpu.py -q --timeout 300 -p no:cacheprovider 2>&1 | tail -4
{
python tools/prof_ops.py xattn 10
ATLAS_B200_ATTN_SKIP_MASKED=0 python tools/prof_ops.py xattn 10
ATLAS_B200_XATTN_STREAM=0 python tools/prof_ops.py xattn 10
} > gpurun_out/n_xattn.log 2>&1; cat gpurun_out/n_xattn.log
cap() {  # name, kernel regex, prof_ops mode
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$2 -s 2 -c 1 -f -o gpurun_out/prof_$1 python tools/prof_ops.py $3 3 > gpurun_out/n_ncu_$1.log 2>&1
  tail -1 gpurun_out/n_ncu_$1.log
}
cap gemm_staged gemm_kernel gemm
cap lanes attention_lanes_kernel attention
cap xstream cross_stream_kernel xattn
cap bwd_dkv_tc attn_bwd_dkv_tc_kernel attn_bwd
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    l = json.load(open("gpurun_out/n_bench.json"))
    print("value", round(l["value"], 1), "ms", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"], 1), "attn", round(l["roofline"]["attention_kernel"]["ms_per_step"], 2), "gemm", round(l["roofline"]["kernel_ms_per_step"], 2), round(l["roofline"]["frac"], 3))
    print("   e2e phases", l["e2e"].get("phases_ms_synchronised"))
    print("   gpu_reference", l.get("gpu_reference", {}).get("value"), l.get("gpu_reference", {}).get("ours_over_reference_e2e"), "cpu", l.get("cpu_baseline", {}).get("value"))
    print("   mips", l["mips"]["value"], "train", round(l["train"].get("value", 0)), "xl", l["train_xl"].get("value"), "refresh", l["refresh"].get("value"), "generate", l["generate"].get("value"))
except Exception as e:
    print("bench parse failed", e)
PY
