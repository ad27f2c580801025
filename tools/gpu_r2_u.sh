#!/bin/bash
# Round 2, visit U: evidence for the padding-compacted encoder - launch list of one timed step (eager launches under ncu), one
# `ncu --set full` capture of the packed three-lane attention kernel.
mkdir -p gpurun_out
ATLAS_B200_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "atlas_b200_timed/" --csv --log-file gpurun_out/u_launches_step.csv python bench.py --steps 1 --warmup 3 --profile-step > gpurun_out/u_ncu_launch.log 2>&1
wc -l gpurun_out/u_launches_step.csv
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attention_lanes_kernel -s 4 -c 1 -f -o gpurun_out/prof_lanes_packed python tools/prof_ops.py packed 3 > gpurun_out/u_ncu_lanes_packed.log 2>&1
tail -2 gpurun_out/u_ncu_lanes_packed.log | cut -c1-300
ls -la gpurun_out/prof_lanes_packed.ncu-rep
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/u_bench.json 2> gpurun_out/u_bench.err
echo "bench rc=$?"; python - <<'PY'
import json
try:
    l = json.load(open("gpurun_out/u_bench.json"))
    r = l["roofline"]
    print("value", round(l["value"], 1), "ms", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"], 1), "attn", round(r["attention_kernel"]["ms_per_step"], 2), "gemm(big)", round(r["kernel_ms_per_step"], 2), r["kernel_launches_per_step"], round(r["achieved"]), round(r["frac"], 3), "all", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r["all_gemm_launches"].items()}, "clocks", l["clocks"].get("sm_mhz"))
    print("   padded_encoder", l.get("padded_encoder", {}).get("value"), "gpu_reference", l.get("gpu_reference", {}).get("value"), l.get("gpu_reference", {}).get("ours_over_reference_e2e"), "cpu", l.get("cpu_baseline", {}).get("value"))
    print("   e2e phases", l["e2e"].get("phases_ms_synchronised"))
except Exception as e:
    print("bench parse failed", e)
PY
