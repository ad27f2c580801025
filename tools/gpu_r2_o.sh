#!/bin/bash
# Round 2, visit O: full suite + bench after the decoder fused-norm / Contriever graph / GC changes.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/o_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/o_suite.log; tail -6 gpurun_out/o_suite.log
for i in 1 2; do
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-xl > gpurun_out/o_bench_$i.json 2> gpurun_out/o_bench_$i.err
python - <<PY
import json
try:
    l = json.load(open("gpurun_out/o_bench_$i.json"))
    print("run $i: value", round(l["value"], 1), "ms", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"], 1), "attn", round(l["roofline"]["attention_kernel"]["ms_per_step"], 2), "gemm", round(l["roofline"]["kernel_ms_per_step"], 2), round(l["roofline"]["frac"], 3), "clocks", l["clocks"].get("sm_mhz"))
    print("   e2e phases", l["e2e"].get("phases_ms_synchronised"))
except Exception as e:
    print("bench parse failed", e)
PY
done
ATLAS_B200_CUDA_GRAPH=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --nvtx --nvtx-include "atlas_b200_timed/" --csv --log-file gpurun_out/o_launches_step.csv python bench.py --steps 1 --warmup 3 --profile-step > gpurun_out/o_ncu_launch.log 2>&1
wc -l gpurun_out/o_launches_step.csv
