#!/bin/bash
# Round 2, visit J: cross-attention stream kernel (correctness + in-step effect), optimiser tests after the scalar fix, e2e phase
# breakdown, launch list of the 256-query search.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_layers_gpu.py tests/test_optim_gpu.py tests/test_backward_gpu.py tests/test_models_gpu.py tests/test_train_gpu.py tests/test_atlas_gpu.py -q --timeout 300 -p no:cacheprovider > gpurun_out/j_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/j_tests.log; tail -8 gpurun_out/j_tests.log
for c in 512 1024 256; do
  ATLAS_B200_XATTN_CHUNK=$c timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-xl > gpurun_out/j_bench_c$c.json 2> gpurun_out/j_bench_c$c.err
  python - <<PY
import json
try:
    l = json.load(open("gpurun_out/j_bench_c$c.json"))
    print("chunk $c: value", l["value"], "ms", l["ms_per_step"], "e2e", l["e2e"]["value"], "attn", l["roofline"]["attention_kernel"]["ms_per_step"], "gemm", l["roofline"]["kernel_ms_per_step"], l["roofline"]["frac"])
    print("   e2e phases", l["e2e"].get("phases_ms_synchronised"))
    print("   train", l["train"].get("value"), l["train"].get("ms_per_step"), "refresh", l["refresh"].get("value"), l["refresh"].get("roofline", {}).get("frac"))
except Exception as e:
    print("bench parse failed", e)
PY
done
ATLAS_B200_XATTN_STREAM=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-xl > gpurun_out/j_bench_nostream.json 2> gpurun_out/j_bench_nostream.err
python - <<PY
import json
l = json.load(open("gpurun_out/j_bench_nostream.json"))
print("no stream kernel: value", l["value"], "ms", l["ms_per_step"], "attn", l["roofline"]["attention_kernel"]["ms_per_step"])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/j_launches_mips.csv python tools/prof_ops.py mips 3 > gpurun_out/j_mips.log 2>&1
tail -1 gpurun_out/j_mips.log; wc -l gpurun_out/j_launches_mips.csv
