#!/bin/bash
# Round 2, visit P (8 GPUs): bench at N = 8 exactly as the driver launches it (weak scaling, in-bench distributed parity check).
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/p_smi.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/p_bench_n8.json 2> gpurun_out/p_bench_n8.err
echo "bench n8 rc=$?"; tail -3 gpurun_out/p_bench_n8.err | cut -c1-200
python - <<'PY'
import json
try:
    l = json.load(open("gpurun_out/p_bench_n8.json"))
    print("N=8: value", round(l["value"], 1), "ms", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"], 1), "parity", l["parity_check"].get("status"), "attn", round(l["roofline"]["attention_kernel"]["ms_per_step"], 2), "gemm", round(l["roofline"]["kernel_ms_per_step"], 2), "clocks", l["clocks"])
    print("   mips", l["mips"]["value"], l["mips"]["ms_per_step"], "train", round(l["train"].get("value", 0)), "xl", str(l.get("train_xl"))[:260])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-xl > gpurun_out/p_bench_n1.json 2> gpurun_out/p_bench_n1.err
python - <<'PY'
import json
l = json.load(open("gpurun_out/p_bench_n1.json"))
print("N=1 same box: value", round(l["value"], 1), "ms", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"], 1))
PY
