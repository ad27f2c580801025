#!/bin/bash
# Round 2, visit M (2 GPUs): NCCL search parity test, bench at N = 2 (in-bench distributed parity check, weak scaling).
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m_smi.txt 2>&1
timeout 600 python -m pytest tests/test_search_gpu.py -q -k "nccl or distributed" --timeout 400 -p no:cacheprovider > gpurun_out/m_nccl_test.log 2>&1
echo "nccl test rc=$?" >> gpurun_out/m_nccl_test.log; tail -4 gpurun_out/m_nccl_test.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/m_bench_n2.json 2> gpurun_out/m_bench_n2.err
echo "bench n2 rc=$?"; tail -3 gpurun_out/m_bench_n2.err | cut -c1-200
python - <<'PY'
import json
try:
    l = json.load(open("gpurun_out/m_bench_n2.json"))
    print("N=2: value", round(l["value"], 1), "ms", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"], 1), "parity", l["parity_check"].get("status"), "attn", round(l["roofline"]["attention_kernel"]["ms_per_step"], 2), "gemm", round(l["roofline"]["kernel_ms_per_step"], 2))
    print("   mips", l["mips"]["value"], l["mips"]["ms_per_step"], "train", round(l["train"].get("value", 0)), "xl", str(l.get("train_xl"))[:160])
except Exception as e:
    print("bench parse failed", e)
PY
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-reference --no-xl > gpurun_out/m_bench_n1.json 2> gpurun_out/m_bench_n1.err
python - <<'PY'
import json
l = json.load(open("gpurun_out/m_bench_n1.json"))
print("N=1 same box: value", round(l["value"], 1), "ms", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"], 1), "phases", l["e2e"].get("phases_ms_synchronised"))
PY
