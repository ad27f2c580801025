#!/bin/bash
# Round 2, visit W (2 GPUs): the final bench.py under torchrun exactly as the driver launches it (N = 2), few steps, xl leg skipped.
mkdir -p gpurun_out
timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 --no-xl > gpurun_out/w_bench_n2.json 2> gpurun_out/w_bench_n2.err
echo "bench n2 rc=$?"; tail -3 gpurun_out/w_bench_n2.err | cut -c1-300
python - <<'PY'
import json
try:
    l = json.load(open("gpurun_out/w_bench_n2.json"))
    r = l["roofline"]
    print("N", l["n_gpus"], "value", round(l["value"], 1), "ms", round(l["ms_per_step"], 2), "e2e", round(l["e2e"]["value"], 1), "gemm", round(r["frac"], 3), "padded_encoder", l.get("padded_encoder", {}).get("value"), "parity", l["parity_check"].get("status"), "clocks", l["clocks"].get("sm_mhz"))
    print("   mips", l["mips"]["value"], "train", round(l["train"].get("value", 0)), "generate", l["generate"].get("value"), "refresh", l["refresh"].get("value"))
except Exception as e:
    print("bench parse failed", e)
PY
