#!/bin/bash
# Round 2, visit A: full GPU suite (incl. the three-lane attention kernel, full-size parity, decode), first bench lines.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/a_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -s --timeout 400 -p no:cacheprovider > gpurun_out/a_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/a_suite.log
tail -30 gpurun_out/a_suite.log
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
echo "bench rc=$?"; tail -5 gpurun_out/a_bench.err; head -c 3000 gpurun_out/a_bench.json
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 --ref-budget-s 45 > gpurun_out/a_ref.json 2> gpurun_out/a_ref.err
echo "ref rc=$?"; tail -3 gpurun_out/a_ref.err; head -c 1500 gpurun_out/a_ref.json
