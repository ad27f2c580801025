#!/bin/bash
# One GPU visit: parity tests, smoke, bench, ncu launch list + full capture of the scan kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
if [ "$1" != "noprof" ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'mips_|topk_merge|cast_f32|widen' -c 40 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
tail -25 gpurun_out/launches.csv
ncu --set full --clock-control none --import-source on -k regex:mips_scan -s 2 -c 2 -o gpurun_out/prof_scan -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
fi
