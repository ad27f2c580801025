#!/bin/bash
# Round 2, visit D: softmax-loop microbenchmark (what can an SM do without MMA / barriers?) + failure details of visit C
mkdir -p gpurun_out
./tools/ubench/softmax_ubench > gpurun_out/d_ubench.log 2>&1
cat gpurun_out/d_ubench.log
timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_fullsize_gpu.py -q -p no:cacheprovider -k "attention or untied or fid_base" --tb=line 2>&1 | grep -E "Error|assert|passed|failed" | cut -c1-300 > gpurun_out/d_tests.log
cat gpurun_out/d_tests.log
