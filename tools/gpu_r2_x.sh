#!/bin/bash
# Round 2, visit X: evidence for the index-refresh path: launch list of one 512-passage embedder batch, packed
# and padded Contriever, eager launches under ncu.
mkdir -p gpurun_out
{
timeout 120 python tools/prof_ops.py refresh 10
ATLAS_B200_BERT_PACKED=0 timeout 120 python tools/prof_ops.py refresh 10
} > gpurun_out/x_refresh_ab.log 2>&1; cat gpurun_out/x_refresh_ab.log | tail -4 | cut -c1-300
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/x_launches_refresh.csv python tools/prof_ops.py refresh 1 > gpurun_out/x_ncu_refresh.log 2>&1
wc -l gpurun_out/x_launches_refresh.csv
