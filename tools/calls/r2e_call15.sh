#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e15
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd2 -s 3 -c 1 -o $O/attn_bwd2 -f python tools/gpu_time_attn.py > $O/ncu_bwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fwd2 -s 3 -c 1 -o $O/attn_fwd2 -f python tools/gpu_time_attn.py > $O/ncu_fwd.log 2>&1
ls -la $O
