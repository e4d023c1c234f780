#!/bin/bash
# GPU call 10: ncu --set full (with source) of the LayerNorm backward kernel and the fused attention backward of the final build
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e10
mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ln_bwd_tma -s 3 -c 1 -o $O/ln_bwd -f python tools/gpu_time_ln.py > $O/ncu_ln.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_bwd2 -s 3 -c 1 -o $O/attn_bwd2 -f python tools/gpu_time_attn.py > $O/ncu_attn.log 2>&1
ls -la $O
