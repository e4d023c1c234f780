#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e18
mkdir -p $O
timeout 300 python tools/gpu_time_sinkhorn.py > $O/time.log 2>&1; cat $O/time.log | tail -8
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_tc_fwd -s 2 -c 1 -o $O/sk_fwd -f python tools/gpu_time_sinkhorn.py > $O/ncu_fwd.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:sinkhorn_tc_bwd -s 2 -c 1 -o $O/sk_bwd -f python tools/gpu_time_sinkhorn.py > $O/ncu_bwd.log 2>&1
ls -la $O
