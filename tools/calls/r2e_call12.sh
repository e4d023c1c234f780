#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e12
mkdir -p $O
timeout 300 python tools/gpu_prof_target_gemm.py > $O/target.log 2>&1; tail -2 $O/target.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 6 -c 3 -o $O/gemm3 -f python tools/gpu_prof_target_gemm.py > $O/ncu.log 2>&1
ls -la $O; tail -3 $O/ncu.log
