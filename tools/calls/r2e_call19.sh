#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e19
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "sinkhorn" > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log; tail -2 $O/tests.log
timeout 300 python tools/gpu_time_sinkhorn.py 2>&1 | grep -v Warn | tee $O/time.log
