#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e17
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "layernorm or ln_" > $O/ln_tests.log 2>&1; echo "rc=$?" >> $O/ln_tests.log; tail -2 $O/ln_tests.log
for i in 1 2; do timeout 300 python tools/gpu_time_ln.py 2>&1 | grep "ln_" | tee -a $O/ln_time.log; done
