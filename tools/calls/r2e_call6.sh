#!/bin/bash
# GPU call 6: attention dropout in the general tcgen05 kernels: parity + timing
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e6
mkdir -p $O
timeout 600 python -m pytest tests/test_model_gpu.py -q -x -k "dropout" > $O/dropout_tests.log 2>&1; echo "rc=$?" >> $O/dropout_tests.log
tail -30 $O/dropout_tests.log
timeout 600 python tools/gpu_time_attn_dropout.py 64 > $O/time_dropout.log 2>&1; cat $O/time_dropout.log | grep -v Warn
