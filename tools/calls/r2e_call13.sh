#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e13
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" > $O/attn_tests.log 2>&1; echo "rc=$?" >> $O/attn_tests.log
tail -2 $O/attn_tests.log
timeout 600 python tools/gpu_infer_sweep.py > $O/h14_sweep.log 2>&1; cat $O/h14_sweep.log
timeout 300 python tools/gpu_h14_train.py > $O/h14_train.log 2>&1; tail -3 $O/h14_train.log
