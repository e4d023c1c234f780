#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e14
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py -q -x -k "attention or dropout" > $O/tests.log 2>&1; echo "rc=$?" >> $O/tests.log
tail -2 $O/tests.log
timeout 300 python tools/gpu_h14_train.py > $O/h14_train.log 2>&1; tail -2 $O/h14_train.log
timeout 600 python tools/gpu_time_attn_dropout.py 64 2>&1 | grep "ViT-B" | tee $O/time_dropout.log
