#!/bin/bash
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e16
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention_tcgen05_fwd_bwd" > $O/attn_tests.log 2>&1; echo "rc=$?" >> $O/attn_tests.log
tail -2 $O/attn_tests.log
for i in 1 2; do timeout 300 python tools/gpu_time_attn.py 2>&1 | grep "bwd" | tee -a $O/attn_time.log; done
timeout 300 python tools/gpu_attn_phases_bwd2.py > $O/phases.log 2>&1
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_training_gpu.py tests/test_fullsize_gpu.py -q -x > $O/model_tests.log 2>&1; echo "rc=$?" >> $O/model_tests.log
tail -2 $O/model_tests.log
