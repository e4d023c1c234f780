#!/bin/bash
# GPU call 8: attn_bwd2 with the dQ product trimmed to the tile's real keys (only change)
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e8
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention_tcgen05_fwd_bwd" > $O/attn_tests.log 2>&1; echo "rc=$?" >> $O/attn_tests.log
tail -2 $O/attn_tests.log
for i in 1 2; do timeout 300 python tools/gpu_time_attn.py 2>&1 | grep bwd | tee -a $O/attn_time.log; done
timeout 300 python tools/gpu_attn_phases_bwd2.py > $O/phases.log 2>&1
