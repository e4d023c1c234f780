#!/bin/bash
# round 2, session 4, GPU call 1: asynchronous-epilogue attention backward (NRV_ATTN_BWD_AEPI) parity + A/B, then the full GPU suite
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv > $O/gpu.txt 2>&1
t0=$(date +%s)
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention_tcgen05_fwd_bwd" > $O/attn_tests_aepi.log 2>&1; echo "rc=$?" >> $O/attn_tests_aepi.log
echo "attn tests done $(( $(date +%s) - t0 )) s"; tail -3 $O/attn_tests_aepi.log
NRV_ATTN_BWD_AEPI=0 timeout 300 python tools/gpu_time_attn.py > $O/attn_time_old.log 2>&1
NRV_ATTN_BWD_AEPI=1 timeout 300 python tools/gpu_time_attn.py > $O/attn_time_aepi.log 2>&1
grep bwd $O/attn_time_old.log $O/attn_time_aepi.log
timeout 300 python tools/gpu_attn_phases_bwd2.py > $O/phases_aepi.log 2>&1
echo "timing done $(( $(date +%s) - t0 )) s"
for i in 1 2; do
  for m in 1 0; do
    NRV_ATTN_BWD_AEPI=$m timeout 400 python bench.py --steps 20 --warmup 8 --no-cpu-baseline 2>$O/bench_err_$m.log | tail -1 > $O/bench_aepi${m}_$i.json
    python -c "import json,sys; d=json.loads(open('$O/bench_aepi${m}_$i.json').read()); print('AEPI=$m', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['achieved']), round(d['e2e']['value']))"
  done
done
echo "bench done $(( $(date +%s) - t0 )) s"
timeout 900 python -m pytest tests -m gpu -x -q --durations=12 > $O/full_suite.log 2>&1; echo "rc=$?" >> $O/full_suite.log
tail -25 $O/full_suite.log
echo "all done $(( $(date +%s) - t0 )) s"
