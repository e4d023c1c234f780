#!/bin/bash
# GPU call 3: AEPI v3 (early hand-back, kvdone/qdone barriers, TMA stores of half accumulators)
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e3
mkdir -p $O
t0=$(date +%s)
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention_tcgen05_fwd_bwd" > $O/attn_tests_aepi.log 2>&1; echo "rc=$?" >> $O/attn_tests_aepi.log
tail -2 $O/attn_tests_aepi.log
for a in 0 1 0 1; do
  echo "AEPI=$a"
  NRV_ATTN_BWD_AEPI=$a timeout 300 python tools/gpu_time_attn.py 2>&1 | grep bwd | tee -a $O/attn_time_a${a}.log
done
timeout 300 python tools/gpu_attn_phases_bwd2.py > $O/phases_aepi.log 2>&1
echo "timing done $(( $(date +%s) - t0 )) s"
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_training_gpu.py tests/test_fullsize_gpu.py -q -x > $O/model_tests_aepi.log 2>&1; echo "rc=$?" >> $O/model_tests_aepi.log
tail -3 $O/model_tests_aepi.log
for i in 1 2 3; do
  for m in 1 0; do
    NRV_ATTN_BWD_AEPI=$m timeout 400 python bench.py --steps 20 --warmup 8 --no-cpu-baseline 2>$O/bench_err.log | tail -1 > $O/bench_a${m}_$i.json
    python -c "import json,sys; d=json.loads(open('$O/bench_a${m}_$i.json').read()); print('AEPI=$m', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['achieved']), round(d['e2e']['value']))"
  done
done
echo "all done $(( $(date +%s) - t0 )) s"
