#!/bin/bash
# GPU call 11: attention forward without per-tile integer divisions; step-level bench of the session's attention changes
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e11
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention" > $O/attn_tests.log 2>&1; echo "rc=$?" >> $O/attn_tests.log
tail -2 $O/attn_tests.log
for i in 1 2; do timeout 300 python tools/gpu_time_attn.py 2>&1 | grep "attn" | tee -a $O/attn_time.log; done
for i in 1 2 3; do
  timeout 400 python bench.py --steps 20 --warmup 8 --no-cpu-baseline 2>$O/bench_err.log | tail -1 > $O/bench_$i.json
  python -c "import json,sys; d=json.loads(open('$O/bench_$i.json').read()); print('NEW', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['achieved']), round(d['e2e']['value']))"
done
