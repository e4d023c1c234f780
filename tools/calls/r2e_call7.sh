#!/bin/bash
# GPU call 7: attn_bwd2 with early accumulator hand-back + dQ product trimmed to the tile's real keys
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e7
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention_tcgen05_fwd_bwd" > $O/attn_tests.log 2>&1; echo "rc=$?" >> $O/attn_tests.log
tail -2 $O/attn_tests.log
for i in 1 2; do timeout 300 python tools/gpu_time_attn.py 2>&1 | grep bwd | tee -a $O/attn_time.log; done
timeout 300 python tools/gpu_attn_phases_bwd2.py > $O/phases.log 2>&1
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_training_gpu.py tests/test_fullsize_gpu.py -q -x > $O/model_tests.log 2>&1; echo "rc=$?" >> $O/model_tests.log
tail -3 $O/model_tests.log
for i in 1 2 3; do
  timeout 400 python bench.py --steps 20 --warmup 8 --no-cpu-baseline 2>$O/bench_err.log | tail -1 > $O/bench_$i.json
  python -c "import json,sys; d=json.loads(open('$O/bench_$i.json').read()); print('NEW', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['achieved']), round(d['e2e']['value']))"
done
timeout 300 python tools/gpu_bench_configs.py simple 2>&1 | grep "ms/step" | tee $O/simple.log
