#!/bin/bash
# GPU call 9: LayerNorm backward with 4 rows (4 warps) per CTA, 4 CTAs per SM, against 8 rows / 2 CTAs
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e9
mkdir -p $O
L=noise-robust-vit_b200/lib
for v in rows4 base rows4 base; do
  cp $L/libnrvit_$v.so $L/libnrvit.so
  echo "== $v"; timeout 300 python tools/gpu_time_ln.py 2>&1 | grep "ln_bwd" | tee -a $O/ln_time_$v.log
done
cp $L/libnrvit_rows4.so $L/libnrvit.so
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "layernorm or ln_" > $O/ln_tests_rows4.log 2>&1; echo "rc=$?" >> $O/ln_tests_rows4.log; tail -2 $O/ln_tests_rows4.log
for i in 1 2; do for v in rows4 base; do
  cp $L/libnrvit_$v.so $L/libnrvit.so
  timeout 400 python bench.py --steps 20 --warmup 8 --no-cpu-baseline 2>$O/bench_err.log | tail -1 > $O/bench_${v}_$i.json
  python -c "import json,sys; d=json.loads(open('$O/bench_${v}_$i.json').read()); print('$v', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['achieved']), round(d['e2e']['value']))"
done; done
