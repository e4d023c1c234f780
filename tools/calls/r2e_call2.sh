#!/bin/bash
# GPU call 2: AEPI with coalesced row stores, dQ-late MMA order; kernel timing in the four combinations + phases + bench A/B
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e2
mkdir -p $O
t0=$(date +%s)
for dq in 0 1; do
  NRV_ATTN_BWD_DQLATE=$dq timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention_tcgen05_fwd_bwd" > $O/attn_tests_aepi_dq$dq.log 2>&1; echo "rc=$?" >> $O/attn_tests_aepi_dq$dq.log
  tail -2 $O/attn_tests_aepi_dq$dq.log
done
NRV_ATTN_BWD_AEPI=0 NRV_ATTN_BWD_DQLATE=1 timeout 600 python -m pytest tests/test_kernels_gpu.py -q -x -k "attention_tcgen05_fwd_bwd" > $O/attn_tests_old_dq1.log 2>&1; echo "rc=$?" >> $O/attn_tests_old_dq1.log
tail -2 $O/attn_tests_old_dq1.log
for a in 0 1; do for dq in 0 1; do
  echo "AEPI=$a DQLATE=$dq"
  NRV_ATTN_BWD_AEPI=$a NRV_ATTN_BWD_DQLATE=$dq timeout 300 python tools/gpu_time_attn.py 2>&1 | grep bwd | tee $O/attn_time_a${a}_dq${dq}.log
done; done
NRV_ATTN_BWD_DQLATE=0 timeout 300 python tools/gpu_attn_phases_bwd2.py > $O/phases_aepi_dq0.log 2>&1
NRV_ATTN_BWD_DQLATE=1 timeout 300 python tools/gpu_attn_phases_bwd2.py > $O/phases_aepi_dq1.log 2>&1
NRV_ATTN_BWD_AEPI=0 NRV_ATTN_BWD_DQLATE=1 timeout 300 python tools/gpu_attn_phases_bwd2.py > $O/phases_old_dq1.log 2>&1
echo "timing done $(( $(date +%s) - t0 )) s"
for i in 1 2; do
  for m in "1 1" "0 0" "1 0" "0 1"; do
    set -- $m
    NRV_ATTN_BWD_AEPI=$1 NRV_ATTN_BWD_DQLATE=$2 timeout 400 python bench.py --steps 20 --warmup 8 --no-cpu-baseline 2>$O/bench_err.log | tail -1 > $O/bench_a$1_dq$2_$i.json
    python -c "import json,sys; d=json.loads(open('$O/bench_a$1_dq$2_$i.json').read()); print('AEPI=$1 DQLATE=$2', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['achieved']), round(d['e2e']['value']))"
  done
done
echo "bench done $(( $(date +%s) - t0 )) s"
NRV_ATTN_BWD_DQLATE=1 timeout 600 python -m pytest tests/test_model_gpu.py tests/test_training_gpu.py tests/test_fullsize_gpu.py -q -x > $O/model_tests_aepi_dq1.log 2>&1; echo "rc=$?" >> $O/model_tests_aepi_dq1.log
tail -3 $O/model_tests_aepi_dq1.log
echo "all done $(( $(date +%s) - t0 )) s"
