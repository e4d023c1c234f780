#!/bin/bash
# GPU call 5: ViT-H/14 inference sweep on the round-2 build (BASELINE configs[4]); launch list of the SimpleViT CIFAR-100 config
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e5
mkdir -p $O
timeout 600 python tools/gpu_infer_sweep.py > $O/h14_sweep.log 2>&1; cat $O/h14_sweep.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1200 --launch-count 500 --csv --log-file $O/launches_simple.csv python tools/gpu_bench_configs.py simple > $O/ncu_simple.log 2>&1
python tools/agg_launches.py $O/launches_simple.csv > $O/launches_simple_by_kernel.txt 2>&1; cat $O/launches_simple_by_kernel.txt
