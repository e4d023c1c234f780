#!/bin/bash
# GPU call 4: launch lists (ncu per-launch durations) of the secondary configs: SimpleViT CIFAR-100 shape, README ViT B=8
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e4
mkdir -p $O
timeout 300 python tools/gpu_bench_configs.py simple readme > $O/configs.log 2>&1; cat $O/configs.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2000 --launch-count 700 --csv --log-file $O/launches_simple.csv python tools/gpu_bench_configs.py simple > $O/ncu_simple.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 --launch-count 500 --csv --log-file $O/launches_readme.csv python tools/gpu_bench_configs.py readme > $O/ncu_readme.log 2>&1
python tools/agg_launches.py $O/launches_simple.csv > $O/launches_simple_by_kernel.txt 2>&1; cat $O/launches_simple_by_kernel.txt
python tools/agg_launches.py $O/launches_readme.csv > $O/launches_readme_by_kernel.txt 2>&1; cat $O/launches_readme_by_kernel.txt
