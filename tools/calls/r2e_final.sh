#!/bin/bash
# final validation of the session: full GPU suite, smoke(), default bench, launch list of the final build, secondary configs
export PYTHONUNBUFFERED=1
O=gpurun_out/r2e_final
mkdir -p $O
t0=$(date +%s)
timeout 600 python -m pytest tests -m gpu -x -q > $O/full_suite.log 2>&1; echo "rc=$?" >> $O/full_suite.log; tail -3 $O/full_suite.log
timeout 300 python __graft_entry__.py smoke > $O/smoke.log 2>&1; echo "rc=$?" >> $O/smoke.log; tail -3 $O/smoke.log
echo "tests done $(( $(date +%s) - t0 )) s"
timeout 600 python bench.py 2>$O/bench_err.log | tail -1 > $O/bench_default.json; python -c "import json; d=json.load(open('$O/bench_default.json')); print('default', round(d['value']), round(d['ms_per_step'],2), d['clocks'], round(d['roofline']['achieved']), round(d['e2e']['value']), d['cpu_baseline']['value'])"
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>>$O/bench_err.log | tail -1 > $O/bench_s20w5.json; python -c "import json; d=json.load(open('$O/bench_s20w5.json')); print('s20w5', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['achieved']), round(d['e2e']['value']))"
echo "bench done $(( $(date +%s) - t0 )) s"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1500 --launch-count 420 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/ncu_bench.log 2>&1
python tools/agg_launches.py $O/launches_bench.csv > $O/launches_by_kernel.txt 2>&1; head -8 $O/launches_by_kernel.txt
timeout 400 python tools/gpu_bench_configs.py simple readme l16 2>&1 | grep "ms/step" | tee $O/secondary.log
echo "all done $(( $(date +%s) - t0 )) s"
