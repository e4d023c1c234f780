#!/bin/bash
# interleaved A/B of two bench.py argument sets on one box:  tools/ab_modes.sh "<args A>" "<args B>" [rounds]
cd "$(dirname "$0")/.."
R=${3:-3}
one() { python bench.py --steps 20 --warmup 8 --no-cpu-baseline $2 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['achieved']))"; }
one warm "$1" > /dev/null
for i in $(seq $R); do one "A[$1]" "$1"; one "B[$2]" "$2"; done
