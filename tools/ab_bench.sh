cd noise-robust-vit_b200/lib; cp libnrvit.so libnrvit_new.so; cd ../..
for i in 1 2; do
  cp noise-robust-vit_b200/lib/libnrvit_new.so noise-robust-vit_b200/lib/libnrvit.so
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('NEW', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['achieved']))"
  cp noise-robust-vit_b200/lib/libnrvit_prev.so noise-robust-vit_b200/lib/libnrvit.so
  python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('OLD', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['achieved']))"
done
cp noise-robust-vit_b200/lib/libnrvit_new.so noise-robust-vit_b200/lib/libnrvit.so
