#!/bin/bash
# quick loop: gather parity subset + bench at 128 and 256
mkdir -p gpurun_out
true
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for mb in 4 5 6; do
CGASM_GATHER_MINB=$mb timeout 600 python bench.py --cells 128 --scatter gather --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench128_gather_mb$mb.json 2> gpurun_out/bench128_gather_mb$mb.err
tail -2 gpurun_out/bench128_gather_mb$mb.err
done
timeout 900 python bench.py --cells 256 --scatter gather --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench256_gather.json 2> gpurun_out/bench256_gather.err
tail -3 gpurun_out/bench256_gather.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench128_gather_mb*.json'))+['gpurun_out/bench256_gather.json']:
    try:
        d=json.loads(open(f).read()); r=d['roofline']
        print(f, 'value %.0f'%d['value'], 'mom %.2f ms'%r['kernel_ms'], 'tra %.2f ms'%r['tracer']['kernel_ms'], 'frac %.3f'%r['frac'], 'setup %.1f'%d['setup_s'])
    except Exception as e: print(f,'ERR',e)
PY
