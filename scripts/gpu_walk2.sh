#!/bin/bash
mkdir -p gpurun_out
for pf in 0 1; do for mb in 3 4; do
CGASM_WALK_PREFETCH=$pf CGASM_WALK_MINB=$mb timeout 600 python bench.py --cells 128 --scatter gather --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench128_walk_pf${pf}_mb$mb.json 2> gpurun_out/bench128_walk_pf${pf}_mb$mb.err
done; done
timeout 900 python bench.py --cells 256 --scatter gather --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench256_walk.json 2> gpurun_out/bench256_walk.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench128_walk_pf*.json'))+['gpurun_out/bench256_walk.json']:
    try:
        d=json.loads(open(f).read()); r=d['roofline']
        print(f, 'value %.0f'%d['value'], 'mom %.2f ms'%r['kernel_ms'], 'tra %.2f ms'%r['tracer']['kernel_ms'], 'frac %.3f'%r['frac'], 'setup %.1f'%d['setup_s'])
    except Exception as e: print(f,'ERR',e)
PY
