#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "gather" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for mb in 2 3 4 5; do
CGASM_DEBUG=1 CGASM_WALK_MINB=$mb timeout 600 python bench.py --cells 128 --scatter gather --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench128_walk_mb$mb.json 2> gpurun_out/bench128_walk_mb$mb.err
grep cgasm gpurun_out/bench128_walk_mb$mb.err | head -1
done
CGASM_GATHER_DIRECT=1 timeout 600 python bench.py --cells 128 --scatter gather --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench128_walk_direct.json 2> gpurun_out/bench128_walk_direct.err
timeout 900 python bench.py --cells 256 --scatter gather --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench256_walk.json 2> gpurun_out/bench256_walk.err
tail -3 gpurun_out/bench256_walk.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench128_walk_*.json'))+['gpurun_out/bench256_walk.json']:
    try:
        d=json.loads(open(f).read()); r=d['roofline']
        print(f, 'value %.0f'%d['value'], 'mom %.2f ms'%r['kernel_ms'], 'tra %.2f ms'%r['tracer']['kernel_ms'], 'frac %.3f'%r['frac'], 'setup %.1f'%d['setup_s'])
    except Exception as e: print(f,'ERR',e)
PY
