#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for pf in 0 1; do for mb in 4 5 6; do
  CGASM_GATHER_PREFETCH=$pf CGASM_GATHER_MINB=$mb timeout 600 python bench.py --cells 128 --scatter gather --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench128_gather_pf${pf}_mb${mb}.json 2> gpurun_out/bench128_gather_pf${pf}_mb${mb}.err
  tail -2 gpurun_out/bench128_gather_pf${pf}_mb${mb}.err
done; done
timeout 900 python bench.py --cells 256 --scatter gather --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench256_gather.json 2> gpurun_out/bench256_gather.err
tail -3 gpurun_out/bench256_gather.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench*_gather*.json')):
    try:
        d=json.loads(open(f).read()); r=d['roofline']
        print(f, 'value %.0f'%d['value'], 'mom %.2f ms'%r['kernel_ms'], 'tra %.2f ms'%r['tracer']['kernel_ms'], 'frac %.3f'%r['frac'], 'setup %.1f'%d['setup_s'], 'e2e', d['e2e'] and round(d['e2e']['value']))
    except Exception as e: print(f,'ERR',e)
PY
