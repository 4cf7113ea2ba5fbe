#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
for cfg in "512 384" "1024 384" "1024 256"; do
  set -- $cfg
  CGASM_TILE_ROWS=$1 CGASM_TILE_THREADS=$2 timeout 600 python bench.py --cells 128 --scatter tiled --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench128_tiled_$1_$2.json 2> gpurun_out/bench128_tiled_$1_$2.err
  tail -3 gpurun_out/bench128_tiled_$1_$2.err
done
timeout 900 python bench.py --cells 256 --scatter tiled --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench256_tiled.json 2> gpurun_out/bench256_tiled.err
tail -3 gpurun_out/bench256_tiled.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tiled_momentum -s 2 -c 1 -o gpurun_out/prof_tiled_mom -f \
  python bench.py --cells 96 --scatter tiled --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tiled.log 2>&1
tail -3 gpurun_out/ncu_tiled.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench*_tiled*.json')):
    try:
        d=json.loads(open(f).read()); r=d['roofline']
        print(f, 'value %.0f'%d['value'], 'mom %.2f ms'%r['kernel_ms'], 'tra %.2f ms'%r['tracer']['kernel_ms'], 'frac %.3f'%r['frac'], 'setup %.1f'%d['setup_s'], 'e2e', d['e2e'] and round(d['e2e']['value']))
    except Exception as e: print(f,'ERR',e)
PY
