#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "tiled or sparsity or element" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
for cfg in "1024 1 0" "1024 1 768" "1024 1 512" "512 1 0"; do
  set -- $cfg
  export CGASM_TILE_ROWS=$1 CGASM_TILE_CLUSTER=$2
  if [ "$3" != "0" ]; then export CGASM_TILE_THREADS=$3; else unset CGASM_TILE_THREADS; fi
  timeout 600 python bench.py --cells 128 --scatter tiled --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench128_tiled_$1_$2_$3.json 2> gpurun_out/bench128_tiled_$1_$2_$3.err
  tail -3 gpurun_out/bench128_tiled_$1_$2_$3.err
done
unset CGASM_TILE_ROWS CGASM_TILE_CLUSTER CGASM_TILE_THREADS
timeout 900 python bench.py --cells 256 --scatter tiled --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench256_tiled.json 2> gpurun_out/bench256_tiled.err
tail -3 gpurun_out/bench256_tiled.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tiled_momentum -s 2 -c 1 -o gpurun_out/prof_tiled_mom4 -f \
  python bench.py --cells 96 --scatter tiled --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_tiled.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench*_tiled*.json')):
    try:
        d=json.loads(open(f).read()); r=d['roofline']
        print(f, 'value %.0f'%d['value'], 'mom %.2f ms'%r['kernel_ms'], 'tra %.2f ms'%r['tracer']['kernel_ms'], 'frac %.3f'%r['frac'], 'setup %.1f'%d['setup_s'], 'e2e', d['e2e'] and round(d['e2e']['value']))
    except Exception as e: print(f,'ERR',e)
PY
