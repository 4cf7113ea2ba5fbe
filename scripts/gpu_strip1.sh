#!/bin/bash
# First GPU contact of the STRIP variant: parity tests, then a small option sweep, then one ncu capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "strip" > gpurun_out/pytest_strip.log 2>&1
echo "pytest strip rc=$?"; tail -5 gpurun_out/pytest_strip.log
run() { # name cells env...
  name=$1; cells=$2; shift 2
  env "$@" timeout 900 python bench.py --cells $cells --scatter strip --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
}
run strip128_n4_mb4 128 CGASM_STRIP_NBUF=4 CGASM_STRIP_MINB=4 CGASM_STRIP_MINB_ADV=4
run strip128_n4_mb3 128 CGASM_STRIP_NBUF=4 CGASM_STRIP_MINB=3 CGASM_STRIP_MINB_ADV=3
run strip128_n4_mb45 128 CGASM_STRIP_NBUF=4 CGASM_STRIP_MINB=4 CGASM_STRIP_MINB_ADV=5
run strip128_n5_mb3 128 CGASM_STRIP_NBUF=5 CGASM_STRIP_MINB=3 CGASM_STRIP_MINB_ADV=4
run strip128_n5_mb2 128 CGASM_STRIP_NBUF=5 CGASM_STRIP_MINB=2 CGASM_STRIP_MINB_ADV=3
run strip256_n4_mb4 256 CGASM_STRIP_NBUF=4 CGASM_STRIP_MINB=4 CGASM_STRIP_MINB_ADV=4
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_strip*.json')):
    try:
        d=json.loads(open(f).read()); r=d['roofline']
        print(f, 'value %.0f'%d['value'], 'mom %.2f ms'%r['kernel_ms'], 'tra %.2f ms'%r['tracer']['kernel_ms'], 'frac %.3f'%r['frac'], 'setup %.1f'%d['setup_s'])
    except Exception as e: print(f,'ERR',e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:strip_momentum -s 2 -c 1 -o gpurun_out/prof_strip_mom -f \
  env CGASM_STRIP_NBUF=4 CGASM_STRIP_MINB=4 python bench.py --cells 96 --scatter strip --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_strip.log 2>&1
tail -2 gpurun_out/ncu_strip.log
