#!/bin/bash
# Staged STRIP kernels: full GPU suite (row blocks changed), v1 strip tests, sweep, ncu capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest full rc=$?"; tail -3 gpurun_out/pytest_gpu.log
CGASM_STRIP_GLOBAL=1 timeout 600 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k strip > gpurun_out/pytest_strip_global.log 2>&1
echo "pytest strip(global) rc=$?"; tail -2 gpurun_out/pytest_strip_global.log
run() { # name cells env...
  name=$1; cells=$2; shift 2
  env CGASM_DEBUG=1 "$@" timeout 900 python bench.py --cells $cells --scatter strip --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
}
run staged128_n4_mb4 128 CGASM_STRIP_NBUF=4 CGASM_STRIP_MINB=4 CGASM_STRIP_MINB_ADV=4
run staged128_n4_mb3 128 CGASM_STRIP_NBUF=4 CGASM_STRIP_MINB=3 CGASM_STRIP_MINB_ADV=3
run staged128_n5_mb3 128 CGASM_STRIP_NBUF=5 CGASM_STRIP_MINB=3 CGASM_STRIP_MINB_ADV=4
run global128_n4_mb4 128 CGASM_STRIP_GLOBAL=1 CGASM_STRIP_NBUF=4 CGASM_STRIP_MINB=4 CGASM_STRIP_MINB_ADV=4
run staged256_n4_mb4 256 CGASM_STRIP_NBUF=4 CGASM_STRIP_MINB=4 CGASM_STRIP_MINB_ADV=4
grep -h "cgasm\]" gpurun_out/bench_staged128_n4_mb4.err gpurun_out/bench_staged256_n4_mb4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_staged*.json')+glob.glob('gpurun_out/bench_global*.json')):
    try:
        d=json.loads(open(f).read()); r=d['roofline']
        print(f, 'value %.0f'%d['value'], 'mom %.2f ms'%r['kernel_ms'], 'tra %.2f ms'%r['tracer']['kernel_ms'], 'frac %.3f'%r['frac'], 'setup %.1f'%d['setup_s'])
    except Exception as e: print(f,'ERR',e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:staged_momentum -s 2 -c 1 -o gpurun_out/prof_staged_mom -f \
  env CGASM_STRIP_NBUF=4 CGASM_STRIP_MINB=4 python bench.py --cells 96 --scatter strip --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_staged.log 2>&1
tail -2 gpurun_out/ncu_staged.log
