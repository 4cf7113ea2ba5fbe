#!/bin/bash
# Staged STRIP kernels after the Morton sampling fix + volatile shared loads: tests, sweep, ncu capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest full rc=$?"; tail -3 gpurun_out/pytest_gpu.log
run() { # name cells scatter env...
  name=$1; cells=$2; sc=$3; shift 3
  env CGASM_DEBUG=1 "$@" timeout 900 python bench.py --cells $cells --scatter $sc --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
}
run s3_staged128_mb4 128 strip CGASM_STRIP_MINB=4 CGASM_STRIP_MINB_ADV=4
run s3_staged128_mb3 128 strip CGASM_STRIP_MINB=3 CGASM_STRIP_MINB_ADV=3
run s3_staged128_n5 128 strip CGASM_STRIP_NBUF=5 CGASM_STRIP_MINB=3 CGASM_STRIP_MINB_ADV=4
run s3_global128 128 strip CGASM_STRIP_GLOBAL=1
run s3_walk128 128 gather
run s3_staged256_mb4 256 strip CGASM_STRIP_MINB=4 CGASM_STRIP_MINB_ADV=4
run s3_staged256_mb3 256 strip CGASM_STRIP_MINB=3 CGASM_STRIP_MINB_ADV=4
grep -h "cgasm\]" gpurun_out/bench_s3_staged128_mb4.err gpurun_out/bench_s3_staged256_mb4.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_s3_*.json')):
    try:
        d=json.loads(open(f).read()); r=d['roofline']
        print(f, 'value %.0f'%d['value'], 'mom %.2f ms'%r['kernel_ms'], 'tra %.2f ms'%r['tracer']['kernel_ms'], 'frac %.3f'%r['frac'], 'setup %.1f'%d['setup_s'])
    except Exception as e: print(f,'ERR',e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:staged_momentum -s 2 -c 1 -o gpurun_out/prof_staged_mom2 -f \
  env CGASM_STRIP_MINB=4 python bench.py --cells 96 --scatter strip --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_staged2.log 2>&1
tail -2 gpurun_out/ncu_staged2.log
