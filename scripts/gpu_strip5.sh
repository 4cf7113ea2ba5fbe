#!/bin/bash
# Staged STRIP kernels with plan L2 prefetch + batched staging ids. Tests (strip), sweep, ncu capture.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "strip" > gpurun_out/pytest_strip.log 2>&1
echo "pytest strip rc=$?"; tail -2 gpurun_out/pytest_strip.log
run() { # name cells scatter env...
  name=$1; cells=$2; sc=$3; shift 3
  env "$@" timeout 900 python bench.py --cells $cells --scatter $sc --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_$name.json 2> gpurun_out/bench_$name.err
}
run s5_n3_mb4_on0 128 strip CGASM_STRIP_NBUF=3 CGASM_STRIP_MINB=4 CGASM_STRIP_ONPF=0 CGASM_STRIP_MINB_ADV=4
run s5_n3_mb4_on1 128 strip CGASM_STRIP_NBUF=3 CGASM_STRIP_MINB=4 CGASM_STRIP_ONPF=1 CGASM_STRIP_MINB_ADV=5
run s5_n3_mb3_on1 128 strip CGASM_STRIP_NBUF=3 CGASM_STRIP_MINB=3 CGASM_STRIP_ONPF=1 CGASM_STRIP_MINB_ADV=3
run s5_n4_mb3_on1 128 strip CGASM_STRIP_NBUF=4 CGASM_STRIP_MINB=3 CGASM_STRIP_ONPF=1 CGASM_STRIP_MINB_ADV=4
run s5_n4_mb3_on0 128 strip CGASM_STRIP_NBUF=4 CGASM_STRIP_MINB=3 CGASM_STRIP_ONPF=0 CGASM_STRIP_MINB_ADV=3
run s5_256_n3_mb4_on0 256 strip CGASM_STRIP_NBUF=3 CGASM_STRIP_MINB=4 CGASM_STRIP_ONPF=0 CGASM_STRIP_MINB_ADV=4
run s5_256_n4_mb3_on1 256 strip CGASM_STRIP_NBUF=4 CGASM_STRIP_MINB=3 CGASM_STRIP_ONPF=1 CGASM_STRIP_MINB_ADV=4
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench_s5_*.json')):
    try:
        d=json.loads(open(f).read()); r=d['roofline']
        print(f, 'value %.0f'%d['value'], 'mom %.2f ms'%r['kernel_ms'], 'tra %.2f ms'%r['tracer']['kernel_ms'], 'frac %.3f'%r['frac'], 'setup %.1f'%d['setup_s'])
    except Exception as e: print(f,'ERR',e)
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:staged_momentum -s 2 -c 1 -o gpurun_out/prof_staged_mom4 -f \
  env CGASM_STRIP_NBUF=3 CGASM_STRIP_MINB=4 python bench.py --cells 96 --scatter strip --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_staged4.log 2>&1
tail -2 gpurun_out/ncu_staged4.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:staged_advdiff -s 2 -c 1 -o gpurun_out/prof_staged_adv4 -f \
  env CGASM_STRIP_NBUF=4 CGASM_STRIP_MINB_ADV=4 python bench.py --cells 96 --scatter strip --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_staged4a.log 2>&1
tail -2 gpurun_out/ncu_staged4a.log
