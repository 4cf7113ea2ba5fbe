#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q -k "gather or sparsity or element" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --cells 128 --scatter gather --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench128_gather.json 2> gpurun_out/bench128_gather.err
tail -3 gpurun_out/bench128_gather.err
timeout 900 python bench.py --cells 256 --scatter gather --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench256_gather.json 2> gpurun_out/bench256_gather.err
tail -3 gpurun_out/bench256_gather.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct --clock-control none -k regex:gather_ -s 4 -c 4 --csv --log-file gpurun_out/launches_gather128.csv \
  python bench.py --cells 128 --scatter gather --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_gather.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/bench*_gather*.json')):
    try:
        d=json.loads(open(f).read()); r=d['roofline']
        print(f, 'value %.0f'%d['value'], 'mom %.2f ms'%r['kernel_ms'], 'tra %.2f ms'%r['tracer']['kernel_ms'], 'frac %.3f'%r['frac'], 'setup %.1f'%d['setup_s'], 'e2e', d['e2e'] and round(d['e2e']['value']))
    except Exception as e: print(f,'ERR',e)
PY
