#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --cells 128 --steps 5 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench128.json 2> gpurun_out/bench128.err; tail -2 gpurun_out/bench128.err
python -c "
import json; d=json.loads(open('gpurun_out/bench128.json').read()); r=d['roofline']; print('128: value %.0f mom %.2f tra %.2f'%(d['value'], r['kernel_ms'], r['tracer']['kernel_ms']))"
