#!/bin/bash
# First GPU contact: parity tests (element-centric variants), short benches, launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt; free -g >> gpurun_out/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "not tiled" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
for v in atomic warpagg coloured; do
  timeout 600 python bench.py --cells 128 --scatter $v --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench128_$v.json 2> gpurun_out/bench128_$v.err
done
timeout 900 python bench.py --cells 256 --scatter atomic --steps 5 --warmup 3 > gpurun_out/bench256_atomic.json 2> gpurun_out/bench256_atomic.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_atomic128.csv \
  python bench.py --cells 128 --scatter atomic --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_bench.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
cat gpurun_out/bench128_*.json gpurun_out/bench256_atomic.json
