#!/bin/bash
# Round-end style check: full GPU tests, smoke, default bench, reference arm, ncu launch list with
# DRAM traffic of the S3 launches (the numbers behind roofline.traffic).
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; tail -2 gpurun_out/bench_reference.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,l1tex__throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:gather_ -s 2 -c 4 --csv --log-file gpurun_out/launches_S3_gather.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_S3.log 2>&1
tail -2 gpurun_out/ncu_S3.log
cat gpurun_out/bench_final.json | cut -c1-300
cat gpurun_out/bench_reference.json | cut -c1-300
