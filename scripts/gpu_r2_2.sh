#!/bin/bash
# Round 2, GPU call 2 (1 GPU): full GPU suite on the fixed dispatch, the new bench.py at N=1 (configs block, honest CPU arm).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_2_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_2_pytest.log
tail -5 gpurun_out/r2_2_pytest.log
CGASM_DEBUG=1 timeout 1200 python bench.py > gpurun_out/r2_2_bench_n1.json 2> gpurun_out/r2_2_bench_n1.err; echo "bench exit $?"; cut -c1-600 gpurun_out/r2_2_bench_n1.json; tail -5 gpurun_out/r2_2_bench_n1.err
