#!/bin/bash
# Round 2, GPU call 16 (1 GPU): full suite (is the surface failure of call 14 order-dependent?), initcheck on the GATHER
# surface test, ncu capture of the CMC expansion kernel.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_16_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_16_pytest.log
tail -8 gpurun_out/r2_16_pytest.log
timeout 900 compute-sanitizer --tool initcheck --print-limit 10 python -m pytest tests/test_surface_gpu.py -m gpu -q -k "cube-parallel and gather" > gpurun_out/r2_16_initcheck.log 2>&1
grep -E "Uninitialized|ERROR SUMMARY|passed|failed" gpurun_out/r2_16_initcheck.log | head -12
timeout 600 ncu --set full --import-source on --clock-control none -k regex:cmc_expand -c 1 -o gpurun_out/r2_cmc_expand python scripts/bench_cmc.py 64 > gpurun_out/r2_16_ncu_cmc.log 2>&1; tail -2 gpurun_out/r2_16_ncu_cmc.log
