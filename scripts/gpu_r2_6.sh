#!/bin/bash
# Round 2, GPU call 6 (1 GPU): COO + C-client GPU tests; full ncu captures (source-level) of the momentum, tracer and fused kernels at 96^3.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_coo_gpu.py tests/test_abi.py -m gpu -x -q > gpurun_out/r2_6_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_6_pytest.log
tail -4 gpurun_out/r2_6_pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:staged_momentum_kernel -s 2 -c 1 -o gpurun_out/r2_prof_mom -f \
  python bench.py --cells 96 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/r2_6_ncu_mom.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:staged_advdiff_kernel -s 2 -c 1 -o gpurun_out/r2_prof_adv -f \
  python bench.py --cells 96 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/r2_6_ncu_adv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:staged_fused_kernel -s 1 -c 1 -o gpurun_out/r2_prof_fused -f \
  python bench.py --cells 96 --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/r2_6_ncu_fused.log 2>&1
ls -la gpurun_out/*.ncu-rep
