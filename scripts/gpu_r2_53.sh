#!/bin/bash
# Round 2, GPU call 53 (1 GPU): diagonal by row sum (a0 -= every flushed column instead of u . sc per pair): full GPU suite,
# A/B against the previous build.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_53_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_53_pytest.log; tail -3 gpurun_out/r2_53_pytest.log
for c in 128 256; do
  for v in base rowsum base rowsum; do
    CGASM_LIB=$PWD/ab/libcgasm_$v.so timeout 300 python scripts/ab_kernels.py $c 2>&1 | tail -1
  done
done | tee gpurun_out/r2_53_ab_rowsum.txt
