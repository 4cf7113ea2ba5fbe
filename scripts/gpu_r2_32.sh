#!/bin/bash
# Round 2, GPU call 32 (1 GPU): searched strips (30 pushes instead of 33 around an interior node of a Kuhn mesh): full suite +
# A/B against the previous build (the same library with CGASM_STRIP_SEARCH=0 and the build before the change).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_32_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_32_pytest.log
tail -4 gpurun_out/r2_32_pytest.log
for c in 128 256; do
  for v in mixed2 search mixed2 search; do
    CGASM_LIB=$PWD/ab/libcgasm_$v.so timeout 300 python scripts/ab_kernels.py $c 2>&1 | tail -1
  done
done | tee gpurun_out/r2_32_ab.txt
CGASM_STRIP_SEARCH=0 CGASM_LIB=$PWD/ab/libcgasm_search.so timeout 300 python scripts/ab_kernels.py 256 2>&1 | tail -1 | sed 's/^/search off: /' | tee -a gpurun_out/r2_32_ab.txt
