#!/bin/bash
# Round 2, GPU call 27 (1 GPU): pipelined staged kernels (strip_pipe.cu): parity (bitwise vs the FIFO kernels) and A/B timing.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "pipelined or shuffled or delaunay or large or strip_matches or fused" > gpurun_out/r2_27_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_27_pytest.log
tail -6 gpurun_out/r2_27_pytest.log
for c in 128 256; do
  for e in 0 1 0 1; do
    CGASM_STRIP_PIPE=$e timeout 300 python scripts/ab_kernels.py $c 2>&1 | tail -1 | sed "s/^/pipe=$e /"
  done
done | tee gpurun_out/r2_27_ab.txt
