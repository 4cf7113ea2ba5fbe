#!/bin/bash
# Round 2, GPU call 45 (1 GPU): bank-group assignment of the staged node records on a Delaunay mesh: A/B by plan (two
# processes, same mesh), unstructured parity tests.
mkdir -p gpurun_out
for v in 0 1 0 1; do
  CGASM_STRIP_BANKS=$v timeout 900 python scripts/ab_classes.py 600000 2>&1 | grep "classes on" | tail -2 | sed "s/^/banks $v: /"
done | tee gpurun_out/r2_45_ab_banks.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "delaunay or occupancy or cube or shuffled or prectangle" 2>&1 | tail -3 | tee gpurun_out/r2_45_pytest.log
