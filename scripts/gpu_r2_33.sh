#!/bin/bash
# Round 2, GPU call 33 (1 GPU): occupancy classes of the staged kernels on a Delaunay mesh (A/B inside one process),
# parity tests of the unstructured meshes.
mkdir -p gpurun_out
CGASM_VERBOSE=1 timeout 900 python scripts/ab_classes.py 600000 2>&1 | grep -v "^cgasm: strip\|^$" | tee gpurun_out/r2_33_ab_classes.txt
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "delaunay or unstructured or cube or shuffled" 2>&1 | tail -3 | tee gpurun_out/r2_33_pytest.log
