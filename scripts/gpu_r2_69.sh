#!/bin/bash
# Round 2, GPU call 69 (1 GPU): the regression test of the upload race.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_surface_gpu.py -m gpu -q -x -k "many_handles" 2>&1 | tail -3 | tee gpurun_out/r2_69_pytest.log
