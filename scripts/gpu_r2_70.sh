#!/bin/bash
# Round 2, GPU call 70 (1 GPU): the full GPU suite of the final tree (upload-race regression test included).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2_70_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_70_pytest.log; tail -3 gpurun_out/r2_70_pytest.log
