#!/bin/bash
# Round 2, GPU call 4 (1 GPU): full GPU suite after the momentum dispatch fix (additive pass, v_* goldens on the CUDA element routines).
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2_4_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2_4_pytest.log
tail -5 gpurun_out/r2_4_pytest.log
timeout 600 python -c "
import sys; sys.path.insert(0,'.')
import bench, json
print(json.dumps(bench.example_configs(0, 128, 2048), indent=1))
" > gpurun_out/r2_4_configs.json 2> gpurun_out/r2_4_configs.err; grep -E "config|gel_s" gpurun_out/r2_4_configs.json | paste - - | cut -c1-220
