#!/bin/bash
# Round 2, GPU call 47 (1 GPU): where the library set-up time goes on scattered numberings.
mkdir -p gpurun_out
CGASM_DEBUG=1 timeout 900 python scripts/prof_setup.py 128 2>&1 | grep -v "^$" | tee gpurun_out/r2_47_setup.txt
nproc | tee -a gpurun_out/r2_47_setup.txt
