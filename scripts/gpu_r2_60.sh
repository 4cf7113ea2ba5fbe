#!/bin/bash
# Round 2, GPU call 60 (1 GPU): stress loop for the rare cube-parallel + GATHER surface failure.
mkdir -p gpurun_out
timeout 600 python scripts/stress_surface.py 400 2>&1 | tail -15 | tee gpurun_out/r2_60_stress.txt
