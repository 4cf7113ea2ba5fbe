#!/bin/bash
# Round 2, GPU call 67 (1 GPU): the stress loop after the uploads drain the legacy stream (cg_upload).
mkdir -p gpurun_out
for s in GATHER GATHER STRIP; do STRESS_SCATTER=$s timeout 300 python scripts/stress_surface.py 1000 2>&1 | grep -E "iterations off|rror|^iteration" | tail -3 | cut -c1-200; done | tee gpurun_out/r2_67_stress.txt
