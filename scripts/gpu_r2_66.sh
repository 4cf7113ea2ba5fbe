#!/bin/bash
# Round 2, GPU call 66 (1 GPU): is it host memory? the stress loop with glibc heap poisoning.
mkdir -p gpurun_out
run() { echo "== $*"; env "$@" timeout 200 python scripts/stress_surface.py 300 2>&1 | grep -E "iterations off|rror|^iteration" | tail -3 | cut -c1-250; }
{ run MALLOC_PERTURB_=165; run MALLOC_PERTURB_=0; run MALLOC_PERTURB_=255 OMP_NUM_THREADS=1; } | tee gpurun_out/r2_66_stress.txt
