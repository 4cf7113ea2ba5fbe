#!/bin/bash
# Round 2, GPU call 49 (1 GPU): library set-up per phase (box, renumbered, Delaunay 10 M tets) and the configs block's
# library_setup_s after the OpenMP thread count is given back behind the single-thread CPU leg.
mkdir -p gpurun_out
timeout 1200 python scripts/prof_setup.py 128 1500000 2>&1 | grep -v "^$" | tee gpurun_out/r2_49_setup.txt
timeout 900 python bench.py --cells 64 --no-e2e --delaunay-points 200000 --cpu-cells 48 > gpurun_out/r2_49_c.json 2> gpurun_out/r2_49_c.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_49_c.json').read().strip().splitlines()[-1])
print("configs library_setup_s", [round(c.get("library_setup_s",-1),2) for c in d["configs"]])
PY
