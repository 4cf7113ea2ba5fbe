#!/bin/bash
# Round 2, GPU call 48 (1 GPU): why the configs block reports 4 s / 17 s of library set-up where the same calls take 0.5 s / 2 s alone.
mkdir -p gpurun_out
show() { python - "$1" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], "setup_s", round(d["setup_s"],2), [round(c.get("library_setup_s",-1),2) for c in d["configs"]])
PY
}
timeout 900 python bench.py --cells 64 --no-cpu-baseline --no-e2e --delaunay-points 200000 > gpurun_out/r2_48_a.json 2> gpurun_out/r2_48_a.err; show gpurun_out/r2_48_a.json
OMP_WAIT_POLICY=passive timeout 900 python bench.py --cells 64 --no-cpu-baseline --no-e2e --delaunay-points 200000 > gpurun_out/r2_48_b.json 2> gpurun_out/r2_48_b.err; show gpurun_out/r2_48_b.json
timeout 900 python bench.py --cells 64 --no-e2e --delaunay-points 200000 --cpu-cells 48 > gpurun_out/r2_48_c.json 2> gpurun_out/r2_48_c.err; show gpurun_out/r2_48_c.json
