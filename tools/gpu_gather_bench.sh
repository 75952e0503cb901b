#!/bin/bash
# gather path: triangle bench + ncu capture of k_cellmat / k_lists (no tests)
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-g}
mkdir -p gpurun_out
timeout 300 python bench.py --mesh-type tri --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/BENCH_tri_gather_$TAG.json 2> gpurun_out/BENCH_tri_gather_$TAG.err; echo "tri gather rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/BENCH_tri_gather_$TAG.json"))
print({k:d[k] for k in ("value","ms_per_step","parts","gpu_launches")}, d["verify"]["ok"], d["roofline"]["frac"], d["config"]["path"])
PY
bash tools/gpu_tri_prof.sh tri_gather_$TAG "k_cellmat|k_lists" 1024
