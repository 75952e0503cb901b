#!/bin/bash
# gather path: GPU parity tests, triangle bench (gather vs scatter), launch-level profile of k_cellmat / k_lists
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-g1}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "gather or config3 or two_handles or check_finite or scatter_path or config1" > gpurun_out/t_gather_$TAG.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/t_gather_$TAG.log
timeout 300 python bench.py --mesh-type tri --steps 10 --warmup 3 --no-e2e --no-cpu > gpurun_out/BENCH_tri_gather_$TAG.json 2> gpurun_out/BENCH_tri_gather_$TAG.err; echo "tri gather rc=$?"; cut -c1-200 gpurun_out/BENCH_tri_gather_$TAG.json
python - <<PY
import json
d=json.load(open("gpurun_out/BENCH_tri_gather_$TAG.json"))
print({k:d[k] for k in ("value","ms_per_step","parts","verify","gpu_launches")}, d["roofline"]["frac"], d["config"]["path"])
PY
bash tools/gpu_tri_prof.sh tri_gather_$TAG "k_cellmat|k_lists" 1024
