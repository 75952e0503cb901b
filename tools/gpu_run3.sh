#!/bin/bash
# quick A/B: parity subset + bench at 2048^2 (no ncu).  usage: gpu_run3.sh <tag>
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-x}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "structured_path or row_blocks" > gpurun_out/t_$TAG.log 2>&1; echo "tests rc=$?"
tail -2 gpurun_out/t_$TAG.log
B="python bench.py --mesh 2048 --power-index 0.6 --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 200 $B > gpurun_out/b_$TAG.json 2> gpurun_out/b_$TAG.err; echo "bench rc=$?"; cut -c1-240 gpurun_out/b_$TAG.json
