#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "partition or gather or config3 or two_handles" > gpurun_out/t_part.log 2>&1; echo "tests rc=$?"; tail -15 gpurun_out/t_part.log
