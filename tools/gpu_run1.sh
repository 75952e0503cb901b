#!/bin/bash
# first GPU run of k_gath: parity, A/B bench at 2048^2, ncu capture
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "structured or row_blocks or newton or config1" > gpurun_out/t_gath_ng2.log 2>&1; echo "tests ng2 rc=$?"
tail -3 gpurun_out/t_gath_ng2.log
GM_GATH_NG=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "structured_path" > gpurun_out/t_gath_ng1.log 2>&1; echo "tests ng1 rc=$?"
tail -3 gpurun_out/t_gath_ng1.log
B="python bench.py --mesh 2048 --power-index 0.6 --steps 5 --warmup 3 --no-cpu --no-e2e"
timeout 200 $B > gpurun_out/b_gath_ng2.json 2> gpurun_out/b_gath_ng2.err; echo "bench ng2 rc=$?"; cat gpurun_out/b_gath_ng2.json
GM_GATH_NG=1 timeout 200 $B > gpurun_out/b_gath_ng1.json 2> gpurun_out/b_gath_ng1.err; echo "bench ng1 rc=$?"; cat gpurun_out/b_gath_ng1.json
GM_CART_KERNEL=sweep timeout 200 $B > gpurun_out/b_sweep.json 2> gpurun_out/b_sweep.err; echo "bench sweep rc=$?"; cat gpurun_out/b_sweep.json
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_gath -c 1 -f -o gpurun_out/prof_gath_r01_2048_v1 python bench.py --mesh 2048 --power-index 0.6 --steps 1 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_gath.log 2>&1; echo "ncu rc=$?"
