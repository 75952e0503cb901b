#!/bin/bash
# full validation of the default kernel: all GPU tests, default bench (4096^2, n=1.4), launch list, ncu full capture of the dominant kernel
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-final}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/t_all_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_all_$TAG.log
timeout 600 python bench.py > gpurun_out/BENCH_default_$TAG.json 2> gpurun_out/BENCH_default_$TAG.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/BENCH_default_$TAG.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/BENCH_reference_$TAG.json 2> gpurun_out/BENCH_reference_$TAG.err; echo "reference rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-verify > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-verify > gpurun_out/ncu1_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 500 ncu --set full --import-source on --clock-control none -k regex:k_mma -c 1 -f -o gpurun_out/prof_mma_4096_$TAG python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-verify > gpurun_out/ncu2_$TAG.log 2>&1; echo "ncu full rc=$?"
timeout 300 python bench.py --mesh-type tri --steps 5 --warmup 3 --no-e2e > gpurun_out/BENCH_tri_$TAG.json 2> gpurun_out/BENCH_tri_$TAG.err; echo "tri rc=$?"
