#!/bin/bash
# full validation of the HEAD build: all GPU tests, smoke, default bench (4096^2, n=1.4), reference arm, launch list + ncu full capture of
# the dominant kernel, triangle bench (gather + scatter) with its ncu capture, quads through the generic gather path
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-final}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/t_all_$TAG.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/t_all_$TAG.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$TAG.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke_$TAG.log
timeout 600 python bench.py > gpurun_out/BENCH_default_$TAG.json 2> gpurun_out/BENCH_default_$TAG.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/BENCH_default_$TAG.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/BENCH_reference_$TAG.json 2> gpurun_out/BENCH_reference_$TAG.err; echo "reference rc=$?"
timeout 300 python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-verify > gpurun_out/plain_$TAG.log 2>&1 && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$TAG.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-verify > gpurun_out/ncu1_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 500 ncu --set full --import-source on --clock-control none -k regex:k_mma -c 1 -f -o gpurun_out/prof_mma_4096_$TAG python bench.py --steps 1 --warmup 1 --no-cpu --no-e2e --no-verify > gpurun_out/ncu2_$TAG.log 2>&1; echo "ncu full rc=$?"
ncu -i gpurun_out/prof_mma_4096_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_mma_4096_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_mma_4096_$TAG.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_mma_4096_${TAG}_src.csv 2>/dev/null
rm -f gpurun_out/prof_mma_4096_$TAG.ncu-rep
timeout 300 python bench.py --mesh-type tri --steps 10 --warmup 3 --no-cpu > gpurun_out/BENCH_tri_$TAG.json 2> gpurun_out/BENCH_tri_$TAG.err; echo "tri gather rc=$?"; cut -c1-200 gpurun_out/BENCH_tri_$TAG.json
timeout 300 python bench.py --mesh-type tri --path scatter --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/BENCH_tri_scatter_$TAG.json 2> gpurun_out/BENCH_tri_scatter_$TAG.err; echo "tri scatter rc=$?"
timeout 300 python bench.py --mesh 1024 --path gather --steps 5 --warmup 3 --no-cpu --no-e2e > gpurun_out/BENCH_quad_gather_1024_$TAG.json 2> gpurun_out/BENCH_quad_gather_1024_$TAG.err; echo "quad gather rc=$?"; cut -c1-200 gpurun_out/BENCH_quad_gather_1024_$TAG.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_tri_$TAG.csv python bench.py --mesh-type tri --steps 2 --warmup 3 --no-cpu --no-e2e --no-verify > gpurun_out/ncu3_$TAG.log 2>&1; echo "tri launch list rc=$?"
bash tools/gpu_tri_prof.sh tri_gather_$TAG "k_cellmat|k_lists" 1024
