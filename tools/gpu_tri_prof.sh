#!/bin/bash
# ncu --set full capture of the generic-path kernels on a triangle mesh (after a plain run has exited 0)
cd "$GRAFT_REPO_ROOT" || exit 1
TAG=${1:-tri}; KREGEX=${2:-k_scatter}; MESH=${3:-512}
mkdir -p gpurun_out
timeout 200 python bench.py --mesh-type tri --mesh $MESH --steps 2 --warmup 3 --no-cpu --no-e2e --no-verify > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
cut -c1-300 gpurun_out/${TAG}_plain.log | tail -1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:$KREGEX -c 2 -f -o gpurun_out/prof_$TAG python bench.py --mesh-type tri --mesh $MESH --steps 1 --warmup 1 --no-cpu --no-e2e --no-verify > gpurun_out/ncu_$TAG.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/prof_${TAG}_src.csv 2>/dev/null
rm -f gpurun_out/prof_$TAG.ncu-rep
