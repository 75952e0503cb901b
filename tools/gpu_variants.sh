#!/bin/bash
# A/B of library builds under build_variants/ (bench only, triangle mesh): each variant replaces the library of the scratch copy
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
cp gemotion.jl_b200/libgemotion_b200.so /tmp/lib_default.so
for f in /tmp/lib_default.so build_variants/lib_*.so; do
  cp "$f" gemotion.jl_b200/libgemotion_b200.so
  n=$(basename "$f" .so)
  timeout 200 python bench.py --mesh-type tri --steps 10 --warmup 3 --no-e2e --no-cpu ${BENCH_ARGS} > gpurun_out/var_$n.json 2> gpurun_out/var_$n.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/var_$n.json"))
    print("$n", round(d["ms_per_step"],3), "ms", round(d["value"],1), "Mcells/s  jac", round(d["parts"]["jacobian_only"]["ms"],3), "res", round(d["parts"]["residual_only"]["ms"],3), "verify", d["verify"]["ok"])
except Exception as e:
    print("$n failed", e)
PY
done
cp /tmp/lib_default.so gemotion.jl_b200/libgemotion_b200.so
