#!/usr/bin/env python
"""End-to-end Newton record (north_star: "end-to-end Newton time is reported separately from assembly").

Runs the restated NLsolve drivers (solver.py) through the B200 operator plug-in on the reference's own runs
(config 1 = examples/simple.jl, the study-1 mesh of conv_study_square.csv:2, a turan-style power-law cavity) and writes
iterations, f/g calls, assembly time (through gm_assemble: H2D + kernels + D2H) and linear-solve time (SciPy SuperLU here,
UMFPACK in the reference) to profiles/r02_newton_e2e.json.  Needs a B200.
    python tools/newton_record.py
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from _pkg import gm  # noqa: E402

RUNS = [
    ("config1 examples/simple.jl 50x50 n=1 newton+BackTracking", 50, cases.SIMPLE, dict(Pr=0.7, Ra=1e3, n=1.0), "newton"),
    ("study1 62x62 wave n=0.6 newton+BackTracking (conv_study_square.csv:2)", 62, cases.WAVE, dict(Pr=100.0, Ra=1e4, n=0.6), "newton"),
    ("turan-style 96x96 n=0.6 Ra=1e4 Pr=100 newton+BackTracking", 96, cases.TURAN, dict(Pr=100.0, Ra=1e4, n=0.6), "newton"),
    ("basak-style 40x40 n=1 trust_region (validation-basak/basak.jl:20-26, CI size)", 40, cases.SIMPLE, dict(Pr=0.7, Ra=1e3, n=1.0), "trust_region"),
]
RUNS.append(("config3 validation-kuehn annulus, 2 x 18 x 128 generated triangles, Pr=0.706 Ra=4.7e4 n=1 newton+BackTracking (generic gather path)",
             None, cases.KUEHN, dict(Pr=0.706, Ra=4.7e4, n=1.0), "newton"))
only = sys.argv[1] if len(sys.argv) > 1 else None   # substring filter: the other runs keep their committed records
recfile = os.path.join(ROOT, "profiles", "r02_newton_e2e.json")
out = []
if only and os.path.exists(recfile):
    out = [r for r in json.load(open(recfile))["runs"] if only not in r["run"]]
for name, N, bc, prm, method in RUNS:
    if only and only not in name:
        continue
    sp = cases.cart(N, bc=bc) if N else cases.spaces(gm.fe.annulus_model(nr=18, nt=128, jitter=0.35, seed=7), bc)
    op = gm.operator.B200FEOperator(sp, **prm)
    t0 = time.perf_counter()
    drv = gm.solver.newton if method == "newton" else gm.solver.trust_region
    r = drv(op, np.zeros(sp.n_total), ftol=1e-8, xtol=1e-50, iterations=400)
    wall = time.perf_counter() - t0
    rec = {"run": name, "cells": sp.model.ncells, "n_free": sp.n_total, "nnz": op.nnz, "method": method, "iterations": r.iterations,
           "f_calls": r.f_calls, "g_calls": r.g_calls, "converged": bool(r.f_converged), "residual_inf": r.residual_norm,
           "t_total_s": wall, "t_assembly_s": r.t_assembly, "t_linsolve_s": r.t_linsolve,
           "assembly_share": r.t_assembly / wall}
    if bc is cases.WAVE:
        rec["Nu_bot_avg"] = gm.solver.nu_bot_avg(sp, r.x)
    op.close()
    out.append(rec)
    print(json.dumps(rec))
json.dump({"what": "Newton end to end through the B200 plug-in; linear solves = SciPy SuperLU on the host", "runs": out},
          open(recfile, "w"), indent=1)
