"""Regenerates tests/golden/square{N}_wave_n06_x.npy: the converged Newton iterate of the
conv-study case of /root/reference/examples/study1/study1.jl:27-43,72-90 (Pr=100, Ra=1e4, n=0.6,
wave BC, zero initial guess, Newton+BackTracking, ftol 1e-8) computed with the CPU oracle.
Usage: python tests/golden/make_square_golden.py 62 [86]      (~2 min for 62, ~6 min for 86)
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(HERE, ".."))
from _pkg import gm  # noqa: E402
from orc_operator import OracleOperator  # noqa: E402

WAVE = dict(V=[1, 2, 3, 4, 5, 6, 7, 8], Ttags=[5, 7, 8, 1, 2],
            Texpr=[lambda x: np.sin(np.pi * x[0]), 0.0, 0.0, 0.0, 0.0])

for N in [int(a) for a in sys.argv[1:]] or [62]:
    m = gm.fe.CartesianModel((0, 1, 0, 1), (N, N))
    sp = gm.fe.build_spaces(m, WAVE["V"], WAVE["Ttags"], WAVE["Texpr"])
    op = OracleOperator(sp, Pr=100.0, Ra=1e4, n=0.6)
    res = gm.solver.newton(op, np.zeros(op.n), ftol=1e-8, xtol=1e-50, iterations=50, show_trace=True)
    print(N, "iterations", res.iterations, "Nu_bot_avg", repr(gm.solver.nu_bot_avg(sp, res.x)))
    np.save(os.path.join(HERE, "square%d_wave_n06_x.npy" % N), res.x)
    with open(os.path.join(HERE, "square%d_wave_n06_trace.txt" % N), "w") as f:
        for t in res.trace:
            f.write("%d %.17g %.17g\n" % t)
