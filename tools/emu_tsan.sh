#!/bin/bash
# ThreadSanitizer run of the host emulation of k_gath (tests/emu): every CUDA thread is a pthread, the CTA
# barrier a pthread barrier, so a missing barrier between two phases shows up as a data race.
# Expected output: races ONLY at the nine staged stores of the flush (lanes of Dirichlet entries write the
# same trash element of a slot -- by design) and the parity numbers of the two cases.
set -e
cd "$(dirname "$0")/.."
g++ -O1 -g -std=c++17 -shared -fPIC -pthread -fsanitize=thread -Wno-unknown-pragmas -o /tmp/_emu_gath_tsan.so tests/emu/emu_gath.cpp
cat > /tmp/_emu_tsan_run.py <<'PY'
import sys, ctypes as C
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import test_gath_emu as T
import cases
from cases import fe
lib = C.CDLL("/tmp/_emu_gath_tsan.so"); lib.emu_gath_run.restype = C.c_int
for (nx, ny, bc, pi, layout, ng) in [(5, 9, cases.TURAN, 1, "csc", 2), (259, 1, cases.SIMPLE, 0, "csr", 1)]:
    sp = cases.cart(nx, ny, bc=bc, domain=(0, 1.5, 0, 1))
    x = fe.splitmix_iterate(sp.n_total, 1234)
    orc, ptr, idx, nz, b, _ = T.run_emu(lib, sp, T.PARAMS[pi], layout, x, ng=ng)
    onz, ob = orc.assemble(x, layout)
    print("parity", nx, ny, layout, cases.block_scaled_err(nz, onz, cases.nz_blocks(sp, ptr, idx)), cases.block_scaled_err(b, ob, cases.res_blocks(sp)))
PY
TSAN_OPTIONS="halt_on_error=0 history_size=4" LD_PRELOAD=$(gcc -print-file-name=libtsan.so) python /tmp/_emu_tsan_run.py 2>&1 | grep -E "SUMMARY|parity" | sort | uniq -c
