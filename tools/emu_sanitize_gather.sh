#!/bin/bash
# ThreadSanitizer and AddressSanitizer runs of the host emulation of the generic gather path's kernels (tests/emu/emu_gather.cpp;
# compute-sanitizer is not available on this pool's GPUs).  Every CUDA thread is a pthread, block / warp barriers are pthread
# barriers, shared memory and the workspaces are heap buffers: a missing barrier between two phases of k_cellmat, a missing
# warp synchronisation between two cells of a list in k_lists or a store outside a list's shared-memory range shows up here.
# Expected output: no SUMMARY line, parity numbers ~1e-16 for every case.
set -e
cd "$(dirname "$0")/.."
TOOL=${1:-thread}   # thread | address
g++ -O1 -g -std=c++17 -shared -fPIC -pthread -fsanitize=$TOOL -Wno-unknown-pragmas -o /tmp/_emu_gather_$TOOL.so tests/emu/emu_gather.cpp
cat > /tmp/_emu_gather_san.py <<'PY'
import sys, ctypes as C
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import test_gather_emu as T
import cases
from cases import fe
lib = C.CDLL(sys.argv[1]); lib.emu_gather_run.restype = C.c_int
for name, layout, pi in (("sqtri", "csc", 0), ("fan14", "csr", 2), ("jquad6x5", "csc", 1), ("sqtri1", "csr", 0)):
    sp = T.MODELS[name]()
    x = fe.splitmix_iterate(sp.n_total, 9)
    orc, ptr, idx, nz, b = T.run_emu(lib, sp, T.PARAMS[pi], layout, x)
    onz, ob = orc.assemble(x, layout)
    print("parity", name, layout, cases.block_scaled_err(nz, onz, cases.nz_blocks(sp, ptr, idx)), cases.block_scaled_err(b, ob, cases.res_blocks(sp)))
PY
if [ "$TOOL" = thread ]; then
  TSAN_OPTIONS="halt_on_error=0 history_size=4" LD_PRELOAD=$(gcc -print-file-name=libtsan.so) python /tmp/_emu_gather_san.py /tmp/_emu_gather_$TOOL.so 2>&1 | grep -E "SUMMARY|parity" | sort | uniq -c
else
  ASAN_OPTIONS="detect_leaks=0 halt_on_error=0" LD_PRELOAD=$(gcc -print-file-name=libasan.so) python /tmp/_emu_gather_san.py /tmp/_emu_gather_$TOOL.so 2>&1 | grep -E "SUMMARY|ERROR|parity" | sort | uniq -c
fi
