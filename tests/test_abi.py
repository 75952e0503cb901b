"""The C-ABI library loads and exports every symbol include/gemotion_b200.h declares (no compute)."""
import os
import re

import pytest


def test_exports_match_header(gm):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "gemotion_b200.h")).read()
    declared = set(re.findall(r"\b(gm_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"gm_desc", "gm_times"}
    assert declared == set(gm.capi.EXPORTS), declared ^ set(gm.capi.EXPORTS)
    if not os.path.exists(gm.capi.LIB_PATH):
        gm.capi.build()
    L = gm.capi.lib()
    for name in declared:
        assert hasattr(L, name)
    assert L.gm_version() == 110


def test_desc_struct_size(gm):
    import ctypes as C
    # layout agreed with the C side (gm_create rejects a mismatching struct_size)
    assert C.sizeof(gm.capi.gm_desc) % 8 == 0


def test_no_cpu_fallback(gm):
    """Without a GPU the product path must fail loudly, never fall back to the oracle."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import cases
    sp = cases.cart(3)
    with pytest.raises(gm.capi.GmError):
        gm.operator.B200FEOperator(sp)


def test_ctypes_structs_match_the_header(gm, tmp_path):
    """sizeof / field offsets of gm_desc, gm_post_desc and gm_times as the C compiler sees include/gemotion_b200.h must equal the
    ctypes mirrors in capi.py (gm_create only checks struct_size at run time, on a GPU box)."""
    import ctypes as C
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "sz.c"
    fields = {"gm_desc": ["struct_size", "ncells", "cartesian", "origin", "coords", "layout", "cell_dofs_u", "n_free", "diri_u", "nq", "w", "dgphi",
                          "device", "rank", "cell_begin", "ghost_lo", "kernel", "n_own_cells", "ghost_cell_ids"],
              "gm_times": ["h2d_ms", "exchange_ms", "launches", "d2h_bytes"],
              "gm_post_desc": ["struct_size", "cell_dofs_psi", "diri_pi", "w2", "dphi4", "cell_nodes", "n_nodes"]}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "include/gemotion_b200.h"', 'int main(void) {']
    for st, fl in fields.items():
        lines.append('  printf("%s %%zu\\n", sizeof(%s));' % (st, st))
        for f in fl:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (st, f, st, f))
    lines += ['  return 0;', '}']
    src.write_text("\n".join(lines))
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", root, "-o", str(exe), str(src)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)]).decode().splitlines())
    mirrors = {"gm_desc": gm.capi.gm_desc, "gm_times": gm.capi.gm_times, "gm_post_desc": gm.capi.gm_post_desc}
    for st, fl in fields.items():
        assert int(out[st]) == C.sizeof(mirrors[st]), st
        for f in fl:
            assert int(out["%s.%s" % (st, f)]) == getattr(mirrors[st], f).offset, (st, f)


def test_julia_struct_mirrors_gm_desc(gm):
    """julia/B200FEOperator.jl cannot be executed here (no Julia); its `struct GmDesc` must at least list the fields of gm_desc in
    the same order with types of the same size and alignment (Julia lays an immutable struct of isbits fields out like C)."""
    import ctypes as C
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    txt = open(os.path.join(root, "julia", "B200FEOperator.jl")).read()
    body = txt[txt.index("struct GmDesc") + len("struct GmDesc"):]
    body = body[:body.index("\nend")]
    body = re.sub(r"#.*", "", body)
    jl = re.findall(r"(\w+)::([\w{},]+)", body)
    size = {"Int32": (4, 4), "Int64": (8, 8), "NTuple{2,Float64}": (16, 8), "NTuple{3,Int64}": (24, 8)}
    off = 0
    cfields = gm.capi.gm_desc._fields_
    assert [n for n, _ in jl] == [f[0] for f in cfields]
    for (name, ty), cf in zip(jl, cfields):
        sz, al = (8, 8) if ty.startswith("Ptr{") else size[ty]
        off = (off + al - 1) // al * al
        assert off == getattr(gm.capi.gm_desc, name).offset and sz == C.sizeof(cf[1]), name
        off += sz
    assert (off + 7) // 8 * 8 == C.sizeof(gm.capi.gm_desc)


def test_header_is_plain_c_and_links(gm, tmp_path):
    """include/gemotion_b200.h must be usable from plain C (the boundary a cgo / ccall / ctypes binding sees): a C11 translation
    unit that takes the address of every declared entry point compiles with -Wall -Werror -pedantic and links against the library."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "gemotion_b200.h")).read()
    names = sorted(set(re.findall(r"\b(gm_[a-z0-9_]+)\s*\(", hdr)) - {"gm_desc", "gm_times"})
    src = tmp_path / "use.c"
    src.write_text('#include "include/gemotion_b200.h"\n#include <stdio.h>\nint main(void) {\n  void *p[] = {%s};\n'
                   '  gm_desc d; d.struct_size = (int32_t)sizeof d; (void)d;\n  printf("%%d %%zu\\n", gm_version(), sizeof p / sizeof p[0]);\n  return 0;\n}\n'
                   % ", ".join("(void *)%s" % n for n in names))
    lib_dir = os.path.join(root, "gemotion.jl_b200")
    exe = tmp_path / "use"
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Werror", "-I", root, "-o", str(exe), str(src),
                           "-L", lib_dir, "-l:libgemotion_b200.so", "-Wl,-rpath," + lib_dir])
