"""Shared problem set-ups of the reference's configs (BASELINE.json `configs`)."""
import numpy as np

from _pkg import gm

fe = gm.fe

ALLWALLS = [1, 2, 3, 4, 5, 6, 7, 8]
SIMPLE = dict(V=ALLWALLS, Ttags=[5, 7, 8, 1, 2], Texpr=[1.0, 0.0, 0.0, 0.5, 0.5])          # examples/simple.jl:32-36
TURAN = dict(V=ALLWALLS, Ttags=[7, 8, 1, 2, 3, 4], Texpr=[0.0, 1.0, 0.0, 1.0, 0.0, 1.0])   # validation-turan/turan.jl:18-23
WAVE = dict(V=ALLWALLS, Ttags=[5, 7, 8, 1, 2],
            Texpr=[lambda x: np.sin(np.pi * x[0]), 0.0, 0.0, 0.0, 0.0])                   # study1/study1.jl:38-43
KUEHN = dict(V=["all"], Ttags=["inner", "outer"], Texpr=[1.0, 0.0])                       # validation-kuehn/kuehn.jl:47-50


def spaces(model, bc):
    return fe.build_spaces(model, bc["V"], bc["Ttags"], bc["Texpr"])


def cart(nx, ny=None, bc=TURAN, domain=(0, 1, 0, 1)):
    return spaces(fe.CartesianModel(domain, (nx, ny or nx)), bc)


def block_scaled_err(a, ref, blocks, floor=0.0):
    """|a-ref| <= tol*max(|ref|, s_B), s_B = max|ref| over the entry's block.
    Returns max over entries of |a-ref| / max(|ref|, s_B).  `floor` (a fraction of the global
    max|ref|) bounds s_B from below for iterates whose block is pure cancellation noise (e.g. the
    pressure residual of a divergence-free smooth iterate is ~1e-17 while its terms are O(1e-2))."""
    err = 0.0
    gmax = float(np.max(np.abs(ref))) if ref.size else 0.0
    for sel in blocks:
        if not np.any(sel):
            continue
        r = ref[sel]
        s = max(np.max(np.abs(r)), floor * gmax)
        if s == 0.0:
            err = max(err, float(np.max(np.abs(a[sel]))))
            continue
        err = max(err, float(np.max(np.abs(a[sel] - r) / np.maximum(np.abs(r), s))))
    return err


def field_of_dof(sp):
    f = np.zeros(sp.n_total, dtype=np.int8)
    o = sp.offsets
    f[o[1]:o[2]] = 1
    f[o[2]:] = 2
    return f


def nz_blocks(sp, ptr, idx):
    """Boolean masks of the 6 touched (field,field) blocks over the nnz array."""
    f = field_of_dof(sp)
    major = np.repeat(np.arange(sp.n_total), np.diff(ptr))
    fm, fi = f[major], f[idx]
    return [(fm == a) & (fi == b) for a in range(3) for b in range(3)]


def res_blocks(sp):
    f = field_of_dof(sp)
    return [f == a for a in range(3)]


def fan_model(valence=14):
    """A vertex of high valence: `valence` triangles around the centre (lists longer than 96 entries, more incident cells than a
    major's 64-byte record holds inline), one ring of 2*valence triangles around them; the outer boundary is tagged "wall"."""
    k = valence
    ang = 2 * np.pi * np.arange(k) / k
    ring1 = np.stack([np.cos(ang), np.sin(ang)], axis=1) * (1.0 + 0.1 * np.cos(3 * ang))[:, None]
    ring2 = np.stack([np.cos(ang + np.pi / k), np.sin(ang + np.pi / k)], axis=1) * 2.0
    coords = np.concatenate([[[0.05, -0.03]], ring1, ring2])
    tris, bnd = [], {}
    for i in range(k):
        a, b = 1 + i, 1 + (i + 1) % k
        c, cprev = 1 + k + i, 1 + k + (i - 1) % k
        tris.append((0, a, b))       # fan
        tris.append((a, b, c))       # ring1 edge -> outer vertex between them
        tris.append((a, cprev, c))   # outer edge (cprev, c) -> ring1 vertex a
        bnd[(min(cprev, c), max(cprev, c))] = 1
    vent = np.zeros(coords.shape[0], dtype=np.int32)
    vent[1 + k:] = 1
    return fe.TriangleModel(coords, tris, vent, bnd, {"wall": [1], "all": [1]})


FAN = dict(V=["wall"], Ttags=["wall"], Texpr=[lambda x: 0.5 + 0.25 * x[0]])
