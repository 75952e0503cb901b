import sys, os
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np
from _pkg import gm
import cases
from oracle.oracle import Oracle
nx, ny = int(sys.argv[1]), int(sys.argv[2])
sp = cases.cart(nx, ny, bc=cases.SIMPLE, domain=(0,1.5,0,1))
prm = dict(Pr=0.7, Ra=1e3, n=1.0)
x = gm.fe.splitmix_iterate(sp.n_total, 1234)
orc = Oracle(sp, **prm); onz, ob = orc.assemble(x, "csc")
h = gm.capi.Handle(sp, layout="csc", path=gm.capi.GM_PATH_CART); h.set_params(**prm); nnz = h.pattern()
for mode in ("res","both","both","both"):
    nz = np.zeros(nnz); b = np.zeros(sp.n_total)
    h.assemble(x, nz if mode=="both" else None, b)
    d = np.abs(b-ob); bad = np.nonzero(d > 1e-10*np.abs(ob).max())[0]
    print(mode, "nbad", bad.size, "of", b.size)
    # locate nodes
    off = sp.offsets
    for k in bad[:12]:
        f = 0 if k < off[1] else (1 if k < off[2] else 2)
        if f == 0:
            node = np.nonzero(sp.U.node_signed == (k//2)+1)[0][0]; xy = sp.U.node_coords[node]
        elif f == 2:
            node = np.nonzero(sp.T.node_signed == (k-off[2])+1)[0][0]; xy = sp.T.node_coords[node]
        else: xy = ((k-off[1])//3,)
        print("  dof", k, "field", f, "xy", np.array(xy)*np.array([nx/1.5, ny])[:len(xy)] if f!=1 else xy, "got", b[k], "ref", ob[k])
import torch
dx = torch.from_numpy(x).cuda()
dnz = torch.full((nnz,), 555.0, dtype=torch.float64, device="cuda")
for it in range(6):
    db = torch.full((sp.n_total,), 777.0, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    h.assemble_device(dx.data_ptr(), dnz.data_ptr(), db.data_ptr(), 3, sync=True)
    bb = db.cpu().numpy()
    bad = np.nonzero(np.abs(bb-ob) > 1e-10*np.abs(ob).max())[0]
    print("device both: nbad", bad.size, [(int(k), float(bb[k])) for k in bad[:6]])
