"""Row-block partition bookkeeping with torch.distributed (gloo, world_size 2 and 3) on CPU: every
dof has exactly one owner, neighbours agree on the interface dof lists (the order the NCCL exchange
packs in), interface traffic matches SURVEY 8e's estimate."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, nx, ny, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    from cases import fe
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = fe.CartesianModel((0, 1, 0, 1), (nx, ny))
    sp = fe.build_rank_spaces(m, cases.TURAN["V"], cases.TURAN["Ttags"], cases.TURAN["Texpr"], rank, world)
    own = fe.owned_dofs(sp)
    # ownership: sizes add up to n and the sets are disjoint
    counts = torch.zeros(sp.n_total, dtype=torch.int32)
    counts[torch.from_numpy(own)] = 1
    dist.all_reduce(counts)
    ok = bool((counts == 1).all())
    # neighbours agree on the interface lists
    lo, hi = fe.interface_dofs(sp, 0), fe.interface_dofs(sp, 1)
    sizes = torch.tensor([lo.size, hi.size], dtype=torch.int64)
    allsz = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(allsz, sizes)
    agree = True
    if rank > 0:
        dist.send(torch.from_numpy(lo.copy()), dst=rank - 1)
    if rank < world - 1:
        buf = torch.zeros(int(allsz[rank + 1][0]), dtype=torch.int64)
        dist.recv(buf, src=rank + 1)
        agree = buf.numpy().size == hi.size and bool(np.array_equal(buf.numpy(), hi))
    res = torch.tensor([int(ok), int(agree)], dtype=torch.int32)
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    if rank == 0:
        q.put((int(res[0]), int(res[1]), [int(s[0]) for s in allsz]))
    dist.destroy_process_group()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world,nx,ny", [(2, 9, 8), (3, 5, 13)])
def test_row_block_ownership_gloo(world, nx, ny):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nx, ny, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    ok, agree, lo_sizes = q.get(timeout=10)
    assert ok == 1, "every dof must have exactly one owner"
    assert agree == 1, "neighbours must enumerate the same interface dofs"
    # one mesh line: (2nx+1) nodes x (ux,uy,T) minus Dirichlet ones (side walls: U and T)
    assert lo_sizes[0] == 0 and all(s == 3 * (2 * nx - 1) for s in lo_sizes[1:])


def _worker_cells(rank, world, port, n, q):
    """Contiguous-cell-range partition of a triangle mesh (gather path, no exchange step): every dof has exactly one owner,
    the ghost cells complete the neighbourhood of every owned dof, and the row oracle on the LOCAL tables reproduces the
    global oracle on the owned dofs."""
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import cases
    from cases import fe
    from oracle.oracle import Oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    m = fe.square_tri_model(n, jitter=0.25, seed=3)
    sp = fe.build_rank_spaces_cells(m, cases.TURAN["V"], cases.TURAN["Ttags"], cases.TURAN["Texpr"], rank, world)
    own = fe.owned_dofs(sp)
    counts = torch.zeros(sp.n_total, dtype=torch.int32)
    counts[torch.from_numpy(own)] = 1
    dist.all_reduce(counts)
    ok = bool((counts == 1).all())
    # owned lists from the local tables == lists of the global problem
    spg = cases.spaces(m, cases.TURAN)
    prm = dict(Pr=100.0, Ra=1e5, n=0.6)
    x = fe.splitmix_iterate(sp.n_total, 11)
    rows = own[:: max(1, own.size // 200)]
    lptr, lidx, lval, lb = Oracle(sp, **prm).rows(rows, x, "csc")
    gptr, gidx, gval, gb = Oracle(spg, **prm).rows(rows, x, "csc")
    same = bool(np.array_equal(lptr, gptr) and np.array_equal(lidx, gidx) and np.allclose(lval, gval, rtol=1e-13, atol=0) and np.allclose(lb, gb, rtol=1e-13, atol=1e-300))
    res = torch.tensor([int(ok), int(same)], dtype=torch.int32)
    dist.all_reduce(res, op=dist.ReduceOp.MIN)
    nown = torch.tensor([own.size], dtype=torch.int64)
    dist.all_reduce(nown)
    if rank == 0:
        q.put((int(res[0]), int(res[1]), int(nown[0]), sp.n_total))
    dist.destroy_process_group()


@pytest.mark.parametrize("world,n", [(2, 7), (3, 10)])
def test_cell_range_ownership_gloo(world, n):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_cells, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    ok, same, nown, ntot = q.get(timeout=10)
    assert ok == 1, "every dof must have exactly one owner"
    assert same == 1, "own + ghost cells must reproduce the global lists of the owned dofs"
    assert nown == ntot
