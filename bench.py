#!/usr/bin/env python
"""bench.py -- Jacobian+residual assembly throughput (Mcells/s, fp64) on B200.

Contract: `python bench.py --gpus N --steps K --warmup W [--impl reference]` prints ONE JSON line.

A "step" is one residual_and_jacobian! assembly (Solver.jl:157/169 hot path) of the whole mesh
at a fixed synthetic iterate.  Workload (BASELINE.json config 5, the mesh the metric is quoted on): 4096x4096 Cartesian quads, turan BC set,
Pr=100, Ra=1e5, n=1.4, on N = 1, 2, 4, 8 GPUs (row blocks for N>1); config 4 = --mesh 2048 --power-index 0.6.  `value` is measured with
inputs resident in HBM (CUDA events on the launching stream); `e2e` goes through gm_assemble with
pinned HOST buffers (iterate H2D + nzval/residual D2H inside the timed region).
`--impl reference`: the CPU oracle (port of the Gridap path; Julia/Gridap cannot run in this image)
on all host cores on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

METRIC = "Jacobian+residual assembly Mcells/s (fp64)"
TURAN = dict(V=[1, 2, 3, 4, 5, 6, 7, 8], Ttags=[7, 8, 1, 2, 3, 4], Texpr=[0.0, 1.0, 0.0, 1.0, 0.0, 1.0])


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.stop_flag = False
        self.proc = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([c.strip() for c in line.split(",")])
                if self.stop_flag:
                    break
        except Exception:
            pass

    def finish(self):
        self.stop_flag = True
        if self.proc:
            try:
                self.proc.terminate()
            except Exception:
                pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for k, nme in enumerate(names):
                    if r[3 + k].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


def build_problem(N, bc=TURAN):
    from _pkg import gm
    model = gm.fe.CartesianModel((0, 1, 0, 1), (N, N))
    return gm.fe.build_spaces(model, bc["V"], bc["Ttags"], bc["Texpr"])


def algorithmic_bytes(sp, nnz):
    """8*nnz + 8*n (read x) + 8*n (write b) + 4*ld*ncells (cell-dof ids); geometry implicit on Cartesian
    meshes, + 16*nverts + 4*nv*ncells (coordinates, cell->vertex) otherwise."""
    m = sp.model
    ld = 3 * sp.ref.nn + 3
    geo = 0 if m.cartesian else 16 * m.nverts + 4 * sp.ref.nv * m.ncells
    return 8 * nnz + 16 * sp.n_total + 4 * ld * m.ncells + geo


def cpu_pass(N, prm, passes, threads=None):
    """Times the CPU oracle (coloured OpenMP scatter over all host cores) on an N x N sample."""
    from oracle.oracle import Oracle
    from _pkg import gm
    if threads:
        os.environ["OMP_NUM_THREADS"] = str(threads)
        Oracle.set_num_threads(threads)  # (the library may already be loaded with torchrun's OMP_NUM_THREADS=1)
    sp = build_problem(N)
    orc = Oracle(sp, **prm)
    x = gm.fe.splitmix_iterate(sp.n_total, 1234)
    orc.pattern("csc")
    nx = N
    cx, cy = np.tile(np.arange(nx), nx), np.repeat(np.arange(nx), nx)
    col = (cx & 1) + 2 * (cy & 1)
    order = np.argsort(col, kind="stable").astype(np.int64)
    cptr = np.concatenate([[0], np.cumsum(np.bincount(col, minlength=4))]).astype(np.int64)
    ts = []
    for _ in range(passes):
        t0 = time.perf_counter()
        orc.assemble(x, "csc", colors=(cptr, order))
        ts.append(time.perf_counter() - t0)
    return sp.model.ncells, ts, Oracle.num_threads()


def run_reference(args, prm):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    N = args.cpu_mesh
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)  # torchrun pins it to 1; this arm uses every host core
    ncells, ts, thr = cpu_pass(N, prm, args.warmup + args.steps)
    ts = ts[args.warmup:]
    ms = 1e3 * float(np.mean(ts))
    val = ncells / (ms * 1e-3) / 1e6
    sample = "%dx%d sub-mesh of the workload (same BCs, params, iterate rule), %d passes" % (N, N, args.steps)
    out = {"impl": "reference", "metric": METRIC, "value": val, "unit": "Mcells/s", "n_gpus": args.gpus,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
           "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
           "config": {"workload": workload_name(args), "sample": sample},
           "cpu_baseline": {"value": val, "unit": "Mcells/s", "cores": thr, "kind": "port", "sample": sample},
           "e2e": {"value": val, "unit": "Mcells/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    _emit(out)


def workload_name(args):
    if args.mesh_type == "tri":
        return ("unit square split into 2x%dx%d jittered triangles (fixed diagonal), turan BCs, Pr=100 Ra=1e5 n=%g, iterate splitmix64(seed=1234)"
                % (args.mesh, args.mesh, args.n))
    return "cartesian %dx%d quads, turan BCs, Pr=100 Ra=1e5 n=%g, iterate splitmix64(seed=1234)" % (args.mesh, args.mesh, args.n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--mesh", type=int, default=0, help="cells per side (default 2048 on 1 GPU)")
    ap.add_argument("--power-index", dest="n", type=float, default=0.0, help="power-law index n (default 0.6; 1.4 for N>1)")
    ap.add_argument("--mesh-type", default="quad", choices=["quad", "tri"], help="tri: generic path (gather; --path scatter for the first generation) on a triangle mesh of 2*mesh^2 cells, 1 GPU")
    ap.add_argument("--layout", default="csc", choices=["csc", "csr"])
    ap.add_argument("--path", default="auto", choices=["auto", "scatter", "cart", "gather"])
    ap.add_argument("--kernel", default="auto", choices=["auto", "mma", "gath", "gath1", "sweep"], help="structured-path kernel generation (A/B)")
    ap.add_argument("--xchunk", type=int, default=0, help="quad columns per x-chunk (0: chosen by the library)")
    ap.add_argument("--cpu-mesh", type=int, default=512)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--force-e2e", action="store_true")
    ap.add_argument("--no-verify", action="store_true", help="skip the sampled row-oracle check (outside the timed region)")
    ap.add_argument("--verify-rows", type=int, default=2000, help="random major dofs checked besides the structural ones")
    args = ap.parse_args()
    if args.mesh == 0 and args.mesh_type == "tri":
        args.mesh = 1024  # 2.1 M triangles
    if args.mesh == 0:
        # config 5 of BASELINE.json: 4096^2, n=1.4 -- the mesh the metric / target is quoted on; it fits one B200
        # (147 GB), so N = 1, 2, 4, 8 is one strong-scaling series.  Config 4: --mesh 2048 --power-index 0.6.
        args.mesh = 4096
    if args.n == 0.0:
        args.n = 1.4
    prm = dict(Pr=100.0, Ra=1e5, n=args.n)
    if args.impl == "reference":
        return run_reference(args, prm)

    import torch
    from _pkg import gm
    capi = gm.capi
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the assembly path has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout = the one JSON line
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    # ---- problem (host tables; in production Gridap supplies them) ----
    # N=1: the whole mesh on one GPU.  N>1: owner-computes row blocks, one block per rank,
    # interface rows summed with NCCL send/recv inside gm_assemble_device.
    t0 = time.perf_counter()
    if args.mesh_type == "tri":
        model = gm.fe.square_tri_model(args.mesh, jitter=0.2, seed=3)
        args.no_cpu = True   # (the CPU baseline leg is defined on the quad workload)
    else:
        model = gm.fe.CartesianModel((0, 1, 0, 1), (args.mesh, args.mesh))
    if world > 1 and args.mesh_type == "tri":
        # contiguous cell ranges + ghost cells (gather path): no exchange step on the data path
        sp = gm.fe.build_rank_spaces_cells(model, TURAN["V"], TURAN["Ttags"], TURAN["Texpr"], rank, world)
    elif world > 1:
        sp = gm.fe.build_rank_spaces(model, TURAN["V"], TURAN["Ttags"], TURAN["Texpr"], rank, world)
    else:
        sp = gm.fe.build_spaces(model, TURAN["V"], TURAN["Ttags"], TURAN["Texpr"])
    x_host = gm.fe.splitmix_iterate(sp.n_total, 1234)
    t_tables = time.perf_counter() - t0
    pathid = {"auto": capi.GM_PATH_AUTO, "scatter": capi.GM_PATH_SCATTER, "cart": capi.GM_PATH_CART, "gather": capi.GM_PATH_GATHER}[args.path]
    kid = {"auto": capi.GM_KERNEL_AUTO, "mma": capi.GM_KERNEL_MMA, "gath": capi.GM_KERNEL_GATH, "gath1": capi.GM_KERNEL_GATH1,
           "sweep": capi.GM_KERNEL_SWEEP}[args.kernel]
    h = capi.Handle(sp, layout=args.layout, device=local, path=pathid, rank=rank, nranks=world, kernel=kid, xchunk=args.xchunk)
    h.set_params(**prm)
    t0 = time.perf_counter()
    nnz = h.pattern()
    t_pattern = time.perf_counter() - t0
    info = h.info()
    nnz_tri_global = None
    if world > 1 and args.mesh_type == "tri":
        # global nnz = lists of the dofs each rank owns (its local pattern also holds the incomplete lists of the ghost cells' other dofs)
        own_rows = gm.fe.owned_dofs(sp)
        lens = np.diff(h.get_pattern(with_idx=False)[0])
        t = torch.tensor([int(lens[own_rows].sum())], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        nnz_tri_global = int(t.item())
    if world > 1 and args.mesh_type != "tri":
        idt = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            idt = torch.frombuffer(bytearray(capi.nccl_unique_id()), dtype=torch.uint8).cuda()
        dist.broadcast(idt, src=0)
        h.comm_init(bytes(idt.cpu().numpy().tobytes()))
    stream = torch.cuda.current_stream()
    h.set_stream(stream.cuda_stream)
    d_x = torch.from_numpy(x_host).cuda()
    d_nz = torch.empty(nnz, dtype=torch.float64, device="cuda")
    d_b = torch.empty(sp.n_total, dtype=torch.float64, device="cuda")
    flags = capi.GM_JACOBIAN | capi.GM_RESIDUAL

    def step():
        h.assemble_device(d_x.data_ptr(), d_nz.data_ptr(), d_b.data_ptr(), flags, sync=False)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    launches_per_step = h.timing()["launches"]
    if world > 1:
        dist.barrier()
    sampler = ClockSampler(local)
    sampler.start()
    time.sleep(0.3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    torch.cuda.synchronize()
    ev[0].record(stream)
    for k in range(args.steps):
        step()
        ev[k + 1].record(stream)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.finish()
    per = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    total_ms = ev[0].elapsed_time(ev[args.steps])
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms = total_ms / args.steps
    ncells_total = sp.model.ncells  # the global mesh (row blocks partition it)
    value = ncells_total / (ms * 1e-3) / 1e6
    # checksum of the result (also the D2H read of the step's result for the record)
    chk = float(d_b.abs().sum().item())
    # a rank's buffers are only defined on the dofs it owns: the end-to-end result is compared with the device-resident one on owned
    # dofs (all of them on small meshes; at the BASELINE sizes the verify sample -- enumerating 10^8 owned ids per rank is not free)
    own_idx = gm.fe.owned_dofs(sp) if (world > 1 and (args.mesh <= 1024 or args.mesh_type == "tri")) else None
    chk_owned = None
    if args.mesh <= 1024:  # sum over OWNED residual entries: must not depend on the number of GPUs
        own = torch.from_numpy(gm.fe.owned_dofs(sp)).cuda()
        t = d_b[own].abs().sum().reshape(1)
        if world > 1:
            dist.all_reduce(t)
        chk_owned = float(t.item())

    # ---- one synchronous step: the NCCL interface-row sum is timed by its own events (gm_times.exchange_ms)
    h.assemble_device(d_x.data_ptr(), d_nz.data_ptr(), d_b.data_ptr(), flags, sync=True)
    timing_sync = h.timing()

    # ---- sampled row oracle (outside the timed region): the kernel that was just timed, at this size, through real NCCL
    verify = None
    verify_rows_idx = None
    if not args.no_verify:
        import rowcheck
        from oracle.oracle import Oracle
        t0 = time.perf_counter()
        rows = rowcheck.sample_rows(sp, nrandom=args.verify_rows)
        verify_rows_idx = rows
        orc = Oracle(sp, **prm)
        verify = rowcheck.check_rows(sp, h, orc, x_host, d_nz.data_ptr(), d_b.data_ptr(), args.layout, rows)
        del orc
        if world > 1:
            bad = torch.tensor([0 if verify["pattern_ok"] else 1, verify["rows"], verify["entries"]], dtype=torch.int64, device="cuda")
            dist.all_reduce(bad)
            errs = torch.tensor([verify["max_err_nz"] if verify["max_err_nz"] is not None else 1e300, verify["max_err_b"]],
                                dtype=torch.float64, device="cuda")
            dist.all_reduce(errs, op=dist.ReduceOp.MAX)
            verify = {"rows": int(bad[1].item()), "entries": int(bad[2].item()), "pattern_ok": int(bad[0].item()) == 0,
                      "max_err_nz": float(errs[0].item()), "max_err_b": float(errs[1].item())}
        verify["tol"] = 1e-12
        verify["ok"] = bool(verify["pattern_ok"] and verify["max_err_nz"] is not None and verify["max_err_nz"] <= 1e-12 and verify["max_err_b"] <= 1e-12)
        verify["what"] = ("complete lists (minor ids bit-equal, values block-scaled) and residual entries of the sampled owned major dofs "
                          "vs the oracle's per-cell restatement; sample = random + chunk/strip/interface boundaries + field tails")
        verify["seconds"] = time.perf_counter() - t0

    # ---- Jacobian-only and residual-only assemblies (a line search calls residual! 1..k times per iteration)
    def timed(fl, reps=5):
        for _ in range(2):
            h.assemble_device(d_x.data_ptr(), d_nz.data_ptr(), d_b.data_ptr(), fl, sync=False)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(reps):
            h.assemble_device(d_x.data_ptr(), d_nz.data_ptr(), d_b.data_ptr(), fl, sync=False)
        e1.record(stream)
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    ms_jac = timed(capi.GM_JACOBIAN)
    ms_res = timed(capi.GM_RESIDUAL)
    parts = {"jacobian_only": {"ms": ms_jac, "Mcells_per_s": sp.model.ncells / (ms_jac * 1e-3) / 1e6},
             "residual_only": {"ms": ms_res, "Mcells_per_s": sp.model.ncells / (ms_res * 1e-3) / 1e6}}

    # ---- e2e through the C ABI with host buffers ----
    e2e = None
    e2e_note = None
    if not args.no_e2e and not args.force_e2e:
        # host buffers of this rank's outputs get page-locked in place; never risk the box: 1.5x headroom in this rank's RAM share
        need = 8 * nnz + 16 * sp.n_total
        try:
            import psutil
            avail = psutil.virtual_memory().available / max(world, 1)
        except Exception:
            avail = 64e9
        if need * 1.5 > avail:
            e2e_note = "skipped: %.0f GB of host memory per rank needed for the outputs, %.0f GB available per rank (use --force-e2e)" % (need / 1e9, avail / 1e9)
            args.no_e2e = True
    if world > 1:  # every rank must take the same branch (gm_assemble is collective)
        flag = torch.tensor([1 if args.no_e2e else 0], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
        if int(flag.item()) and not args.no_e2e:
            args.no_e2e = True
            e2e_note = "skipped: another rank lacks host memory for pinned buffers"

    if not args.no_e2e:
        # the caller's own arrays (Julia: A.nzval, b, x), page-locked once with gm_host_register as INTEGRATION.md says
        xn, nzn, bn = x_host, np.empty(nnz), np.zeros(sp.n_total)
        pinned = []
        try:
            for arr in (xn, nzn, bn):
                capi.host_register(arr)
                pinned.append(arr)
        except capi.GmError as exc:
            e2e_note = "host buffers not page-locked (%s): copies run through the driver's staging" % exc
        del d_nz  # the host path uses the library's own staging buffer; free the HBM-resident output first
        d_nz = None
        torch.cuda.empty_cache()
        h.assemble(xn, nzn, bn)  # warm-up (allocates the device staging buffer)
        torch.cuda.synchronize()
        ts = []
        for _ in range(args.e2e_steps):
            t0 = time.perf_counter()
            h.assemble(xn, nzn, bn)
            ts.append(time.perf_counter() - t0)
        tt = h.timing()
        e_ms = 1e3 * float(np.mean(ts))
        if world > 1:
            t = torch.tensor([e_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e_ms = float(t.item())
        e2e = {"value": ncells_total / (e_ms * 1e-3) / 1e6, "unit": "Mcells/s", "ms_per_step": e_ms,
               "h2d_bytes_per_step": int(tt["h2d_bytes"]), "d2h_bytes_per_step": int(tt["d2h_bytes"]),
               "h2d_ms": tt["h2d_ms"], "kernel_ms": tt["kernel_ms"], "d2h_ms": tt["d2h_ms"], "steps": args.e2e_steps,
               "checksum_match": None}
        cmp_idx = own_idx if own_idx is not None else (verify_rows_idx if world > 1 else None)
        if world == 1 or cmp_idx is not None:
            dev_b = d_b.cpu().numpy() if cmp_idx is None else d_b[torch.from_numpy(np.ascontiguousarray(cmp_idx)).cuda()].cpu().numpy()
            host_b = bn if cmp_idx is None else bn[cmp_idx]
            ref_sum = float(np.abs(dev_b).sum())
            e2e["checksum_match"] = bool(abs(float(np.abs(host_b).sum()) - ref_sum) <= 1e-9 * max(1.0, ref_sum))
        # residual-only end to end (iterate H2D + residual D2H): the call a line search repeats
        ts = []
        for _ in range(args.e2e_steps + 1):
            t0 = time.perf_counter()
            h.assemble(xn, None, bn)
            ts.append(time.perf_counter() - t0)
        r_ms = 1e3 * float(np.mean(ts[1:]))
        if world > 1:
            t = torch.tensor([r_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            r_ms = float(t.item())
        tr = h.timing()
        e2e["residual_only"] = {"value": ncells_total / (r_ms * 1e-3) / 1e6, "unit": "Mcells/s", "ms_per_step": r_ms,
                                "h2d_bytes_per_step": int(tr["h2d_bytes"]), "d2h_bytes_per_step": int(tr["d2h_bytes"])}
        for arr in pinned:
            capi.host_unregister(arr)
        del nzn

    timing_last = timing_sync
    # symmetric teardown on every rank BEFORE rank 0 goes on alone (CPU baseline): the library's NCCL
    # communicator and torch's process group are both destroyed collectively
    del d_nz, d_b, d_x
    h.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
        if rank != 0:
            return
    pk, pk_kind = peaks()
    N_ = args.mesh
    nnz_global = 684 * N_ * N_ - 1232 * N_ + 527 if world > 1 else nnz  # turan BC closed form, tested at small N
    if nnz_tri_global is not None:
        nnz_global = nnz_tri_global
    balg = algorithmic_bytes(sp, nnz_global) / world  # per-GPU share
    achieved = balg / (ms * 1e-3) / 1e9
    roof = {"bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
            "traffic": None, "peak_kind": pk_kind + " (MEASURED_PEAKS.json hbm_gbs, burst copy)",
            "algorithmic_bytes_per_launch": balg, "bytes_per_cell": balg * world / sp.model.ncells,
            "kernel": "whole device step (all numeric kernels of one assembly)"}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            tr_ = json.load(open(prof)).get(("tri%d" % args.mesh) if args.mesh_type == "tri" else str(args.mesh))
            if (args.mesh_type == "tri") != (info["active_path"] == 3) or (args.mesh_type != "tri" and info["active_path"] != 2):
                tr_ = None  # (the committed captures are of the default path of each mesh type)
            # DRAM bytes of the dominant kernel from the committed `ncu --set full` capture of the 1-GPU run of this mesh
            # (profile-time figure, not re-measured in this run); a rank of an N-GPU run moves 1/N of it
            roof["traffic"] = tr_ / world if tr_ else None
            roof["traffic_note"] = "dram__bytes_read+write of the 1-GPU ncu capture in profiles/, divided by n_gpus"
        except Exception:
            pass
    cpu = None
    if not args.no_cpu:
        N = args.cpu_mesh
        ncs, ts, thr = cpu_pass(N, prm, 3, threads=os.cpu_count() or 1)
        cv = ncs / float(np.mean(ts[1:])) / 1e6
        cpu = {"value": cv, "unit": "Mcells/s", "cores": thr, "kind": "port",
               "sample": "%dx%d sub-mesh of the workload, 2 timed passes of the CPU oracle (OpenMP coloured scatter)" % (N, N)}
    out = {"metric": METRIC, "value": value, "unit": "Mcells/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
           "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": workload_name(args), "ncells": sp.model.ncells, "n_free": sp.n_total, "nnz": nnz_global, "nnz_local": nnz,
                      "layout": args.layout, "path": {1: "scatter", 2: "cart", 3: "gather"}[info["active_path"]], "kernel": args.kernel,
                      "l2": "outputs (%.1f GB) exceed L2 by >100x; no flush needed" % (8 * nnz / 1e9),
                      "parallelism": (("%d contiguous cell ranges + ghost cells, no exchange step" % world) if args.mesh_type == "tri" else
                                      ("%d owner-computes row blocks, NCCL send/recv interface-row sum" % world)) if world > 1 else "single",
                      "exchange_ms": timing_last.get("exchange_ms"),
                      "pattern_s": t_pattern, "tables_s": t_tables, "device_bytes": info["device_bytes"]},
           "roofline": roof, "verify": verify, "parts": parts, "cpu_baseline": cpu, "e2e": e2e, "e2e_note": e2e_note, "gpu_launches": int(launches_per_step * args.steps),
           "clocks": clocks, "ms_per_step_each": per, "checksum_abs_b": chk, "checksum_owned_abs_b": chk_owned}
    _emit(out)


def _emit(out):
    """The ONE JSON line of the contract, on the process's original stdout."""
    line = (json.dumps(out) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, line)
    else:
        sys.stdout.write(line.decode())
        sys.stdout.flush()


_REAL_STDOUT = None

if __name__ == "__main__":
    # Libraries print to file descriptor 1 behind Python's back (NCCL: "NCCL version ..." at the first communicator); the contract is
    # ONE JSON line on stdout, so everything else written to fd 1 goes to stderr and the line is written to the saved descriptor.
    try:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)
    except OSError:
        _REAL_STDOUT = None
    main()
