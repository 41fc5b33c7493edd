#!/usr/bin/env python
"""Benchmark of the THCM Newton-step hot path (BASELINE.json metric: "Newton-step s (RHS+Jacobian+GMRES) & SpMV HBM
GB/s at 1/2/4/8 B200").

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K --warmup W  # restated reference CPU path (oracle/)

A "step" = one Newton step on the synthetic 1-degree global ocean grid 360x152x24 (BASELINE configs[3]; Mixing = 0):
F(x) (residual) + J(x) (Jacobian into the static maximal-graph CSR) + 6x6 block-diagonal preconditioner + right-
preconditioned FGMRES with a FIXED number of iterations (--gmres-iters, default 50, one cycle) -- with only identity /
block-diagonal preconditioning the singular THCM Jacobian does not reach 1e-4 in any fixed budget (SURVEY.md section 7,
hard part 9), so the work per step is pinned instead of the tolerance.  N > 1 = strong scaling: the same grid
block-partitioned in lon x lat like the reference's Decomp2D, one rank per GPU (torchrun), NCCL halo exchange + dot
all-reduces.  One JSON line on stdout (rank 0).
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

GRID = (360, 152, 24)
PARS = {"COMB": 1.0, "WIND": 1.0, "TEMP": 10.0, "SALT": 1.0}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--grid", type=int, nargs=3, default=list(GRID))
    ap.add_argument("--weak", action="store_true",
                    help="weak scaling (BASELINE configs[4]): the grid grows with the GPU count -- 1: 360x152x24, 2: 510x214x24, "
                         "4: 720x304x24 (1.31 M cells per GPU each), 8: 720x304x32 (the 0.5-degree grid, 0.88 M cells per GPU)")
    ap.add_argument("--gmres-iters", type=int, default=50)
    ap.add_argument("--precon", type=int, default=1)
    ap.add_argument("--ortho", default="dgks", choices=["mgs", "dgks"],
                    help="GMRES orthogonalisation: mgs = src/gmressolver template, dgks = batched Gram-Schmidt as Belos uses in Ocean::solve")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-mixing", action="store_true", help="skip the Mixing = 1 timing (second instance, same grid and state)")
    ap.add_argument("--no-b1", action="store_true", help="skip the timing of the B1 Fortran-symbol boundary (rhs_ + matrix_ with host buffers)")
    return ap.parse_args()


def workload_name(n, m, l, iters):
    return (f"THCM ocean-only synthetic global {n}x{m}x{l} (1 deg = BASELINE configs[3]), Mixing=0: Newton step = residual + "
            f"Jacobian(graph CSR) + 6x6 block-diag precon + FGMRES({iters}) fixed {iters} iterations")


# ---------------------------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi DURING the timed region (B200_PROFILING.md)
# ---------------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device=0):
        self.rows, self.proc, self.device = [], None, device

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.device}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---------------------------------------------------------------------------------------------------------------
# restated reference CPU path (oracle/): one Newton step on ONE block of the 8-rank Decomp2D partition, with the
# reference's own ghost layers, per worker process -- the way one reference MPI rank works (THCM.C:957-1199)
# ---------------------------------------------------------------------------------------------------------------
def _ref_block(n, m, l, nblocks, b):
    """Block b of the reference's decomposition incl. its 2 ghost layers (TRIOS_Domain.C:201-315)."""
    t1, npM, npN, r_min = nblocks, nblocks, 1, 100
    while t1 > 0:
        t2 = nblocks // t1
        r = abs(m // t1 - n // t2)
        if t1 * t2 == nblocks and r <= r_min:
            r_min, npM, npN = r, t1, t2
        t1 -= 1
    pidN = b % npN
    pidM = (b - pidN) // npN
    n0, m0 = n // npN + (1 if pidN < n % npN else 0), m // npM + (1 if pidM < m % npM else 0)
    i0, j0 = pidN * (n // npN) + min(n % npN, pidN), pidM * (m // npM) + min(m % npM, pidM)
    return dict(i0=i0, j0=j0, n0=n0, m0=m0, npN=npN, npM=npM)


def _block_inverses(rp, col, val, ncell):
    """6x6 block-diagonal preconditioner of the restated path: the in-cell blocks of the CSR Jacobian, inverted with LAPACK (numpy),
    identity for singular blocks -- the same preconditioner the GPU arm builds (blockdiag_build_kernel)."""
    import numpy as np
    rows = np.repeat(np.arange(len(rp) - 1), np.diff(rp))
    m = (rows // 6) == (col // 6)
    D = np.zeros((ncell, 6, 6))
    np.add.at(D, (rows[m] // 6, rows[m] % 6, col[m] % 6), val[m])
    try:
        return np.linalg.inv(D)
    except np.linalg.LinAlgError:
        out = np.empty_like(D)
        for q in range(ncell):
            try:
                out[q] = np.linalg.inv(D[q])
            except np.linalg.LinAlgError:
                out[q] = np.eye(6)
        return out


_REF_STATE = {}


def _ref_setup(args):
    """Per-process model of one Decomp2D block incl. the reference's 2 ghost layers (what one reference MPI rank holds)."""
    n, m, l, nblocks, b = args
    key = (n, m, l, nblocks, b)
    if key in _REF_STATE:
        return _REF_STATE[key]
    import numpy as np
    import cases
    from cases import PAR_INDEX as P
    from oracle.oracle import OracleTHCM
    s_glob, landm = cases.global_synth(n, m, l)
    blk = _ref_block(n, m, l, nblocks, b)
    g = 2  # numGhosts (TRIOS_Domain.H:365)
    i0, j0, n0, m0 = blk["i0"], blk["j0"], blk["n0"], blk["m0"]
    xper = blk["npN"] > 1
    ia, ib = (i0 - g, i0 + n0 + g) if xper else (i0, i0 + n0)
    ja, jb = max(j0 - g, 0), min(j0 + m0 + g, m)
    nl, ml = ib - ia, jb - ja
    ii = np.arange(ia - 1, ib + 1) % n if xper else np.clip(np.arange(ia - 1, ib + 1), -1, n)
    # local mask incl. the dummy frame (init_ turns the frame into LAND itself, usrc.F90:100-107)
    lm = landm[:, ja:jb + 2, :][:, :, (ii + 1) % (n + 2)].copy()
    dx, dy = (s_glob.xmax - s_glob.xmin) / n, (s_glob.ymax - s_glob.ymin) / m
    s = cases.Settings.from_degrees(nl, ml, l, 0, 1, 0, 1, periodic=(not xper), hdim=5000.0, qz=2.25)
    s.xmin, s.xmax = s_glob.xmin + ia * dx, s_glob.xmin + ib * dx      # Grid::SubGrid, TRIOS_Domain.C:77-88
    s.ymin, s.ymax = s_glob.ymin + ja * dy, s_glob.ymin + jb * dy
    o = OracleTHCM(s, lm)
    for k, v in PARS.items():
        o.setpar(P[k], v)
    x = cases.consistent_state(s, lm, scale=0.05)
    st = dict(o=o, x=x, graph=o.graph(), cells=nl * ml * l, owned=n0 * m0 * l)
    _REF_STATE[key] = st
    return st


def _ref_worker(args):
    """One Newton step of the restated reference path on one block: residual, Jacobian (dense Al/An -> CRS -> graph), 6x6 block-diagonal
    preconditioner, FGMRES(iters) through the reference's own GMRESSolver.H.  Returns (block, cells, seconds by stage)."""
    n, m, l, nblocks, b, iters = args
    import numpy as np
    from oracle.oracle import kref_gmres
    st = _ref_setup((n, m, l, nblocks, b))
    o, x, (rp, col) = st["o"], st["x"], st["graph"]
    t0 = time.perf_counter(); B = o.rhs(x)
    t1 = time.perf_counter(); val, _miss = o.jacobian_graph(x, (rp, col))
    t2 = time.perf_counter(); minv = _block_inverses(rp, col, val, o.ndim // 6)
    kref_gmres(rp, col, val, B, np.zeros(o.ndim), tol=0.0, maxit=iters - 1, restart=iters, prec_kind=1, minv=minv)
    t3 = time.perf_counter()
    return dict(block=b, cells=st["cells"], stages=(t1 - t0, t2 - t1, t3 - t2))


def _ref_process(conn, n, m, l, nblocks, blocks, iters):
    """One worker = one host core: owns a fixed subset of the blocks (their models stay resident between steps, like the ranks of an MPI
    run) and runs them back to back on every "step" command."""
    try:
        while True:
            cmd = conn.recv()
            if cmd != "step":
                break
            out = [_ref_worker((n, m, l, nblocks, b, iters)) for b in blocks]
            conn.send(out)
    except EOFError:
        pass


class ReferenceRunner:
    """The restated reference CPU path on the box's host cores.  The grid is split into 32 blocks with the reference's Decomp2D rule
    (each with its 2 ghost layers: what `mpirun -np 32` of the reference holds); one process per usable core owns a fixed share of the
    blocks and works through ALL of them every step -- nothing is extrapolated: a step's time is the wall clock of the whole pass."""

    def __init__(self, n, m, l, iters, max_workers=None, nblocks=32, blocks=None):
        import multiprocessing as mp
        import psutil
        self.n, self.m, self.l, self.iters, self.nblocks = n, m, l, iters, nblocks
        self.blocks = list(range(nblocks)) if blocks is None else list(blocks)
        cores = len(os.sched_getaffinity(0))
        mem_gb = psutil.virtual_memory().available / 2**30
        per_block_gb = 6.5 * (8 / nblocks) * (n * m * l) / (360 * 152 * 24) + 0.3
        fit = max(1, int(mem_gb // (per_block_gb * max(1, len(self.blocks)) / max(1, min(cores, len(self.blocks))))))
        self.workers = max(1, min(len(self.blocks), cores, fit, max_workers or len(self.blocks)))
        ctx = mp.get_context("spawn")
        self.procs = []
        for w in range(self.workers):
            parent, child = ctx.Pipe()
            pr = ctx.Process(target=_ref_process, args=(child, n, m, l, nblocks, self.blocks[w::self.workers], iters), daemon=True)
            pr.start()
            self.procs.append((pr, parent))
        self.cells = None

    def step(self):
        t0 = time.perf_counter()
        for _, conn in self.procs:
            conn.send("step")
        res = [r for _, conn in self.procs for r in conn.recv()]
        wall = time.perf_counter() - t0
        self.cells = res[0]["cells"]
        stages = [sum(r["stages"][q] for r in res) / self.workers for q in range(3)]   # core-seconds / cores
        return wall, stages

    def close(self):
        for pr, conn in self.procs:
            try:
                conn.send("stop")
            except Exception:
                pass
        for pr, _ in self.procs:
            pr.join(timeout=10)
            if pr.is_alive():
                pr.terminate()

    def sample(self):
        whole = len(self.blocks) == self.nblocks
        return (f"{'all' if whole else len(self.blocks)} of the {self.nblocks} Decomp2D blocks of the {self.n}x{self.m}x{self.l} grid (each ~{self.cells} cells incl. "
                f"2 ghost layers) on {self.workers} processes, every step: residual + Jacobian via the dense Al/An restatement (g++ -O3), 6x6 "
                f"block-diagonal preconditioner, FGMRES({self.iters}) through the reference's own GMRESSolver.H (modified Gram-Schmidt); "
                + ("time = wall clock of the whole pass, nothing extrapolated" if whole else
                   f"time = wall clock of this sample x {self.nblocks / len(self.blocks):g} (stated, bounded sample)"))


def reference_run(n, m, l, iters, steps, warmup, max_workers=None, blocks=None):
    r = ReferenceRunner(n, m, l, iters, max_workers=max_workers, blocks=blocks)
    try:
        for _ in range(warmup):
            r.step()
        walls, stages = [], [0.0, 0.0, 0.0]
        for _ in range(steps):
            w, st = r.step()
            walls.append(w)
            stages = [a + b for a, b in zip(stages, st)]
        scale = r.nblocks / len(r.blocks)
        value = sum(walls) / len(walls) * scale
        return dict(value=value, cores=r.workers, sample=r.sample(), scale=scale,
                    stages_s=dict(rhs=stages[0] / steps * scale, jacobian=stages[1] / steps * scale, gmres_incl_precon_build=stages[2] / steps * scale))
    finally:
        r.close()


# ---------------------------------------------------------------------------------------------------------------
def main():
    a = parse()
    if a.weak:
        a.grid = list({1: (360, 152, 24), 2: (510, 214, 24), 4: (720, 304, 24), 8: (720, 304, 32)}.get(a.gpus, tuple(a.grid)))
    scaling = "weak" if a.weak else "strong"
    n, m, l = a.grid
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    iters = a.gmres_iters
    # `config` names the WORKLOAD and is identical in both arms; how each arm runs it goes to `impl_config`
    config = {"workload": workload_name(n, m, l, iters), "grid": [n, m, l], "mixing": 0, "gmres_iters": iters, "gmres_restart": iters,
              "precon": "6x6 block-diagonal", "state": "0.05 * N(0,1), seed 20261017, zero on LAND / Dirichlet unknowns",
              "parameters": PARS, "gpus": a.gpus}

    if a.impl == "reference":
        if rank != 0:
            return 0
        r = reference_run(n, m, l, iters, a.steps, a.warmup)
        line = {"impl": "reference", "metric": "newton_step_seconds", "value": r["value"], "unit": "s", "n_gpus": a.gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": r["value"] * 1e3, "higher_is_better": False, "scaling": scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "impl_config": {"what": "restated reference CPU path (oracle/, g++ -O3 -ffp-contract=off) + the reference's own GMRESSolver.H",
                                "orthogonalisation": "modified Gram-Schmidt (GMRESSolver.H:177-181)", "parallelism": r["sample"]},
                "cpu_baseline": {"value": r["value"], "unit": "s", "cores": r["cores"], "kind": "port", "sample": r["sample"], "stages_s": r["stages_s"]},
                "e2e": {"value": r["value"], "unit": "s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
        print(json.dumps(line))
        return 0

    # libraries (NCCL, torch) may print to stdout: keep the real stdout for the ONE JSON line, send the rest to stderr
    json_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        os.write(json_fd, (json.dumps(obj) + "\n").encode())

    import numpy as np
    import torch
    import cases
    from cases import PAR_INDEX as P  # noqa: F401
    import iemic_b200

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the THCM B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        comm = dist.group.WORLD
    balance = 0 if os.environ.get("THCM_BALANCE", "1") == "0" else 1
    s, landm = cases.global_synth(n, m, l, rank=rank, nranks=world, device=local_rank, balance=balance)
    kry_compact = os.environ.get("THCM_KRYLOV_COMPACT", "1") != "0"
    impl_config = {
        "parallelism": f"lon x lat block partition over {a.gpus} GPU(s), one process per GPU; SpMV halo: LL-format stores into the neighbours' "
                       "buffers over NVLink, polled by the SpMV kernel itself; dots: reduction kernels with a fused LL all-reduce over peer memory",
        "decomposition": ("Decomp2D rank grid, cut lines placed by ocean-cell count (thcmb_settings.balance = 1)" if balance and world > 1
                          else "Decomp2D, uniform cut lines (TRIOS_Domain.C:258-273)"),
        "orthogonalisation": ("batched classical Gram-Schmidt + DGKS criterion (Belos 'DGKS', Ocean.C:977-1024): first update and second projection "
                              "share one sweep over the basis; the second update rides in the head kernel of the next Arnoldi step with its norm from "
                              "Pythagoras (two all-reduces per iteration)" if a.ortho == "dgks" else "modified Gram-Schmidt (GMRESSolver.H:177-181)"),
        "krylov_space": "ocean cells only (LAND rows are identity rows, b = 0 there)" if kry_compact else "full-length vectors",
        "l2": "working set (Jacobian 1.6 GB + Krylov basis) >> 126 MB L2; no flush needed"}
    t = iemic_b200.THCM(s, landm, comm)
    for k, v in PARS.items():
        t.setParameter(k, v)
    t.set_ortho(a.ortho)
    xg = cases.consistent_state(s, landm, scale=0.05)
    x_local = xg[t.local_gids()]
    xd = torch.from_numpy(x_local).cuda()
    dx = t.new_vector()
    x_pin = torch.from_numpy(x_local).pin_memory()
    dx_pin = torch.empty(t.ndim, dtype=torch.float64).pin_memory()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
        t.sync()

    def step_dev():
        return t.newton_step_dev(xd, dx, tol=0.0, maxit=iters - 1, restart=iters, precon=a.precon)

    def step_e2e():
        return t.newton_step(x_pin, dx_pin, tol=0.0, maxit=iters - 1, restart=iters, precon=a.precon)

    def timed(fn, k):
        """k steps bracketed by barrier + synchronize, device time via CUDA events on the library's stream, max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record(t.stream)
        for _ in range(k):
            res = fn()
        e1.record(t.stream)
        barrier()
        ms = e0.elapsed_time(e1)
        wall = (time.perf_counter() - w0) * 1e3
        if world > 1:
            tt = torch.tensor([ms, wall], dtype=torch.float64, device="cuda")
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
            ms, wall = tt.tolist()
        return ms, wall, res

    for _ in range(max(a.warmup, 3)):
        step_dev()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = t.launch_count()
    ms, wall, (res, _fn) = timed(step_dev, a.steps)
    launches = t.launch_count() - l0
    for _ in range(2):
        step_e2e()
    ms_e2e, wall_e2e, (res_e, fnorm) = timed(step_e2e, a.steps)
    clocks = sampler.stop() if rank == 0 else None

    # ---- per-kernel device time inside one (untimed) profiled step: the roofline evidence ----
    t.profile(True)
    step_dev()
    prof = t.profile_report()
    t.profile(False)
    # extra launches of the assembly kernels alone (they run once per step) for a stable average
    F = t.new_vector()
    t.profile(True)
    for _ in range(10):
        t.evaluate(xd, F, True)
    prof_asm = t.profile_report()
    t.profile(False)
    prof.update({k: v for k, v in prof_asm.items() if k.startswith("thcm_assemble")})
    # the full-length operator application (Model::applyMatrix outside a solve)
    yv = t.new_vector()
    t.applyMatrix(xd, yv)
    t.profile(True)
    for _ in range(10):
        t.applyMatrix(xd, yv)
    prof_full = t.profile_report()
    t.profile(False)

    if os.environ.get("THCM_BENCH_RANK_TABLES"):   # per-rank kernel tables (load-balance diagnosis on multi-GPU boxes)
        os.makedirs(os.environ["THCM_BENCH_RANK_TABLES"], exist_ok=True)
        with open(os.path.join(os.environ["THCM_BENCH_RANK_TABLES"], f"kernels_g{world}_rank{rank}.json"), "w") as fh:
            json.dump({"rank": rank, "ms_per_step": ms / a.steps, "ocean_cells": int(t.n_ocean_cells()) if hasattr(t, "n_ocean_cells") else None,
                       "ndim": int(t.ndim), "nnz": int(t.nnz), "kernels": {k: [c_, tot / c_] for k, (c_, tot) in prof.items()}}, fh)
    if rank != 0:
        return 0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    ncell_loc, ndim_loc, nnz_loc = t.ndim // 6, t.ndim, t.nnz
    # ocean / LAND bookkeeping of this rank's block (SURVEY 8d accounting rule: report the streamed-bytes AND the graph-equivalent figure)
    rp_, _ = t.graph()
    cellg = t.local_gids()[::6] // 6
    ci, cj, ck = cellg % n, (cellg // n) % m, cellg // (n * m)
    land = landm[ck + 1, cj + 1, ci + 1] != 0
    lens = np.diff(rp_).reshape(-1, 6).sum(axis=1)
    nnz_ocean, ncell_ocean = int(lens[~land].sum()), int((~land).sum())
    ntile_all, ntile_active = t.tile_counts()
    nnz_active = int(nnz_loc * ntile_active / max(ntile_all, 1))     # entries of the tiles the Jacobian kernels revisit (tiles are 32 cells)
    nk = 6 * ncell_ocean if kry_compact else ndim_loc      # length of the Krylov vectors
    nvavg = (iters + 1) / 2                                 # basis vectors of an orthogonalisation pass, averaged over a cycle
    pyth_head = kry_compact and a.ortho == "dgks" and a.precon == 1 and not os.environ.get("THCM_EXPLICIT_NORM") and not os.environ.get("THCM_NO_FUSED_HEAD")
    alg_bytes = {  # ALGORITHMIC bytes per launch of the format being timed (minimum compulsory traffic)
        # compact SpMV: values + compact column ids of the ocean rows, row pointers, cell map, x and y once
        "spmv_csr": (nnz_ocean * 12 + 6 * ncell_ocean * 4 + ncell_ocean * 4 + nk * 16) if kry_compact else (nnz_ocean * 12 + 6 * ncell_ocean * 4 + ncell_loc + ndim_loc * 16),
        "thcm_assemble<JAC_GRAPH>": int(ncell_loc * ntile_active / max(ntile_all, 1)) * 49 + 8 * nnz_active,
        "thcm_assemble<RHS>": ncell_loc * 145,
        "mgs_step": 32 * nk, "dot": 16 * nk,
        "multi_dot": int(8 * nk * (nvavg + 1)),             # nv basis vectors + w, each once
        "multi_axpy": int(8 * nk * (nvavg + 2)),            # first update + second projection: nv basis vectors + w read + w written
        "second_update": 0,                                 # the explicit second update + norm: launched every iteration, leaves at once unless
                                                            # the guard of the Pythagorean norm tripped (then 8 nk (nv + 2))
        "axpby": 24 * nk, "axpy_negdev": 24 * nk, "scale_invsqrt": 16 * nk, "copy": 16 * nk, "fill": 8 * nk,
        # head of an Arnoldi step: w in, v and z out, 36 doubles of the block inverse per cell -- plus, whenever the DGKS criterion asked
        # for it (every iteration of this benchmark: counted), the second Gram-Schmidt update against the nv basis vectors
        "blockdiag_apply": (24 * 6 + 36 * 8) * (nk // 6) + (int(8 * nk * nvavg) if pyth_head else 0),
        "blockdiag_build": ncell_ocean * (36 * 8 + 6 * 4) + ncell_loc * 36 * 8 + ncell_loc,   # in-cell entries + row pointers of the ocean rows, inverses out
    }
    graph_equiv = {"spmv_csr": nnz_loc * 12 + ndim_loc * 20, "thcm_assemble<JAC_GRAPH>": ncell_loc * 49 + 8 * nnz_loc,
                   "multi_dot": int(8 * ndim_loc * (nvavg + 1)), "multi_axpy": int(8 * ndim_loc * (nvavg + 2))}
    symbol = {"spmv_csr": "spmv_compact_kernel" if kry_compact else "spmv_csr_kernel", "multi_dot": "multi_dot_kernel",
              "multi_axpy": "fused2_axpy_dot_kernel", "thcm_assemble<JAC_GRAPH>": "thcm_jac_tma_kernel", "thcm_assemble<RHS>": "thcm_rhs_tma_kernel",
              "blockdiag_apply": "scale_precon_push_kernel", "blockdiag_build": "blockdiag_build_kernel", "second_update": "multi_axpy_dot_kernel"}
    step_ms = ms / a.steps
    kernels = {}
    for name, (cnt, tot) in prof.items():
        avg = tot / cnt
        e = {"launches_per_step": cnt if not name.startswith("thcm_assemble") else 1, "avg_ms": avg, "symbol": symbol.get(name)}
        if name in alg_bytes:
            e["alg_bytes"] = alg_bytes[name]
            e["gbs"] = alg_bytes[name] / (avg * 1e-3) / 1e9
            e["frac_of_peak"] = e["gbs"] / peak
            if name in graph_equiv and graph_equiv[name] != alg_bytes[name]:
                e["graph_equivalent_bytes"] = graph_equiv[name]
                e["graph_equivalent_frac_of_peak"] = graph_equiv[name] / (avg * 1e-3) / 1e9 / peak
        e["share_of_step"] = e["launches_per_step"] * avg / step_ms
        kernels[name] = e
    if "spmv_csr" in prof_full:
        cnt, tot = prof_full["spmv_csr"]
        fb = nnz_ocean * 12 + 6 * ncell_ocean * 4 + ncell_loc + ndim_loc * 16
        kernels["spmv_full_length"] = {"symbol": "spmv_csr_kernel", "avg_ms": tot / cnt, "alg_bytes": fb, "gbs": fb / (tot / cnt * 1e-3) / 1e9,
                                       "frac_of_peak": fb / (tot / cnt * 1e-3) / 1e9 / peak, "graph_equivalent_bytes": graph_equiv["spmv_csr"],
                                       "graph_equivalent_frac_of_peak": graph_equiv["spmv_csr"] / (tot / cnt * 1e-3) / 1e9 / peak,
                                       "note": "Model::applyMatrix on full-length vectors (identity rows of LAND cells answered without streaming them)"}
    dom = max((k for k in kernels if "share_of_step" in kernels[k]), key=lambda k: kernels[k]["share_of_step"])
    # DRAM traffic per launch from the committed `ncu --set full` captures (profiles/ncu_traffic.json, keyed by kernel symbol; captures at
    # nv ~ 25 = the average of a 50-iteration cycle), null when that kernel was not captured at this size
    traffic, traffic_src = None, None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if a.gpus == 1 and [n, m, l] == list(GRID):
            for kname, e in kernels.items():
                ent = tr.get(e.get("symbol") or "", None)
                if ent:
                    e["ncu_dram_bytes"] = ent["dram_bytes"]
                    e["ncu_capture"] = ent.get("capture")
            traffic = kernels[dom].get("ncu_dram_bytes")
            traffic_src = kernels[dom].get("ncu_capture")
    except Exception:
        pass
    roof = {"kernel": dom, "symbol": kernels[dom].get("symbol"), "bound": "hbm", "achieved": kernels[dom].get("gbs"), "peak": peak, "unit": "GB/s",
            "frac": kernels[dom].get("frac_of_peak"), "traffic": traffic, "traffic_capture": traffic_src, "peak_source": peak_src,
            "alg_bytes_per_launch": kernels[dom].get("alg_bytes"), "avg_launch_ms": kernels[dom]["avg_ms"],
            "share_of_step": kernels[dom]["share_of_step"],
            "graph_equivalent_frac": kernels[dom].get("graph_equivalent_frac_of_peak")}
    line = {"metric": "newton_step_seconds", "value": step_ms * 1e-3, "unit": "s", "n_gpus": a.gpus, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": False, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "impl_config": impl_config, "clocks": clocks,
            "e2e": {"value": ms_e2e / a.steps * 1e-3, "unit": "s", "h2d_bytes_per_step": 8 * t.ndim, "d2h_bytes_per_step": 8 * t.ndim + 8,
                    "wall_ms_per_step": wall_e2e / a.steps, "api": "thcmb_newton_step (pinned host state in, host update out)"},
            "gpu_launches": int(launches), "roofline": roof, "kernels": kernels,
            "spmv_hbm_gbs": kernels.get("spmv_csr", {}).get("gbs"), "spmv_frac_of_peak": kernels.get("spmv_csr", {}).get("frac_of_peak"),
            "assembly_ms": kernels.get("thcm_assemble<JAC_GRAPH>", {}).get("avg_ms"), "residual_ms": kernels.get("thcm_assemble<RHS>", {}).get("avg_ms"),
            "spmv_ms": kernels.get("spmv_csr", {}).get("avg_ms"), "gmres": {"iters": res.iters, "resid": res.resid, "fnorm": fnorm},
            "wall_ms_per_step": wall / a.steps, "ndim_global": 6 * n * m * l, "nnz_local": int(nnz_loc), "ocean_cells_local": ncell_ocean,
            "cells_local": int(ncell_loc), "ns_per_cell_per_gpu": step_ms * 1e6 / (n * m * l / max(a.gpus, 1))}
    # north-star target: FP64 Jacobian assembly + SpMV as a fraction of the HBM roofline -- (a) bytes actually streamed by the formats
    # being timed, (b) graph-equivalent bytes (the full maximal graph incl. identity rows: what the Epetra-equivalent matrix would move)
    ka, ks = kernels.get("thcm_assemble<JAC_GRAPH>"), kernels.get("spmv_csr")
    if ka and ks:
        tms = (ka["avg_ms"] + ks["avg_ms"]) * 1e-3
        gb_s = (ka["alg_bytes"] + ks["alg_bytes"]) / tms / 1e9
        gb_g = (graph_equiv["thcm_assemble<JAC_GRAPH>"] + graph_equiv["spmv_csr"]) / tms / 1e9
        line["assembly_plus_spmv"] = {"ms": tms * 1e3, "streamed_bytes": ka["alg_bytes"] + ks["alg_bytes"], "streamed_gbs": gb_s,
                                      "streamed_frac_of_peak": gb_s / peak, "streamed_frac_of_nominal_8TBs": gb_s / 8000.0,
                                      "graph_equivalent_bytes": graph_equiv["thcm_assemble<JAC_GRAPH>"] + graph_equiv["spmv_csr"],
                                      "graph_equivalent_gbs": gb_g, "graph_equivalent_frac_of_peak": gb_g / peak,
                                      "graph_equivalent_frac_of_nominal_8TBs": gb_g / 8000.0}
    t.close()
    if a.gpus == 1 and not a.no_mixing:
        try:
            line["mixing1"] = mixing_timing(n, m, l, xg, iters, a, peak)
        except Exception as ex:
            line["mixing1"] = {"failed": str(ex)}
    if a.gpus == 1 and not a.no_b1 and [n, m, l] == list(GRID):
        try:
            line["e2e_b1"] = b1_boundary_timing(s, landm, x_local)
        except Exception as ex:
            line["e2e_b1"] = {"failed": str(ex)}
    if a.gpus == 1 and not a.no_cpu_baseline:
        try:
            # bounded sample (about 15-25 s of CPU work): as many of the 32 blocks as there are cores, one untimed pass (it builds the
            # per-block models: a timed first pass reported 9.7 s where the reference arm measures 5.0 s) + one timed pass, scaled by 32 / blocks
            ncores = len(os.sched_getaffinity(0))
            r = reference_run(n, m, l, iters, 1, 1, blocks=range(min(32, max(1, ncores))))
            line["cpu_baseline"] = {"value": r["value"], "unit": "s", "cores": r["cores"], "kind": "port", "sample": r["sample"], "stages_s": r["stages_s"]}
        except Exception as ex:  # the baseline must never take the benchmark line down
            line["cpu_baseline"] = {"value": None, "unit": "s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
    emit(line)
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


def mixing_timing(n, m, l, xg, iters, a, peak):
    """The same grid, state and parameters with Mixing = 1 (the default of the reference's run configurations: implicit vertical mixing /
    convective adjustment, mix_imp.f:489-492, whose forward-difference Jacobian block costs six extra evaluations of the mixing term per
    T / S row, mix_imp.f:729-815): device time of the residual kernel, the Jacobian pair and the whole fixed-work Newton step."""
    import torch
    import cases
    import iemic_b200
    s, landm = cases.global_synth(n, m, l, vmix=1)
    t = iemic_b200.THCM(s, landm, None)
    for k, v in PARS.items():
        t.setParameter(k, v)
    t.set_ortho(a.ortho)
    xd = torch.from_numpy(xg).cuda()
    dx, F = t.new_vector(), t.new_vector()
    for _ in range(3):
        t.newton_step_dev(xd, dx, tol=0.0, maxit=iters - 1, restart=iters, precon=a.precon)
    torch.cuda.synchronize(); t.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(t.stream)
    for _ in range(a.steps):
        res, _ = t.newton_step_dev(xd, dx, tol=0.0, maxit=iters - 1, restart=iters, precon=a.precon)
    e1.record(t.stream)
    torch.cuda.synchronize(); t.sync()
    step_ms = e0.elapsed_time(e1) / a.steps
    t.profile(True)
    for _ in range(10):
        t.evaluate(xd, F, True)
    prof = t.profile_report()
    t.profile(False)
    ncell = t.ndim // 6
    ntile_all, ntile_active = t.tile_counts()
    jac_bytes = int(ncell * ntile_active / max(ntile_all, 1)) * 49 + 8 * int(t.nnz * ntile_active / max(ntile_all, 1))
    out = {"what": "Mixing = 1 (implicit vertical mixing / convective adjustment + forward-difference Jacobian block), same grid, state, parameters",
           "ms_per_step": step_ms, "gmres": {"iters": res.iters, "resid": res.resid}, "vmix_flags": t.vmix_flags()}
    for key, name, nbytes in (("residual", "thcm_assemble<RHS>", ncell * 145), ("jacobian", "thcm_assemble<JAC_GRAPH>", jac_bytes)):
        if name in prof:
            cnt, tot = prof[name]
            out[key + "_ms"] = tot / cnt
            out[key + "_frac_of_peak"] = nbytes / (tot / cnt * 1e-3) / 1e9 / peak
    t.close()
    return out


def b1_boundary_timing(s, landm, x):
    """The reference boundary itself: rhs_ + matrix_ through the gfortran-mangled symbols with caller-owned PAGEABLE host
    buffers, the way THCM.C:1001 / :1066 call them -- state H2D, kernels, residual resp. the Fortran-order thresholded CRS D2H."""
    import numpy as np
    import iemic_b200
    f = iemic_b200.FortranABI()
    f.global_initialize(s)
    f.init(s, landm)
    for k, v in PARS.items():
        f.setparcs(k, v)
    x = np.ascontiguousarray(x)
    out = {"rhs_ms": [], "matrix_ms": []}
    nnz = 0
    for _ in range(4):
        t0 = time.perf_counter(); f.rhs_inplace(x)
        t1 = time.perf_counter(); nnz = f.matrix_inplace(x)
        t2 = time.perf_counter()
        out["rhs_ms"].append((t1 - t0) * 1e3); out["matrix_ms"].append((t2 - t1) * 1e3)
    ndim = f.ndim
    f.finalize()
    rhs_ms, mat_ms = sorted(out["rhs_ms"][1:])[1], sorted(out["matrix_ms"][1:])[1]
    return {"api": "rhs_(un, B) + matrix_(un) (B1 Fortran symbols, pageable caller-owned buffers)", "rhs_ms": rhs_ms, "matrix_ms": mat_ms,
            "value": (rhs_ms + mat_ms) * 1e-3, "unit": "s", "h2d_bytes": 2 * 8 * ndim, "d2h_bytes": 8 * ndim + 4 * (ndim + 1) + 12 * int(nnz) + 8 * ndim,
            "crs_entries": int(nnz), "note": "wall clock, median of 3 after one warm-up; the D2H of the thresholded CRS into pageable memory dominates matrix_"}


if __name__ == "__main__":
    sys.exit(main())
