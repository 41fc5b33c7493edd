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


class ReferenceRunner:
    """The restated reference CPU path on the box's host cores.  The grid is split into 32 blocks with the reference's Decomp2D rule
    (each with its 2 ghost layers: what `mpirun -np 32` of the reference holds); a pool of one process per usable core works through
    ALL 32 blocks every step -- nothing is extrapolated: a step's time is the wall clock of the whole pass."""

    def __init__(self, n, m, l, iters, max_workers=None, nblocks=32, blocks=None):
        import multiprocessing as mp
        import psutil
        self.n, self.m, self.l, self.iters, self.nblocks = n, m, l, iters, nblocks
        self.blocks = list(range(nblocks)) if blocks is None else list(blocks)
        cores = len(os.sched_getaffinity(0))
        mem_gb = psutil.virtual_memory().available / 2**30
        per_worker_gb = 6.5 * (8 / nblocks) * (n * m * l) / (360 * 152 * 24) * max(1, len(self.blocks) / max(1, min(cores, len(self.blocks)))) + 0.5
        self.workers = max(1, min(len(self.blocks), cores, int(mem_gb // per_worker_gb), max_workers or len(self.blocks)))
        self.pool = mp.get_context("spawn").Pool(self.workers)
        self.cells = None

    def step(self):
        t0 = time.perf_counter()
        res = self.pool.map(_ref_worker, [(self.n, self.m, self.l, self.nblocks, b, self.iters) for b in self.blocks], chunksize=1)
        wall = time.perf_counter() - t0
        self.cells = res[0]["cells"]
        stages = [sum(r["stages"][q] for r in res) / self.workers for q in range(3)]   # core-seconds / cores
        return wall, stages

    def close(self):
        self.pool.close()
        self.pool.join()

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
    config = {"workload": workload_name(n, m, l, iters), "grid": [n, m, l], "gmres_iters": iters, "precon": "6x6 block-diagonal",
              "parallelism": f"lon x lat block partition over {a.gpus} GPU(s) (Decomp2D), halo: one P2P push kernel over NVLink peer memory, dots: reduction kernels with a fused LL all-reduce over peer memory", "l2": "working set (Jacobian 1.6 GB + Krylov basis) >> 126 MB L2; no flush needed"}

    if a.impl == "reference":
        if rank != 0:
            return 0
        r = reference_run(n, m, l, iters, a.steps, a.warmup)
        line = {"impl": "reference", "metric": "newton_step_seconds", "value": r["value"], "unit": "s", "n_gpus": a.gpus, "steps": a.steps,
                "warmup": a.warmup, "ms_per_step": r["value"] * 1e3, "higher_is_better": False, "scaling": scaling, "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
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
    config["decomposition"] = ("Decomp2D rank grid, cut lines placed by ocean-cell count (thcmb_settings.balance = 1)" if balance and world > 1
                               else "Decomp2D, uniform cut lines (TRIOS_Domain.C:258-273)")
    t = iemic_b200.THCM(s, landm, comm)
    for k, v in PARS.items():
        t.setParameter(k, v)
    t.set_ortho(a.ortho)
    config["gmres_orthogonalisation"] = ("batched Gram-Schmidt + DGKS criterion (Belos 'DGKS', Ocean.C:977-1024); on one GPU the first "
                                         "update and the second projection share one sweep over the basis (3 basis reads per iteration)"
                                         if a.ortho == "dgks" else "modified Gram-Schmidt (GMRESSolver.H:177-181)")
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

    if rank != 0:
        return 0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak, peak_src = (peaks["hbm_gbs"], "measured (MEASURED_PEAKS.json hbm_gbs)") if "hbm_gbs" in peaks else (6650.0, "fallback (B200_PROFILING.md)")
    ncell_loc, ndim_loc, nnz_loc = t.ndim // 6, t.ndim, t.nnz
    spmv_nnz, spmv_rows_streamed = nnz_loc, ndim_loc
    kry_compact = os.environ.get("THCM_KRYLOV_COMPACT", "1") != "0"
    if os.environ.get("THCM_SPMV_SKIP_LAND") == "1" or kry_compact:
        # identity rows of LAND cells are not streamed (y = x): count the entries of the other rows ("a land-compressed
        # format would legitimately move fewer bytes; report both")
        import numpy as _np
        rp_, _ = t.graph()
        cellg = t.local_gids()[::6] // 6
        ci, cj, ck = cellg % n, (cellg // n) % m, cellg // (n * m)
        land = landm[ck + 1, cj + 1, ci + 1] != 0
        lens = _np.diff(rp_).reshape(-1, 6).sum(axis=1)
        spmv_nnz, spmv_rows_streamed = int(lens[~land].sum()), int(6 * (~land).sum())
        config["spmv"] = f"identity rows of LAND cells not streamed: {spmv_nnz} of {nnz_loc} entries, graph-equivalent bytes {nnz_loc * 12 + ndim_loc * 20}"
    nk = ndim_loc          # length of the Krylov vectors
    if kry_compact:
        nk = spmv_rows_streamed
        config["krylov"] = f"ocean-only Krylov space: vectors of {nk} of {ndim_loc} unknowns (LAND rows are identity rows, b = 0 there)"
    alg_bytes = {  # algorithmic bytes per launch
        # bytes actually streamed by the format being timed (SURVEY 8d accounting rule): explicit CRS = values + column ids +
        # row pointers + x + y; with THCM_SPMV_PATTERN=1 the column ids shrink to a 2-byte pattern id per row
        "spmv_csr": ((spmv_nnz * 8 + spmv_rows_streamed * 6 + ndim_loc * 16) if os.environ.get("THCM_SPMV_PATTERN") == "1"
                     else (spmv_nnz * 12 + spmv_rows_streamed * 4 + (nk if kry_compact else ndim_loc) * 16)) + (ndim_loc // 6 if spmv_rows_streamed != ndim_loc else 0),
        "thcm_assemble<JAC_GRAPH>": ncell_loc * 49 + 8 * nnz_loc,
        "thcm_assemble<RHS>": ncell_loc * 145,
        "mgs_step": 32 * nk, "dot": 16 * nk,
        # batched Gram-Schmidt: a pass over nv basis vectors + w; nv averages (iters+1)/2 over a cycle
        "multi_dot": int(8 * nk * ((iters + 1) / 2 * (1 + 1 / 8) + 0)), "multi_axpy": int(8 * nk * ((iters + 1) / 2 + 2)), "axpby": 24 * nk, "axpy_negdev": 24 * nk,
        "scale_invsqrt": 16 * nk, "copy": 16 * nk, "fill": 8 * nk,
        "blockdiag_apply": (36 + 12) * 8 * (nk // 6), "blockdiag_build": ncell_loc * 36 * 8 + 12 * nnz_loc,
    }
    step_ms = ms / a.steps
    kernels = {}
    for name, (cnt, tot) in prof.items():
        avg = tot / cnt
        e = {"launches_per_step": cnt if not name.startswith("thcm_assemble") else 1, "avg_ms": avg}
        if name in alg_bytes:
            e["alg_bytes"] = alg_bytes[name]
            e["gbs"] = alg_bytes[name] / (avg * 1e-3) / 1e9
            e["frac_of_peak"] = e["gbs"] / peak
        e["share_of_step"] = e["launches_per_step"] * avg / step_ms
        kernels[name] = e
    dom = max(kernels, key=lambda k: kernels[k]["share_of_step"])
    # DRAM traffic per launch from the committed `ncu --set full` capture (profiles/summarize.py), null when not captured
    traffic = None
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        if a.gpus == 1 and [n, m, l] == list(GRID):
            traffic = tr.get(dom, {}).get("dram_bytes")
            for kname, e in kernels.items():
                if kname in tr:
                    e["ncu_dram_bytes"] = tr[kname]["dram_bytes"]
    except Exception:
        pass
    roof = {"kernel": dom, "bound": "hbm", "achieved": kernels[dom].get("gbs"), "peak": peak, "unit": "GB/s",
            "frac": kernels[dom].get("frac_of_peak"), "traffic": traffic, "peak_source": peak_src,
            "alg_bytes_per_launch": kernels[dom].get("alg_bytes"), "avg_launch_ms": kernels[dom]["avg_ms"],
            "share_of_step": kernels[dom]["share_of_step"]}
    line = {"metric": "newton_step_seconds", "value": step_ms * 1e-3, "unit": "s", "n_gpus": a.gpus, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": step_ms, "higher_is_better": False, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config, "clocks": clocks,
            "e2e": {"value": ms_e2e / a.steps * 1e-3, "unit": "s", "h2d_bytes_per_step": 8 * t.ndim, "d2h_bytes_per_step": 8 * t.ndim + 8,
                    "wall_ms_per_step": wall_e2e / a.steps},
            "gpu_launches": int(launches), "roofline": roof, "kernels": kernels,
            "spmv_hbm_gbs": kernels.get("spmv_csr", {}).get("gbs"), "spmv_frac_of_peak": kernels.get("spmv_csr", {}).get("frac_of_peak"),
            "assembly_ms": kernels.get("thcm_assemble<JAC_GRAPH>", {}).get("avg_ms"), "residual_ms": kernels.get("thcm_assemble<RHS>", {}).get("avg_ms"),
            "spmv_ms": kernels.get("spmv_csr", {}).get("avg_ms"), "gmres": {"iters": res.iters, "resid": res.resid, "fnorm": fnorm},
            "wall_ms_per_step": wall / a.steps, "ndim_global": 6 * n * m * l, "nnz_local": int(nnz_loc),
            "ns_per_cell_per_gpu": step_ms * 1e6 / (n * m * l / max(a.gpus, 1))}
    # north-star target: FP64 Jacobian assembly + SpMV as a fraction of the HBM roofline (algorithmic bytes of both / time of both)
    ka, ks = kernels.get("thcm_assemble<JAC_GRAPH>"), kernels.get("spmv_csr")
    if ka and ks:
        gbs = (ka["alg_bytes"] + ks["alg_bytes"]) / ((ka["avg_ms"] + ks["avg_ms"]) * 1e-3) / 1e9
        line["assembly_plus_spmv"] = {"ms": ka["avg_ms"] + ks["avg_ms"], "alg_bytes": ka["alg_bytes"] + ks["alg_bytes"], "gbs": gbs,
                                      "frac_of_peak": gbs / peak, "frac_of_nominal_8TBs": gbs / 8000.0}
    if a.gpus == 1 and not a.no_cpu_baseline:
        try:
            # bounded sample (about 15-25 s of CPU work): as many of the 32 blocks as there are cores, one pass, scaled by 32 / blocks
            ncores = len(os.sched_getaffinity(0))
            r = reference_run(n, m, l, iters, 1, 0, blocks=range(min(32, max(1, ncores))))
            line["cpu_baseline"] = {"value": r["value"], "unit": "s", "cores": r["cores"], "kind": "port", "sample": r["sample"], "stages_s": r["stages_s"]}
        except Exception as ex:  # the baseline must never take the benchmark line down
            line["cpu_baseline"] = {"value": None, "unit": "s", "cores": 0, "kind": "port", "sample": f"failed: {ex}"}
    emit(line)
    t.close()
    if world > 1:
        torch.distributed.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
