"""CPU laboratory (scipy) for the device block preconditioner of SURVEY section 8f N1.  Test infrastructure / design record: uses the
oracle, never the product.  Builds the blocks of the ocean-only Jacobian, the water-column recurrences and the depth-averaged saddle
point of the reference's block Gauss-Seidel (src/trios/TRIOS_BlockPreconditioner.C:1479-1611) with vectorised numpy, and lets the
sub-solves be swapped (exact LU / cell blocks / line solves / inner Krylov) to find what the CUDA version must contain.

    python scripts/n1_lab.py --grid 4 --state newton --comb 0.1 --newton 2 --saddle celldiag --ats exact --iters 60
"""
import argparse
import os
import sys
import time

import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
from cases import PAR_INDEX as P  # noqa: E402
from oracle.oracle import OracleTHCM  # noqa: E402


class Blocks:
    """Ocean-only Jacobian split into the reference's blocks + the water-column structure."""

    def __init__(self, s, landm, J):
        N, M, L = s.N, s.M, s.L
        self.N, self.M, self.L = N, M, L
        land = (landm[1:-1, 1:-1, 1:-1] != 0).reshape(-1)
        self.oc = np.repeat(~land, 6)
        Jc = J[self.oc][:, self.oc].tocsr()
        self.Jc = Jc
        n = Jc.shape[0]
        nc = n // 6
        self.n, self.nc = n, nc
        cells = np.nonzero(~land)[0]
        ci, cj, ck = cells % N, (cells // N) % M, cells // (N * M)
        self.ci, self.cj, self.ck = ci, cj, ck
        colid, col_of_cell = np.unique(cj * N + ci, return_inverse=True)
        self.colid, self.col_of_cell = colid, col_of_cell
        self.ncol = ncol = len(colid)
        var = np.tile(np.arange(6), nc)
        self.iUV, self.iW, self.iP, self.iTS = (np.nonzero(m)[0] for m in (var < 2, var == 2, var == 3, var >= 4))

        def blk(r, c):
            return Jc[r][:, c].tocsr()
        iUV, iW, iP, iTS = self.iUV, self.iW, self.iP, self.iTS
        self.Auv, self.Guv, self.Gw = blk(iUV, iUV), blk(iUV, iP), blk(iW, iP)
        self.Duv, self.Dw = blk(iP, iUV), blk(iP, iW)
        self.BTSuv, self.BTSw, self.ATS = blk(iTS, iUV), blk(iTS, iW), blk(iTS, iTS)
        self.BwTS, self.Buvw, self.Aww, self.BuvTS = blk(iW, iTS), blk(iUV, iW), blk(iW, iW), blk(iUV, iTS)
        # cell index of the cell above / below (compact), -1 when none
        full2c = -np.ones(N * M * L, int)
        full2c[cells] = np.arange(nc)
        up = np.where(ck < L - 1, full2c[np.minimum(cells + N * M, N * M * L - 1)], -1)
        dn = np.where(ck > 0, full2c[np.maximum(cells - N * M, 0)], -1)
        self.up, self.dn = up, dn
        self.top = up < 0
        assert (ck[self.top] == L - 1).all(), "ocean columns reach the surface"
        self.levels = [np.nonzero(ck == k)[0] for k in range(L)]

        def entries(Mx, rows, cols):
            out = np.zeros(len(rows))
            ok = cols >= 0
            out[ok] = np.asarray(Mx[rows[ok], cols[ok]]).ravel()
            return out
        allc = np.arange(nc)
        self.g0, self.g1 = entries(self.Gw, allc, allc), entries(self.Gw, allc, up)
        self.a_, self.cdn = entries(self.Dw, allc, allc), entries(self.Dw, allc, dn)
        Pi, Om = np.zeros(nc), np.zeros(nc)
        Pi[self.top] = Om[self.top] = 1.0
        for k in range(L - 2, -1, -1):
            c = self.levels[k]
            Pi[c] = -self.g1[c] * Pi[up[c]] / self.g0[c]
            Om[c] = -Om[up[c]] * self.cdn[up[c]] / self.a_[c]
        self.Pi, self.Om = Pi, Om
        self.PiM = sp.csr_matrix((Pi, (allc, col_of_cell)), shape=(nc, ncol))
        self.OmM = sp.csr_matrix((Om, (col_of_cell, allc)), shape=(ncol, nc))
        nt = np.nonzero(~self.top)[0]
        assert abs((self.OmM @ self.Dw)[:, nt]).max() < 1e-10 and abs((self.Gw @ self.PiM)[nt]).max() < 1e-10
        self.GPi, self.OD = (self.Guv @ self.PiM).tocsr(), (self.OmM @ self.Duv).tocsr()

    def p_tilde(self, bw):
        p = np.zeros(self.nc)
        for k in range(self.L - 2, -1, -1):
            c = self.levels[k]
            p[c] = (bw[c] - self.g1[c] * p[self.up[c]]) / self.g0[c]
        return p

    def w_solve(self, rhs, bw):
        w = np.zeros(self.nc)
        for k in range(self.L - 1):
            c = self.levels[k]
            below = np.where(self.dn[c] >= 0, w[np.maximum(self.dn[c], 0)], 0.0)
            w[c] = (rhs[c] - self.cdn[c] * below) / self.a_[c]
        w[self.top] = bw[self.top]
        return w


def cell_blocks(A, nb):
    """block diagonal (nb x nb cell blocks) of a sparse matrix, as a sparse matrix"""
    A = A.tocoo()
    keep = A.row // nb == A.col // nb
    return sp.csr_matrix((A.data[keep], (A.row[keep], A.col[keep])), shape=A.shape)


def blockdiag_inverse(Ah, nb):
    """inverse of a block-diagonal sparse matrix with nb x nb blocks (dense batched inverse)"""
    n = Ah.shape[0] // nb
    A = Ah.tocoo()
    blocks = np.zeros((n, nb, nb))
    blocks[A.row // nb, A.row % nb, A.col % nb] = A.data
    inv = np.linalg.inv(blocks)
    r = (np.arange(n)[:, None, None] * nb + np.arange(nb)[None, :, None]) + np.zeros((1, 1, nb), int)
    c = (np.arange(n)[:, None, None] * nb + np.arange(nb)[None, None, :]) + np.zeros((1, nb, 1), int)
    return sp.csr_matrix((inv.ravel(), (r.ravel(), c.ravel())), shape=Ah.shape)


def line_blocks(A, line_of_unknown):
    A = A.tocoo()
    keep = line_of_unknown[A.row] == line_of_unknown[A.col]
    return sp.csr_matrix((A.data[keep], (A.row[keep], A.col[keep])), shape=A.shape)


class Counter:
    def __init__(self):
        self.n = {}

    def add(self, k, v=1):
        self.n[k] = self.n.get(k, 0) + v


def inner_gmres(A, b, M=None, tol=1e-2, maxit=20, cnt=None, key=""):
    its = [0]

    def cb(_):
        its[0] += 1
    x, _ = spla.gmres(A, b, M=M, rtol=tol, restart=maxit, maxiter=1, callback=cb, callback_type="pr_norm")
    if cnt is not None:
        cnt.add(key + "_solves")
        cnt.add(key + "_its", its[0])
    return x


class BlockGS:
    def __init__(self, B, saddle="exact", ats="exact", chat="exact", chat_tol=1e-3, chat_it=400, ats_tol=1e-2, ats_it=15, uv_sweeps=0,
                 upper=False):
        self.B, self.cnt = B, Counter()
        self.saddle_mode, self.ats_mode, self.chat_mode, self.upper = saddle, ats, chat, upper
        self.chat_tol, self.chat_it, self.ats_tol, self.ats_it, self.uv_sweeps = chat_tol, chat_it, ats_tol, ats_it, uv_sweeps
        nuv = len(B.iUV)
        if saddle == "exact":
            self.Klu = spla.splu(sp.bmat([[B.Auv, B.GPi], [B.OD, 1e-9 * sp.eye(B.ncol)]]).tocsc())
        else:
            if saddle == "celldiag":
                Ah = cell_blocks(B.Auv, 2)
            elif saddle == "vline":
                Ah = line_blocks(B.Auv, np.repeat(B.col_of_cell, 2))
            else:
                raise ValueError(saddle)
            self.Ahlu = spla.splu(Ah.tocsc())
            # explicit Chat = OD Ah^-1 GPi (sparse: Ah^-1 is block diagonal for celldiag)
            if saddle == "celldiag":
                Ai = blockdiag_inverse(Ah, 2)
                self.Ahinv = Ai
                self.Chat = (B.OD @ Ai @ B.GPi).tocsr()
            else:
                self.Chat = sp.csr_matrix(B.OD @ self.Ahlu.solve(B.GPi.toarray()))
            self.Chat.eliminate_zeros()
            if chat == "exact":
                self.Chat_pinv = np.linalg.pinv(self.Chat.toarray(), rcond=1e-10) if B.ncol < 6000 else None
                if self.Chat_pinv is None:
                    self.Chat_lu = spla.splu((self.Chat + 1e-9 * abs(self.Chat).max() * sp.eye(B.ncol)).tocsc())
        if ats == "exact":
            self.ATSlu = spla.splu(B.ATS.tocsc())
        elif ats == "vline":
            self.ATSlu = spla.splu(line_blocks(B.ATS, np.repeat(B.col_of_cell, 2)).tocsc())
        elif ats == "celldiag":
            self.ATSlu = spla.splu(cell_blocks(B.ATS, 2).tocsc())
        elif ats.startswith("gmres"):  # gmres-vline / gmres-celldiag
            pre = ats.split("-")[1]
            lu = spla.splu((line_blocks(B.ATS, np.repeat(B.col_of_cell, 2)) if pre == "vline" else cell_blocks(B.ATS, 2)).tocsc())
            self.ATSpre = spla.LinearOperator(B.ATS.shape, matvec=lu.solve)

    def chat_solve(self, r):
        if self.chat_mode == "exact":
            return self.Chat_pinv @ r if self.Chat_pinv is not None else self.Chat_lu.solve(r)
        raise ValueError

    def saddle(self, r_uv, r_p):
        B = self.B
        if self.saddle_mode == "exact":
            z = self.Klu.solve(np.concatenate([r_uv, r_p]))
            return z[:len(B.iUV)], z[len(B.iUV):]
        pb = self.chat_solve(B.OD @ self.Ahlu.solve(r_uv) - r_p)
        y = self.Ahlu.solve(r_uv - B.GPi @ pb)
        for _ in range(self.uv_sweeps):   # Jacobi-type defect correction on the uv solve (keeps pb)
            y = y + self.Ahlu.solve(r_uv - B.GPi @ pb - B.Auv @ y)
        return y, pb

    def ats_solve(self, r):
        if self.ats_mode in ("exact", "vline", "celldiag"):
            return self.ATSlu.solve(r)
        return inner_gmres(self.B.ATS, r, M=self.ATSpre, tol=self.ats_tol, maxit=self.ats_it, cnt=self.cnt, key="ats")

    def __call__(self, bv):
        B = self.B
        buv, bw, bp, bTS = bv[B.iUV], bv[B.iW], bv[B.iP], bv[B.iTS]
        pt = B.p_tilde(bw)
        yuv, pb = self.saddle(buv - B.Guv @ pt, B.OmM @ (bp - B.Dw @ np.where(B.top, bw, 0.0)))
        yw = B.w_solve(bp - B.Duv @ yuv, bw)
        out = np.zeros(B.n)
        out[B.iUV], out[B.iW], out[B.iP] = yuv, yw, pt + B.PiM @ pb
        out[B.iTS] = self.ats_solve(bTS - B.BTSuv @ yuv - B.BTSw @ yw)
        return out


def fgmres(A, rhs, Mv, m, tol=1e-10, verbose=False):
    nb = np.linalg.norm(rhs)
    V, Z, H, hist = [rhs / nb], [], np.zeros((m + 1, m)), []
    y = np.zeros(0)
    for i in range(m):
        Z.append(Mv(V[i]))
        w = A @ Z[i]
        for _ in range(2):
            for k in range(i + 1):
                hk = w @ V[k]
                H[k, i] += hk
                w -= hk * V[k]
        H[i + 1, i] = np.linalg.norm(w)
        V.append(w / H[i + 1, i])
        e1 = np.zeros(i + 2)
        e1[0] = nb
        y = np.linalg.lstsq(H[:i + 2, :i + 1], e1, rcond=None)[0]
        hist.append(np.linalg.norm(H[:i + 2, :i + 1] @ y - e1) / nb)
        if verbose:
            print(f"   it {i + 1}: {hist[-1]:.3e}", flush=True)
        if hist[-1] < tol:
            break
    return sum(yj * zj for yj, zj in zip(y, Z)), np.array(hist)


def make_case(grid):
    if grid == 4:
        return cases.global4deg()
    if grid == 2:
        return cases.global_synth(180, 76, 16)
    if grid == 1:
        return cases.global_synth(360, 152, 24)
    if grid == 16:
        return cases.gateway16()
    raise ValueError(grid)


def assemble(o, x):
    F = -o.rhs(x)
    val, _ = o.jacobian_graph(x)
    rp, col = o.graph()
    return F, sp.csr_matrix((val, col, rp), shape=(o.ndim, o.ndim))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", type=int, default=4)
    ap.add_argument("--state", default="random")     # zero | random | newton
    ap.add_argument("--scale", type=float, default=0.01)
    ap.add_argument("--comb", type=float, default=1.0)
    ap.add_argument("--newton", type=int, default=2)
    ap.add_argument("--mixing", type=int, default=0)
    ap.add_argument("--saddle", default="exact")
    ap.add_argument("--ats", default="exact")
    ap.add_argument("--chat", default="exact")
    ap.add_argument("--uv-sweeps", type=int, default=0)
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--tol", type=float, default=1e-6)
    ap.add_argument("--ats-tol", type=float, default=1e-2)
    ap.add_argument("--ats-it", type=int, default=15)
    ap.add_argument("--chat-tol", type=float, default=1e-3)
    ap.add_argument("--chat-it", type=int, default=400)
    ap.add_argument("--hist", type=int, default=0)
    a = ap.parse_args()
    kw = {"vmix": a.mixing} if a.mixing else {}
    s, landm = make_case(a.grid)
    if a.mixing:
        s.vmix = a.mixing
    o = OracleTHCM(s, landm)
    for k, v in {"COMB": a.comb, "WIND": 1.0, "TEMP": 10.0, "SALT": 1.0}.items():
        o.setpar(P[k], v)
    if a.state == "zero":
        x = np.zeros(o.ndim)
    elif a.state == "random":
        x = cases.consistent_state(s, landm, scale=a.scale)
    else:
        x = np.zeros(o.ndim)
    nsteps = a.newton if a.state == "newton" else 0
    cache = f"/tmp/n1_state_g{a.grid}_c{a.comb}_n{a.newton}_m{a.mixing}.npy"
    if nsteps and os.path.exists(cache):
        x, nsteps0 = np.load(cache), nsteps
    else:
        nsteps0 = 0
    for step in range(nsteps0, nsteps + 1):
        last = step == nsteps
        t0 = time.time()
        F, J = assemble(o, x)
        B = Blocks(s, landm, J)
        b = -F[B.oc]
        t1 = time.time()
        if last:
            if nsteps:
                np.save(cache, x)
            Mv = BlockGS(B, saddle=a.saddle, ats=a.ats, chat=a.chat, uv_sweeps=a.uv_sweeps, ats_tol=a.ats_tol, ats_it=a.ats_it,
                         chat_tol=a.chat_tol, chat_it=a.chat_it)
        else:
            Mv = BlockGS(B)   # exact sub-solves for the states on the way
        t2 = time.time()
        sol, hist = fgmres(B.Jc, b, Mv, a.iters if last else 100, tol=a.tol if last else 1e-8)
        first = lambda t: (int(np.nonzero(hist < t)[0][0]) + 1) if (hist < t).any() else None  # noqa: E731
        print(f"step {step}: |x|max {abs(x).max():.3g} |F| {np.linalg.norm(F):.4e}  its {len(hist)} resid {hist[-1]:.2e} "
              f"(true {np.linalg.norm(b - B.Jc @ sol) / np.linalg.norm(b):.2e}) 1e-2@{first(1e-2)} 1e-4@{first(1e-4)} 1e-6@{first(1e-6)} "
              f"[asm {t1 - t0:.0f}s setup {t2 - t1:.0f}s solve {time.time() - t2:.0f}s] inner {Mv.cnt.n}", flush=True)
        if last and a.hist:
            print("   hist:", " ".join(f"{h:.1e}" for h in hist[::a.hist]))
        dx = np.zeros(o.ndim)
        dx[B.oc] = sol
        x = x + dx


if __name__ == "__main__":
    main()
