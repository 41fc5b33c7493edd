"""CPU prototype (scipy) for SURVEY section 8f N1: the reference's block Gauss-Seidel ocean preconditioner
(src/trios/TRIOS_BlockPreconditioner.C:1479-1611, SolveLower1) built ALGEBRAICALLY from this library's Jacobian, to find out what a
device version has to contain before any kernel is written.  Test infrastructure / design record: uses the oracle, never the product.

    python scripts/n1_block_gs_prototype.py [state amplitude, default 0.01] [exact|celldiag|vline] [iterations, default 80]

Steps of M^-1 b with b = (b_uv, b_w, b_p, b_TS) on the ocean-only (compact) system:
  1. p~  : hydrostatic rows G_w p = b_w per water column, top-down recurrence with p_top = 0
  2. (y_uv, pbar): depth-averaged saddle point  [A_uv  G_uv Pi ; Om D_uv  0] = (b_uv - G_uv p~ ; Om b_p)
       Pi = barotropic pressure shape per column (null vector of the hydrostatic rows), Om = the weights of the continuity rows that
       cancel w (both derived from the matrix itself by recurrences)
       exact: sparse LU of the saddle system;  celldiag / vline: A_uv replaced by its 2x2 cell blocks / its vertical lines, explicit
       Schur complement S = Om D_uv Ahat^-1 G_uv Pi (dense pseudo-inverse here)
  3. p = p~ + Pi pbar;  w bottom-up from the continuity rows;  T,S: A_TS y = b_TS - B_TSuv y_uv - B_TSw y_w (sparse LU)
Findings (4-degree real mask, 135 234 ocean unknowns): exact sub-solves converge in 2 iterations at the zero
state, reach 1e-2 in 5 iterations and stall near 1e-3 for a random state of amplitude 0.01, 4e-2 after 80 iterations at amplitude 0.05."""
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

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.01
mode = sys.argv[2] if len(sys.argv) > 2 else "exact"
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 80

s, landm = cases.global4deg()
N, M, L = s.N, s.M, s.L
o = OracleTHCM(s, landm)
for k, v in {"COMB": 1.0, "WIND": 1.0, "TEMP": 10.0, "SALT": 1.0}.items():
    o.setpar(P[k], v)
x = cases.consistent_state(s, landm, scale=scale)
F = -o.rhs(x)
val, _ = o.jacobian_graph(x)
rp, col = o.graph()
J = sp.csr_matrix((val, col, rp), shape=(o.ndim, o.ndim))
land = (landm[1:-1, 1:-1, 1:-1] != 0).reshape(-1)
oc = np.repeat(~land, 6)
Jc = J[oc][:, oc].tocsr()
b = -F[oc]
n = Jc.shape[0]
nc = n // 6
cells = np.nonzero(~land)[0]
ci, cj, ck = cells % N, (cells // N) % M, cells // (N * M)
_, col_of_cell = np.unique(cj * N + ci, return_inverse=True)
ncol = col_of_cell.max() + 1
var = np.tile(np.arange(6), nc)
iUV, iW, iP, iTS = (np.nonzero(m)[0] for m in (var < 2, var == 2, var == 3, var >= 4))


def blk(r, c):
    return Jc[r][:, c].tocsr()


Auv, Guv, Gw, Duv, Dw = blk(iUV, iUV), blk(iUV, iP), blk(iW, iP), blk(iP, iUV), blk(iP, iW)
BTSuv, BTSw, ATS = blk(iTS, iUV), blk(iTS, iW), blk(iTS, iTS)
order = np.lexsort((ck, col_of_cell))
colstart = np.searchsorted(col_of_cell[order], np.arange(ncol + 1))
top = np.zeros(nc, bool)
up, dn = -np.ones(nc, int), -np.ones(nc, int)
for q in range(ncol):
    cs = order[colstart[q]:colstart[q + 1]]
    top[cs[-1]] = True
    up[cs[:-1]] = cs[1:]
    dn[cs[1:]] = cs[:-1]


def entry(Mx, r, c):
    s_, e_ = Mx.indptr[r], Mx.indptr[r + 1]
    idx = np.nonzero(Mx.indices[s_:e_] == c)[0]
    return Mx.data[s_ + idx[0]] if len(idx) else 0.0


g0 = np.array([entry(Gw, c, c) for c in range(nc)])
g1 = np.array([entry(Gw, c, up[c]) if up[c] >= 0 else 0.0 for c in range(nc)])
a_ = np.array([entry(Dw, c, c) for c in range(nc)])
cdn = np.array([entry(Dw, c, dn[c]) if dn[c] >= 0 else 0.0 for c in range(nc)])
Pi, Om = np.zeros(nc), np.zeros(nc)
for q in range(ncol):
    cs = order[colstart[q]:colstart[q + 1]]
    Pi[cs[-1]] = Om[cs[-1]] = 1.0
    for c in cs[-2::-1]:
        Pi[c] = -g1[c] * Pi[up[c]] / g0[c]          # g0 p_c + g1 p_up = 0
        Om[c] = -Om[up[c]] * cdn[up[c]] / a_[c]     # Om_c a_c + Om_up cdn_up = 0
PiM = sp.csr_matrix((Pi, (np.arange(nc), col_of_cell)), shape=(nc, ncol))
OmM = sp.csr_matrix((Om, (col_of_cell, np.arange(nc))), shape=(ncol, nc))
assert abs((OmM @ Dw)[:, np.nonzero(~top)[0]]).max() < 1e-12 and abs((Gw @ PiM)[np.nonzero(~top)[0]]).max() < 1e-12


def p_tilde(bw):
    p = np.zeros(nc)
    for q in range(ncol):
        for c in order[colstart[q]:colstart[q + 1]][-2::-1]:
            p[c] = (bw[c] - g1[c] * p[up[c]]) / g0[c]
    return p


def w_solve(rhs, bw):
    w = np.zeros(nc)
    for q in range(ncol):
        cs = order[colstart[q]:colstart[q + 1]]
        for c in cs[:-1]:
            w[c] = (rhs[c] - (cdn[c] * w[dn[c]] if dn[c] >= 0 else 0.0)) / a_[c]
        w[cs[-1]] = bw[cs[-1]]
    return w


GPi, OD = (Guv @ PiM).tocsr(), (OmM @ Duv).tocsr()
if mode == "exact":
    Klu = spla.splu(sp.bmat([[Auv, GPi], [OD, 1e-9 * sp.eye(ncol)]]).tocsc())

    def saddle(r_uv, r_p):
        z = Klu.solve(np.concatenate([r_uv, r_p]))
        return z[:len(iUV)], z[len(iUV):]
else:
    A = Auv.tocoo()
    keep = (A.row // 2 == A.col // 2) if mode == "celldiag" else (np.repeat(col_of_cell, 2)[A.row] == np.repeat(col_of_cell, 2)[A.col])
    Ahlu = spla.splu(sp.csr_matrix((A.data[keep], (A.row[keep], A.col[keep])), shape=Auv.shape).tocsc())
    Sinv = np.linalg.pinv(OD @ Ahlu.solve(GPi.toarray()), rcond=1e-10)

    def saddle(r_uv, r_p):
        pb = Sinv @ (OD @ Ahlu.solve(r_uv) - r_p)
        return Ahlu.solve(r_uv - GPi @ pb), pb
ATSlu = spla.splu(ATS.tocsc())


def prec(bv):
    buv, bw, bp, bTS = bv[iUV], bv[iW], bv[iP], bv[iTS]
    pt = p_tilde(bw)
    yuv, pb = saddle(buv - Guv @ pt, OmM @ (bp - Dw @ np.where(top, bw, 0.0)))
    yw = w_solve(bp - Duv @ yuv, bw)
    out = np.zeros(n)
    out[iUV], out[iW], out[iP] = yuv, yw, pt + PiM @ pb
    out[iTS] = ATSlu.solve(bTS - BTSuv @ yuv - BTSw @ yw)
    return out


def fgmres(A, rhs, Mv, m):
    nb = np.linalg.norm(rhs)
    V, Z, H, hist = [rhs / nb], [], np.zeros((m + 1, m)), []
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
        if hist[-1] < 1e-10:
            break
    return sum(yj * zj for yj, zj in zip(y, Z)), np.array(hist)


t0 = time.time()
sol, hist = fgmres(Jc, b, prec, iters)
first = lambda t: (int(np.nonzero(hist < t)[0][0]) + 1) if (hist < t).any() else None  # noqa: E731
print(f"state amplitude {scale}, {mode}: {len(hist)} iterations, residual {hist[-1]:.3e} (true {np.linalg.norm(b - Jc @ sol) / np.linalg.norm(b):.3e}), "
      f"1e-2 at {first(1e-2)}, 1e-4 at {first(1e-4)}, {time.time() - t0:.0f} s")
print("history:", " ".join(f"{h:.1e}" for h in hist[::5]))
