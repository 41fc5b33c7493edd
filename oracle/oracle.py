"""ctypes front-end of the CPU oracles (TEST INFRASTRUCTURE -- never imported by the product).

* ``OracleTHCM``  : oracle/thcm_oracle.cpp, the dense Al/An restatement of the reference's THCM path.
* ``kref_gmres`` / ``kref_idrs`` : oracle/_ref/libkrylov_ref.so, the reference's UNMODIFIED
  GMRESSolver.H / IDRSolver.H compiled from /root/reference (prebuilt .so travels to the GPU box).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


class OracleSettings(C.Structure):
    _fields_ = [("hdim", C.c_double), ("qz", C.c_double), ("alphaT", C.c_double), ("alphaS", C.c_double),
                ("ymin_glob", C.c_double), ("ymax_glob", C.c_double),
                ("periodic", C.c_int), ("ih", C.c_int), ("vmix", C.c_int), ("tap", C.c_int), ("rho_mixing", C.c_int),
                ("coriolis_on", C.c_int), ("TRES", C.c_int), ("SRES", C.c_int), ("iza", C.c_int), ("ite", C.c_int),
                ("its", C.c_int), ("coupled_T", C.c_int), ("coupled_S", C.c_int), ("forcing_type", C.c_int)]


def build(force=False):
    """Compiles the oracles (gcc only; the reference Krylov templates only when /root/reference exists)."""
    so = os.path.join(_HERE, "libthcm_oracle.so")
    src = os.path.join(_HERE, "thcm_oracle.cpp")
    need = force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "fdlibm_tanh.h")))
    kso = os.path.join(_HERE, "_ref", "libkrylov_ref.so")
    ksrc = os.path.join(_HERE, "krylov_ref.cpp")
    if os.path.isdir("/root/reference/src/gmressolver") and (force or not os.path.exists(kso) or os.path.getmtime(kso) < os.path.getmtime(ksrc)):
        need = True
    if need:
        subprocess.run(["make", "-C", _HERE, "-B", "all"], check=True, capture_output=True)


_lib = None
_klib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(os.path.join(_HERE, "libthcm_oracle.so"))
        vp, i, d, ll = C.c_void_p, C.c_int, C.c_double, C.c_long
        L.oracle_create.restype = vp
        L.oracle_create.argtypes = [i, i, i, d, d, d, d, C.POINTER(OracleSettings), vp]
        for name, res, args in [("oracle_destroy", None, [vp]), ("oracle_ndim", i, [vp]), ("oracle_setpar", None, [vp, i, d]),
                                ("oracle_getpar", d, [vp, i]), ("oracle_rhs", None, [vp, vp, vp]), ("oracle_matrix", None, [vp, vp]),
                                ("oracle_fillcolb", None, [vp]), ("oracle_nnz", i, [vp]), ("oracle_bad_columns", ll, [vp]),
                                ("oracle_get_crs", None, [vp, vp, vp, vp]), ("oracle_get_cob", None, [vp, vp]),
                                ("oracle_get_forcing", None, [vp, vp]), ("oracle_get_landm", None, [vp, vp]),
                                ("oracle_get_grid", None, [vp] * 9), ("oracle_set_field", None, [vp, i, vp]),
                                ("oracle_graph", i, [i, i, i, i, vp, vp]),
                                ("oracle_scatter_to_graph", ll, [i, vp, vp, vp, vp, vp, vp]),
                                ("oracle_spmv", None, [i, vp, vp, vp, vp, vp]), ("oracle_matavec", None, [i, vp, vp, vp, vp, vp]),
                                ("oracle_average_block", None, [vp, vp]), ("oracle_compute_scaling", i, [vp, vp, vp, vp]),
                                ("oracle_intcond_scaling", i, [vp, vp, vp]),
                                ("oracle_set_vmix_fix", None, [vp, i]), ("oracle_vmix_fun", None, [vp, vp, vp]),
                                ("oracle_vmix_flags", None, [vp, vp]),
                                ("oracle_set_atmos_parameters", None, [vp, vp]), ("oracle_set_seaice_parameters", None, [vp, vp]),
                                ("oracle_salt_advection", None, [vp, vp, vp]), ("oracle_salt_diffusion", None, [vp, vp, vp]),
                                ("oracle_stochastic_forcing", None, [vp, vp, vp, vp]), ("oracle_set_internal_forcing", None, [vp, vp, vp]),
                                ("oracle_get_coupling_state", None, [vp, vp, vp]), ("oracle_get_field", None, [vp, i, vp]),
                                ("oracle_set_landmask", None, [vp, vp, i, i]), ("oracle_setsres", None, [vp, i])]:
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def klib():
    global _klib
    if _klib is None:
        build()
        path = os.path.join(_HERE, "_ref", "libkrylov_ref.so")
        if not os.path.exists(path):
            raise FileNotFoundError(path + " (built from /root/reference by oracle/Makefile; must travel prebuilt)")
        L = C.CDLL(path)
        vp, i, d = C.c_void_p, C.c_int, C.c_double
        L.kref_gmres.restype = i
        L.kref_gmres.argtypes = [i, vp, vp, vp, i, i, vp, vp, vp, d, i, i, i, vp, i, vp, vp, vp, vp]
        L.kref_idrs.restype = i
        L.kref_idrs.argtypes = [i, vp, vp, vp, i, i, vp, vp, vp, d, i, i, vp, vp, i, vp, vp, vp, vp]
        _klib = L
    return _klib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OracleTHCM:
    """Single-domain THCM exactly as the reference's Fortran sees it after ``init_`` (usrc.F90:6-139)."""

    def __init__(self, s, landm):
        """s: any object with the thcmb_settings fields (iemic_b200.Settings); landm int32[L+2,M+2,N+2]."""
        self.L_ = lib()
        yg = (s.ymin_glob, s.ymax_glob) if (getattr(s, "ymin_glob", 0.0) or getattr(s, "ymax_glob", 0.0)) else (s.ymin, s.ymax)
        os_ = OracleSettings(hdim=s.hdim, qz=s.qz, alphaT=s.alphaT, alphaS=s.alphaS, ymin_glob=yg[0], ymax_glob=yg[1],
                             periodic=s.periodic, ih=s.ih, vmix=s.vmix, tap=s.tap, rho_mixing=s.rho_mixing,
                             coriolis_on=s.coriolis_on, TRES=s.TRES, SRES=s.SRES, iza=s.iza, ite=s.ite, its=s.its,
                             coupled_T=s.coupled_T, coupled_S=s.coupled_S, forcing_type=s.forcing_type)
        landm = np.ascontiguousarray(landm, dtype=np.int32)
        self.n, self.m, self.l, self.periodic = s.N, s.M, s.L, int(s.periodic)
        self.h = self.L_.oracle_create(s.N, s.M, s.L, s.xmin, s.xmax, s.ymin, s.ymax, C.byref(os_), _p(landm))
        if not self.h:
            raise RuntimeError("oracle_create failed")
        self.ndim = self.L_.oracle_ndim(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L_.oracle_destroy(self.h)
            self.h = None

    def setpar(self, idx, val):
        self.L_.oracle_setpar(self.h, int(idx), float(val))

    def getpar(self, idx):
        return self.L_.oracle_getpar(self.h, int(idx))

    def average_block(self):
        """m_scaling::average_block (scaling.F90:29-64) on the Jacobian of the last matrix() call; (6,6), [row, col]."""
        db = np.zeros(36)
        self.L_.oracle_average_block(self.h, _p(db))
        return db.reshape(6, 6).T.copy()

    def compute_scaling(self, db):
        """m_scaling::compute (scaling.F90:70-105): (row_scaling, col_scaling) as THCM (not Trilinos) defines them."""
        dbf = np.ascontiguousarray(np.asarray(db, dtype=np.float64).T).reshape(-1)   # column-major for the Fortran layout
        rs, cs = np.empty(self.ndim), np.empty(self.ndim)
        ok = self.L_.oracle_compute_scaling(self.h, _p(dbf), _p(rs), _p(cs))
        return rs, cs, bool(ok)

    def intcond_scaling(self):
        """thcm_utils.F90:285-309: (values, 1-based S-row indices) of the integral-condition coefficients."""
        val = np.empty(self.n * self.m * self.l); ind = np.empty(self.n * self.m * self.l, dtype=np.int32)
        k = self.L_.oracle_intcond_scaling(self.h, _p(val), _p(ind))
        return val[:k].copy(), ind[:k].copy()

    def set_vmix_fix(self, fix):
        self.L_.oracle_set_vmix_fix(self.h, int(fix))

    def vmix_flags(self):
        out = np.zeros(4, dtype=np.int32)
        self.L_.oracle_vmix_flags(self.h, _p(out))
        return dict(zip(("flag", "temp", "salt", "fix"), out.tolist()))

    def vmix_fun(self, un):
        """Divergence of the diffusive tracer flux (mix_imp.f:231-562) with the current vmix_temp / vmix_salt flags."""
        un = np.ascontiguousarray(un, dtype=np.float64)
        mix = np.empty(self.ndim)
        self.L_.oracle_vmix_fun(self.h, _p(un), _p(mix))
        return mix

    def rhs(self, un):
        """Fortran-sign residual B = -Au - mix + Frc (usrc.F90:523-603)."""
        un = np.ascontiguousarray(un, dtype=np.float64)
        B = np.empty(self.ndim)
        self.L_.oracle_rhs(self.h, _p(un), _p(B))
        return B

    def matrix(self, un):
        """(begA, jcoA, coA) 1-based Fortran-order thresholded CRS + coB (usrc.F90:449-521)."""
        un = np.ascontiguousarray(un, dtype=np.float64)
        self.L_.oracle_matrix(self.h, _p(un))
        nnz = self.L_.oracle_nnz(self.h)
        beg = np.empty(self.ndim + 1, dtype=np.int32); jco = np.empty(nnz, dtype=np.int32); co = np.empty(nnz)
        self.L_.oracle_get_crs(self.h, _p(beg), _p(jco), _p(co))
        cob = np.empty(self.ndim)
        self.L_.oracle_get_cob(self.h, _p(cob))
        return beg, jco, co, cob

    def bad_columns(self):
        return self.L_.oracle_bad_columns(self.h)

    FIELDS = ("taux", "tauy", "tatm", "emip", "spert", "adapted_emip", "qatm", "albe", "patm", "qsa", "msi", "gsi")

    def set_field(self, name, f):
        """m_inserts::insert_* (inserts.F90): surface field [M, N] (i fastest); no recompute until the next setpar."""
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        assert f.size == self.n * self.m
        self.L_.oracle_set_field(self.h, self.FIELDS.index(name), _p(f))

    def set_landmask(self, landm, periodic=None, reinit=1):
        """SUBROUTINE set_landmask (usrc.F90:353-418)."""
        lm = np.ascontiguousarray(landm, dtype=np.int32)
        assert lm.shape == (self.l + 2, self.m + 2, self.n + 2)
        if periodic is not None:
            self.periodic = int(periodic)
        self.L_.oracle_set_landmask(self.h, _p(lm), self.periodic, int(reinit))

    def setsres(self, sres):
        """SUBROUTINE setsres (usrc.F90:434-446): THCM.C:1059-1070 brackets matrix_ with it for the mask test."""
        self.L_.oracle_setsres(self.h, int(sres))

    def get_field(self, name):
        f = np.empty((self.m, self.n))
        self.L_.oracle_get_field(self.h, self.FIELDS.index(name), _p(f))
        return f

    def coupling_state(self):
        """Constants of m_usr / m_atm / m_ice the surface diagnostics of probe.F90 use, and suno(1..m)."""
        v = np.empty(17); suno = np.empty(self.m)
        self.L_.oracle_get_coupling_state(self.h, _p(v), _p(suno))
        names = ("QTnd", "QSnd", "Ooa", "Os", "nus", "lvsc", "qdim", "eta", "dqso", "eo0", "albe0", "albed", "zeta", "a0", "Lf", "Qvar", "Q0")
        d = dict(zip(names, v.tolist()))
        d["suno"] = suno
        return d

    def salt_advection(self, un):
        """m_integrals::salt_advection (integrals.F90:17-51): per-cell integrand [l, m, n]; skipped cells stay 0."""
        un = np.ascontiguousarray(un, dtype=np.float64)
        out = np.zeros(self.n * self.m * self.l)
        self.L_.oracle_salt_advection(self.h, _p(un), _p(out))
        return out

    def salt_diffusion(self, un):
        """m_integrals::salt_diffusion (integrals.F90:53-88)."""
        un = np.ascontiguousarray(un, dtype=np.float64)
        out = np.zeros(self.n * self.m * self.l)
        self.L_.oracle_salt_diffusion(self.h, _p(un), _p(out))
        return out

    def stochastic_forcing(self):
        """get_stochastic_forcing (forcing.F90:235-280): (begF, jcoF, coF), 1-based."""
        beg = np.zeros(self.ndim + 1, dtype=np.int32); jco = np.zeros(self.n * self.m, dtype=np.int32); co = np.zeros(self.n * self.m)
        self.L_.oracle_stochastic_forcing(self.h, _p(beg), _p(jco), _p(co))
        return beg, jco, co

    def set_internal_forcing(self, temp, salt):
        t = np.ascontiguousarray(temp, dtype=np.float64).reshape(-1); s_ = np.ascontiguousarray(salt, dtype=np.float64).reshape(-1)
        assert t.size == s_.size == self.n * self.m * self.l
        self.L_.oracle_set_internal_forcing(self.h, _p(t), _p(s_))

    def set_atmos_parameters(self, pars18):
        """usrc.F90:254-310 with the 18 doubles of Atmosphere::CommPars; re-runs forcing + lin."""
        p = np.ascontiguousarray(pars18, dtype=np.float64)
        assert p.size == 18
        self.L_.oracle_set_atmos_parameters(self.h, _p(p))

    def set_seaice_parameters(self, pars7):
        """usrc.F90:313-350 with the 7 doubles of SeaIce::CommPars; re-runs forcing + lin."""
        p = np.ascontiguousarray(pars7, dtype=np.float64)
        assert p.size == 7
        self.L_.oracle_set_seaice_parameters(self.h, _p(p))

    def forcing(self):
        f = np.empty(self.ndim)
        self.L_.oracle_get_forcing(self.h, _p(f))
        return f

    def landm(self):
        out = np.empty((self.l + 2, self.m + 2, self.n + 2), dtype=np.int32)
        self.L_.oracle_get_landm(self.h, _p(out))
        return out

    def grid(self):
        n, m, l = self.n, self.m, self.l
        x, xu = np.empty(n + 1), np.empty(n + 1); y, yv = np.empty(m + 2), np.empty(m + 1)
        z, zw, dfzT, dfzW = np.empty(l + 1), np.empty(l + 1), np.empty(l + 1), np.empty(l + 1)
        self.L_.oracle_get_grid(self.h, _p(x), _p(y), _p(z), _p(xu), _p(yv), _p(zw), _p(dfzT), _p(dfzW))
        return dict(x=x, y=y, z=z, xu=xu, yv=yv, zw=zw, dfzT=dfzT, dfzW=dfzW)

    def graph(self):
        """Maximal graph (THCM.C:2300-2580), 0-based CSR, columns ascending."""
        nnz = self.L_.oracle_graph(self.n, self.m, self.l, self.periodic, None, None)
        rowptr = np.empty(self.ndim + 1, dtype=np.int32); col = np.empty(nnz, dtype=np.int32)
        self.L_.oracle_graph(self.n, self.m, self.l, self.periodic, _p(rowptr), _p(col))
        return rowptr, col

    def jacobian_graph(self, un, graph=None):
        """Jacobian values in graph order: THCM.C:1052-1162 (zero, ReplaceGlobalValues per row).  Returns (val, n_missing)."""
        beg, jco, co, _ = self.matrix(un)
        rowptr, col = graph if graph is not None else self.graph()
        val = np.empty(len(col))
        miss = self.L_.oracle_scatter_to_graph(self.ndim, _p(beg), _p(jco), _p(co), _p(rowptr), _p(col), _p(val))
        return val, miss


def spmv(rowptr, col, val, x):
    y = np.empty(len(rowptr) - 1)
    x = np.ascontiguousarray(x, dtype=np.float64)
    lib().oracle_spmv(len(rowptr) - 1, _p(rowptr), _p(col), _p(val), _p(x), _p(y))
    return y


def matavec(beg, jco, co, x):
    y = np.empty(len(beg) - 1)
    x = np.ascontiguousarray(x, dtype=np.float64)
    lib().oracle_matavec(len(beg) - 1, _p(beg), _p(jco), _p(co), _p(x), _p(y))
    return y


def kref_gmres(rowptr, col, val, b, x0, tol=1e-4, maxit=500, restart=400, prec_kind=0, minv=None, flexible=True, hist_cap=8192):
    """The reference's GMRESSolver::solve on a CSR operator.  Returns dict(rc, x, hist, iters, resid, n_matvec);
    hist[0] is the initial scaled residual, hist[k] the residual after inner iteration k."""
    n = len(b)
    x = np.array(x0, dtype=np.float64)
    hist = np.zeros(hist_cap); nh = C.c_int(); it = C.c_int(); fr = C.c_double(); nmv = C.c_long()
    # the reference's flexible path reads Z[j], which is only filled when preconditioning is on (GMRESSolver.H:160-164,
    # 237-238): 'no preconditioner' is therefore run as prec = identity (prec_kind 0), never with the flag off
    flags = 1 | (4 if flexible else 0)
    mv = np.ascontiguousarray(minv) if minv is not None else np.zeros(1)
    rc = klib().kref_gmres(n, _p(rowptr), _p(col), _p(val), prec_kind, 6, _p(mv), _p(np.ascontiguousarray(b)), _p(x), tol, maxit, restart,
                           flags, _p(hist), hist_cap, C.byref(nh), C.byref(it), C.byref(fr), C.byref(nmv))
    return dict(rc=rc, x=x, hist=hist[:nh.value].copy(), iters=it.value, resid=fr.value, n_matvec=nmv.value)


def kref_idrs(rowptr, col, val, b, x0, P_raw, tol=1e-8, maxit=500, s=4, prec_kind=0, minv=None, hist_cap=8192):
    n = len(b)
    x = np.array(x0, dtype=np.float64)
    hist = np.zeros(hist_cap); nh = C.c_int(); it = C.c_int(); fr = C.c_double(); nmv = C.c_long()
    mv = np.ascontiguousarray(minv) if minv is not None else np.zeros(1)
    P = np.ascontiguousarray(P_raw, dtype=np.float64)
    rc = klib().kref_idrs(n, _p(rowptr), _p(col), _p(val), prec_kind, 6, _p(mv), _p(np.ascontiguousarray(b)), _p(x), tol, maxit, s, _p(P),
                          _p(hist), hist_cap, C.byref(nh), C.byref(it), C.byref(fr), C.byref(nmv))
    return dict(rc=rc, x=x, hist=hist[:nh.value].copy(), iters=it.value, resid=fr.value, n_matvec=nmv.value)
