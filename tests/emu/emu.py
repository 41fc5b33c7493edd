"""ctypes front-end of tests/emu/libthcm_emu.so: the library's device functions compiled for the host (a unit-test
harness for `pytest -m "not gpu"`, see emu_cell.cpp).  Built on demand with g++."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(os.path.dirname(_HERE))
_CSRC = os.path.join(_ROOT, "i-emic_b200", "csrc")
_lib = None


def build(force=False):
    so = os.path.join(_HERE, "libthcm_emu.so")
    srcs = [os.path.join(_HERE, "emu_cell.cpp"), os.path.join(_CSRC, "thcm_host.cpp"), os.path.join(_CSRC, "thcm_probe.cpp")]
    deps = srcs + [os.path.join(_CSRC, f) for f in ("thcm_cell.cuh", "thcm_internal.h", "thcm_slots.h", "thcm_tanh.h")] + \
        [os.path.join(_ROOT, "include", "thcm_b200.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        cuda_inc = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"
        subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared", "-I" + cuda_inc, "-o", so] + srcs,
                       check=True, capture_output=True)
    return so


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        vp, i, d, ll = C.c_void_p, C.c_int, C.c_double, C.c_longlong
        for name, res, args in [("emu_create", vp, [vp, vp]), ("emu_destroy", None, [vp]), ("emu_set_par", None, [vp, i, d]),
                                ("emu_get_par", d, [vp, i]), ("emu_ndim", i, [vp]), ("emu_gnnz", ll, [vp]), ("emu_halo_size", i, [vp]),
                                ("emu_block", None, [vp, vp]), ("emu_graph", None, [vp, vp, vp]), ("emu_halo_gids", None, [vp, vp]),
                                ("emu_local_gids", None, [vp, vp]), ("emu_get_forcing", None, [vp, vp, i]), ("emu_get_cob", None, [vp, vp]),
                                ("emu_plan_sizes", None, [vp, vp, vp, vp]), ("emu_plan", None, [vp, vp, vp, vp]),
                                ("emu_jacobian", None, [vp, vp, vp, vp]), ("emu_crs", ll, [vp, vp, vp, vp, vp, vp]),
                                ("emu_rhs", None, [vp, vp, vp, vp]), ("emu_check_staging", ll, [vp, vp, vp]),
                                ("emu_check_tiles", ll, [vp, vp, vp]), ("emu_fast_tiles", ll, [vp, vp]),
                                ("emu_vmix_control", None, [vp, i, i]), ("emu_set_vmix_fix", None, [vp, i]),
                                ("emu_set_field", None, [vp, i, vp]), ("emu_set_atmos", None, [vp, vp]), ("emu_set_seaice", None, [vp, vp]),
                                ("emu_set_internal_forcing", None, [vp, vp, vp]), ("emu_probe_field", i, [vp, i, vp]),
                                ("emu_probe_suno", None, [vp, vp]), ("emu_compute_evap", None, [vp, vp, vp]),
                                ("emu_salflux", None, [vp, vp, vp, vp, vp, vp]), ("emu_temflux", None, [vp, vp, vp]),
                                ("emu_derivatives", None, [vp, vp, vp]), ("emu_salt_advection", None, [vp, vp, vp]),
                                ("emu_salt_diffusion", None, [vp, vp, vp]), ("emu_stochastic_forcing", None, [vp, vp, vp, vp]),
                                ("emu_getdeps", None, [vp, vp]), ("emu_loadbal", None, [vp, vp]),
                                ("emu_grid", None, [vp] * 9), ("emu_set_landmask", None, [vp, vp, i, i]), ("emu_setsres", None, [vp, i]),
                                ("emu_ocean_cells", i, [vp]), ("emu_cell_maps", None, [vp, vp, vp])]:
            fn = getattr(L, name)
            fn.restype, fn.argtypes = res, args
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class EmuTHCM:
    def __init__(self, settings, landm):
        self.L_ = lib()
        landm = np.ascontiguousarray(landm, dtype=np.int32)
        self.h = self.L_.emu_create(C.byref(settings), _p(landm))
        assert self.h
        self.ndim = self.L_.emu_ndim(self.h)
        self.nnz = self.L_.emu_gnnz(self.h)
        self.nhalo = self.L_.emu_halo_size(self.h)

    def __del__(self):
        if getattr(self, "h", None):
            self.L_.emu_destroy(self.h)
            self.h = None

    def setpar(self, idx, v):
        self.L_.emu_set_par(self.h, idx, float(v))

    def getpar(self, idx):
        return self.L_.emu_get_par(self.h, idx)

    FIELDS = ("taux", "tauy", "tatm", "emip", "spert", "adapted_emip", "qatm", "albe", "patm", "qsa", "msi", "gsi")

    def set_field(self, name, f):
        f = np.ascontiguousarray(f, dtype=np.float64).reshape(-1)
        self.L_.emu_set_field(self.h, self.FIELDS.index(name), _p(f))

    def set_atmos_parameters(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64); assert p.size == 18
        self.L_.emu_set_atmos(self.h, _p(p))

    # ---- host diagnostics of the B1 boundary (thcm_probe.cpp); nm = (M, N) of the single-block domain ----
    def _nm(self):
        b = self.block()
        return b["m0"], b["n0"]

    def set_internal_forcing(self, t, s_):
        t = np.ascontiguousarray(t, dtype=np.float64).reshape(-1); s_ = np.ascontiguousarray(s_, dtype=np.float64).reshape(-1)
        self.L_.emu_set_internal_forcing(self.h, _p(t), _p(s_))

    def probe_field(self, name):
        out = np.full(self._nm(), np.nan)
        ok = self.L_.emu_probe_field(self.h, self.FIELDS.index(name), _p(out))
        return out if ok else None

    def suno(self):
        out = np.empty(self._nm()); self.L_.emu_probe_suno(self.h, _p(out)); return out

    def compute_evap(self, un):
        un = np.ascontiguousarray(un, dtype=np.float64); out = np.empty(self._nm())
        self.L_.emu_compute_evap(self.h, _p(un), _p(out)); return out

    def salflux(self, un):
        un = np.ascontiguousarray(un, dtype=np.float64)
        sf, qa, qs = np.empty(self._nm()), np.empty(self._nm()), np.empty(self._nm()); corr = np.zeros(1)
        self.L_.emu_salflux(self.h, _p(un), _p(sf), _p(corr), _p(qa), _p(qs))
        return sf, corr[0], qa, qs

    def temflux(self, un):
        un = np.ascontiguousarray(un, dtype=np.float64)
        six = np.zeros((6,) + self._nm())
        self.L_.emu_temflux(self.h, _p(un), _p(six))
        return dict(zip(("totflux", "swflux", "shflux", "lhflux", "siflux", "simask"), six))

    def derivatives(self, un):
        un = np.ascontiguousarray(un, dtype=np.float64)
        four = np.zeros((4,) + self._nm())
        self.L_.emu_derivatives(self.h, _p(un), _p(four))
        return tuple(four)

    def salt_advection(self, un):
        un = np.ascontiguousarray(un, dtype=np.float64); out = np.zeros(self.ndim // 6)
        self.L_.emu_salt_advection(self.h, _p(un), _p(out)); return out

    def salt_diffusion(self, un):
        un = np.ascontiguousarray(un, dtype=np.float64); out = np.zeros(self.ndim // 6)
        self.L_.emu_salt_diffusion(self.h, _p(un), _p(out)); return out

    def stochastic_forcing(self):
        m, n = self._nm()
        beg = np.zeros(self.ndim + 1, dtype=np.int32); jco = np.zeros(n * m, dtype=np.int32); co = np.zeros(n * m)
        self.L_.emu_stochastic_forcing(self.h, _p(beg), _p(jco), _p(co))
        return beg, jco, co

    def set_landmask(self, landm, periodic, reinit=1):
        lm = np.ascontiguousarray(landm, dtype=np.int32)
        self.L_.emu_set_landmask(self.h, _p(lm), int(periodic), int(reinit))
        self.nnz = self.L_.emu_gnnz(self.h)

    def cell_maps(self):
        """(ocell[n_ocean], ccell[ncell]): ocean cells of the block in cell order and the inverse map (-1 = LAND)."""
        no = self.L_.emu_ocean_cells(self.h)
        ocell = np.zeros(no, dtype=np.int32); ccell = np.zeros(self.ndim // 6, dtype=np.int32)
        self.L_.emu_cell_maps(self.h, _p(ocell), _p(ccell))
        return ocell, ccell

    def setsres(self, sres):
        self.L_.emu_setsres(self.h, int(sres))

    def grid(self, N, M, L):
        a = dict(x=np.empty(N), xu=np.empty(N + 1), y=np.empty(M), yv=np.empty(M + 1), z=np.empty(L), zw=np.empty(L + 1),
                 dfzT=np.empty(L), dfzW=np.empty(L + 1))
        self.L_.emu_grid(self.h, *[_p(a[k]) for k in ("x", "xu", "y", "yv", "z", "zw", "dfzT", "dfzW")])
        return a

    def ocean_block_atmosphere(self, albed, pdist, colT, colQ, colA, colP):
        """Ocean::getBlock(Atmosphere) (Ocean.C:1603-1730): (beg, jco, co) over all ocean rows."""
        m, n = self._nm()
        cap = 6 * n * m
        beg = np.zeros(self.ndim + 1, dtype=np.int32); jco = np.zeros(cap, dtype=np.int32); co = np.zeros(cap)
        a = [np.ascontiguousarray(c, dtype=np.int32) for c in (colT, colQ, colA, colP)]
        pd = np.ascontiguousarray(pdist, dtype=np.float64)
        f = self.L_.emu_ocean_block_atmosphere
        f.restype = C.c_int; f.argtypes = [C.c_void_p, C.c_double] + [C.c_void_p] * 8
        nnz = f(self.h, float(albed), _p(pd), *[_p(x) for x in a], _p(beg), _p(jco), _p(co))
        return beg, jco[:nnz].copy(), co[:nnz].copy()

    def ocean_block_seaice(self, un, colQ, colM, colG):
        """Ocean::getBlock(SeaIce) (Ocean.C:1733-1810)."""
        m, n = self._nm()
        cap = 6 * n * m
        beg = np.zeros(self.ndim + 1, dtype=np.int32); jco = np.zeros(cap, dtype=np.int32); co = np.zeros(cap)
        a = [np.ascontiguousarray(c, dtype=np.int32) for c in (colQ, colM, colG)]
        un = np.ascontiguousarray(un, dtype=np.float64)
        f = self.L_.emu_ocean_block_seaice
        f.restype = C.c_int; f.argtypes = [C.c_void_p] * 8
        nnz = f(self.h, _p(un), *[_p(x) for x in a], _p(beg), _p(jco), _p(co))
        return beg, jco[:nnz].copy(), co[:nnz].copy()

    def getdeps(self):
        out = np.empty(7); self.L_.emu_getdeps(self.h, _p(out)); return out

    def loadbal_weights(self):
        out = np.empty(self._nm()); self.L_.emu_loadbal(self.h, _p(out)); return out

    def set_seaice_parameters(self, p):
        p = np.ascontiguousarray(p, dtype=np.float64); assert p.size == 7
        self.L_.emu_set_seaice(self.h, _p(p))

    @staticmethod
    def block_only(s, landm):
        """(i0, j0, n0, m0, npN, npM) of rank s.rank without building a model."""
        out = np.zeros(6, dtype=np.int32)
        lm = np.ascontiguousarray(landm, dtype=np.int32)
        L = lib()
        L.emu_block_only.restype = C.c_int
        L.emu_block_only.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        assert L.emu_block_only(C.byref(s), _p(lm), _p(out)) == 0
        return tuple(out.tolist())

    def block(self):
        out = np.zeros(9, dtype=np.int32)
        self.L_.emu_block(self.h, _p(out))
        return dict(zip(("i0", "j0", "n0", "m0", "npN", "npM", "pidN", "pidM", "hk"), out.tolist()))

    def graph(self):
        rowptr = np.empty(self.ndim + 1, dtype=np.int32); col = np.empty(self.nnz, dtype=np.int32)
        self.L_.emu_graph(self.h, _p(rowptr), _p(col))
        return rowptr, col

    def local_gids(self):
        g = np.empty(self.ndim, dtype=np.int32); self.L_.emu_local_gids(self.h, _p(g)); return g

    def halo_gids(self):
        g = np.empty(max(self.nhalo, 1), dtype=np.int32); self.L_.emu_halo_gids(self.h, _p(g)); return g[:self.nhalo]

    def forcing(self, masked=True):
        f = np.empty(self.ndim); self.L_.emu_get_forcing(self.h, _p(f), int(masked)); return f

    def cob(self):
        f = np.empty(self.ndim); self.L_.emu_get_cob(self.h, _p(f)); return f

    def plan(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.L_.emu_plan_sizes(self.h, C.byref(a), C.byref(b), C.byref(c))
        peers = np.zeros((max(a.value, 1), 5), dtype=np.int32); send = np.zeros(max(b.value, 1), dtype=np.int32); recv = np.zeros(max(c.value, 1), dtype=np.int32)
        self.L_.emu_plan(self.h, _p(peers), _p(send), _p(recv))
        return peers[:a.value], send[:b.value], recv[:c.value]

    def _halo(self, halo):
        return np.ascontiguousarray(halo, dtype=np.float64) if halo is not None else np.zeros(max(self.nhalo, 1))

    def jacobian(self, un, halo=None):
        un = np.ascontiguousarray(un, dtype=np.float64); h = self._halo(halo)
        val = np.empty(self.nnz); self.L_.emu_jacobian(self.h, _p(un), _p(h), _p(val)); return val

    def crs(self, un, halo=None):
        un = np.ascontiguousarray(un, dtype=np.float64); h = self._halo(halo)
        beg = np.empty(self.ndim + 1, dtype=np.int32); jco = np.empty(self.nnz, dtype=np.int32); co = np.empty(self.nnz)
        nnz = self.L_.emu_crs(self.h, _p(un), _p(h), _p(beg), _p(jco), _p(co))
        return beg, jco[:nnz].copy(), co[:nnz].copy()

    def check_staging(self, un, halo=None):
        un = np.ascontiguousarray(un, dtype=np.float64); h = self._halo(halo)
        return self.L_.emu_check_staging(self.h, _p(un), _p(h))

    def check_tiles(self, un, halo=None):
        un = np.ascontiguousarray(un, dtype=np.float64); h = self._halo(halo)
        return self.L_.emu_check_tiles(self.h, _p(un), _p(h))

    def fast_tiles(self):
        tot = C.c_longlong()
        return self.L_.emu_fast_tiles(self.h, C.byref(tot)), tot.value

    def vmix_control(self, un_global):
        """vmix_control with the GLOBAL field norms (what the library all-reduces), mix_imp.f:139-169."""
        t = np.sqrt(np.sum(np.asarray(un_global)[4::6] ** 2)) > 1.0e-12
        s = np.sqrt(np.sum(np.asarray(un_global)[5::6] ** 2)) > 1.0e-12
        self.L_.emu_vmix_control(self.h, int(t), int(s))

    def rhs(self, un, halo=None):
        un = np.ascontiguousarray(un, dtype=np.float64); h = self._halo(halo)
        B = np.empty(self.ndim); self.L_.emu_rhs(self.h, _p(un), _p(h), _p(B)); return B
