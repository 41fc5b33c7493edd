"""Host-side mirror of the reference's THCM / Ocean interface over the C ABI (include/thcm_b200.h).

* :class:`THCM`   -- src/ocean/THCM.{H,C}: ``evaluate(soln, rhs, computeJac)``, ``getJacobian``,
  ``setParameter/getParameter``, the static maximal graph, the domain decomposition queries.
* :class:`Ocean`  -- the ``Model`` API (src/utils/Model.H:54-117, src/ocean/Ocean.C:1070-1391):
  ``computeRHS``, ``computeJacobian``, ``applyMatrix``, ``solve``, ``getState/getRHS/getSolution``,
  ``setPar/getPar``.
* :class:`FortranABI` -- the gfortran-mangled symbols ``THCM.C`` binds (``init_``, ``rhs_``,
  ``matrix_``, ...) with host numpy buffers, exactly as the reference's C++ would call them.

PyTorch only provides device memory (``torch.float64`` CUDA tensors) and ``torch.distributed`` for
exchanging the NCCL id; every kernel is in ``libthcm_b200.so``.  No CPU fallback exists.
"""
import ctypes as C
import math
import os

import numpy as np

from .params import par_index

_HERE = os.path.dirname(os.path.abspath(__file__))


def lib_path():
    return os.path.join(_HERE, "libthcm_b200.so")


class Settings(C.Structure):
    """thcmb_settings (include/thcm_b200.h).  Angles in radians."""
    _fields_ = [("N", C.c_int), ("M", C.c_int), ("L", C.c_int),
                ("xmin", C.c_double), ("xmax", C.c_double), ("ymin", C.c_double), ("ymax", C.c_double),
                ("hdim", C.c_double), ("qz", C.c_double), ("periodic", C.c_int),
                ("ih", C.c_int), ("vmix", C.c_int), ("tap", C.c_int), ("rho_mixing", C.c_int), ("coriolis_on", C.c_int),
                ("TRES", C.c_int), ("SRES", C.c_int), ("iza", C.c_int), ("ite", C.c_int), ("its", C.c_int),
                ("coupled_T", C.c_int), ("coupled_S", C.c_int), ("forcing_type", C.c_int),
                ("alphaT", C.c_double), ("alphaS", C.c_double),
                ("rank", C.c_int), ("nranks", C.c_int), ("device", C.c_int),
                ("ymin_glob", C.c_double), ("ymax_glob", C.c_double), ("balance", C.c_int)]

    PI = 3.14159265358979323846  # src/trios/THCMdefs.H:17

    @classmethod
    def from_degrees(cls, n, m, l, xmin, xmax, ymin, ymax, periodic=False, hdim=4000.0, qz=1.0, **kw):
        """Bounds in degrees are converted like THCM.C:203-206 (value * PI_ / 180.0)."""
        s = cls()
        s.N, s.M, s.L = n, m, l
        s.xmin, s.xmax = xmin * cls.PI / 180.0, xmax * cls.PI / 180.0
        s.ymin, s.ymax = ymin * cls.PI / 180.0, ymax * cls.PI / 180.0
        s.hdim, s.qz, s.periodic = hdim, qz, int(periodic)
        s.ih, s.vmix, s.tap, s.rho_mixing, s.coriolis_on = 0, 0, 1, 0, 1
        s.TRES, s.SRES, s.iza, s.ite, s.its = 1, 1, 2, 1, 1
        s.coupled_T = s.coupled_S = s.forcing_type = 0
        s.alphaT, s.alphaS = 1.0e-04, 7.6e-04
        s.rank, s.nranks, s.device = 0, 1, 0
        for k, v in kw.items():
            if not hasattr(s, k):
                raise KeyError(k)
            setattr(s, k, v)
        return s

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


class KrylovResult(C.Structure):
    _fields_ = [("status", C.c_int), ("iters", C.c_int), ("resid", C.c_double), ("nhist", C.c_int),
                ("n_matvec", C.c_longlong)]


_lib = None


def last_error():
    """The message of the last argument / runtime error the handle API reported (thcmb_last_error)."""
    return load_library().thcmb_last_error().decode()


def load_library(path=None):
    """Loads libthcm_b200.so.  Fails loudly when it has not been built (there is no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = path or lib_path()
    if not os.path.exists(path):
        raise ImportError(f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`. "
                          "The THCM B200 path has no CPU fallback.")
    L = C.CDLL(path, mode=C.RTLD_GLOBAL)
    vp, ip, dp, ll, i, d = C.c_void_p, C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_longlong, C.c_int, C.c_double
    sp = C.POINTER(Settings)
    sig = {
        "thcmb_default_settings": (None, [sp]),
        "thcmb_create": (vp, [sp, vp]), "thcmb_destroy": (None, [vp]), "thcmb_last_error": (C.c_char_p, []),
        "thcmb_local_block": (None, [vp, ip, ip, ip, ip, ip, ip]), "thcmb_ndim_local": (i, [vp]),
        "thcmb_graph_nnz": (ll, [vp]), "thcmb_get_graph": (None, [vp, vp, vp]), "thcmb_halo_size": (i, [vp]),
        "thcmb_halo_gids": (None, [vp, vp]), "thcmb_local_gids": (None, [vp, vp]),
        "thcmb_set_par": (None, [vp, i, d]), "thcmb_get_par": (d, [vp, i]),
        "thcmb_get_forcing": (None, [vp, vp]), "thcmb_get_cob": (None, [vp, vp]),
        "thcmb_apply_mass_dev": (i, [vp, vp, vp]), "thcmb_theta_rhs_dev": (i, [vp, d, d, vp, vp, vp, vp]), "thcmb_theta_jacobian_dev": (i, [vp, d, d]),
        "thcmb_set_landmask": (i, [vp, vp, i]), "thcmb_enable_intcond": (None, [vp, i, i, i]), "thcmb_set_intcond_correction": (d, [vp, vp]),
        "thcmb_fix_pressure_points": (None, [vp, i]), "thcmb_intcond_row": (i, [vp]),
        "thcmb_insert_field": (None, [vp, i, vp]), "thcmb_set_atmos_parameters": (None, [vp, vp]),
        "thcmb_set_seaice_parameters": (None, [vp, vp]),
        "thcmb_nccl_unique_id": (i, [vp]), "thcmb_nccl_init": (i, [vp, vp]),
        "thcmb_p2p_local_handle": (i, [vp, vp]), "thcmb_p2p_open": (i, [vp, vp]), "thcmb_set_ortho": (None, [vp, i]),
        "thcmb_set_vmix_fix": (None, [vp, i]), "thcmb_get_vmix_flags": (None, [vp, vp]),
        "thcmb_recompute_scaling": (i, [vp, vp, vp, vp]), "thcmb_intcond_coeff": (d, [vp, vp]),
        "thcmb_halo_exchange": (i, [vp, vp]), "thcmb_residual_dev": (i, [vp, vp, vp]), "thcmb_rhs_dev": (i, [vp, vp, vp]),
        "thcmb_jacobian_dev": (i, [vp, vp]), "thcmb_jacobian_values": (vp, [vp]), "thcmb_graph_rowptr_dev": (vp, [vp]),
        "thcmb_graph_col_dev": (vp, [vp]), "thcmb_scatter_values_dev": (i, [vp, ll, vp, vp, vp]), "thcmb_fortran_context": (vp, []), "thcmb_jacobian_crs_dev": (ll, [vp, vp, vp, vp, vp]),
        "thcmb_tile_counts": (None, [vp, vp, vp]),
        "thcmb_ocean_block_atmosphere": (i, [vp, d, vp, vp, vp, vp, vp, vp, vp, vp]),
        "thcmb_ocean_block_seaice": (i, [vp, vp, vp, vp, vp, vp, vp, vp]), "thcmb_spmv_dev": (i, [vp, vp, vp]), "thcmb_csr_spmv_dev": (i, [vp, i, vp, vp, vp, vp, vp]),
        "thcmb_dot": (d, [vp, i, vp, vp]), "thcmb_nrm2": (d, [vp, i, vp]), "thcmb_axpby": (i, [vp, i, d, vp, d, vp]),
        "thcmb_scale": (i, [vp, i, d, vp]), "thcmb_build_precon": (i, [vp, i]), "thcmb_apply_precon_dev": (i, [vp, vp, vp]),
        "thcmb_gmres": (i, [vp, vp, vp, d, i, i, i, vp, i, C.POINTER(KrylovResult)]),
        "thcmb_idrs": (i, [vp, vp, vp, d, i, i, vp, vp, i, C.POINTER(KrylovResult)]),
        "thcmb_newton_step": (i, [vp, vp, vp, d, i, i, i, dp, C.POINTER(KrylovResult)]),
        "thcmb_newton_step_dev": (i, [vp, vp, vp, d, i, i, i, dp, C.POINTER(KrylovResult)]),
        "thcmb_profile": (None, [vp, i]), "thcmb_kernel_count": (i, []), "thcmb_kernel_name": (C.c_char_p, [i]),
        "thcmb_profile_report": (i, [vp, i, ip, dp]),
        "thcmb_device_alloc": (vp, [vp, ll]), "thcmb_device_free": (None, [vp, vp]),
        "thcmb_h2d": (i, [vp, vp, vp, ll]), "thcmb_d2h": (i, [vp, vp, vp, ll]), "thcmb_sync": (i, [vp]),
        "thcmb_stream": (vp, [vp]), "thcmb_launch_count": (ll, [vp]), "thcmb_last_stage_ms": (d, [vp, C.c_char_p]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L


def lib():
    return load_library()


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _dev_ptr(t):
    """Device pointer of a torch CUDA float64/int32 tensor (or a raw int)."""
    if isinstance(t, int):
        return C.c_void_p(t)
    if not t.is_cuda:
        raise ValueError("expected a CUDA tensor: the THCM B200 path has no CPU fallback")
    if not t.is_contiguous():
        raise ValueError("expected a contiguous tensor")
    return C.c_void_p(t.data_ptr())


class THCM:
    """Mirror of the reference's THCM class (src/ocean/THCM.H) on one rank of the 2-D decomposition."""

    def __init__(self, settings, landm, comm=None):
        """settings: Settings; landm: int32[L+2, M+2, N+2] GLOBAL mask (the array THCM.C:389 reads from
        m_global::get_landm).  comm: None or an initialised torch.distributed process group; when
        settings.nranks > 1 it is used once to broadcast the NCCL unique id."""
        import torch
        self._torch = torch
        self.L_ = load_library()
        self.settings = settings
        landm = np.ascontiguousarray(landm, dtype=np.int32)
        assert landm.shape == (settings.L + 2, settings.M + 2, settings.N + 2), landm.shape
        self.landm_global = landm
        self.ctx = self.L_.thcmb_create(C.byref(settings), _np_ptr(landm))
        if not self.ctx:
            raise RuntimeError("thcmb_create failed: " + self.L_.thcmb_last_error().decode())
        self.device = torch.device("cuda", settings.device)
        self.ndim = self.L_.thcmb_ndim_local(self.ctx)
        self.nnz = self.L_.thcmb_graph_nnz(self.ctx)
        self.stream = torch.cuda.ExternalStream(self.L_.thcmb_stream(self.ctx), device=self.device)
        if settings.nranks > 1:
            self._init_nccl(comm)

    @classmethod
    def from_parameter_list(cls, params, comm=None, rank=0, nranks=1, device=0, balance=0, data_dir=None):
        """THCM::THCM(Teuchos::ParameterList&, comm) (THCM.C:180-795): `params` is the THCM list of the reference's ocean_params.xml (a
        paramlist.ParameterList, or the path of an XML file whose root or "THCM" sublist it is).  Land mask from "Land Mask" / "Topography",
        integral condition when "Restoring Salinity Profile" is 0, "Fix Pressure Points", then the "Starting Parameters"."""
        from . import paramlist as pl
        if isinstance(params, (str, os.PathLike)):
            params = pl.read_xml(params)
        if "THCM" in params and isinstance(params["THCM"], pl.ParameterList):
            params = params["THCM"]
        su = pl.thcm_setup(params, rank=rank, nranks=nranks, device=device, balance=balance, data_dir=data_dir)
        t = cls(su["settings"], su["landm"], comm)
        t.paramList_ = su["params"]
        t.scalingType_ = su["scaling"]
        if su["spert"] is not None:
            t.insertSurfaceField("emip_pert", su["spert"])        # spert of init_ (THCM.C:566-611); used from the next parameter change on
        if su["integral_condition"] is not None:
            t.enableIntegralCondition(*su["integral_condition"])
        if su["fix_pressure_points"]:
            t.fixPressurePoints(True)
        start = t.paramList_["Starting Parameters"]
        for name in list(start):                                   # THCM.C:781-792: NaN = "keep stpnt's value", which the list then reports
            if isinstance(start[name], float) and math.isnan(start[name]):
                start[name] = t.getParameter(name)
            else:
                t.setParameter(name, start[name])
        return t

    def getParameters(self):
        """THCM::getParameters (THCM.C:2789-2790): the validated list the model was built from, every starting parameter with its value."""
        if not hasattr(self, "paramList_"):
            raise RuntimeError("this THCM was not built from a parameter list")
        return self.paramList_

    def setParameters(self, new_params):
        """THCM::setParameters (THCM.C:2792-2801): only the "Starting Parameters" may change after construction; NaN entries are skipped."""
        from . import paramlist as pl
        pl.validate_parameters(new_params, pl.thcm_default_parameters())
        for name, value in new_params.get("Starting Parameters", {}).items():
            if not (isinstance(value, float) and math.isnan(value)):
                self.setParameter(name, value)
                if hasattr(self, "paramList_"):
                    self.paramList_["Starting Parameters"][name] = float(value)

    def _init_nccl(self, comm):
        import torch.distributed as dist
        idbuf = np.zeros(128, dtype=np.uint8)
        if self.settings.rank == 0:
            rc = self.L_.thcmb_nccl_unique_id(_np_ptr(idbuf))
            if rc != 0:
                raise RuntimeError("ncclGetUniqueId failed")
        obj = [idbuf.tobytes() if self.settings.rank == 0 else None]
        dist.broadcast_object_list(obj, src=0, group=comm)
        idbuf = np.frombuffer(obj[0], dtype=np.uint8).copy()
        rc = self.L_.thcmb_nccl_init(self.ctx, _np_ptr(idbuf))
        if rc != 0:
            raise RuntimeError("thcmb_nccl_init failed: " + self.L_.thcmb_last_error().decode())
        # fused reduction + all-reduce over NVLink peer memory (CUDA IPC mailboxes); THCM_P2P=0 keeps plain NCCL
        self.p2p = False
        if os.environ.get("THCM_P2P", "1") != "0":
            h = np.zeros(64, dtype=np.uint8)
            ok = self.L_.thcmb_p2p_local_handle(self.ctx, _np_ptr(h)) == 0
            allh = [None] * self.settings.nranks
            dist.all_gather_object(allh, (ok, h.tobytes()), group=comm)
            if all(o for o, _ in allh):
                buf = np.frombuffer(b"".join(b for _, b in allh), dtype=np.uint8).copy()
                ok = self.L_.thcmb_p2p_open(self.ctx, _np_ptr(buf)) == 0
            flags = [None] * self.settings.nranks
            dist.all_gather_object(flags, bool(ok), group=comm)
            if not all(flags):
                raise RuntimeError("P2P mailbox setup failed on some rank (set THCM_P2P=0 to use NCCL all-reduce): "
                                   + self.L_.thcmb_last_error().decode())
            self.p2p = True

    def close(self):
        if getattr(self, "ctx", None):
            self.L_.thcmb_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- decomposition / graph (TRIOS_Domain.H:114-247, THCM.C:2300-2580) ----
    def local_block(self):
        v = [C.c_int() for _ in range(6)]
        self.L_.thcmb_local_block(self.ctx, *[C.byref(x) for x in v])
        return dict(zip(("i0", "j0", "n0", "m0", "npN", "npM"), [x.value for x in v]))

    def graph(self):
        rowptr = np.empty(self.ndim + 1, dtype=np.int32)
        col = np.empty(self.nnz, dtype=np.int32)
        self.L_.thcmb_get_graph(self.ctx, _np_ptr(rowptr), _np_ptr(col))
        return rowptr, col

    def local_gids(self):
        g = np.empty(self.ndim, dtype=np.int32)
        self.L_.thcmb_local_gids(self.ctx, _np_ptr(g))
        return g

    def halo_gids(self):
        g = np.empty(self.L_.thcmb_halo_size(self.ctx), dtype=np.int32)
        self.L_.thcmb_halo_gids(self.ctx, _np_ptr(g))
        return g

    # ---- parameters (THCM.C:1945-1969) ----
    def setParameter(self, name, value):
        self.L_.thcmb_set_par(self.ctx, par_index(name), float(value))

    def getParameter(self, name):
        return self.L_.thcmb_get_par(self.ctx, par_index(name))

    def getForcing(self):
        f = np.empty(self.ndim)
        self.L_.thcmb_get_forcing(self.ctx, _np_ptr(f))
        return f

    # ---- coupled mode (coupled_T / coupled_S = 1): what Ocean::synchronize feeds THCM (Ocean.C:1451-1560, THCM.C:1336-1568)
    SURFACE_FIELDS = ("taux", "tauy", "atmosphere_t", "emip", "emip_pert", "adapted_emip", "atmosphere_q", "atmosphere_a",
                      "atmosphere_p", "seaice_q", "seaice_m", "seaice_g")

    def insertSurfaceField(self, name, field):
        """m_inserts::insert_<name> (inserts.F90) with the GLOBAL [M, N] field; takes effect at the next parameter change."""
        f = np.ascontiguousarray(field, dtype=np.float64).reshape(-1)
        if f.size != self.settings.N * self.settings.M:
            raise ValueError("surface fields are global N*M arrays")
        self.L_.thcmb_insert_field(self.ctx, self.SURFACE_FIELDS.index(name), _np_ptr(f))

    def setAtmosphereParameters(self, pars18):
        """set_atmos_parameters (usrc.F90:254-310): the 18 doubles of Atmosphere::CommPars; re-runs forcing + lin."""
        p = np.ascontiguousarray(pars18, dtype=np.float64)
        if p.size != 18:
            raise ValueError("Atmosphere::CommPars has 18 members")
        self.L_.thcmb_set_atmos_parameters(self.ctx, _np_ptr(p))

    def setSeaIceParameters(self, pars7):
        """set_seaice_parameters (usrc.F90:313-350): the 7 doubles of SeaIce::CommPars; re-runs forcing + lin."""
        p = np.ascontiguousarray(pars7, dtype=np.float64)
        if p.size != 7:
            raise ValueError("SeaIce::CommPars has 7 members")
        self.L_.thcmb_set_seaice_parameters(self.ctx, _np_ptr(p))

    def oceanBlockAtmosphere(self, albed, pdist, colT, colQ, colA, colP):
        """Ocean::getBlock(Atmosphere) (Ocean.C:1603-1730): d F_ocean / d x_atmosphere as (beg, jco, co), CRS over all ocean rows."""
        n, m = self.settings.N, self.settings.M
        cols = [np.ascontiguousarray(c, dtype=np.int32).reshape(-1) for c in (colT, colQ, colA, colP)]
        pd = None if pdist is None else np.ascontiguousarray(pdist, dtype=np.float64).reshape(-1)
        beg = np.zeros(self.ndim + 1, dtype=np.int32); jco = np.zeros(6 * n * m, dtype=np.int32); co = np.zeros(6 * n * m)
        f = self.L_.thcmb_ocean_block_atmosphere
        nnz = f(self.ctx, float(albed), None if pd is None else _np_ptr(pd), *[_np_ptr(c) for c in cols], _np_ptr(beg), _np_ptr(jco), _np_ptr(co))
        return beg, jco[:nnz].copy(), co[:nnz].copy()

    def oceanBlockSeaIce(self, state_host, colQ, colM, colG):
        """Ocean::getBlock(SeaIce) (Ocean.C:1733-1810) at the host state (THCM::getDerivatives)."""
        n, m = self.settings.N, self.settings.M
        cols = [np.ascontiguousarray(c, dtype=np.int32).reshape(-1) for c in (colQ, colM, colG)]
        un = np.ascontiguousarray(state_host, dtype=np.float64)
        beg = np.zeros(self.ndim + 1, dtype=np.int32); jco = np.zeros(6 * n * m, dtype=np.int32); co = np.zeros(6 * n * m)
        f = self.L_.thcmb_ocean_block_seaice
        nnz = f(self.ctx, _np_ptr(un), *[_np_ptr(c) for c in cols], _np_ptr(beg), _np_ptr(jco), _np_ptr(co))
        return beg, jco[:nnz].copy(), co[:nnz].copy()

    def getMassDiagonal(self):
        """coB of fillcolB (assemble.F90:18-54), THCM::evaluateB."""
        f = np.empty(self.ndim)
        self.L_.thcmb_get_cob(self.ctx, _np_ptr(f))
        return f

    # ---- hot path ----
    def new_vector(self):
        return self._torch.zeros(self.ndim, dtype=self._torch.float64, device=self.device)

    def _pre(self):
        # the library runs on its own stream: order it after work queued on torch's current stream
        self._torch.cuda.current_stream(self.device).synchronize()

    def sync(self):
        self.L_.thcmb_sync(self.ctx)

    def evaluate(self, soln, rhs=None, computeJac=False):
        """THCM::evaluate (THCM.C:957-1199): rhs <- F(soln) = A(u)u + mix - Frc (C++ sign), Jacobian values
        into the static graph.  soln / rhs: CUDA float64 tensors of the owned unknowns."""
        self._pre()
        if rhs is not None:
            self.L_.thcmb_residual_dev(self.ctx, _dev_ptr(soln), _dev_ptr(rhs))
        if computeJac:
            self.L_.thcmb_jacobian_dev(self.ctx, _dev_ptr(soln))
        self.sync()
        return True

    def rhs_fortran_sign(self, soln, out):
        self._pre()
        self.L_.thcmb_rhs_dev(self.ctx, _dev_ptr(soln), _dev_ptr(out))
        self.sync()

    def getJacobian(self):
        """(rowptr, col, values) of the local Jacobian: numpy graph + a CUDA tensor copy of the values."""
        torch = self._torch
        rowptr, col = self.graph()
        val = torch.from_numpy(self.jacobian_values_host()).to(self.device)
        return rowptr, col, val

    def jacobian_values_host(self):
        tmp = np.empty(self.nnz)
        self.L_.thcmb_d2h(self.ctx, _np_ptr(tmp), C.c_void_p(self.L_.thcmb_jacobian_values(self.ctx)), self.nnz * 8)
        return tmp

    def jacobian_crs(self, soln):
        """Fortran-order thresholded CRS (begA, jcoA, coA), 1-based, as matrix_ fills it (assemble.F90:57-139)."""
        torch = self._torch
        self._pre()
        beg = torch.empty(self.ndim + 1, dtype=torch.int32, device=self.device)
        jco = torch.empty(self.nnz, dtype=torch.int32, device=self.device)
        co = torch.empty(self.nnz, dtype=torch.float64, device=self.device)
        torch.cuda.current_stream(self.device).synchronize()
        nnz = self.L_.thcmb_jacobian_crs_dev(self.ctx, _dev_ptr(soln), _dev_ptr(beg), _dev_ptr(jco), _dev_ptr(co))
        return beg, jco[:nnz], co[:nnz]

    def applyMatrix(self, v, out):
        self._pre()
        self.L_.thcmb_spmv_dev(self.ctx, _dev_ptr(v), _dev_ptr(out))
        self.sync()

    def dot(self, x, y):
        self._pre()
        return self.L_.thcmb_dot(self.ctx, x.numel(), _dev_ptr(x), _dev_ptr(y))

    def norm(self, x):
        self._pre()
        return self.L_.thcmb_nrm2(self.ctx, x.numel(), _dev_ptr(x))

    def update(self, y, a, x, b):
        """y = a*x + b*y (Vector::update of GMRESSolverDecl.H:12-14)."""
        self._pre()
        self.L_.thcmb_axpby(self.ctx, y.numel(), float(a), _dev_ptr(x), float(b), _dev_ptr(y))
        self.sync()

    def buildPreconditioner(self, kind=1):
        self.L_.thcmb_build_precon(self.ctx, int(kind))
        self.sync()

    def applyPrecon(self, v, out):
        self._pre()
        self.L_.thcmb_apply_precon_dev(self.ctx, _dev_ptr(v), _dev_ptr(out))
        self.sync()

    def csr_spmv(self, rowptr, col, val, x, y):
        """y = A x for a CSR matrix in device memory (int32 rowptr / col, float64 val): the plain kernel that streams every row."""
        self._pre()
        self.L_.thcmb_csr_spmv_dev(self.ctx, rowptr.numel() - 1, _dev_ptr(rowptr), _dev_ptr(col), _dev_ptr(val), _dev_ptr(x), _dev_ptr(y))
        self.sync()

    def gmres(self, b, x, tol=1e-4, maxit=500, restart=400, prec=True, flexible=True, hist_cap=4096, ortho="mgs", full_space=False):
        """GMRESSolver::solve (src/gmressolver/GMRESSolver.H:81-255).  Returns (KrylovResult, history).  By default the Krylov vectors
        hold the ocean cells only when b and x vanish on LAND (identity rows); full_space=True keeps full-length vectors."""
        self._pre()
        hist = np.zeros(hist_cap)
        res = KrylovResult()
        flags = (1 if prec else 0) | (4 if flexible else 0) | (8 if ortho == "dgks" else 0) | (16 if full_space else 0)
        self.L_.thcmb_gmres(self.ctx, _dev_ptr(b), _dev_ptr(x), tol, maxit, restart, flags, _np_ptr(hist), hist_cap, C.byref(res))
        return res, hist[:res.nhist].copy()

    def idrs(self, b, x, P_raw, tol=1e-8, maxit=500, s=4, hist_cap=4096):
        """IDRSolver::solve (src/idrsolver/IDRSolver.H:109-340); P_raw: (s, ndim) CUDA tensor."""
        self._pre()
        hist = np.zeros(hist_cap)
        res = KrylovResult()
        self.L_.thcmb_idrs(self.ctx, _dev_ptr(b), _dev_ptr(x), tol, maxit, s, _dev_ptr(P_raw), _np_ptr(hist), hist_cap, C.byref(res))
        return res, hist[:res.nhist].copy()

    def newton_step(self, un_host, dx_host, tol=1e-4, maxit=500, restart=400, precon=1):
        """End-to-end Newton step from pinned host buffers (torch CPU tensors or numpy arrays)."""
        res = KrylovResult()
        fn = C.c_double()
        up = un_host.data_ptr() if hasattr(un_host, "data_ptr") else un_host.ctypes.data
        dp = dx_host.data_ptr() if hasattr(dx_host, "data_ptr") else dx_host.ctypes.data
        self.L_.thcmb_newton_step(self.ctx, C.c_void_p(up), C.c_void_p(dp), tol, maxit, restart, precon, C.byref(fn), C.byref(res))
        return res, fn.value

    def recomputeScaling(self):
        """THCM::RecomputeScaling (THCM.C:1781-1834) on the stored Jacobian: (rowScaling, colScaling, averaged 6x6 block)."""
        self._pre()
        rs, cs, db = np.empty(self.ndim), np.empty(self.ndim), np.empty(36)
        rc = self.L_.thcmb_recompute_scaling(self.ctx, _np_ptr(rs), _np_ptr(cs), _np_ptr(db))
        return rs, cs, db.reshape(6, 6).T.copy(), rc == 0

    def getIntCondCoeff(self):
        """THCM::getIntCondCoeff (THCM.C:2608-2637): (coefficient vector on the owned rows, local volume)."""
        coeff = np.empty(self.ndim)
        vol = self.L_.thcmb_intcond_coeff(self.ctx, _np_ptr(coeff))
        return coeff, vol

    def enableIntegralCondition(self, Nic=-1, Mic=-1, sign=-1):
        """The salinity integral condition of THCM::evaluate (SRES = 0; THCM.C:653-697, 1013-1026, 2180-2256); returns the global row."""
        self.L_.thcmb_enable_intcond(self.ctx, int(Nic), int(Mic), int(sign))
        return self.L_.thcmb_intcond_row(self.ctx)

    def setIntCondCorrection(self, vec):
        """THCM::setIntCondCorrection (THCM.C:2078-2097)."""
        self._pre()
        return self.L_.thcmb_set_intcond_correction(self.ctx, _dev_ptr(vec))

    def setLandMask(self, landm, init=True):
        """THCM::setLandMask(global mask, init) (THCM.C:1362-1392): with init the instance follows the new GLOBAL mask [l+2, m+2, n+2]
        (set_landmask_ with reinit = 1, usrc.F90:353-418); a preconditioner built before must be rebuilt."""
        lm = np.ascontiguousarray(landm, dtype=np.int32)
        assert lm.shape == self.landm_global.shape, lm.shape
        if self.L_.thcmb_set_landmask(self.ctx, _np_ptr(lm), 1 if init else 0) != 0:
            raise ValueError(last_error())
        self.landm_global = lm

    def fixPressurePoints(self, on=True):
        """"Fix Pressure Points" (THCM.C:749-757, 2258-2296)."""
        self.L_.thcmb_fix_pressure_points(self.ctx, 1 if on else 0)

    def set_vmix_fix(self, fix):
        """m_mix::set_vmix_fix (THCM.C:2639-2647): 0 lets the next rhs / matrix call re-decide the Mixing = 2 partition."""
        self.L_.thcmb_set_vmix_fix(self.ctx, int(fix))

    def vmix_flags(self):
        out = np.zeros(4, dtype=np.int32)
        self.L_.thcmb_get_vmix_flags(self.ctx, _np_ptr(out))
        return dict(zip(("flag", "temp", "salt", "fix"), out.tolist()))

    def set_ortho(self, mode):
        """Orthogonalisation of the Newton-step GMRES: 'mgs' (GMRESSolver.H) or 'dgks' (batched, Belos-style)."""
        self.L_.thcmb_set_ortho(self.ctx, 1 if mode == "dgks" else 0)

    def newton_step_dev(self, un, dx, tol=1e-4, maxit=500, restart=400, precon=1):
        """Newton step with the state already in HBM (CUDA tensors)."""
        self._pre()
        res = KrylovResult()
        fn = C.c_double()
        self.L_.thcmb_newton_step_dev(self.ctx, _dev_ptr(un), _dev_ptr(dx), tol, maxit, restart, precon, C.byref(fn), C.byref(res))
        self.sync()
        return res, fn.value

    def profile(self, on):
        self.L_.thcmb_profile(self.ctx, int(on))

    def profile_report(self):
        """{kernel name: (launches, total device ms)} since profile(True)."""
        out = {}
        for kid in range(self.L_.thcmb_kernel_count()):
            n, ms = C.c_int(), C.c_double()
            self.L_.thcmb_profile_report(self.ctx, kid, C.byref(n), C.byref(ms))
            if n.value:
                out[self.L_.thcmb_kernel_name(kid).decode()] = (n.value, ms.value)
        return out

    def tile_counts(self):
        """(tiles of 32 cells in this rank's block, tiles the Jacobian kernels revisit on every assembly = not all-LAND)."""
        a, b = C.c_int(), C.c_int()
        self.L_.thcmb_tile_counts(self.ctx, C.byref(a), C.byref(b))
        return a.value, b.value

    def launch_count(self):
        return self.L_.thcmb_launch_count(self.ctx)


class Ocean:
    """The Model API of the reference (src/utils/Model.H:54-117) as implemented by Ocean (Ocean.C)."""

    def __init__(self, settings, landm, comm=None, solver_params=None, thcm=None):
        self.thcm = thcm if thcm is not None else THCM(settings, landm, comm)
        t = self.thcm
        self.state_ = t.new_vector()
        self.rhs_ = t.new_vector()
        self.sol_ = t.new_vector()
        sp = dict(tol=1e-4, maxit=500, restart=400, precon=1)  # run/ocean/solver_params.xml: FGMRES 1e-4, 500 its
        sp.update(solver_params or {})
        self.solver_params = sp
        self.jac_valid = False
        self.precon_valid = False

    @classmethod
    def from_parameter_list(cls, params, comm=None, **kw):
        """Ocean::Ocean(comm, Teuchos::ParameterList&) (Ocean.C:71-200): `params` is the reference's ocean list (ocean_params.xml; a
        paramlist.ParameterList or a path) with its "THCM" sublist and, as the reference's drivers merge it in, the "Belos Solver" sublist
        (solver_params.xml).  kw: rank, nranks, device, balance, data_dir (THCM.from_parameter_list)."""
        from . import paramlist as pl
        if isinstance(params, (str, os.PathLike)):
            params = pl.read_xml(params)
        params = pl.validate_parameters_and_set_defaults(params, pl.ocean_default_init_parameters())
        t = THCM.from_parameter_list(params["THCM"], comm=comm, **kw)
        o = cls(t.settings, None, comm, solver_params=pl.solver_parameters(params["Belos Solver"]), thcm=t)
        o.params_ = params
        return o

    def getState(self, mode="V"):
        return self.state_ if mode == "V" else self.state_.clone()

    def getRHS(self, mode="V"):
        return self.rhs_ if mode == "V" else self.rhs_.clone()

    def getSolution(self, mode="V"):
        return self.sol_ if mode == "V" else self.sol_.clone()

    def setPar(self, name, value):
        self.thcm.setParameter(name, value)
        self.jac_valid = False

    def getPar(self, name):
        return self.thcm.getParameter(name)

    def computeRHS(self):      # Ocean.C:1277-1295
        self.thcm.evaluate(self.state_, self.rhs_, False)

    def computeJacobian(self):  # Ocean.C:1297-1309
        self.thcm.evaluate(self.state_, None, True)
        self.jac_valid, self.precon_valid = True, False

    def computeMassMat(self):
        return self.thcm.getMassDiagonal()

    def applyMatrix(self, v, out):  # Ocean.C:1369-1374
        self.thcm.applyMatrix(v, out)

    def applyMassMat(self, v, out):  # Ocean.C:1448-1457: out = diag(B) v
        t = self.thcm
        t._pre()
        if t.L_.thcmb_apply_mass_dev(t.ctx, _dev_ptr(v), _dev_ptr(out)) != 0:
            raise ValueError(last_error())
        t.sync()

    # ---- the small queries of the Model API (Model.H:82-100, Ocean.H) ----
    def npar(self):
        return 30                      # _NPAR_ (THCMdefs.H)

    def int2par(self, ind):
        from .params import PAR_NAMES
        return PAR_NAMES[ind]

    def name(self):
        return "ocean"

    def dof(self):
        return 6                       # _NUN_

    def getLandMask(self):
        """The GLOBAL land mask [l+2, m+2, n+2] the model works on (Ocean::getLandMask's `global_borders`)."""
        return self.thcm.landm_global.copy()

    def setLandMask(self, landm, init=True):   # Ocean::setLandMask (Ocean.C) -> THCM::setLandMask
        self.thcm.setLandMask(landm, init)
        self.jac_valid = self.precon_valid = False

    def buildPreconditioner(self):  # Ocean.C:1377-1391
        if not self.precon_valid:
            self.thcm.buildPreconditioner(self.solver_params["precon"])
            self.precon_valid = True

    def applyPrecon(self, v, out):
        self.buildPreconditioner()
        self.thcm.applyPrecon(v, out)

    def getBlock(self, other):
        """Ocean::getBlock(std::shared_ptr<Atmosphere>) / getBlock(std::shared_ptr<SeaIce>) (Ocean.C:1603-1810): the coupling block of the
        coupled model's Jacobian, (beg, jco, co).  `other` stands for the other model: .kind "atmosphere" | "seaice" and
        .interface_row(i, j, XX) (1-based unknown XX, 0-based row, -1 = none); an atmosphere also gives .da (CommPars::da) and .pdist."""
        s = self.thcm.settings
        ii, jj = np.meshgrid(np.arange(s.N), np.arange(s.M))
        rows = lambda XX: np.array([other.interface_row(int(i), int(j), XX) for i, j in zip(ii.ravel(), jj.ravel())], dtype=np.int32)  # noqa: E731
        if other.kind == "atmosphere":     # ATMOS_TT_ 1, ATMOS_QQ_ 2, ATMOS_AA_ 3, ATMOS_PP_ 4 (AtmosphereDefinitions.H:27-33)
            return self.thcm.oceanBlockAtmosphere(other.da, getattr(other, "pdist", None), rows(1), rows(2), rows(3), rows(4))
        if other.kind == "seaice":         # SEAICE_QQ_ 2, SEAICE_MM_ 3, SEAICE_GG_ 5 (SeaIceDefinitions.H:19-24)
            return self.thcm.oceanBlockSeaIce(self.state_.cpu().numpy(), rows(2), rows(3), rows(5))
        raise ValueError("getBlock: unknown model kind " + str(other.kind))

    def solve(self, rhs=None):  # Ocean.C:1070-1147: sol_ = J^{-1} rhs with preconditioned FGMRES, zero initial guess
        b = self.rhs_ if rhs is None else rhs
        self.buildPreconditioner()
        self.sol_.zero_()
        sp = self.solver_params
        res, hist = self.thcm.gmres(b, self.sol_, tol=sp["tol"], maxit=sp["maxit"], restart=sp["restart"],
                                    prec=sp["precon"] != 0, flexible=True)
        self.last_solve = res
        self.last_history = hist
        return res.status


class ThetaOcean(Ocean):
    """src/transient/ThetaModel.H:18-165 over the Ocean mirror: the implicit theta step
    M u_n + dt theta F(u_{n+1}) + dt (1 - theta) F(u_n) - M u_{n+1} = 0 with J_theta = J - M / (theta dt)."""

    def __init__(self, settings, landm, theta=1.0, comm=None, solver_params=None):
        super().__init__(settings, landm, comm, solver_params)
        if theta < 0 or theta > 1:
            raise ValueError(f"ThetaModel: incorrect theta {theta}")   # ThetaModel.H:93-97
        self.theta_ = float(theta)
        self.timestep_ = 1.0e-3
        self.oldState_ = self.thcm.new_vector()
        self.oldRhs_ = self.thcm.new_vector()

    def initStep(self, timestep):   # ThetaModel.H:65-75
        self.timestep_ = float(timestep)
        self.oldState_.copy_(self.state_)
        Ocean.computeRHS(self)
        self.oldRhs_.copy_(self.rhs_)

    def setState(self, state):      # ThetaModel.H:77-81
        if state is not self.state_:
            self.state_.copy_(state)

    def computeRHS(self):           # ThetaModel.H:87-113
        Ocean.computeRHS(self)
        t = self.thcm
        t._pre()
        t.L_.thcmb_theta_rhs_dev(t.ctx, self.theta_, self.timestep_, _dev_ptr(self.state_), _dev_ptr(self.oldState_),
                                 _dev_ptr(self.oldRhs_), _dev_ptr(self.rhs_))
        t.sync()

    def computeJacobian(self):      # ThetaModel.H:118-149
        Ocean.computeJacobian(self)
        t = self.thcm
        t.L_.thcmb_theta_jacobian_dev(t.ctx, self.theta_, self.timestep_)

    def solve(self, rhs=None):      # ThetaModel.H:153-165: J_theta x = b / (theta dt)
        if self.theta_ == 0.0:
            raise ValueError("theta = 0 divides by the mass matrix, which is singular for THCM (w and p rows)")
        b = (self.rhs_ if rhs is None else rhs).clone()
        t = self.thcm
        t._pre()
        t.L_.thcmb_scale(t.ctx, t.ndim, 1.0 / self.timestep_ / self.theta_, _dev_ptr(b))
        return Ocean.solve(self, b)


class FortranABI:
    """The gfortran symbols of the B1 boundary (THCM.C:49-176) bound with host numpy buffers."""
    _owner = None

    def __init__(self):
        L = load_library()
        self.L_ = L
        ip, dp, vp = C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_void_p
        L.init_.restype = None
        L.init_.argtypes = [ip] * 4 + [dp] * 6 + [ip] * 6 + [vp] + [vp] * 5
        L.rhs_.restype = None; L.rhs_.argtypes = [vp, vp]
        L.matrix_.restype = None; L.matrix_.argtypes = [vp]
        L.fillcolb_.restype = None; L.fillcolb_.argtypes = []
        L.finalize_.restype = None; L.finalize_.argtypes = []
        L.setparcs_.restype = None; L.setparcs_.argtypes = [ip, dp]
        L.getparcs_.restype = None; L.getparcs_.argtypes = [ip, dp]
        L.get_forcing_.restype = None; L.get_forcing_.argtypes = [vp]
        L.set_landmask_.restype = None; L.set_landmask_.argtypes = [vp, ip, ip]
        # module procedures: fetched with getattr (a literal L.__m_... inside a class body would be name-mangled by Python)
        self._get_array_sizes = getattr(L, "__m_mat_MOD_get_array_sizes")
        self._get_array_sizes.restype = None; self._get_array_sizes.argtypes = [ip, ip]
        self._set_pointers = getattr(L, "__m_mat_MOD_set_pointers")
        self._set_pointers.restype = None; self._set_pointers.argtypes = [ip, ip] + [vp] * 7
        self._global_initialize = getattr(L, "__m_global_MOD_initialize")
        self._global_initialize.restype = None
        self._global_initialize.argtypes = [ip] * 3 + [dp] * 6 + [ip] * 13 + [C.c_char_p] * 5

    def global_initialize(self, s, maskfile=b"", itopo=1, flat=False, spertmaskfile=b""):
        """m_global::initialize (global.F90:65-160, THCM.C:325-337).  maskfile = "Land Mask" with "Read Land Mask" (a path, or a name below
        $THCM_DATA_DIR/mkmask); without one the next global_get_landm builds the idealised "Topography" case itopo (topo.F90:129-330).
        spertmaskfile = "Salinity Perturbation Mask" with "Read Salinity Perturbation Mask"."""
        i, d = C.c_int, C.c_double
        self._gdims = (s.N, s.M, s.L)
        a = [i(s.N), i(s.M), i(s.L), d(s.xmin), d(s.xmax), d(s.ymin), d(s.ymax), d(s.hdim), d(s.qz), i(s.periodic), i(itopo), i(int(flat)),
             i(1 if maskfile else 0), i(s.TRES), i(s.SRES), i(s.iza), i(s.ite), i(s.its), i(1 if spertmaskfile else 0), i(s.coupled_T),
             i(s.coupled_S), i(s.forcing_type)]
        self._global_initialize(*[C.byref(x) for x in a], maskfile, spertmaskfile, b"", b"", b"")

    def global_get_landm(self):
        """m_global::get_landm (global.F90:299-318): runs topofit -- the mask file or the Topography case -- and returns the GLOBAL
        mask [l+2, m+2, n+2] THCM.C distributes over the sub-domains (THCM.C:372-400)."""
        n, m, l = self._gdims
        out = np.empty((l + 2, m + 2, n + 2), dtype=np.int32)
        self._call("__m_global_MOD_get_landm", _np_ptr(out))
        return out

    def global_get_spert(self):
        """m_global::get_spert (global.F90:587-608): the salinity perturbation mask [m, n] THCM.C hands to init_."""
        n, m, _ = self._gdims
        out = np.empty((m, n))
        self._call("__m_global_MOD_get_spert", _np_ptr(out))
        return out

    def init(self, s, landm):
        """init_ for a single-rank domain (usrc.F90:6-139), followed by get_array_sizes / set_pointers like THCM.C:619-638."""
        i, d = C.c_int, C.c_double
        n, m, l = s.N, s.M, s.L
        self.n, self.m, self.l = n, m, l
        landm = np.ascontiguousarray(landm, dtype=np.int32)
        z = np.zeros(n * m)
        a = [i(n), i(m), i(l), i(n * m * l), d(s.xmin), d(s.xmax), d(s.ymin), d(s.ymax), d(s.alphaT), d(s.alphaS), i(s.ih), i(s.vmix),
             i(s.tap), i(s.rho_mixing), i(s.coriolis_on), i(s.periodic)]
        self.L_.init_(*[C.byref(x) for x in a], _np_ptr(landm), _np_ptr(z), _np_ptr(z), _np_ptr(z), _np_ptr(z), _np_ptr(z))
        FortranABI._owner = id(self)   # the library holds ONE instance per process (THCM.H:76-84): the last init_ owns it
        nrows, nnz = i(), i()
        self._get_array_sizes(C.byref(nrows), C.byref(nnz))
        self.ndim = nrows.value
        # the reference allocates ndim*(6*27+1) entries (mat.F90:56-68); rows never hold more than 24
        cap = self.ndim * 24 + 1
        self.begA = np.zeros(self.ndim + 1, dtype=np.int32)
        self.jcoA = np.zeros(cap, dtype=np.int32)
        self.coA = np.zeros(cap)
        self.coB = np.zeros(self.ndim)
        self.begF = np.zeros(self.ndim + 1, dtype=np.int32); self.jcoF = np.zeros(self.ndim, dtype=np.int32); self.coF = np.zeros(self.ndim)
        capi = i(cap)
        self._set_pointers(C.byref(nrows), C.byref(capi), _np_ptr(self.begA), _np_ptr(self.jcoA), _np_ptr(self.coA),
                                         _np_ptr(self.coB), _np_ptr(self.begF), _np_ptr(self.jcoF), _np_ptr(self.coF))

    def set_landmask(self, landm, periodic, reinit=1):
        """set_landmask_ (usrc.F90:353-418; THCM.C:1357)."""
        lm = np.ascontiguousarray(landm, dtype=np.int32)
        assert lm.shape == (self.l + 2, self.m + 2, self.n + 2)
        self.L_.set_landmask_(_np_ptr(lm), C.byref(C.c_int(int(periodic))), C.byref(C.c_int(int(reinit))))

    def setparcs(self, idx, val):
        self.L_.setparcs_(C.byref(C.c_int(par_index(idx))), C.byref(C.c_double(val)))

    def getparcs(self, idx):
        v = C.c_double()
        self.L_.getparcs_(C.byref(C.c_int(par_index(idx))), C.byref(v))
        return v.value

    def rhs(self, un):
        un = np.ascontiguousarray(un, dtype=np.float64)
        B = np.empty_like(un)
        self.L_.rhs_(_np_ptr(un), _np_ptr(B))
        return B

    def rhs_inplace(self, un):
        """rhs_ into a reused caller-owned buffer (what THCM.C:1001 does every call): returns that buffer."""
        if getattr(self, "_B", None) is None or self._B.shape != un.shape:
            self._B = np.empty_like(un)
        self.L_.rhs_(_np_ptr(un), _np_ptr(self._B))
        return self._B

    def matrix_inplace(self, un):
        """matrix_ into the buffers handed to set_pointers (THCM.C:1066), no copies on the Python side: returns the entry count."""
        self.L_.matrix_(_np_ptr(un))
        return int(self.begA[self.ndim] - 1)

    def matrix(self, un):
        un = np.ascontiguousarray(un, dtype=np.float64)
        self.L_.matrix_(_np_ptr(un))
        nnz = self.begA[self.ndim] - 1
        return self.begA.copy(), self.jcoA[:nnz].copy(), self.coA[:nnz].copy(), self.coB.copy()

    def get_forcing(self):
        f = np.empty(self.ndim)
        self.L_.get_forcing_(_np_ptr(f))
        return f

    def insert(self, name, field):
        """m_inserts::insert_<name> (THCM.C:85-98): local n*m surface field."""
        f = np.ascontiguousarray(field, dtype=np.float64).reshape(-1)
        assert f.size == self.n * self.m
        fn = getattr(self.L_, "__m_inserts_MOD_insert_" + name)
        fn.restype = None; fn.argtypes = [C.c_void_p]
        fn(_np_ptr(f))

    def set_atmos_parameters(self, pars18):
        p = np.ascontiguousarray(pars18, dtype=np.float64); assert p.size == 18
        self.L_.set_atmos_parameters_.restype = None; self.L_.set_atmos_parameters_.argtypes = [C.c_void_p]
        self.L_.set_atmos_parameters_(_np_ptr(p))

    def set_seaice_parameters(self, pars7):
        p = np.ascontiguousarray(pars7, dtype=np.float64); assert p.size == 7
        self.L_.set_seaice_parameters_.restype = None; self.L_.set_seaice_parameters_.argtypes = [C.c_void_p]
        self.L_.set_seaice_parameters_(_np_ptr(p))

    # ---- diagnostics / setup symbols (thcm_probe.cpp; THCM.C:136-170, Ocean.C:42-49) ----
    def _call(self, name, *args):
        fn = getattr(self.L_, name)
        fn.restype = None; fn.argtypes = [C.c_void_p] * len(args)
        fn(*args)

    def probe(self, name):
        """m_probe::get_<name> for atmosphere_t/q/p, emip, adapted_emip, emip_pert, taux, tauy, suno: [m, n]."""
        out = np.full((self.m, self.n), np.nan)
        self._call("__m_probe_MOD_get_" + name, _np_ptr(out))
        return out

    def compute_evap(self, un):
        un = np.ascontiguousarray(un, dtype=np.float64); out = np.empty((self.m, self.n))
        self._call("__m_probe_MOD_compute_evap", _np_ptr(out), _np_ptr(un))
        return out

    def get_salflux(self, un):
        un = np.ascontiguousarray(un, dtype=np.float64)
        sf, qa, qs = (np.empty((self.m, self.n)) for _ in range(3)); corr = np.zeros(1)
        self._call("__m_probe_MOD_get_salflux", _np_ptr(un), _np_ptr(sf), _np_ptr(corr), _np_ptr(qa), _np_ptr(qs))
        return sf, corr[0], qa, qs

    def get_temflux(self, un):
        un = np.ascontiguousarray(un, dtype=np.float64)
        f = [np.zeros((self.m, self.n)) for _ in range(6)]
        self._call("__m_probe_MOD_get_temflux", _np_ptr(un), *[_np_ptr(a) for a in f])
        return dict(zip(("totflux", "swflux", "shflux", "lhflux", "siflux", "simask"), f))

    def get_derivatives(self, un):
        un = np.ascontiguousarray(un, dtype=np.float64)
        f = [np.zeros((self.m, self.n)) for _ in range(4)]
        self._call("__m_probe_MOD_get_derivatives", _np_ptr(un), *[_np_ptr(a) for a in f])
        return tuple(f)

    def salt_integrals(self, un):
        """m_integrals::salt_advection / salt_diffusion (THCM.C:2133, 2155): the two per-cell integrands."""
        un = np.ascontiguousarray(un, dtype=np.float64)
        a, d = np.zeros(self.ndim // 6), np.zeros(self.ndim // 6)
        self._call("__m_integrals_MOD_salt_advection", _np_ptr(un), _np_ptr(a))
        self._call("__m_integrals_MOD_salt_diffusion", _np_ptr(un), _np_ptr(d))
        return a, d

    def get_stochastic_forcing(self):
        self.L_.get_stochastic_forcing_.restype = None; self.L_.get_stochastic_forcing_.argtypes = []
        self.L_.get_stochastic_forcing_()
        return self.begF.copy(), self.jcoF.copy(), self.coF.copy()

    def set_internal_forcing(self, temp, salt):
        t = np.ascontiguousarray(temp, dtype=np.float64).reshape(-1); s_ = np.ascontiguousarray(salt, dtype=np.float64).reshape(-1)
        self._call("__m_usr_MOD_set_internal_forcing", _np_ptr(t), _np_ptr(s_))

    def getdeps(self):
        v = [C.c_double() for _ in range(7)]
        fn = self.L_.getdeps_; fn.restype = None; fn.argtypes = [C.POINTER(C.c_double)] * 7
        fn(*[C.byref(x) for x in v])
        return np.array([x.value for x in v])

    def get_landm(self):
        out = np.empty((self.l + 2, self.m + 2, self.n + 2), dtype=np.int32)
        self._call("__m_thcm_utils_MOD_get_landm", _np_ptr(out))
        return out

    def average_block(self):
        """m_scaling::average_block on the Jacobian of the last matrix_ call (THCM.C:1798); (6,6) [row, col]."""
        db = np.zeros(36)
        getattr(self.L_, "__m_scaling_MOD_average_block")(_np_ptr(db))
        return db.reshape(6, 6).T.copy()

    def compute_scaling(self, db):
        """m_scaling::compute (THCM.C:1807)."""
        dbf = np.ascontiguousarray(np.asarray(db, dtype=np.float64).T).reshape(-1)
        rs, cs = np.empty(self.ndim), np.empty(self.ndim)
        getattr(self.L_, "__m_scaling_MOD_compute")(_np_ptr(dbf), _np_ptr(rs), _np_ptr(cs))
        return rs, cs

    def intcond_scaling(self):
        """m_thcm_utils::intcond_scaling (THCM.C:2619)."""
        val = np.empty(self.ndim // 6); ind = np.empty(self.ndim // 6, dtype=np.int32); n = C.c_int()
        getattr(self.L_, "__m_thcm_utils_MOD_intcond_scaling")(_np_ptr(val), _np_ptr(ind), C.byref(n))
        return val[:n.value].copy(), ind[:n.value].copy()

    def finalize(self):
        self.L_.finalize_()
        FortranABI._owner = None

    def __del__(self):   # the library page-locks the CRS arrays of set_pointers until finalize_: never let them go while registered
        if FortranABI._owner == id(self):
            try:
                self.finalize()
            except Exception:
                pass
