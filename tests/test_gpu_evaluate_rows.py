"""GPU tests of the parts of the path that sit around the core kernels (first run on a B200 in round 2: profiles/r02/pytest_gpu_r02a_ungated.log):

(1) The row replacements THCM::evaluate applies above the Fortran core (THCM.C:1013-1041, 1164-1172, 2180-2296) on the device
API: salinity integral condition (SRES = 0 -- the configuration of the reference's own test/ocean/ocean_params.xml) and the
pressure Dirichlet rows.  Checked against a numpy restatement built on the oracle's residual, Jacobian and
m_thcm_utils::intcond_scaling coefficients.
(2) set_landmask_ and init_ on an MPI sub-domain through the Fortran symbols.
(3) GMRES on the ocean-only (cell-compacted) Krylov space (the default) against the full-length solve; the Newton step through host
buffers against the device-resident one.
(4) The SpMV that does not stream the identity rows of LAND cells.
(5) The reference's own src/tests/test_ocean.C restated over the C++ mirror."""
import numpy as np
import pytest

import cases
from cases import PAR_INDEX as P

import os

pytestmark = [pytest.mark.gpu]
PARS = dict(cases.DEFAULT_PARS, NLES=1.0)


def reference_evaluate(o, x, sign, correction, rows_pfix):
    """THCM::evaluate on one rank: (F, dense J as a LinearOperator-like callable, cob, rowintcon)."""
    from oracle.oracle import spmv
    F = -o.rhs(x)
    val, _ = o.jacobian_graph(x)
    rowptr, col = o.graph()
    _, _, _, cob = o.matrix(x)
    cv, ci = o.intcond_scaling()
    coeff = np.zeros(o.ndim); coeff[ci - 1] = cv
    n, m, l = o.n, o.m, o.l
    rowic = 6 * (((l - 1) * m + (m - 1)) * n + (n - 1)) + 5
    F = F.copy(); cob = cob.copy()
    F[rowic] = sign * (coeff @ x - correction)
    cob[rowic] = 0.0
    val = val.copy()
    val[rowptr[rowic]:rowptr[rowic + 1]] = 0.0
    for r in rows_pfix:
        F[r] = 0.0; cob[r] = 0.0
        sl = slice(rowptr[r], rowptr[r + 1])
        val[sl] = np.where(col[sl] == r, 1.0, 0.0)

    def apply(v):
        y = spmv(rowptr, col, val, v)
        y[rowic] = sign * (coeff @ v)
        return y
    return F, apply, cob, rowic, coeff


@pytest.mark.parametrize("name,pfix", [("natl8", False), ("natl8", True), ("box_np", False)])
def test_integral_condition_and_pressure_rows(name, pfix):
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the THCM B200 path has no CPU fallback")
    import iemic_b200
    from oracle.oracle import OracleTHCM
    mk = {"natl8": cases.natl8, "box_np": lambda **kw: cases.box(6, 7, 4, False, seed=2, land_frac=0.0, **kw)}[name]
    s, landm = mk(SRES=0)
    o = OracleTHCM(s, landm)
    t = iemic_b200.THCM(s, landm)
    for k, v in PARS.items():
        o.setpar(P[k], v)
        t.setParameter(k, v)
    n, m, l = s.N, s.M, s.L
    sign = -1
    rowic = t.enableIntegralCondition(-1, -1, sign)
    rows_pfix = []
    if pfix:
        t.fixPressurePoints(True)
        rows_pfix = [6 * (((l - 1) * m + (m - 1)) * n + (n - 1 - q)) + 3 for q in range(2)]
    x0 = cases.consistent_state(s, landm, scale=0.05, seed=5)
    corr = t.setIntCondCorrection(torch.from_numpy(x0).cuda())
    x = cases.consistent_state(s, landm, scale=0.05, seed=6)
    Fo, apply, cob, rowic_o, coeff = reference_evaluate(o, x, sign, 0.0, rows_pfix)
    assert rowic == rowic_o
    assert abs(corr - coeff @ x0) <= 1e-13 * np.abs(coeff * x0).sum()
    Fo[rowic] = sign * (coeff @ x - corr)
    xd = torch.from_numpy(x).cuda()
    F = t.new_vector()
    t.evaluate(xd, F, True)
    Fg = F.cpu().numpy()
    other = np.ones(o.ndim, bool); other[rowic] = False
    assert np.array_equal(Fg[other], Fo[other])
    assert abs(Fg[rowic] - Fo[rowic]) <= 1e-12 * np.abs(coeff * x).sum()
    assert np.array_equal(t.getMassDiagonal(), cob)
    rng = np.random.default_rng(3)
    y = t.new_vector()
    for _ in range(3):
        v = rng.standard_normal(o.ndim)
        t.applyMatrix(torch.from_numpy(v).cuda(), y)
        yo = apply(v)
        assert np.linalg.norm(y.cpu().numpy() - yo) <= 1e-13 * np.linalg.norm(yo)
    # a parameter change keeps the replaced rows' mass entries at zero
    t.setParameter("COMB", 0.9)
    assert t.getMassDiagonal()[rowic] == 0.0
    t.close()


def test_integral_condition_needs_sres_zero_and_an_ocean_point():
    import ctypes as C
    import iemic_b200
    L = iemic_b200.load_library()
    s, landm = cases.natl8()          # SRES = 1
    t = iemic_b200.THCM(s, landm)
    assert L.thcmb_intcond_row(t.ctx) == -1
    t.close()


@pytest.mark.parametrize("name", ["natl8", "gateway16", "global4deg", "box_p33"])
def test_spmv_skipping_land_rows(name):
    """The SpMV answers y = x on the identity rows of LAND cells without streaming them: it must equal the full CSR product over the
    graph arrays (thcmb_csr_spmv_dev on the same device arrays streams every row) bit for bit -- 1.0 * x + 0.0 * ... = x."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the THCM B200 path has no CPU fallback")
    import iemic_b200
    mk = {"natl8": cases.natl8, "gateway16": cases.gateway16, "global4deg": cases.global4deg,
          "box_p33": lambda **kw: cases.box(33, 5, 3, True, seed=6, land_frac=0.2, **kw)}[name]
    s, landm = mk()
    x = cases.random_state(s, landm, scale=0.2)
    rng = np.random.default_rng(4)
    t = iemic_b200.THCM(s, landm)
    for k, v in PARS.items():
        t.setParameter(k, v)
    t.evaluate(torch.from_numpy(x).cuda(), None, True)
    rp, col = t.graph()
    rpd, cold = torch.from_numpy(rp).cuda(), torch.from_numpy(col).cuda()
    vald = torch.from_numpy(t.jacobian_values_host()).cuda()
    y, y_full = t.new_vector(), t.new_vector()
    for _ in range(3):
        v = torch.from_numpy(rng.standard_normal(t.ndim)).cuda()
        t.applyMatrix(v, y)
        t.csr_spmv(rpd, cold, vald, v, y_full)
        assert np.array_equal(y.cpu().numpy(), y_full.cpu().numpy())
    t.close()


@pytest.mark.parametrize("name", ["natl8", "gateway16"])
def test_set_landmask_through_the_fortran_symbol(name):
    """set_landmask_ (usrc.F90:353-418) on the device path: the static per-cell data, tile descriptors and forcing are rebuilt; the
    host part is verified on the CPU (tests/test_emu_parity.py::test_set_landmask_rebuilds_everything)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the THCM B200 path has no CPU fallback")
    import iemic_b200
    from oracle.oracle import OracleTHCM
    s, landm = {"natl8": cases.natl8, "gateway16": cases.gateway16}[name]()
    o = OracleTHCM(s, landm)
    f = iemic_b200.FortranABI()
    f.global_initialize(s)
    f.init(s, landm)
    for k, v in PARS.items():
        o.setpar(P[k], v)
        f.setparcs(k, v)
    n, m, l = s.N, s.M, s.L
    new = landm.copy()
    ocean = np.argwhere(new[1:l + 1, 1:m + 1, 1:n + 1] == 0)
    rng = np.random.default_rng(9)
    for k, j, i in ocean[rng.choice(len(ocean), size=max(3, len(ocean) // 20), replace=False)]:
        new[k + 1, j + 1, i + 1] = 1
    o.set_landmask(new, s.periodic, 1)
    f.set_landmask(new, s.periodic, 1)
    x = cases.random_state(s, o.landm(), scale=0.2, zero_on_land=False)
    assert np.array_equal(f.rhs(x), o.rhs(x))
    beg, jco, co, cob = f.matrix(x)
    bo, jo, cf, cobo = o.matrix(x)
    assert np.array_equal(beg, bo) and np.array_equal(jco, jo) and np.array_equal(co, cf) and np.array_equal(cob, cobo)
    assert np.array_equal(f.get_forcing(), o.forcing())
    f.finalize()


def test_fortran_symbols_on_a_sub_domain_of_an_mpi_run():
    """init_ with the bounds of ONE Decomp2D block while m_global holds the global domain (THCM.C:328-338, 566-611): the idealised
    forcing uses the global latitude bounds, the flux correction goes through the thcm_forcing_integral_ callback (here the library's
    one-rank default).  Host logic verified on the CPU: test_emu_parity.py::test_sub_domain_with_global_latitude_bounds."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the THCM B200 path has no CPU fallback")
    import copy
    import iemic_b200
    from oracle.oracle import OracleTHCM
    rad = np.pi / 180.0
    s, landm = cases.box(6, 5, 4, False, seed=8, land_frac=0.25, TRES=0, SRES=0)
    s.xmin, s.xmax, s.ymin, s.ymax = 300 * rad, 340 * rad, 22 * rad, 54 * rad
    sg = copy.copy(s)
    sg.ymin, sg.ymax = 10 * rad, 74 * rad
    s.ymin_glob, s.ymax_glob = sg.ymin, sg.ymax
    o = OracleTHCM(s, landm)
    f = iemic_b200.FortranABI()
    f.global_initialize(sg)          # m_global: the whole domain
    f.init(s, landm)                 # usrc init: this rank's block
    for k, v in dict(PARS, CMPR=0.3, FPER=0.2).items():
        o.setpar(P[k], v)
        f.setparcs(k, v)
    x = cases.random_state(s, landm, scale=0.1)
    assert np.array_equal(f.rhs(x), o.rhs(x))
    beg, jco, co, cob = f.matrix(x)
    bo, jo, cf, cobo = o.matrix(x)
    assert np.array_equal(beg, bo) and np.array_equal(jco, jo) and np.array_equal(co, cf)
    assert np.array_equal(f.get_forcing(), o.forcing())
    f.finalize()


def test_reference_ocean_tests_over_the_cpp_mirror():
    """tests/cpp/test_ocean_mirror.cpp: the reference's src/tests/test_ocean.C (Initialization, MassMat, ComputeJacobian + the integral
    condition of its SRES = 0 configuration) re-stated over include/thcm_model.hpp; every EXPECT must hold."""
    import os
    import subprocess
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the THCM B200 path has no CPU fallback")
    exe = os.path.join(cases.ROOT, "tests", "cpp", "_bin", "test_ocean_mirror")
    if not os.path.exists(exe):
        pytest.fail(exe + " is missing: build it with `make -C tests/cpp`")
    r = subprocess.run([exe, os.path.join(cases.MASKS, "mask_natl8")], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]
    assert "0 failed" in r.stdout


@pytest.mark.parametrize("name", ["natl8", "gateway16", "global4deg"])
@pytest.mark.parametrize("ortho", ["mgs", "dgks"])
def test_gmres_on_the_ocean_only_krylov_space(name, ortho, monkeypatch):
    """THCM_KRYLOV_COMPACT=1: the Krylov vectors hold the ocean cells only (LAND rows are identity rows and b vanishes there): same
    residual history (1e-10) and iteration count as the full-length solve, same solution, zero on LAND; a right-hand side that does
    not vanish on LAND falls back to the full space.  The reduction itself is validated on the CPU
    (tests/test_emu_parity.py::test_ocean_only_krylov_space_is_exact)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the THCM B200 path has no CPU fallback")
    import iemic_b200
    from oracle.oracle import OracleTHCM
    mk = {"natl8": cases.natl8, "gateway16": cases.gateway16, "global4deg": cases.global4deg}[name]
    s, landm = mk()
    o = OracleTHCM(s, landm)
    for k, v in PARS.items():
        o.setpar(P[k], v)
    x = cases.consistent_state(s, landm, scale=0.05)
    b = o.rhs(x)
    out = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("THCM_KRYLOV_COMPACT", mode)
        t = iemic_b200.THCM(s, landm)
        for k, v in PARS.items():
            t.setParameter(k, v)
        t.evaluate(torch.from_numpy(x).cuda(), None, True)
        t.buildPreconditioner(1)
        sol = t.new_vector()
        res, hist = t.gmres(torch.from_numpy(b).cuda(), sol, tol=1e-9, maxit=40, restart=40, ortho=ortho)
        out[mode] = (res.iters, np.array(hist), sol.cpu().numpy().copy())
        if mode == "1":     # a right-hand side with LAND entries must not use the compact space (and still give the full answer)
            b2 = b.copy(); b2[:] += 1.0e-3
            sol2 = t.new_vector()
            res2, hist2 = t.gmres(torch.from_numpy(b2).cuda(), sol2, tol=1e-9, maxit=10, restart=40, ortho=ortho)
            land = np.repeat((landm[1:-1, 1:-1, 1:-1] != 0).reshape(-1), 6)
            assert res2.iters > 0 and np.abs(sol2.cpu().numpy()[land]).max() > 0.0   # the full space was used (compact would leave zeros)
        t.close()
    (i0, h0, s0), (i1, h1, s1) = out["0"], out["1"]
    k = min(len(h0), len(h1))
    assert abs(i0 - i1) <= 1 and k > 5
    assert np.abs(h0[:k] - h1[:k]).max() <= 1e-10
    assert np.linalg.norm(s0 - s1) <= 1e-9 * np.linalg.norm(s0)
    land = np.repeat((landm[1:-1, 1:-1, 1:-1] != 0).reshape(-1), 6)
    assert np.all(s1[land] == 0.0)


@pytest.mark.parametrize("name", ["gateway16", "global4deg"])
@pytest.mark.parametrize("compact", ["0", "1"])
def test_newton_step_from_host_buffers_equals_the_device_resident_step(name, compact, monkeypatch):
    """thcmb_newton_step (H2D state, step, D2H update) must return the dx of thcmb_newton_step_dev bit for bit -- in particular on the
    ocean-only Krylov space, whose gather / scatter buffers must not alias the update (round-1 advisor finding)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the THCM B200 path has no CPU fallback")
    import iemic_b200
    monkeypatch.setenv("THCM_KRYLOV_COMPACT", compact)
    s, landm = {"gateway16": cases.gateway16, "global4deg": cases.global4deg}[name]()
    t = iemic_b200.THCM(s, landm)
    for k, v in PARS.items():
        t.setParameter(k, v)
    t.set_ortho("dgks")
    x = cases.consistent_state(s, landm, scale=0.05)
    dx_dev = t.new_vector()
    res_d, fn_d = t.newton_step_dev(torch.from_numpy(x).cuda(), dx_dev, tol=1e-8, maxit=24, restart=25, precon=1)
    dx_host = np.full(t.ndim, np.nan)
    res_h, fn_h = t.newton_step(x, dx_host, tol=1e-8, maxit=24, restart=25, precon=1)
    assert fn_d == fn_h and res_d.iters == res_h.iters and res_d.resid == res_h.resid
    assert np.array_equal(dx_host, dx_dev.cpu().numpy())
    assert np.linalg.norm(dx_host) > 0
    land = np.repeat((landm[1:-1, 1:-1, 1:-1] != 0).reshape(-1), 6)
    assert np.all(dx_host[land] == 0.0)
    t.close()


def test_ocean_coupling_blocks_through_the_model_mirror():
    """Ocean::getBlock(Atmosphere) / getBlock(SeaIce) (Ocean.C:1603-1810) through the Python mirror of the Model API, against finite
    differences of the CUDA residual with respect to the atmosphere temperature / the sea-ice mask at one surface point (the residual is
    affine in both; the full finite-difference check of every column runs on the CPU: tests/test_probe_host.py)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the THCM B200 path has no CPU fallback")
    import iemic_b200
    s, landm = cases.natl8(coupled_T=1, coupled_S=1)
    oc = iemic_b200.Ocean(s, landm)
    t = oc.thcm
    for k, v in dict(PARS, SUNP=1.0).items():
        oc.setPar(k, v)
    fields, atmos, seaice = cases.coupled_inputs(s)
    fields["msi"] = np.random.default_rng(12).random(fields["msi"].shape)
    names = {"tatm": "atmosphere_t", "qatm": "atmosphere_q", "albe": "atmosphere_a", "patm": "atmosphere_p", "qsa": "seaice_q",
             "msi": "seaice_m", "gsi": "seaice_g", "emip": "emip", "adapted_emip": "adapted_emip", "spert": "emip_pert"}
    for k, f in fields.items():
        t.insertSurfaceField(names[k], f)
    t.setAtmosphereParameters(atmos)
    t.setSeaIceParameters(seaice)
    n, m, l = s.N, s.M, s.L
    x = cases.random_state(s, landm, scale=0.2)
    oc.getState().copy_(torch.from_numpy(x))

    class Atmos:
        kind, da, pdist = "atmosphere", atmos[13], None
        @staticmethod
        def interface_row(i, j, XX):
            return 3 * (j * n + i) + XX - 1 if XX <= 3 else 3 * n * m

    class Ice:
        kind = "seaice"
        @staticmethod
        def interface_row(i, j, XX):
            return 4 * (j * n + i) + XX - 1

    import scipy.sparse as sp
    beg, jco, co = oc.getBlock(Atmos)
    A = sp.csr_matrix((co, jco, beg), shape=(t.ndim, 3 * n * m + 1)).tocsc()
    beg, jco, co = oc.getBlock(Ice)
    B = sp.csr_matrix((co, jco, beg), shape=(t.ndim, 4 * n * m + 4)).tocsc()
    j, i = np.argwhere(landm[l, 1:-1, 1:-1] == 0)[7]
    oc.computeRHS()
    F0 = oc.getRHS("C").cpu().numpy()
    for field, mat, col in (("tatm", A, 3 * (j * n + i)), ("msi", B, 4 * (j * n + i) + 2)):
        f = fields[field].copy(); f[j, i] += 1e-3
        t.insertSurfaceField(names[field], f); oc.setPar("COMB", PARS["COMB"])
        oc.computeRHS()
        want = (oc.getRHS("C").cpu().numpy() - F0) / 1e-3
        t.insertSurfaceField(names[field], fields[field]); oc.setPar("COMB", PARS["COMB"])
        got = mat[:, col].toarray().ravel()
        assert np.abs(got).max() > 0 and np.abs(got - want).max() <= 1e-9 * max(np.abs(mat).max(), 1.0), field
    t.close()


def test_transient_run_on_the_device_reproduces_the_reference_golden_norm():
    """The reference's own regression number for the ocean time stepper (src/tests/trns_ocean.C:63-64: || state || = 37.03750142 +- 1e-4,
    30 Newton steps after ten adaptive theta steps from rest; Mixing = 1, salinity integral condition) through the DEVICE path: the
    theta residual (theta_rhs_kernel over the CUDA residual), the theta Jacobian (theta_jac_kernel over the CUDA Jacobian values) and the
    integral-condition row of the ThetaOcean mirror, driven by the restated AdaptiveTransient / Newton loop (tests/transient_twin.py).
    Only the direct linear solve runs on the host (the solver is not what is pinned here)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the THCM B200 path has no CPU fallback")
    import scipy.sparse as sp
    import iemic_b200
    import transient_twin as tt
    s, landm = cases.natl8(**tt.SETTINGS)
    p = tt.TIMESTEPPER
    model = iemic_b200.ThetaOcean(s, landm, theta=p["theta"])
    t = model.thcm
    for k, v in tt.PARAMETERS.items():
        t.setParameter(k, v)
    rowic = t.enableIntegralCondition(-1, -1, -1)
    assert t.setIntCondCorrection(t.new_vector()) == 0.0                 # Ocean.C:145-148 at the zero state
    coeff, _ = t.getIntCondCoeff()
    rowptr, col = t.graph()
    nd = t.ndim
    solve = tt.NullSpaceSolver()
    dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()   # noqa: E731

    def theta_rhs(y):
        model.setState(dev(y))
        model.computeRHS()
        return model.getRHS("V").cpu().numpy()

    x = np.zeros(nd)
    dt, time_, steps, total = p["dt"], 0.0, 0, 0
    tmax = p["tmax_years"] / (737.2685 / 365.0)
    while time_ < tmax and steps < p["nsteps"]:
        model.setState(dev(x))
        model.initStep(dt)
        y = x.copy()
        Fx = theta_rhs(y)
        converged = False
        for k in range(p["max_newton"]):
            model.setState(dev(y))
            model.computeJacobian()                                       # J - M / (theta dt) on the device
            val = t.jacobian_values_host()
            val[rowptr[rowic]:rowptr[rowic + 1]] = 0.0
            A = sp.csr_matrix((val, col, rowptr), shape=(nd, nd)).tolil()
            A[rowic, :] = -1.0 * coeff                                    # the dense row the device applies inside its SpMV (THCM.C:2180-2229)
            dx = solve(A.tocsr(), Fx / (p["theta"] * dt))
            normdx = np.abs(dx).max()
            y = y - dx
            Fx = theta_rhs(y)
            if normdx < p["newton_tol"] and np.linalg.norm(Fx) < p["newton_tol"]:
                converged = True
                break
            if normdx > 1e2:
                break
        if not converged:
            assert dt > p["dt_min"]
            dt = max(dt / p["decrease"], p["dt_min"])
            continue
        steps += 1
        time_ += dt
        x = y
        if k < p["min_wanted"]:
            dt = min(dt * p["increase"], p["dt_max"])
        elif k > p["max_wanted"]:
            dt = max(dt / p["decrease"], p["dt_min"])
        total += k
    assert steps == 10 and total == tt.GOLDEN_NEWTON_STEPS
    assert abs(np.linalg.norm(x) - tt.GOLDEN_NORM) < 1e-7, np.linalg.norm(x)
    t.close()


def test_models_built_from_the_reference_parameter_lists(tmp_path):
    """THCM / Ocean from the reference's Teuchos XML lists (paramlist.py; THCM.C:186-795, Ocean.C:985-1012): the model the constructor
    builds from tests/golden/params/natl8_integral_condition.xml (mask by file name, Mixing = 1, SRES = 0 -> integral condition,
    starting parameters) evaluates to the same bits as the hand-built one, and the basin the default run's "Topography" = 1 builds
    reproduces the stored steady state of that run as a root (tests/test_oracle_pins.py)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the THCM B200 path has no CPU fallback")
    import iemic_b200
    from iemic_b200 import paramlist as pl
    here = os.path.dirname(os.path.abspath(__file__))
    os.symlink(os.path.join(here, "golden", "masks"), tmp_path / "mkmask")
    xml = os.path.join(here, "golden", "params", "natl8_integral_condition.xml")
    oc = iemic_b200.Ocean.from_parameter_list(xml, data_dir=tmp_path)
    assert oc.solver_params == dict(tol=1e-6, restart=120, maxit=360, precon=1)
    s, landm = cases.natl8(vmix=1, SRES=0)
    t = iemic_b200.THCM(s, landm)
    rowic = t.enableIntegralCondition(-1, -1, -1)
    start = pl.read_xml(xml)["THCM"]["Starting Parameters"]
    for k, v in start.items():
        t.setParameter(k, v)
    assert oc.thcm.L_.thcmb_intcond_row(oc.thcm.ctx) == rowic
    for k in ("Combined Forcing", "SPL1", "Rossby-Number", "Wind Forcing"):
        assert oc.getPar(k) == t.getParameter(k)
    x = cases.consistent_state(s, landm, scale=0.05, seed=6)
    oc.getState("V").copy_(torch.from_numpy(x).cuda())
    oc.computeRHS()
    oc.computeJacobian()
    F = t.new_vector()
    t.evaluate(torch.from_numpy(x).cuda(), F, True)
    assert np.array_equal(oc.getRHS("V").cpu().numpy(), F.cpu().numpy())
    assert np.array_equal(oc.thcm.jacobian_values_host(), t.jacobian_values_host())
    assert np.array_equal(oc.thcm.getMassDiagonal(), t.getMassDiagonal())
    # Ocean::applyMassMat (Ocean.C:1448-1457) on the device: out = diag(B) v, the integral-condition row included (B = 0 there)
    v = torch.from_numpy(np.random.default_rng(2).standard_normal(t.ndim)).cuda()
    out = torch.full_like(v, float("nan"))
    oc.applyMassMat(v, out)
    B = oc.thcm.getMassDiagonal()
    assert np.array_equal(out.cpu().numpy(), B * v.cpu().numpy()) and B[rowic] == 0.0 and (B != 0).any()
    with pytest.raises(ValueError):
        oc.applyMassMat(v, v)
    assert oc.npar() == 30 and oc.int2par(19) == "Combined Forcing" and oc.dof() == 6 and np.array_equal(oc.getLandMask(), landm)
    # the list the model reports (test_parameterlist.C:326-370): overrides kept, everything else at its default, and every starting
    # parameter with its current value instead of the NaN placeholder (THCM.C:781-792)
    cur, dflt = oc.thcm.getParameters(), pl.thcm_default_init_parameters()
    given = pl.read_xml(xml)["THCM"]
    for k, v in cur.items():
        if k != "Starting Parameters":
            assert v == given.get(k, dflt[k]), k
    sp = cur["Starting Parameters"]
    assert len(sp) == 30 and not any(np.isnan(v) for v in sp.values())
    assert all(sp[k] == oc.getPar(k) for k in sp) and sp["Combined Forcing"] == 0.25 and sp["Rossby-Number"] == oc.getPar("Rossby-Number") != 0.0
    upd = pl.ParameterList("THCM")
    upd.sublist("Starting Parameters").update({"Combined Forcing": 0.5, "Wind Forcing": float("nan")})
    wind = oc.getPar("Wind Forcing")
    oc.thcm.setParameters(upd)                                             # THCM::setParameters: NaN entries are skipped
    assert oc.getPar("Combined Forcing") == 0.5 and oc.getPar("Wind Forcing") == wind and sp["Combined Forcing"] == 0.5
    with pytest.raises(pl.InvalidParameter):
        oc.thcm.setParameters(pl.ParameterList("THCM", {"Mixing": 2}))    # only the starting parameters may change after construction
    t.close(); oc.thcm.close()
    # the default run: Topography = 1 basin, Forcing Type 2; its stored steady state is a root
    from test_oracle_pins import DEFAULT_RUN_STATE
    t2 = iemic_b200.THCM.from_parameter_list(os.path.join(here, "golden", "params", "basin16_topography1.xml"))
    xs = np.fromfile(DEFAULT_RUN_STATE)
    F2, F0 = t2.new_vector(), t2.new_vector()
    t2.evaluate(torch.from_numpy(xs).cuda(), F2, False)
    t2.evaluate(torch.zeros(t2.ndim, dtype=torch.float64, device="cuda"), F0, False)
    assert float(F2.norm()) < 1e-10 * float(F0.norm())
    t2.close()


def test_set_land_mask_on_the_model_mirror():
    """Ocean::setLandMask / THCM::setLandMask(mask, init = true) (THCM.C:1362-1392) on a handle: the instance follows the new global mask
    exactly like one created on it -- residual, Jacobian values, mass diagonal, forcing bit for bit -- with the integral condition kept on
    its cell; a mask that turns that cell into land is refused."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the THCM B200 path has no CPU fallback")
    import iemic_b200
    s, landm = cases.box(9, 7, 4, False, seed=2, land_frac=0.0, SRES=0, vmix=1)
    n, m, l = s.N, s.M, s.L
    new = landm.copy()
    new[1:l + 1, 2:4, 3] = 1                      # two full-depth land columns
    new[1:3, 5, 6] = 1                            # and a sea mount: the two deepest levels (k = 1 is the bottom; no ocean below land)
    oc = iemic_b200.Ocean(s, landm)
    oc.thcm.enableIntegralCondition(-1, -1, -1)
    ref = iemic_b200.THCM(s, new)
    ref.enableIntegralCondition(-1, -1, -1)
    for k, v in PARS.items():
        oc.setPar(k, v)
        ref.setParameter(k, v)
    x0 = cases.consistent_state(s, landm, scale=0.05, seed=6)
    oc.getState("V").copy_(torch.from_numpy(x0).cuda())
    oc.computeRHS(); oc.computeJacobian(); oc.buildPreconditioner()      # state of the old mask everywhere, then the change
    oc.setLandMask(new)
    assert np.array_equal(oc.getLandMask(), new)
    x = cases.consistent_state(s, new, scale=0.05, seed=7)
    oc.getState("V").copy_(torch.from_numpy(x).cuda())
    oc.computeRHS(); oc.computeJacobian()
    F = ref.new_vector()
    ref.evaluate(torch.from_numpy(x).cuda(), F, True)
    assert np.array_equal(oc.getRHS("V").cpu().numpy(), F.cpu().numpy())
    assert np.array_equal(oc.thcm.jacobian_values_host(), ref.jacobian_values_host())
    assert np.array_equal(oc.thcm.getMassDiagonal(), ref.getMassDiagonal())
    assert np.array_equal(oc.thcm.getForcing(), ref.getForcing())
    # the ocean-only Krylov space follows the mask as well: one preconditioned solve gives the same iterates on both
    b = torch.from_numpy(np.random.default_rng(1).standard_normal(ref.ndim) * (ref.getMassDiagonal() != 0)).cuda()
    oc.buildPreconditioner(); ref.buildPreconditioner(1)
    sol = ref.new_vector()
    r1, h1 = ref.gmres(b, sol, tol=1e-6, maxit=30, restart=30, prec=True, flexible=True)
    oc.solver_params.update(tol=1e-6, maxit=30, restart=30)
    oc.solve(b)
    assert np.array_equal(np.asarray(h1), np.asarray(oc.last_history))
    bad = new.copy()
    bad[1:l + 1, m, n] = 1                          # the integral-condition cell (N-1, M-1) becomes land
    with pytest.raises(ValueError, match="integral-condition cell"):
        oc.setLandMask(bad)
    ref.close(); oc.thcm.close()
