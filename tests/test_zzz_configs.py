"""BASELINE.json configs as parity cases at their FULL size where the oracle finishes in seconds: configs[1] = the synthetic
2-degree global grid 180x76x16 (Jacobian assembly + SpMV sweep).  configs[0] (4 degree, real mask) is `global4deg` in
test_emu_parity.py / test_gpu_parity.py, configs[2] the coupled-mode tests there, configs[3] (1 degree) the property test
`test_full_size_properties_1deg`; configs[4] (0.5 degree, 8 GPUs) is a bench configuration (`bench.py --grid 720 304 32`)."""
import os

import numpy as np
import pytest

import cases
from cases import PAR_INDEX as P

PARS = dict(cases.DEFAULT_PARS, NLES=1.0)


@pytest.fixture(scope="module")
def two_degree():
    from oracle.oracle import OracleTHCM
    s, landm = cases.global_synth(180, 76, 16)
    o = OracleTHCM(s, landm)
    for k, v in PARS.items():
        o.setpar(P[k], v)
    x = cases.random_state(s, landm, scale=0.1)
    B = o.rhs(x)
    val, missing = o.jacobian_graph(x)
    assert missing == 0
    return s, landm, o, x, B, val


def test_two_degree_device_functions_bit_exact(two_degree):
    """The library's device functions (compiled for the host) against the oracle on the full 2-degree grid."""
    from emu.emu import EmuTHCM
    s, landm, o, x, B, val = two_degree
    e = EmuTHCM(s, landm)
    for k, v in PARS.items():
        e.setpar(P[k], v)
    assert np.array_equal(e.rhs(x), B)
    assert np.array_equal(e.jacobian(x), val)
    assert e.check_tiles(x) == 0
    rp, col = e.graph()
    ro, co = o.graph()
    assert np.array_equal(rp, ro) and np.array_equal(col, co)
    assert 22_000_000 < len(col) <= 104 * s.N * s.M * s.L           # ~104 entries per cell before edge clipping


@pytest.mark.gpu
def test_two_degree_assembly_and_spmv_sweep(two_degree):
    """configs[1] on the B200 through the C ABI: residual and Jacobian bit-exact, a sweep of SpMVs within 1e-13 (2-norm)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.fail("no CUDA device: the THCM B200 path has no CPU fallback")
    import iemic_b200
    from oracle.oracle import spmv
    s, landm, o, x, B, val = two_degree
    t = iemic_b200.THCM(s, landm)
    for k, v in PARS.items():
        t.setParameter(k, v)
    xd = torch.from_numpy(x).cuda()
    F = t.new_vector()
    t.evaluate(xd, F, True)
    assert np.array_equal(F.cpu().numpy(), -B)
    assert np.array_equal(t.jacobian_values_host(), val)
    rowptr, col = o.graph()
    rng = np.random.default_rng(2)
    y = t.new_vector()
    for _ in range(4):
        v = rng.standard_normal(t.ndim)
        t.applyMatrix(torch.from_numpy(v).cuda(), y)
        yo = spmv(rowptr, col, val, v)
        assert np.linalg.norm(y.cpu().numpy() - yo) <= 1e-13 * np.linalg.norm(yo)
    t.close()


@pytest.mark.skipif(os.environ.get("THCM_SLOW_TESTS") != "1", reason="needs ~40 GB of RAM and ~3 min (the oracle's dense Al/An at 1 degree); set THCM_SLOW_TESTS=1")
@pytest.mark.parametrize("variant", ["plain", "mixing+coupled"])
def test_one_degree_device_functions_bit_exact(variant):
    """BASELINE configs[3] at FULL size (360 x 152 x 24, 7.88 M unknowns, 134.9 M graph entries): the library's device functions
    (compiled for the host) against the oracle, residual and Jacobian bit for bit -- plain, and with Mixing = 1 plus the coupled
    ocean block (configs[2]'s terms).  Last run: round 1 (r01h), all True."""
    from emu.emu import EmuTHCM
    from oracle.oracle import OracleTHCM
    flags = dict(vmix=1, coupled_T=1, coupled_S=1) if variant != "plain" else {}
    s, landm = cases.global_synth(360, 152, 24, **flags)
    o, e = OracleTHCM(s, landm), EmuTHCM(s, landm)
    for k, v in dict(PARS, SUNP=1.0).items():
        o.setpar(P[k], v)
        e.setpar(P[k], v)
    if flags:
        fields, atmos, seaice = cases.coupled_inputs(s)
        cases.apply_coupled(o, fields, atmos, seaice)
        cases.apply_coupled(e, fields, atmos, seaice)
    x = cases.random_state(s, landm, scale=0.1)
    assert np.array_equal(e.rhs(x), o.rhs(x))
    vo, missing = o.jacobian_graph(x)
    assert missing == 0 and len(vo) == 134866080
    assert np.array_equal(e.jacobian(x), vo)


def test_two_degree_device_functions_with_mixing_and_coupling():
    """The same 2-degree grid with Mixing = 2 (vmix_control decides the partition) and the coupled ocean block: bit-exact."""
    from emu.emu import EmuTHCM
    from oracle.oracle import OracleTHCM
    s, landm = cases.global_synth(180, 76, 16, vmix=2, coupled_T=1, coupled_S=1)
    o, e = OracleTHCM(s, landm), EmuTHCM(s, landm)
    for k, v in dict(PARS, NLES=0.0, SUNP=1.0).items():
        o.setpar(P[k], v)
        e.setpar(P[k], v)
    fields, atmos, seaice = cases.coupled_inputs(s)
    cases.apply_coupled(o, fields, atmos, seaice)
    cases.apply_coupled(e, fields, atmos, seaice)
    x = cases.random_state(s, landm, scale=0.1)
    e.vmix_control(x)
    assert np.array_equal(e.rhs(x), o.rhs(x))
    vo, missing = o.jacobian_graph(x)
    assert missing == 0 and np.array_equal(e.jacobian(x), vo)
    assert np.abs(o.vmix_fun(x)).max() > 0
