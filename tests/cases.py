"""Shared test cases: grids, masks, parameters and seeded states.

Every case is a (Settings, landm, params) triple that can be handed unchanged to the oracle
(oracle.oracle.OracleTHCM), to the CPU emulation of the device functions (tests/emu) and to the CUDA
library (iemic_b200.THCM / FortranABI)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import iemic_b200  # noqa: E402
from iemic_b200 import Settings, read_mask, synthetic_global_mask, all_ocean_mask, PAR_INDEX  # noqa: E402

MASKS = os.path.join(ROOT, "tests", "golden", "masks")
SEED = 20261017

# parameter block of test/ocean/ocean_params.xml ("Starting Parameters") + SURVEY 8d
DEFAULT_PARS = {"COMB": 1.0, "WIND": 1.0, "TEMP": 10.0, "SALT": 1.0}


def natl8(**kw):      # test/ocean/ocean_params.xml: 8x8x4, North Atlantic box, non-periodic
    s = Settings.from_degrees(8, 8, 4, 286, 350, 10, 74, periodic=False, hdim=4000.0, qz=1.0, **kw)
    return s, read_mask(os.path.join(MASKS, "mask_natl8"), 8, 8, 4)


def test6x6x4(**kw):
    s = Settings.from_degrees(6, 6, 4, 286, 350, 10, 74, periodic=False, hdim=4000.0, qz=1.0, **kw)
    return s, read_mask(os.path.join(MASKS, "test6x6x4"), 6, 6, 4)


def gateway16(**kw):  # test/ocean/reft_ocean_params.xml: 16x16x16 periodic, mask_gateway
    s = Settings.from_degrees(16, 16, 16, 0, 360, -60, 60, periodic=True, hdim=4000.0, qz=1.0, **kw)
    return s, read_mask(os.path.join(MASKS, "mask_gateway"), 16, 16, 16)


def global4deg(**kw):  # run/ocean/global/ocean_params.xml:21-42 with data/mkmask/mask_global_96x38x12
    s = Settings.from_degrees(96, 38, 12, 0, 359.99, -85.5, 85.5, periodic=True, hdim=5000.0, qz=2.25, **kw)
    return s, read_mask(os.path.join(MASKS, "mask_global_96x38x12"), 96, 38, 12)


def global_synth(n, m, l, **kw):  # 2, 1, 0.5 degree synthetic grids
    base = read_mask(os.path.join(MASKS, "mask_global_96x38x12"), 96, 38, 12)
    s = Settings.from_degrees(n, m, l, 0, 359.99, -85.5, 85.5, periodic=True, hdim=5000.0, qz=2.25, **kw)
    return s, synthetic_global_mask(base, n, m, l, periodic=True)


def box(n, m, l, periodic, seed=0, land_frac=0.0, **kw):
    """Random-topography box: stresses every branch of `boundaries` (isolated cells, steps, seams)."""
    xmax = 359.99 if periodic else 350
    s = Settings.from_degrees(n, m, l, 0 if periodic else 286, xmax, -60 if periodic else 10, 60 if periodic else 74,
                              periodic=periodic, hdim=4000.0, qz=1.8 if seed % 2 else 1.0, **kw)
    landm = all_ocean_mask(n, m, l, periodic=False)
    if land_frac > 0:
        rng = np.random.default_rng(1000 + seed)
        depth = np.where(rng.random((m, n)) < land_frac, rng.integers(0, l + 1, size=(m, n)), 0)  # land levels from the bottom
        for k in range(1, l + 1):
            landm[k, 1:m + 1, 1:n + 1][depth >= k] = 1
    if periodic:
        both = (landm[:, :, 1] == 0) & (landm[:, :, n] == 0)
        landm[:, :, 0][both] = 3
        landm[:, :, n + 1][both] = 3
    return s, landm


def random_state(s, landm, scale=0.01, seed=SEED, zero_on_land=True):
    """un = scale * N(0,1), zero on LAND cells."""
    rng = np.random.default_rng(seed)
    n, m, l = s.N, s.M, s.L
    un = scale * rng.standard_normal(6 * n * m * l)
    if zero_on_land:
        land = (landm[1:l + 1, 1:m + 1, 1:n + 1] != 0).reshape(-1)
        un.reshape(-1, 6)[land, :] = 0.0
    return un


def dirichlet_mask(s, landm):
    """True for unknowns whose row `boundaries` turns into an identity row (boundary.F90): every unknown of a
    non-ocean cell, u and v next to north / east / north-east land, w under a land (or surface) lid."""
    n, m, l = s.N, s.M, s.L
    lm = landm
    c = lm[1:-1, 1:-1, 1:-1]
    d = np.zeros((l, m, n, 6), bool)
    d[c != 0] = True
    uv = (lm[1:-1, 2:, 1:-1] == 1) | (lm[1:-1, 1:-1, 2:] == 1) | (lm[1:-1, 2:, 2:] == 1)
    d[..., 0] |= uv
    d[..., 1] |= uv
    top = lm[2:, 1:-1, 1:-1].copy()
    top[-1] = 1  # the surface lid: landm(:,:,l+1) = LAND (usrc.F90:107)
    d[..., 2] |= top == 1
    return d.reshape(-1)


def consistent_state(s, landm, scale=0.01, seed=SEED):
    """Random state on the constraint manifold: all Dirichlet unknowns are 0, as in any state the solver produces."""
    un = random_state(s, landm, scale=scale, seed=seed)
    un[dirichlet_mask(s, landm)] = 0.0
    return un


def smooth_state(s, value=1.234):  # test_ocean.C:141
    return np.full(6 * s.N * s.M * s.L, value)


def apply_pars(obj, pars, setter="setpar"):
    for k, v in pars.items():
        getattr(obj, setter)(PAR_INDEX[k], v)


# ---- coupled mode (BASELINE configs[2]: the ocean block of the coupled ocean + atmosphere + sea-ice model) ----
def coupled_inputs(s, seed=7):
    """Seeded stand-ins for what Ocean::synchronize hands THCM in a coupled run (Ocean.C:1451-1560): the atmosphere's
    temperature / humidity / albedo / precipitation, the sea-ice heat flux / mask / salt-flux correction on the GLOBAL
    surface grid, and the two CommPars structs (AtmosLocal.H / SeaIce.H defaults order of magnitude)."""
    rng = np.random.default_rng(seed)
    n, m = s.N, s.M
    yy = np.linspace(-1.0, 1.0, m)[:, None] * np.ones((1, n))
    fields = {
        "tatm": 0.3 * np.cos(1.3 * yy) + 0.05 * rng.standard_normal((m, n)),
        "qatm": 0.2 * rng.standard_normal((m, n)),
        "albe": rng.random((m, n)),
        "patm": 0.1 * rng.standard_normal((m, n)),
        "qsa": 0.5 * rng.standard_normal((m, n)),
        "msi": (rng.random((m, n)) < 0.3).astype(np.float64),   # 0 / 1 sea-ice mask, ~30 % ice covered
        "gsi": 0.01 * rng.standard_normal((m, n)),
        "emip": rng.standard_normal((m, n)),
        "adapted_emip": rng.standard_normal((m, n)),
        "spert": rng.standard_normal((m, n)),
    }
    # Atmosphere::CommPars: tdim qdim nuq eta dqso dqsi dqdt Eo0 Ei0 Cs t0o t0i a0 da tauf tauc comb albf
    atmos = np.array([1.0, 0.01, 5.9e-9 * 1e9, 4.87e-2, 5.0e-4, 5.3e-4, 6.4e-4, 2.0e-3, 9.0e-4, 1.1e-2, 15.0, -5.0, 0.3, 0.5,
                      4.7, 4.7, 1.0, 0.1])
    # SeaIce::CommPars: zeta a0 Lf s0 rhoo Qvar Q0
    seaice = np.array([7.3e-3 * 1e3, -0.0575, 3.347e5, 35.0, 1.024e3, 8.9, -100.0 * 1e-2])
    return fields, atmos, seaice


def apply_coupled(obj, fields, atmos, seaice):
    """obj: OracleTHCM / EmuTHCM (set_field, set_atmos_parameters, set_seaice_parameters)."""
    for k, f in fields.items():
        obj.set_field(k, f)
    obj.set_atmos_parameters(atmos)
    obj.set_seaice_parameters(seaice)
