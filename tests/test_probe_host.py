"""CPU tests of the host-side diagnostics / setup symbols of the B1 boundary (i-emic_b200/csrc/thcm_probe.cpp compiled into
tests/emu): m_probe, m_integrals, get_stochastic_forcing, set_internal_forcing, getdeps, loadbal_weights against the oracle
(oracle/thcm_oracle.cpp for the routines that need usol / forcing, oracle/probe_oracle.py = numpy restatement of probe.F90)."""
import numpy as np
import pytest

import cases
from cases import PAR_INDEX as P
from oracle.oracle import OracleTHCM
from oracle import probe_oracle as po
from emu.emu import EmuTHCM

PARS = dict(cases.DEFAULT_PARS, NLES=1.0, SUNP=1.0)
CASES = {"natl8": cases.natl8, "gateway16": cases.gateway16,
         "box_p": lambda **kw: cases.box(7, 6, 5, True, seed=3, land_frac=0.3, **kw),
         "box_np": lambda **kw: cases.box(6, 7, 4, False, seed=2, land_frac=0.3, **kw)}


def setup(name, coupled=True, **kw):
    if coupled:
        kw = dict(kw, coupled_T=1, coupled_S=1)
    s, landm = CASES[name](**kw)
    o, e = OracleTHCM(s, landm), EmuTHCM(s, landm)
    for k, v in PARS.items():
        o.setpar(P[k], v); e.setpar(P[k], v)
    fields, atmos, seaice = cases.coupled_inputs(s)
    cases.apply_coupled(o, fields, atmos, seaice)
    cases.apply_coupled(e, fields, atmos, seaice)
    return s, landm, o, e


def close(a, b, tol=1e-13):
    a, b = np.asarray(a), np.asarray(b)
    return np.abs(a - b).max() <= tol * (np.abs(b).max() + 1e-300)


@pytest.mark.parametrize("name", list(CASES))
def test_salt_integrals_bit_exact(name):
    s, landm, o, e = setup(name, coupled=False)
    x = cases.random_state(s, landm, scale=0.3, zero_on_land=False)
    assert np.array_equal(o.salt_advection(x), e.salt_advection(x))
    assert np.array_equal(o.salt_diffusion(x), e.salt_diffusion(x))
    assert np.abs(o.salt_advection(x)).max() > 0
    # THCM::integralChecks (THCM.C:2133-2140) / test_ocean.C:247-252: the flux-form integrand telescopes -- on the constraint
    # manifold (no flow through walls, lid and bottom) its plain sum over the cells vanishes
    xc = cases.consistent_state(s, landm, scale=0.1)
    adv = e.salt_advection(xc)
    assert np.abs(adv).sum() > 1e-3 and abs(adv.sum()) < 1e-12 * np.abs(adv).sum() * adv.size


@pytest.mark.parametrize("name", ["natl8", "box_p"])
@pytest.mark.parametrize("coupled", [True, False])
def test_surface_probes_match_probe_F90(name, coupled):
    s, landm, o, e = setup(name, coupled=coupled, SRES=0 if not coupled else 1)
    x = cases.random_state(s, landm, scale=0.2)
    cs = o.coupling_state()
    assert np.array_equal(e.getdeps(), [cs["Ooa"], cs["Os"], cs["nus"], cs["eta"], cs["lvsc"], cs["qdim"], o.getpar(P["COMB"]) * o.getpar(P["SALT"]) * cs["QSnd"]])
    assert np.array_equal(e.suno(), np.repeat(cs["suno"][:, None], s.N, axis=1))
    assert close(e.compute_evap(x), po.compute_evap(o, x, coupled))
    sf, corr, qa, qs = e.salflux(x)
    sfo, corro, qao, qso = po.get_salflux(o, x, coupled, s.SRES)
    assert close(sf, sfo, 1e-12) and abs(corr - corro) <= 1e-12 * (abs(corro) + 1e-300) and close(qa, qao) and close(qs, qso)
    tf = e.temflux(x)
    tfo = po.get_temflux(o, x, coupled, s.TRES)
    for k in tfo:
        assert close(tf[k], tfo[k]), k
    d = e.derivatives(x)
    do = po.get_derivatives(o, x, coupled, coupled)
    for a, b in zip(d, do):
        assert close(a, b)
    if coupled:
        assert np.abs(d[0]).max() > 0 and np.abs(d[2]).max() > 0 and np.all(d[3][landm[s.L, 1:-1, 1:-1] == 0] == -1.0)
    # plain getters (probe.F90:11-175, 440-490): emip comes back masked, q / p only in coupled mode
    land = landm[s.L, 1:s.M + 1, 1:s.N + 1]
    assert np.array_equal(e.probe_field("tatm"), o.get_field("tatm"))
    assert np.array_equal(e.probe_field("emip"), o.get_field("emip") * (1 - land))
    assert np.array_equal(e.probe_field("taux"), o.get_field("taux"))
    if coupled:
        assert np.array_equal(e.probe_field("qatm"), o.get_field("qatm")) and np.array_equal(e.probe_field("patm"), o.get_field("patm"))
    else:
        assert e.probe_field("qatm") is None and e.probe_field("patm") is None


@pytest.mark.parametrize("name,flags", [("natl8", dict(SRES=0)), ("box_p", dict()), ("box_np", dict(SRES=0, its=0))])
def test_stochastic_forcing_and_internal_forcing(name, flags):
    s, landm, o, e = setup(name, coupled=False, **flags)
    for obj in (o, e):
        obj.setpar(P["SPER"], 0.3)
    bo, jo, co = o.stochastic_forcing()
    be, je, ce = e.stochastic_forcing()
    assert np.array_equal(bo, be) and np.array_equal(jo, je) and np.array_equal(co, ce)
    assert bo[-1] == s.N * s.M + 1 and np.count_nonzero(co) > 0
    assert o.getpar(P["SPER"]) == e.getpar(P["SPER"]) == 0.3
    # m_usr::set_internal_forcing (usr.F90:267-300) -> w-row forcing (forcing.F90:199-209), effective at the next setpar
    rng = np.random.default_rng(5)
    t3, s3 = rng.standard_normal((s.L, s.M, s.N)), rng.standard_normal((s.L, s.M, s.N))
    o.set_internal_forcing(t3, s3); e.set_internal_forcing(t3, s3)
    for obj in (o, e):
        obj.setpar(P["COMB"], 0.7)
    fo = o.forcing()
    assert np.array_equal(fo, e.forcing(masked=False))
    assert np.count_nonzero(fo.reshape(-1, 6)[:, 2]) > 0
    x = cases.random_state(s, landm, scale=0.1)
    assert np.array_equal(o.rhs(x), e.rhs(x))


def test_loadbal_weights_count_ocean_cells():
    s, landm, o, e = setup("natl8", coupled=False)
    w = e.loadbal_weights()
    assert np.array_equal(w, (landm[:, 1:-1, 1:-1] == 0).sum(axis=0) / s.L)   # thcm_utils.F90:335-351 (no extra mixing weights)


def test_wind_forcing_round_trip_like_the_reference_test():
    """src/tests/test_ocean.C:346-380 (TEST(Ocean, WindForcing)): with "Wind Forcing Type" = 3 the inserted wind stress survives the
    parameter change and the Jacobian computation, and comes back unchanged."""
    s, landm = cases.natl8(iza=3)
    o, e = OracleTHCM(s, landm), EmuTHCM(s, landm)
    rng = np.random.default_rng(11)
    taux, tauy = rng.uniform(-1, 1, (s.M, s.N)), rng.uniform(-1, 1, (s.M, s.N))
    for obj in (o, e):
        obj.set_field("taux", taux); obj.set_field("tauy", tauy)
        obj.setpar(P["COMB"], 0.2)
    x = cases.smooth_state(s)                       # getState('V')->PutScalar(1.234)
    assert np.array_equal(o.jacobian_graph(x)[0], e.jacobian(x))
    tx, ty = e.probe_field("taux"), e.probe_field("tauy")
    assert np.linalg.norm(tx) >= 1e-2 and np.linalg.norm(tx - taux) <= 1e-7
    assert np.linalg.norm(ty) >= 1e-2 and np.linalg.norm(ty - tauy) <= 1e-7
    # ... and it is what drives the u, v rows of the forcing (forcing.F90:40-45)
    f = e.forcing(masked=False).reshape(s.L, s.M, s.N, 6)
    sigma = 0.2 * o.getpar(P["WIND"]) * o.getpar(P["AL_T"])
    assert np.array_equal(f[s.L - 1, : s.M - 1, :, 0], (sigma * taux)[: s.M - 1])
    assert np.array_equal(o.forcing(), e.forcing(masked=True))     # after matrix(): identity rows zeroed (boundary.F90)


def test_salt_and_temperature_forcing_round_trip_like_the_reference_test():
    """src/tests/test_ocean.C:383-416 (TEST(Ocean, SaltTempForcing)): "Levitus T" = "Levitus S" = 2 (fields set externally): emip comes
    back zero on the land mask, tatm unchanged, after a parameter change and a Jacobian computation."""
    s, landm = cases.natl8(ite=2, its=2)
    o, e = OracleTHCM(s, landm), EmuTHCM(s, landm)
    rng = np.random.default_rng(12)
    emip, tatm = rng.uniform(-1, 1, (s.M, s.N)), rng.uniform(-1, 1, (s.M, s.N))
    for obj in (o, e):
        obj.set_field("emip", emip); obj.set_field("tatm", tatm)
        obj.setpar(P["COMB"], 0.05)
    x = cases.smooth_state(s)
    assert np.array_equal(o.jacobian_graph(x)[0], e.jacobian(x))
    land = landm[s.L, 1:s.M + 1, 1:s.N + 1]
    em, ta = e.probe_field("emip"), e.probe_field("tatm")
    assert np.linalg.norm(ta) >= 1e-2 and np.linalg.norm(ta - tatm) <= 1e-7
    assert np.linalg.norm(em) >= 1e-2 and np.array_equal(em, emip * (1 - land))
    assert np.array_equal(o.forcing(), e.forcing(masked=False))
    assert np.array_equal(o.rhs(x), e.rhs(x))


@pytest.mark.parametrize("name", ["natl8", "box_np"])
def test_ocean_coupling_blocks_are_the_derivatives_of_the_residual(name):
    """Ocean::getBlock(Atmosphere) / getBlock(SeaIce) (Ocean.C:1603-1810, SURVEY 8f N4: the ocean's side of the coupled Jacobian) against
    finite differences of the library's own residual with respect to the fields the other models hand in -- the residual is affine in
    every one of them, so the quotient is exact up to rounding.  F = -B (THCM.C:1011)."""
    import scipy.sparse as sp
    mk = {"natl8": cases.natl8, "box_np": lambda **kw: cases.box(6, 7, 4, False, seed=2, land_frac=0.3, **kw)}[name]
    s, landm = mk(coupled_T=1, coupled_S=1)
    e = EmuTHCM(s, landm)
    for k, v in dict(cases.DEFAULT_PARS, NLES=1.0, SUNP=1.0).items():
        e.setpar(P[k], v)
    fields, atmos, seaice = cases.coupled_inputs(s)
    rng = np.random.default_rng(12)
    fields["msi"] = rng.random(fields["msi"].shape)           # a fractional mask exercises the (1 - M) factors
    cases.apply_coupled(e, fields, atmos, seaice)
    n, m = s.N, s.M
    x = cases.random_state(s, landm, scale=0.2)
    sr = np.arange(n * m)
    colT, colQ, colA = 3 * sr, 3 * sr + 1, 3 * sr + 2         # FIND_ROW_ATMOS0 with dof = 3 on the surface level
    colP = np.full(n * m, 3 * n * m)                          # the auxiliary precipitation row (dim - aux)
    pdist = 0.5 + rng.random((m, n))
    F0 = -e.rhs(x)

    def refresh():
        e.setpar(P["COMB"], cases.DEFAULT_PARS["COMB"])       # forcing + lin are re-run at the next parameter change (inserts.F90)

    def fd_column(field, pattern, delta=1e-3):
        f = fields[field].copy()
        e.set_field(field, f + delta * pattern); refresh()
        col = (-e.rhs(x) - F0) / delta
        e.set_field(field, f); refresh()
        return col

    beg, jco, co = e.ocean_block_atmosphere(atmos[13], pdist, colT, colQ, colA, colP)
    A = sp.csr_matrix((co, jco, beg), shape=(e.ndim, 3 * n * m + 1)).tocsc()
    surf_ocean = np.argwhere(landm[s.L, 1:-1, 1:-1] == 0)
    assert len(surf_ocean) > 3 and A.nnz > 0
    scale = np.abs(A).max()
    for j, i in surf_ocean[rng.choice(len(surf_ocean), size=4, replace=False)]:
        unit = np.zeros((m, n)); unit[j, i] = 1.0
        for field, col in (("tatm", colT), ("qatm", colQ), ("albe", colA)):
            want = fd_column(field, unit)
            got = A[:, col[j * n + i]].toarray().ravel()
            assert np.abs(got - want).max() <= 1e-9 * scale, (field, i, j)
    # P is the atmosphere's nondimensional global precipitation anomaly: the field it hands the ocean is
    # pdist * (Po0 + eta * qdim * P) (AtmosLocal.C:1117-1124), hence the eta * qdim inside nus (usrc.F90:266)
    want = fd_column("patm", atmos[3] * atmos[1] * pdist)
    assert np.abs(A[:, 3 * n * m].toarray().ravel() - want).max() <= 1e-9 * scale

    beg, jco, co = e.ocean_block_seaice(x, 4 * sr + 1, 4 * sr + 2, 4 * sr + 3)
    B = sp.csr_matrix((co, jco, beg), shape=(e.ndim, 4 * n * m)).tocsc()
    scale = np.abs(B).max()
    assert B.nnz > 0
    for j, i in surf_ocean[rng.choice(len(surf_ocean), size=4, replace=False)]:
        unit = np.zeros((m, n)); unit[j, i] = 1.0
        for field, off in (("qsa", 1), ("msi", 2), ("gsi", 3)):
            want = fd_column(field, unit)
            got = B[:, 4 * (j * n + i) + off].toarray().ravel()
            assert np.abs(got - want).max() <= 1e-9 * max(scale, 1.0), (field, i, j)
