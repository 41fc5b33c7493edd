"""CPU tests of the library's device functions and host setup (tests/emu = thcm_cell.cuh + thcm_host.cpp compiled for
the host) against the oracle: bit-exact residual, Fortran-order CRS (pattern, row order, values), graph-order Jacobian,
mass diagonal, forcing -- on one block and on every block of 2/4/8-rank decompositions with halos filled from the global
state (which checks the decomposition, the static graph with halo columns and the halo plan without a GPU)."""
import numpy as np
import pytest
import scipy.sparse as sp

import cases
from cases import PAR_INDEX as P
from oracle.oracle import OracleTHCM
from emu.emu import EmuTHCM

CASES = {
    "natl8": cases.natl8,
    "test6x6x4": cases.test6x6x4,
    "gateway16": cases.gateway16,
    "global4deg": cases.global4deg,
    "box_p": lambda **kw: cases.box(7, 6, 5, True, seed=3, land_frac=0.3, **kw),
    "box_np": lambda **kw: cases.box(6, 7, 4, False, seed=2, land_frac=0.3, **kw),
    "box_p_open": lambda **kw: cases.box(9, 5, 3, True, seed=4, land_frac=0.0, **kw),
    "box_tiny": lambda **kw: cases.box(3, 2, 2, True, seed=5, land_frac=0.2, **kw),
}
PARS = dict(cases.DEFAULT_PARS, NLES=1.0)


def setup(name, pars=PARS, **kw):
    s, landm = CASES[name](**kw)
    o, e = OracleTHCM(s, landm), EmuTHCM(s, landm)
    for k, v in pars.items():
        o.setpar(P[k], v)
        e.setpar(P[k], v)
    return s, landm, o, e


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("state", ["random_raw", "consistent", "smooth"])
def test_single_block_bit_exact(name, state):
    s, landm, o, e = setup(name)
    x = {"random_raw": lambda: cases.random_state(s, landm, scale=0.3, zero_on_land=False),
         "consistent": lambda: cases.consistent_state(s, landm, scale=0.01),
         "smooth": lambda: cases.smooth_state(s)}[state]()
    assert np.array_equal(o.rhs(x), e.rhs(x))
    assert e.check_staging(x) == 0   # the kernels' unconditional-load staging == the usol ghost rules
    assert e.check_tiles(x) == 0     # the pipelined kernel's TMA line plan + descriptor fix-up == the usol ghost rules
    bo, jo, co, cob = o.matrix(x)
    be, je, ce = e.crs(x)
    assert np.array_equal(bo, be) and np.array_equal(jo, je) and np.array_equal(co, ce)   # pattern, row order, values
    vo, missing = o.jacobian_graph(x)
    assert missing == 0
    ro, cco = o.graph()
    re_, ce_ = e.graph()
    assert np.array_equal(ro, re_) and np.array_equal(cco, ce_)
    assert np.array_equal(vo, e.jacobian(x))
    assert np.array_equal(cob, e.cob())
    assert np.array_equal(o.forcing(), e.forcing(masked=True))


def test_parameters_and_forcing_follow_setpar():
    s, landm, o, e = setup("natl8", pars={})
    for idx in range(1, 31):
        assert o.getpar(idx) == e.getpar(idx)
    f0 = e.forcing(masked=False)
    assert np.all(f0 == 0.0)
    for k, v in PARS.items():
        o.setpar(P[k], v)
        e.setpar(P[k], v)
    assert np.count_nonzero(e.forcing(masked=False)) > 0
    x = cases.random_state(s, landm)
    o.rhs(x)  # boundaries zeroes Frc rows lazily inside rhs/matrix (boundary.F90:167,256,...)
    assert np.array_equal(o.forcing(), e.forcing(masked=True))


@pytest.mark.parametrize("name,flags", [("natl8", dict(TRES=0, SRES=0)), ("gateway16", dict(ih=1)),
                                         ("box_p", dict(coriolis_on=0, forcing_type=2)), ("box_np", dict(forcing_type=1, SRES=0))])
def test_model_flags(name, flags):
    s, landm, o, e = setup(name, **flags)
    x = cases.random_state(s, landm, scale=0.1)
    assert np.array_equal(o.rhs(x), e.rhs(x))
    assert np.array_equal(o.jacobian_graph(x)[0], e.jacobian(x))


@pytest.mark.parametrize("name", ["natl8", "gateway16", "box_p", "box_np", "global4deg"])
@pytest.mark.parametrize("vmix,rho_mixing,xes", [(1, 0, 1.0), (1, 1, 0.0), (2, 0, 0.0)])
def test_tracer_mixing_bit_exact(name, vmix, rho_mixing, xes):
    """Mixing = 1, 2 (mix_imp.f: implicit vertical mixing / convective adjustment + its forward-difference Jacobian): the
    device functions, compiled for the host (tanh = thcm_tanh.h, the oracle's = oracle/fdlibm_tanh.h: the same specified algorithm), against the oracle's whole-field vmix_fun / vmix_jac."""
    s, landm, o, e = setup(name, pars=dict(cases.DEFAULT_PARS, NLES=xes), vmix=vmix, rho_mixing=rho_mixing)
    x = cases.random_state(s, landm, scale=0.3, zero_on_land=False)
    e.vmix_control(x)
    assert np.array_equal(o.rhs(x), e.rhs(x))             # (Mixing = 2: the oracle decides its partition inside rhs)
    assert np.abs(o.vmix_fun(x)).max() > 0
    assert np.array_equal(o.jacobian_graph(x)[0], e.jacobian(x))
    bo, jo, co, _ = o.matrix(x)
    be, je, ce = e.crs(x)
    assert np.array_equal(bo, be) and np.array_equal(jo, je) and np.array_equal(co, ce)


@pytest.mark.parametrize("name", ["natl8", "gateway16", "box_p", "global4deg"])
@pytest.mark.parametrize("alpc,pvc,rho_mixing,xes", [(0.5, None, 0, 1.0), (0.0, 0.0, 0, 0.0), (0.7, None, 1, 0.0)])
def test_consistent_vertical_mixing_bit_exact(name, alpc, pvc, rho_mixing, xes):
    """ALPC != 1 ("consistent" vertical mixing, mix_imp.f:478-487: Ftzt = tprstb(drhodzt) * eps * dtdzt / (drhodzt - 1e-20) with
    eps = (1 - ALPC) * ENER * PE_V), alone (P_VC = 0) and together with the implicit mixing: residual, every Jacobian value incl. the
    forward-difference block, CRS pattern / order / values against the oracle's whole-field vmix_fun / vmix_jac."""
    pars = dict(cases.DEFAULT_PARS, NLES=xes, ALPC=alpc)
    if pvc is not None:
        pars["P_VC"] = pvc
    s, landm, o, e = setup(name, pars=pars, vmix=1, rho_mixing=rho_mixing)
    x = cases.random_state(s, landm, scale=0.3, zero_on_land=False)
    ref = OracleTHCM(s, landm)                      # the same run with ALPC = 1: the new term must actually be live
    for k, v in dict(pars, ALPC=1.0).items():
        ref.setpar(P[k], v)
    assert np.abs(o.vmix_fun(x) - ref.vmix_fun(x)).max() > 0
    assert np.array_equal(o.rhs(x), e.rhs(x))
    assert np.array_equal(o.jacobian_graph(x)[0], e.jacobian(x))
    bo, jo, co, _ = o.matrix(x)
    be, je, ce = e.crs(x)
    assert np.array_equal(bo, be) and np.array_equal(jo, je) and np.array_equal(co, ce)


def test_neutral_physics_leaves_the_reference_graph():
    """Why MIXP / MKAP != 0 stay refused: with neutral physics or GM the forward-difference block of vmix_jac has T,S entries on
    stencil positions outside the maximal graph of THCM.C:2320-2549 -- THCM::evaluate's ReplaceGlobalValues (THCM.C:1095-1104) fails on
    them, so the reference's own hot path cannot run these modes either; with the vertical schemes every entry is inside."""
    s, landm = cases.natl8(vmix=1)
    x = cases.random_state(s, landm, scale=0.3)
    for pars, outside in [(dict(ALPC=0.5), False), (dict(MIXP=0.5), True), (dict(MKAP=0.5), True)]:
        o = OracleTHCM(s, landm)
        for k, v in dict(cases.DEFAULT_PARS, **pars).items():
            o.setpar(P[k], v)
        _, missing = o.jacobian_graph(x)
        assert (missing > 0) == outside, (pars, missing)


def test_mixing_partition_follows_the_fields():
    """Mixing = 2 (vmix_control, mix_imp.f:139-169): a zero temperature field switches the temperature mixing off, and --
    because the reference only re-partitions when the TEMPERATURE flag is set (:163) -- leaves the pair count at zero."""
    s, landm, o, e = setup("natl8", vmix=2)
    x = cases.random_state(s, landm, scale=0.3)
    x[4::6] = 0.0
    e.vmix_control(x)
    assert np.array_equal(o.rhs(x), e.rhs(x))
    assert o.vmix_flags() == {"flag": 2, "temp": 0, "salt": 1, "fix": 1}
    assert np.array_equal(o.jacobian_graph(x)[0], e.jacobian(x))


@pytest.mark.parametrize("name", ["natl8", "gateway16", "box_p", "box_np", "box_p_open"])
@pytest.mark.parametrize("nranks", [2, 4, 8])
@pytest.mark.parametrize("balance", [0, 1])
def test_decomposed_blocks_reproduce_the_global_answer(name, nranks, balance):
    """Every rank's block, with its halo filled from the global state, must give exactly the owned rows of the 1-rank
    answer (the multi-GPU path reproduces the global result on any GPU count)."""
    s, landm, o, _ = setup(name)
    x = cases.random_state(s, landm, scale=0.3, zero_on_land=False)
    B = o.rhs(x)
    vo, _ = o.jacobian_graph(x)
    ro, cco = o.graph()
    bo, jo, co, cob = o.matrix(x)
    seen = np.zeros(o.ndim, int)
    for rank in range(nranks):
        sr, _ = CASES[name](rank=rank, nranks=nranks, balance=balance)   # 1: ocean-weighted cut lines instead of the reference's uniform ones
        e = EmuTHCM(sr, landm)
        for k, v in PARS.items():
            e.setpar(P[k], v)
        gid = e.local_gids()
        seen[gid] += 1
        hg = e.halo_gids()
        halo = np.where(hg >= 0, x[np.maximum(hg, 0)], np.nan)   # unused halo slots stay NaN: must never be read
        xl = x[gid]
        assert e.check_staging(xl, np.nan_to_num(halo, nan=12345.0)) == 0
        assert e.check_tiles(xl, np.nan_to_num(halo, nan=12345.0)) == 0
        assert np.array_equal(e.rhs(xl, halo), B[gid])
        # graph rows: same global columns in the same (ascending) order, same values
        rp, col = e.graph()
        val = e.jacobian(xl, halo)
        gcol = np.where(col < e.ndim, gid[np.minimum(col, e.ndim - 1)], hg[np.maximum(col - e.ndim, 0)])
        for r in range(0, e.ndim, max(1, e.ndim // 997)):
            g = gid[r]
            assert np.array_equal(gcol[rp[r]:rp[r + 1]], cco[ro[g]:ro[g + 1]])
            assert np.array_equal(val[rp[r]:rp[r + 1]], vo[ro[g]:ro[g + 1]])
        lens_ok = np.array_equal(np.diff(rp), np.diff(ro)[gid])
        assert lens_ok
        # all values at once (rows are contiguous per owned row)
        idx = np.concatenate([np.arange(ro[g], ro[g + 1]) for g in gid[:: max(1, e.ndim // 5000)]])
        idl = np.concatenate([np.arange(rp[r], rp[r + 1]) for r in range(0, e.ndim, max(1, e.ndim // 5000))])
        assert np.array_equal(val[idl], vo[idx])
        # Fortran-order CRS with global 1-based columns
        be, je, ce = e.crs(xl, halo)
        for r in range(0, e.ndim, max(1, e.ndim // 499)):
            g = gid[r]
            assert np.array_equal(je[be[r] - 1:be[r + 1] - 1], jo[bo[g] - 1:bo[g + 1] - 1])
            assert np.array_equal(ce[be[r] - 1:be[r + 1] - 1], co[bo[g] - 1:bo[g + 1] - 1])
        assert np.array_equal(e.cob(), cob[gid])
    assert np.all(seen == 1)   # the blocks tile the global index space exactly once


@pytest.mark.parametrize("nranks", [2, 4, 8])
@pytest.mark.parametrize("name", ["gateway16", "box_np", "global4deg"])
@pytest.mark.parametrize("balance", [0, 1])
def test_halo_plan_is_consistent(name, nranks, balance):
    """What rank p packs for rank q is exactly what q unpacks, slot by slot (global ids match)."""
    _, landm = CASES[name]()
    emus = []
    for rank in range(nranks):
        sr, _ = CASES[name](rank=rank, nranks=nranks, balance=balance)
        emus.append(EmuTHCM(sr, landm))
    plans = [e.plan() for e in emus]
    for p, e in enumerate(emus):
        peers, send, recv = plans[p]
        gid = e.local_gids().reshape(-1, 6)[:, 0] // 6       # global cell id of every owned cell
        hg = e.halo_gids().reshape(-1, 6)
        for (q, so, sc, ro_, rc) in peers:
            qpeers, qsend, qrecv = plans[q]
            row = [r for r in qpeers if r[0] == p]
            assert len(row) == 1
            _, qso, qsc, qro, qrc = row[0]
            assert sc == qrc and rc == qsc
            sent_cells = gid[send[so:so + sc]]
            qhg = emus[q].halo_gids().reshape(-1, 6)
            want = qhg[qrecv[qro:qro + qrc], 0] // 6
            ok = qhg[qrecv[qro:qro + qrc], 0] >= 0
            assert np.array_equal(sent_cells[ok], want[ok])
        # every halo slot that the graph references is received from exactly one peer
        used = np.unique(np.nonzero(hg[:, 0] >= 0)[0])
        assert set(used).issubset(set(recv.tolist()))
        assert len(set(recv.tolist())) == len(recv)


def test_ocean_weighted_cut_lines_balance_the_global_grids():
    """thcmb_settings.balance = 1: same rank grid as the reference's Decomp2D, rectangular blocks that tile the domain, cut lines
    placed by OCEAN-cell count.  On the synthetic 1-degree mask the slowest of 8 ranks owns 1.62x the mean number of ocean cells with
    the reference's uniform cuts and <= 1.05x with the weighted ones."""
    s0, landm = cases.global_synth(360, 152, 24)
    ocean = (landm[1:-1, 1:-1, 1:-1] == 0)
    for nranks, worst_uniform in ((2, 1.15), (4, 1.29), (8, 1.60)):
        out = {}
        for balance in (0, 1):
            cnt, cover = [], np.zeros((152, 360), int)
            for r in range(nranks):
                s, _ = cases.global_synth(360, 152, 24, rank=r, nranks=nranks, balance=balance)
                i0, j0, n0, m0, npN, npM = EmuTHCM.block_only(s, landm)
                cover[j0:j0 + m0, i0:i0 + n0] += 1
                cnt.append(ocean[:, j0:j0 + m0, i0:i0 + n0].sum())
            assert np.all(cover == 1)
            out[balance] = max(cnt) / np.mean(cnt)
        assert out[0] >= worst_uniform and out[1] <= 1.05, out


def test_decomp2d_matches_reference_rule():
    """TRIOS_Domain.C:201-315 incl. the r_min = 100 quirk; SURVEY.md section 8e block shapes."""
    def blocks(n, m, l, nranks):
        out = []
        for r in range(nranks):
            s = cases.Settings.from_degrees(n, m, l, 0, 359.99, -85.5, 85.5, periodic=True, rank=r, nranks=nranks)
            e = EmuTHCM(s, cases.all_ocean_mask(n, m, l, True))
            out.append(e.block())
        return out
    b8 = blocks(360, 152, 2, 8)
    assert {(b["npN"], b["npM"]) for b in b8} == {(4, 2)} and {(b["n0"], b["m0"]) for b in b8} == {(90, 76)}
    b2 = blocks(360, 152, 2, 2)
    assert {(b["npN"], b["npM"]) for b in b2} == {(2, 1)} and {(b["n0"], b["m0"]) for b in b2} == {(180, 152)}
    b4 = blocks(360, 152, 2, 4)
    assert {(b["npN"], b["npM"]) for b in b4} == {(4, 1)} and {(b["n0"], b["m0"]) for b in b4} == {(90, 152)}
    # BASELINE configs[4]: 0.5 degree on 8 GPUs = 4 x 2 blocks of 180 x 152 columns (0.88 M cells per GPU at L = 32)
    b05 = blocks(720, 304, 2, 8)
    assert {(b["npN"], b["npM"]) for b in b05} == {(4, 2)} and {(b["n0"], b["m0"]) for b in b05} == {(180, 152)}
    # remainders go to the first ranks (TRIOS_Domain.C:267-273)
    b3 = blocks(10, 7, 2, 3)
    assert sum(b["n0"] * b["m0"] for b in b3) == 70


# ---- coupled mode (coupled_T / coupled_S = 1, usrc.F90:742-783, forcing.F90:66-164, inserts.F90) ----
@pytest.mark.parametrize("name", ["natl8", "gateway16", "box_p", "box_np"])
@pytest.mark.parametrize("flags", [dict(coupled_T=1, coupled_S=1), dict(coupled_T=1, coupled_S=0), dict(coupled_T=0, coupled_S=1, SRES=0)])
def test_coupled_mode_bit_exact(name, flags):
    s, landm, o, e = setup(name, pars=dict(PARS, SUNP=1.0), **flags)
    fields, atmos, seaice = cases.coupled_inputs(s)
    cases.apply_coupled(o, fields, atmos, seaice)
    cases.apply_coupled(e, fields, atmos, seaice)
    x = cases.random_state(s, landm, scale=0.1)
    fo = o.forcing()
    assert np.array_equal(fo, e.forcing(masked=False))
    assert np.count_nonzero(fo.reshape(-1, 6)[:, 4:]) > 0
    assert np.array_equal(o.rhs(x), e.rhs(x))
    bo, jo, co, cob = o.matrix(x)
    be, je, ce = e.crs(x)
    assert np.array_equal(bo, be) and np.array_equal(jo, je) and np.array_equal(co, ce)
    vo, missing = o.jacobian_graph(x)
    assert missing == 0                        # the T,S / S,T centre entries lie inside the maximal graph
    assert np.array_equal(vo, e.jacobian(x))
    # the coupling really changes the operator: the sea-ice mask puts TT,SS / SS,TT entries on the surface level
    s2, landm2, o2, e2 = setup(name, pars=dict(PARS, SUNP=1.0))
    assert not np.array_equal(o2.jacobian_graph(x)[0], vo)
    # a later parameter change keeps nus / lvsc frozen (usrc.F90:297-304) and re-runs forcing + lin
    for obj in (o, e):
        obj.setpar(P["COMB"], 0.5)
    assert np.array_equal(o.rhs(x), e.rhs(x))
    assert np.array_equal(o.jacobian_graph(x)[0], e.jacobian(x))


def test_insert_masks_the_freshwater_fields():
    # inserts.F90:179,198,217: emip / adapted_emip / emip_pert are multiplied by (1 - landm(i,j,l)) on the way in
    s, landm, o, e = setup("natl8", pars=dict(PARS, HMTP=0.3, SPER=0.2), its=0, SRES=0)
    fields, _, _ = cases.coupled_inputs(s)
    for k in ("emip", "adapted_emip", "spert", "tatm", "qatm", "msi"):
        o.set_field(k, fields[k]); e.set_field(k, fields[k])
    o.setpar(P["SALT"], 1.0); e.setpar(P["SALT"], 1.0)
    x = cases.random_state(s, landm)
    o.rhs(x)
    assert np.array_equal(o.forcing(), e.forcing(masked=True))
    assert np.array_equal(o.rhs(x), e.rhs(x))


@pytest.mark.parametrize("name", ["gateway16", "box_np"])
@pytest.mark.parametrize("nranks", [2, 4])
def test_coupled_mode_on_decomposed_blocks(name, nranks):
    """Coupled mode on the 2-D decomposition: every rank holds the GLOBAL surface fields, restricts msi to its columns and
    must reproduce the owned rows of the 1-rank residual, Jacobian and forcing."""
    flags = dict(coupled_T=1, coupled_S=1)
    s, landm, o, _ = setup(name, pars=dict(PARS, SUNP=1.0), **flags)
    fields, atmos, seaice = cases.coupled_inputs(s)
    cases.apply_coupled(o, fields, atmos, seaice)
    x = cases.random_state(s, landm, scale=0.3, zero_on_land=False)
    fo = o.forcing()            # as `forcing` leaves it (before rhs zeroes the identity rows)
    B = o.rhs(x)
    vo, _ = o.jacobian_graph(x)
    ro, _ = o.graph()
    for rank in range(nranks):
        sr, _ = CASES[name](rank=rank, nranks=nranks, **flags)
        e = EmuTHCM(sr, landm)
        for k, v in dict(PARS, SUNP=1.0).items():
            e.setpar(P[k], v)
        cases.apply_coupled(e, fields, atmos, seaice)
        gid = e.local_gids()
        hg = e.halo_gids()
        halo = np.where(hg >= 0, x[np.maximum(hg, 0)], np.nan)
        xl = x[gid]
        assert np.array_equal(e.forcing(masked=False), fo[gid])
        assert np.array_equal(e.rhs(xl, halo), B[gid])
        rp, _ = e.graph()
        val = e.jacobian(xl, halo)
        idx = np.concatenate([np.arange(ro[g], ro[g + 1]) for g in gid])
        assert np.array_equal(val, vo[idx])


@pytest.mark.parametrize("name", ["natl8", "gateway16", "box_p"])
@pytest.mark.parametrize("vmix", [0, 1])
def test_set_landmask_rebuilds_everything(name, vmix):
    """SUBROUTINE set_landmask (usrc.F90:353-418; THCM::setLandMask, the mask fixing of Ocean.C:496-566): a new mask with extra land
    -- one cell placed so that the land-inversion fix has to act -- and reinit = 1: residual, Jacobian, forcing and mass
    diagonal afterwards equal a fresh oracle's, bit for bit."""
    s, landm, o, e = setup(name, vmix=vmix)
    n, m, l = s.N, s.M, s.L
    new = landm.copy()
    ocean = np.argwhere(new[1:l + 1, 1:m + 1, 1:n + 1] == 0)
    rng = np.random.default_rng(9)
    for k, j, i in ocean[rng.choice(len(ocean), size=max(3, len(ocean) // 20), replace=False)]:
        new[k + 1, j + 1, i + 1] = 1                      # LAND somewhere in the column: everything below must follow
    for obj in (o, e):
        obj.set_landmask(new, s.periodic, 1)
    fixed = o.landm()
    assert (fixed[1:l + 1, 1:m + 1, 1:n + 1] != new[1:l + 1, 1:m + 1, 1:n + 1]).any()   # the inversion fix acted
    x = cases.random_state(s, fixed, scale=0.2, zero_on_land=False)
    assert np.array_equal(o.rhs(x), e.rhs(x))
    assert e.check_tiles(x) == 0
    bo, jo, co, cob = o.matrix(x)
    be, je, ce = e.crs(x)
    assert np.array_equal(bo, be) and np.array_equal(jo, je) and np.array_equal(co, ce)
    assert np.array_equal(o.jacobian_graph(x)[0], e.jacobian(x))
    assert np.array_equal(cob, e.cob())
    assert np.array_equal(o.forcing(), e.forcing(masked=True))
    # and a fresh model on the fixed mask agrees with the re-masked one
    o2 = OracleTHCM(s, fixed)
    for k, v in PARS.items():
        o2.setpar(P[k], v)
    assert np.array_equal(o2.rhs(x), e.rhs(x))


@pytest.mark.parametrize("forcing_type,TRES,SRES", [(0, 1, 1), (2, 1, 1), (1, 0, 0), (0, 0, 1)])
def test_sub_domain_with_global_latitude_bounds(forcing_type, TRES, SRES):
    """What one MPI rank of the reference hands the Fortran symbols (THCM.C:566-611): a sub-domain with its OWN bounds, while the
    idealised forcing profiles keep using the GLOBAL latitude bounds of m_global (forcing.F90:418-449)."""
    rad = np.pi / 180.0
    s, landm = cases.box(6, 5, 4, False, seed=8, land_frac=0.25, forcing_type=forcing_type, TRES=TRES, SRES=SRES)
    s.xmin, s.xmax, s.ymin, s.ymax = 300 * rad, 340 * rad, 22 * rad, 54 * rad      # the block
    s.ymin_glob, s.ymax_glob = 10 * rad, 74 * rad                                  # the domain it belongs to
    o, e = OracleTHCM(s, landm), EmuTHCM(s, landm)
    for k, v in dict(PARS, CMPR=0.3, FPER=0.2).items():
        o.setpar(P[k], v)
        e.setpar(P[k], v)
    x = cases.random_state(s, landm, scale=0.1)
    assert np.array_equal(o.forcing(), e.forcing(masked=False))
    assert np.array_equal(o.rhs(x), e.rhs(x))
    assert np.array_equal(o.jacobian_graph(x)[0], e.jacobian(x))
    # the profiles really are the global ones: a model that believes the block is the whole domain gets another forcing
    s2, _ = cases.box(6, 5, 4, False, seed=8, land_frac=0.25, forcing_type=forcing_type, TRES=TRES, SRES=SRES)
    s2.xmin, s2.xmax, s2.ymin, s2.ymax = s.xmin, s.xmax, s.ymin, s.ymax
    e2 = EmuTHCM(s2, landm)
    for k, v in dict(PARS, CMPR=0.3, FPER=0.2).items():
        e2.setpar(P[k], v)
    assert not np.array_equal(e2.forcing(masked=False), e.forcing(masked=False))


def test_setsres_toggles_the_restoring_terms_like_thcm_evaluate():
    """SUBROUTINE setsres (usrc.F90:434-446) re-runs forcing + lin with the other restoring flag and back again, the way
    THCM::evaluate brackets matrix_ for its mask test (THCM.C:1059-1070)."""
    s, landm, o, e = setup("natl8", SRES=0)
    x = cases.random_state(s, landm, scale=0.1)
    B0 = o.rhs(x)
    assert np.array_equal(B0, e.rhs(x))
    for obj in (o, e):
        obj.setsres(1)
    vo, _ = o.jacobian_graph(x)
    assert np.array_equal(vo, e.jacobian(x))
    assert np.array_equal(o.forcing(), e.forcing(masked=True))
    for obj in (o, e):
        obj.setsres(0)
    assert np.array_equal(o.rhs(x), e.rhs(x)) and np.array_equal(o.rhs(x), B0)
    assert not np.array_equal(o.jacobian_graph(x)[0], vo)        # the restoring term sits on the S diagonal of the surface cells


@pytest.mark.parametrize("name", ["natl8", "gateway16", "global4deg"])
def test_ocean_only_krylov_space_is_exact(name):
    """Design validation for the cell-compacted Krylov space: rows of LAND cells are identity rows and the Newton
    right-hand side vanishes there, so GMRES on the system restricted to the ocean cells (the library's cell maps) is the SAME iteration:
    the reference's own GMRES template gives the same residual history and, scattered back, the same solution."""
    from oracle.oracle import kref_gmres, spmv
    s, landm, o, e = setup(name)
    x = cases.consistent_state(s, landm, scale=0.05)
    val, _ = o.jacobian_graph(x)
    rp, col = o.graph()
    b = o.rhs(x)
    n = o.ndim
    ocell, ccell = e.cell_maps()
    land = landm[1:-1, 1:-1, 1:-1].reshape(-1) != 0
    assert np.array_equal(ccell >= 0, ~land) and np.array_equal(ocell, np.nonzero(~land)[0])
    assert np.all(b.reshape(-1, 6)[land] == 0.0)                                  # F is masked by (1 - landm) (usrc.F90:580-591)
    J = sp.csr_matrix((val, col, rp), shape=(n, n))
    orow = (6 * ocell[:, None] + np.arange(6)[None, :]).reshape(-1)
    Jl = J[np.repeat(land, 6)]
    assert (Jl != sp.identity(n, format="csr")[np.repeat(land, 6)]).nnz == 0      # LAND rows: exactly the identity
    assert abs(J[orow][:, np.repeat(land, 6)]).sum() == 0.0                        # no ocean row couples to a LAND unknown (boundary.F90)
    Jc = J[orow][:, orow].tocsr()
    Jc.sort_indices()
    tol, maxit, restart = 1e-10, 40, 40
    kf = kref_gmres(rp, col, val, b, np.zeros(n), tol=tol, maxit=maxit, restart=restart, prec_kind=0)
    kc = kref_gmres(Jc.indptr.astype(np.int32), Jc.indices.astype(np.int32), Jc.data, b[orow], np.zeros(len(orow)), tol=tol, maxit=maxit,
                    restart=restart, prec_kind=0)
    k = min(len(kf["hist"]), len(kc["hist"]))
    assert k > 10 and abs(kf["iters"] - kc["iters"]) <= 1
    assert np.abs(kf["hist"][:k] - kc["hist"][:k]).max() <= 1e-12
    xs = np.zeros(n); xs[orow] = kc["x"]
    assert np.linalg.norm(xs - kf["x"]) <= 1e-10 * np.linalg.norm(kf["x"])
    assert len(orow) < n and (name != "global4deg" or len(orow) < 0.55 * n)        # 51.5 % of the unknowns on the real 4-degree mask


@pytest.mark.parametrize("name", ["gateway16", "global4deg", "box_p"])
def test_compact_spmv_index_maps(name):
    """The index arithmetic of spmv_compact_kernel / gather_cells / scatter_cells (thcm_linalg.cu), replayed in numpy on the library's
    cell maps and graph: compact row i -> full row 6*ocell[i/6] + i%6, full column c -> 6*ccell[c/6] + c%6 (skipped on LAND)."""
    s, landm, o, e = setup(name)
    x = cases.consistent_state(s, landm, scale=0.05)
    val, _ = o.jacobian_graph(x)
    rp, col = e.graph()
    ocell, ccell = e.cell_maps()
    n = e.ndim
    rng = np.random.default_rng(2)
    xf = rng.standard_normal(n)
    xf.reshape(-1, 6)[ccell < 0] = 0.0                       # a vector of the compact space: zero on LAND
    xc = xf.reshape(-1, 6)[ocell].reshape(-1)                # gather_cells
    yc = np.zeros(6 * len(ocell))
    for i in range(0, 6 * len(ocell), max(1, len(ocell) // 400)):
        ci, r = divmod(i, 6)
        row = 6 * ocell[ci] + r
        c = col[rp[row]:rp[row + 1]]
        cc = ccell[c // 6]
        keep = cc >= 0
        yc[i] = np.sum(val[rp[row]:rp[row + 1]][keep] * xc[6 * cc[keep] + c[keep] % 6])
        full = np.sum(val[rp[row]:rp[row + 1]] * xf[c])
        assert abs(yc[i] - full) <= 1e-13 * (abs(full) + 1e-300) + 1e-300
    back = np.zeros(n)                                       # scatter_cells
    back.reshape(-1, 6)[ccell >= 0] = xc.reshape(-1, 6)[ccell[ccell >= 0]]
    assert np.array_equal(back, xf)
