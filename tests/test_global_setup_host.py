"""CPU tests of m_global's setup symbols in the product library (host code of the B1 boundary, no kernel involved): what THCM.C gets
back from m_global::get_landm / get_spert after m_global::initialize (THCM.C:325-400) for the land-mask options of the reference's
parameter list -- "Read Land Mask" + "Land Mask" (readmask, topo.F90:41-127) and the idealised "Topography" cases 1..4 + "Flat Bottom"
(depth3land, topo.F90:129-330; the reference's default run uses Topography = 1) -- and "Read Salinity Perturbation Mask"
(read_spertm, forcing.F90:372-402).  Checked against numpy restatements written from those lines."""
import os
import subprocess
import sys

import numpy as np
import pytest

import iemic_b200
from iemic_b200 import masks

OCEAN, LAND, PERIO = masks.OCEAN, masks.LAND, masks.PERIO
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "masks")
PI = 3.14159265358979323846


def settings(n, m, l, xmin=286.0, xmax=350.0, ymin=10.0, ymax=74.0, periodic=False, **kw):
    return iemic_b200.Settings.from_degrees(n, m, l, xmin, xmax, ymin, ymax, periodic=periodic, **kw)


def depth3land(s, itopo, flat=False):
    """topo.F90:129-330 with depth = 0 (no bathymetry data): [l+2, m+2, n+2]."""
    n, m, l = s.N, s.M, s.L
    lm = np.full((l + 2, m + 2, n + 2), LAND, dtype=np.int32)
    lm[1:l + 1, 1:m + 1, 1:n + 1] = OCEAN
    if itopo == 2:
        x = (np.arange(1, n + 1) - 0.5) * (s.xmax - s.xmin) / n + s.xmin
        y = (np.arange(1, m + 1) - 0.5) * (s.ymax - s.ymin) / m + s.ymin
        ph1, ph2, ph3, ph4 = 250 * PI / 180, 315 * PI / 180, 10 * PI / 180, 65 * PI / 180
        thd, thsa, thn, tha = -60 * PI / 180, -35 * PI / 180, 10 * PI / 180, 30 * PI / 180
        am, af = (x < ph2) & (x > ph1), (x < ph4) & (x > ph3)
        for xs, ys in ((am, (y < 0) & (y > thd)), (af, (y < thn) & (y > thsa)), (am, (y < s.ymax) & (y > tha)), (af, (y < s.ymax) & (y > tha))):
            lm[1:l + 1, 1:m + 1, 1:n + 1][:, np.outer(ys, xs)] = LAND
    elif itopo == 3:
        lm[1:l + 1, 1:17, 18:21] = LAND
    elif itopo == 4:
        lm[1:l + 1, 6:m + 1, 22:25] = LAND
    if flat:
        lm[1:l] = lm[l]
    if s.periodic:
        both = (lm[:, :, 1] == OCEAN) & (lm[:, :, n] == OCEAN)
        lm[:, :, 0][both] = PERIO
        lm[:, :, n + 1][both] = PERIO
    return lm


@pytest.mark.parametrize("itopo,dims,kw", [
    (1, (16, 16, 16), dict(xmin=300.0, xmax=340.0, ymin=20.0, ymax=60.0)),       # run/ocean/ocean_params.xml: the reference's default run
    (1, (12, 7, 3), dict(xmin=0.0, xmax=360.0, ymin=-80.0, ymax=80.0, periodic=True)),
    (2, (96, 38, 12), dict(xmin=0.0, xmax=360.0, ymin=-85.5, ymax=85.5, periodic=True)),
    (2, (45, 30, 2), dict(xmin=0.0, xmax=360.0, ymin=-75.0, ymax=75.0)),
    (3, (24, 16, 4), dict()),
    (4, (40, 30, 16), dict(xmin=100.0, xmax=260.0, ymin=-60.0, ymax=60.0)),      # the resolution topo.F90:268 names
])
@pytest.mark.parametrize("flat", [False, True])
def test_topography_cases_of_depth3land(itopo, dims, kw, flat):
    s = settings(*dims, **kw)
    f = iemic_b200.FortranABI()
    f.global_initialize(s, itopo=itopo, flat=flat)
    got = f.global_get_landm()
    want = depth3land(s, itopo, flat)
    assert np.array_equal(got, want)
    assert (got[:, 0, :] == LAND).all() and (got[:, -1, :] == LAND).all() and (got[0] == LAND).all() and (got[-1] == LAND).all()
    if itopo == 2 and dims[0] == 96:
        assert 0.2 < (got[1:-1, 1:-1, 1:-1] == LAND).mean() < 0.5     # four continents, not an empty or a full basin
    if kw.get("periodic"):
        assert (got[1:-1, 1:-1, 0] == PERIO).any()
    # the same mask again on the next call (every get_landm runs topofit, global.F90:308)
    assert np.array_equal(f.global_get_landm(), want)


@pytest.mark.parametrize("name,dims,periodic", [("mask_natl8", (8, 8, 4), False), ("mask_gateway", (16, 16, 16), True),
                                                 ("mask_global_96x38x12", (96, 38, 12), True), ("test6x6x4", (6, 6, 4), False)])
@pytest.mark.parametrize("by_name", [False, True])
def test_read_land_mask(name, dims, periodic, by_name, monkeypatch, tmp_path):
    """readmask: the file as a path, or as a bare name below <data dir>/mkmask (global.F90 locate_file) -- equal to the Python reader
    (which the oracle-checked cases use), land inversion fix included; "Flat Bottom" copies the surface level down (topo.F90:105-109)."""
    path = os.path.join(GOLDEN, name)
    s = settings(*dims, periodic=periodic)
    f = iemic_b200.FortranABI()
    if by_name:
        os.makedirs(tmp_path / "mkmask")
        os.symlink(path, tmp_path / "mkmask" / name)
        monkeypatch.setenv("THCM_DATA_DIR", str(tmp_path))
        f.global_initialize(s, maskfile=name.encode())
    else:
        f.global_initialize(s, maskfile=path.encode())
    want = masks.read_mask(path, *dims)
    assert np.array_equal(f.global_get_landm(), want)
    f.global_initialize(s, maskfile=path.encode(), flat=True)
    flat = want.copy()
    flat[1:dims[2]] = flat[dims[2]]
    assert np.array_equal(f.global_get_landm(), flat)


def test_salinity_perturbation_mask(tmp_path):
    """get_spert: SRES everywhere without a mask file; with one, (1 - digit) * (1 - landm(i,j,l)) from rows j = m+1 .. 0 of n+2 digits."""
    n, m, l = 8, 8, 4
    s = settings(n, m, l, SRES=0)
    f = iemic_b200.FortranABI()
    path = os.path.join(GOLDEN, "mask_natl8")
    f.global_initialize(s, maskfile=path.encode())
    landm = f.global_get_landm()
    assert np.array_equal(f.global_get_spert(), np.zeros((m, n)))
    s.SRES = 1
    f.global_initialize(s, maskfile=path.encode())
    f.global_get_landm()
    assert np.array_equal(f.global_get_spert(), np.ones((m, n)))
    rng = np.random.default_rng(5)
    dum = rng.integers(0, 2, size=(m + 2, n + 2))
    pert = tmp_path / "pertmask.txt"
    with open(pert, "w") as fh:
        for j in range(m + 1, -1, -1):
            fh.write("".join(str(v) for v in dum[j]) + "\n")
    f.global_initialize(s, maskfile=path.encode(), spertmaskfile=str(pert).encode())
    f.global_get_landm()
    want = (1 - dum[1:m + 1, 1:n + 1]) * (1 - landm[l, 1:m + 1, 1:n + 1])
    got = f.global_get_spert()
    assert np.array_equal(got, want.astype(float)) and 0 < got.sum() < n * m


def test_topography_from_data_is_refused_loudly():
    """"Topography" = 0 fits bathymetry data that does not ship with the reference (its own depth3land stops): no silent all-ocean mask."""
    code = ("import iemic_b200\n"
            "s = iemic_b200.Settings.from_degrees(8, 8, 4, 286.0, 350.0, 10.0, 74.0)\n"
            "f = iemic_b200.FortranABI(); f.global_initialize(s, itopo=0); f.global_get_landm()\n")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], cwd=root, capture_output=True, text=True)
    assert r.returncode != 0 and "Topography" in r.stderr
