"""CPU tests of the parameter-list layer of the mirror (i-emic_b200/paramlist.py): the reference's Teuchos XML dialect, the
validate-and-default rules of THCM::getDefaultInitParameters / Ocean::getDefaultInitParameters, and the host part of the THCM
constructor (THCM.C:186-400, 640-795) -- checked against the hand-built cases the parity tests use, and (here, where /root/reference
exists) on every ocean_params.xml the reference ships."""
import glob
import math
import os

import numpy as np
import pytest

import cases
import iemic_b200
from iemic_b200 import paramlist as pl

HERE = os.path.dirname(os.path.abspath(__file__))
PARAMS = os.path.join(HERE, "golden", "params")
DATA = os.path.join(HERE, "golden")              # golden/masks stands in for <data dir>/mkmask
REFERENCE = "/root/reference"


@pytest.fixture
def data_dir(tmp_path):
    os.symlink(os.path.join(DATA, "masks"), tmp_path / "mkmask")
    return tmp_path


def same_settings(a, b, skip=("rank", "nranks", "device", "balance", "ymin_glob", "ymax_glob")):
    da, db = a.as_dict(), b.as_dict()
    return {k: (da[k], db[k]) for k in da if k not in skip and da[k] != db[k]}


def test_xml_dialect_and_round_trip():
    p = pl.read_xml(os.path.join(PARAMS, "natl8_integral_condition.xml"))
    assert p.name == "Ocean" and p["Save state"] is False
    t = p["THCM"]
    assert t["Periodic"] is False and t["Read Land Mask"] is True            # bool written as "0" / "1" (as the reference's files do)
    assert t["Global Grid-Size n"] == 8 and isinstance(t["Global Bound xmin"], float) and t["Land Mask"] == "mask_natl8"
    assert list(t["Starting Parameters"]) == ["Combined Forcing", "Salinity Forcing", "Temperature Forcing", "SPL1", "SPL2"]
    q = pl.from_xml_string(pl.to_xml_string(p))
    assert q == p and q["THCM"].name == "THCM"
    # numbers are read like `istringstream >> value`: leading number, rest ignored (test/ocean/continuation_params.xml: value="1.0-2")
    lenient = pl.from_xml_string('<ParameterList name="x"><Parameter name="a" type="double" value="1.0-2"/>'
                                 '<Parameter name="b" type="int" value=" 12 "/><Parameter name="c" type="double" value="2.0e3"/></ParameterList>')
    assert lenient == {"a": 1.0, "b": 12, "c": 2000.0}
    with pytest.raises(pl.InvalidParameter):
        pl.from_xml_string('<ParameterList name="x"><Parameter name="a" type="int" value="x1"/></ParameterList>')
    with pytest.raises(pl.InvalidParameter):
        pl.from_xml_string('<ParameterList name="x"><Parameter name="a" type="bool" value="maybe"/></ParameterList>')


def test_validation_follows_teuchos():
    d = pl.thcm_default_init_parameters()
    assert len(d["Starting Parameters"]) == 30 and all(math.isnan(v) for v in d["Starting Parameters"].values())
    assert d["Mixing"] == 1 and d["Rho Mixing"] is True and d["Topography"] == 1 and d["Salinity Integral Sign"] == -1
    p = pl.ParameterList("THCM", {"Mixing": 2})
    pl.validate_parameters_and_set_defaults(p, d)
    assert p["Mixing"] == 2 and p["Global Grid-Size n"] == 16 and len(p["Starting Parameters"]) == 30
    with pytest.raises(pl.InvalidParameter, match="not a valid parameter name"):      # e.g. parameterfiles/ocean_params.xml's stale key
        pl.validate_parameters_and_set_defaults(pl.ParameterList("THCM", {"Coupled Atmosphere": 0}), d)
    with pytest.raises(pl.InvalidParameter, match="has type"):
        pl.validate_parameters_and_set_defaults(pl.ParameterList("THCM", {"Depth hdim": 4000}), d)      # int where a double is declared
    sp = pl.ParameterList("THCM")
    sp.sublist("Starting Parameters")["Combined Forcings"] = 1.0
    with pytest.raises(pl.InvalidParameter):
        pl.validate_parameters_and_set_defaults(sp, d)
    o = pl.validate_parameters_and_set_defaults(pl.ParameterList("Ocean"), pl.ocean_default_init_parameters())
    assert o["Belos Solver"]["FGMRES iterations"] == 500 and o["THCM"]["Mixing"] == 1


def test_setup_equals_the_hand_built_test_case(data_dir):
    p = pl.read_xml(os.path.join(PARAMS, "natl8_integral_condition.xml"))
    su = pl.thcm_setup(p["THCM"], data_dir=data_dir)
    s, landm = cases.natl8(vmix=1, SRES=0)
    assert not same_settings(su["settings"], s)
    assert np.array_equal(su["landm"], landm)
    assert su["integral_condition"] == (7, 7, -1) and su["fix_pressure_points"] is False and su["scaling"] == "THCM"
    assert su["starting_parameters"] == [("Combined Forcing", 0.25), ("Salinity Forcing", 1.0), ("Temperature Forcing", 10.0),
                                         ("SPL1", 2000.0), ("SPL2", 0.01)]
    assert pl.solver_parameters(p["Belos Solver"]) == dict(tol=1e-6, restart=120, maxit=360, precon=1)


def test_setup_of_the_default_run():
    from test_oracle_pins import default_run_case
    su = pl.thcm_setup(pl.read_xml(os.path.join(PARAMS, "basin16_topography1.xml")))
    s, landm, pars = default_run_case()
    assert not same_settings(su["settings"], s)
    assert np.array_equal(su["landm"], landm)
    assert su["integral_condition"] is None
    assert {iemic_b200.par_index(k): v for k, v in su["starting_parameters"]} == {iemic_b200.par_index(k): v for k, v in pars.items()}


def test_constructor_rules(data_dir):
    base = pl.read_xml(os.path.join(PARAMS, "natl8_integral_condition.xml"))["THCM"]

    def variant(**kw):
        q = pl.from_xml_string(pl.to_xml_string(base))
        q.update(kw)
        return q
    su = pl.thcm_setup(variant(**{"Restoring Salinity Profile": 1, "Coupled Salinity": 1}), data_dir=data_dir)   # THCM.C:253-259
    assert su["settings"].SRES == 0 and su["settings"].coupled_S == 1 and su["integral_condition"] == (7, 7, -1)
    with pytest.raises(pl.InvalidParameter, match="integral sign"):
        pl.thcm_setup(variant(**{"Salinity Integral Sign": 2}), data_dir=data_dir)
    with pytest.raises(pl.InvalidParameter, match="land point"):                                                  # THCM.C:662-690
        pl.thcm_setup(variant(**{"Integral row coordinate i": 0, "Integral row coordinate j": 0}), data_dir=data_dir)
    with pytest.raises(pl.InvalidParameter, match="outside"):
        pl.thcm_setup(variant(**{"Global Bound ymax": 94.0}), data_dir=data_dir)
    with pytest.raises(FileNotFoundError):
        pl.thcm_setup(variant(**{"Land Mask": "no_such_mask"}), data_dir=data_dir)
    for kw in ({"Wind Forcing Type": 0}, {"Levitus T": 0}, {"Levitus S": 0, "Restoring Salinity Profile": 1}):
        with pytest.raises(pl.InvalidParameter, match="not ship"):                                                # data files outside the tree
            pl.thcm_setup(variant(**kw), data_dir=data_dir)
    pl.thcm_setup(variant(**{"Levitus S": 0}), data_dir=data_dir)          # with SRES = 0 the flux comes through the HDF5 interface: accepted
    su = pl.thcm_setup(variant(**{"Integral row coordinate i": 4, "Integral row coordinate j": 3, "Salinity Integral Sign": 1,
                                  "Fix Pressure Points": True}), data_dir=data_dir)
    assert su["integral_condition"] == (4, 3, 1) and su["fix_pressure_points"] is True


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="the reference tree is only present in the build container")
def test_every_ocean_list_the_reference_ships():
    """Each ocean_params.xml of the reference either sets up (its mask found below the reference's data directory) or is refused for a
    reason the reference itself would stop for (a stale parameter name, a mask or forcing data file that is not in its tree)."""
    files = sorted(glob.glob(os.path.join(REFERENCE, "**", "*ocean*.xml"), recursive=True))
    ocean = [f for f in files if "THCM" in pl.read_xml(f)]
    assert len(ocean) >= 20
    ok, refused = 0, []
    for f in ocean:
        try:
            su = pl.thcm_setup(pl.read_xml(f)["THCM"], data_dir=os.path.join(REFERENCE, "data"))
            s = su["settings"]
            assert su["landm"].shape == (s.L + 2, s.M + 2, s.N + 2)
            ok += 1
        except (pl.InvalidParameter, FileNotFoundError) as e:
            refused.append((os.path.relpath(f, REFERENCE), str(e)))
    # run/ocean/global asks for Levitus / Trenberth data and for mask.glo2, none of which is in the tree; parameterfiles/ holds a stale key
    assert ok >= 20 and all("Coupled Atmosphere" in m or "mask.glo2" in m or "do not ship" in m or "does not ship" in m for _, m in refused), refused
    # the three configurations the parity tests build by hand (tests/cases.py, tests/transient_twin.py)
    su = pl.thcm_setup(pl.read_xml(os.path.join(REFERENCE, "test/ocean/ocean_params.xml"))["THCM"], data_dir=os.path.join(REFERENCE, "data"))
    s, landm = cases.natl8(vmix=1, SRES=0)
    assert not same_settings(su["settings"], s) and np.array_equal(su["landm"], landm)
    su = pl.thcm_setup(pl.read_xml(os.path.join(REFERENCE, "test/ocean/reft_ocean_params.xml"))["THCM"], data_dir=os.path.join(REFERENCE, "data"))
    from test_oracle_pins import reft_case
    s, landm, pars, _ = reft_case()
    assert not same_settings(su["settings"], s) and np.array_equal(su["landm"], landm)
    start = {iemic_b200.par_index(k): v for k, v in su["starting_parameters"]}
    assert start == {iemic_b200.par_index(k): (0.0 if k == "COMB" else v) for k, v in pars.items()}     # the continuation then moves COMB to 0.02
    import transient_twin as tt
    su = pl.thcm_setup(pl.read_xml(os.path.join(REFERENCE, "test/ocean/test_oceantransient.xml"))["THCM"], data_dir=os.path.join(REFERENCE, "data"))
    s, landm = cases.natl8(**tt.SETTINGS)
    assert not same_settings(su["settings"], s) and np.array_equal(su["landm"], landm)
    assert {iemic_b200.par_index(k): v for k, v in su["starting_parameters"]} == {iemic_b200.par_index(k): v for k, v in tt.PARAMETERS.items()}


def test_cpp_mirror_reads_the_same_lists(data_dir):
    """include/thcm_paramlist.hpp (thcm_b200::ParameterList, the XML reader, setupFromParameterList): tests/cpp/test_paramlist.cpp,
    host-only -- defaults, the mask by name, the integral-condition rules, type checking, the XML dialect, the committed fixtures; and the
    C++ reader against the Python reader, entry by entry, on the fixtures and (here) on every ocean list of the reference."""
    import subprocess
    cpp = os.path.join(HERE, "cpp")
    subprocess.run(["make", "-C", cpp, "all"], check=True, capture_output=True)
    exe = os.path.join(cpp, "_bin", "test_paramlist")
    r = subprocess.run([exe, str(data_dir), PARAMS], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("PASS"), r.stdout + r.stderr

    def canonical(p, prefix=""):
        out = []
        for k, v in p.items():
            if isinstance(v, pl.ParameterList):
                out += canonical(v, prefix + k + "/")
            elif isinstance(v, bool):
                out.append(f"{prefix}{k} = bool {'true' if v else 'false'}")
            elif isinstance(v, int):
                out.append(f"{prefix}{k} = int {v}")
            elif isinstance(v, float):
                out.append(f"{prefix}{k} = double {v!r}")
            else:
                out.append(f"{prefix}{k} = string {v}")
        return out

    files = sorted(glob.glob(os.path.join(PARAMS, "*.xml")))
    if os.path.isdir(REFERENCE):
        files += sorted(glob.glob(os.path.join(REFERENCE, "**", "*ocean*.xml"), recursive=True))
    assert len(files) >= 2
    for f in files:
        r = subprocess.run([exe, "--dump", f], capture_output=True, text=True)
        assert r.returncode == 0, (f, r.stdout)
        got = []
        for line in r.stdout.splitlines():     # doubles: compare the values, not their spelling
            head, _, val = line.partition(" = double ")
            got.append(f"{head} = double {float(val)!r}" if val else line)
        assert got == canonical(pl.read_xml(f)), f
