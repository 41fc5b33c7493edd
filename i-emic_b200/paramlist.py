"""The reference's parameter lists (Teuchos::ParameterList XML files such as run/ocean/ocean_params.xml and solver_params.xml) for the
host-side mirror: read the XML, validate it against the defaults of ``THCM::getDefaultInitParameters`` (src/ocean/THCM.C:2697-2770) /
``Ocean::getDefaultInitParameters`` (src/ocean/Ocean.C:2262-2300) the way ``validateParametersAndSetDefaults`` does (unknown names and
wrong types are errors, missing entries take the default), and translate the THCM list into what the library is created from --
``Settings``, the global land mask, the integral-condition / pressure-point options and the starting parameters -- following the THCM
constructor statement by statement (THCM.C:186-340, 640-795).  No CUDA call is made here: ``thcm_setup`` runs on any host.
"""
import math
import os
import re
import xml.etree.ElementTree as ET

from .params import PAR_NAMES, par_index


class ParameterList(dict):
    """A named, nested dict: values are bool / int / float / str or another ParameterList (a sublist)."""

    def __init__(self, name="ANONYMOUS", *a, **kw):
        super().__init__(*a, **kw)
        self.name = name

    def sublist(self, name):
        if name not in self:
            self[name] = ParameterList(name)
        if not isinstance(self[name], ParameterList):
            raise TypeError(f'"{name}" is a parameter, not a sublist, of "{self.name}"')
        return self[name]


class InvalidParameter(ValueError):
    """Teuchos::Exceptions::InvalidParameterName / InvalidParameterType."""


_BOOL = {"true": True, "1": True, "false": False, "0": False}
_LEADING_INT = re.compile(r"\s*[+-]?\d+")
_LEADING_FLOAT = re.compile(r"\s*[+-]?(\d+\.?\d*([eE][+-]?\d+)?|\.\d+([eE][+-]?\d+)?|inf|nan)", re.IGNORECASE)


def _value(typ, text, where):
    t = typ.strip().lower()
    try:
        if t == "bool":
            return _BOOL[text.strip().lower()]
        # numbers are extracted the way `std::istringstream >> value` does: the longest leading number, trailing characters ignored
        # (the reference's test/ocean/continuation_params.xml holds value="1.0-2", which Teuchos reads as 1.0)
        if t in ("int", "long", "unsigned int", "short", "long long"):
            return int(_LEADING_INT.match(text).group(0))
        if t in ("double", "float"):
            return float(_LEADING_FLOAT.match(text).group(0))
        if t in ("string", "char"):
            return text
    except (KeyError, ValueError, AttributeError):
        raise InvalidParameter(f'{where}: cannot read "{text}" as {typ}') from None
    raise InvalidParameter(f"{where}: unsupported parameter type {typ!r}")


def _from_element(el, where):
    if el.tag != "ParameterList":
        raise InvalidParameter(f"{where}: expected <ParameterList>, found <{el.tag}>")
    pl = ParameterList(el.get("name", "ANONYMOUS"))
    for ch in el:
        name = ch.get("name")
        if ch.tag == "ParameterList":
            pl[name] = _from_element(ch, f"{where}/{name}")
        elif ch.tag == "Parameter":
            if name is None or ch.get("type") is None or ch.get("value") is None:
                raise InvalidParameter(f"{where}: a <Parameter> needs name, type and value")
            pl[name] = _value(ch.get("type"), ch.get("value"), f"{where}/{name}")
        # (Teuchos ignores anything else, e.g. <Validators>)
    return pl


def read_xml(path):
    """Teuchos::updateParametersFromXmlFile into an empty list."""
    return _from_element(ET.parse(path).getroot(), os.path.basename(path))


def from_xml_string(text):
    return _from_element(ET.fromstring(text), "<string>")


def to_xml_string(pl, indent=0):
    pad = "  " * indent
    out = [f'{pad}<ParameterList name="{pl.name}">']
    for k, v in pl.items():
        if isinstance(v, ParameterList):
            out.append(to_xml_string(v, indent + 1))
        else:
            typ = "bool" if isinstance(v, bool) else "int" if isinstance(v, int) else "double" if isinstance(v, float) else "string"
            val = ("true" if v else "false") if isinstance(v, bool) else repr(v) if isinstance(v, float) else str(v)
            out.append(f'{pad}  <Parameter name="{k}" type="{typ}" value="{val}"/>')
    out.append(f"{pad}</ParameterList>")
    return "\n".join(out)


def thcm_default_parameters():
    """THCM::getDefaultParameters (THCM.C:2772-2787): the 30 continuation parameters by XML name, NaN = "keep stpnt's value"."""
    pl = ParameterList("THCM Default Parameters")
    sp = pl.sublist("Starting Parameters")
    for i in range(1, 31):
        sp[PAR_NAMES[i]] = math.nan
    return pl


def thcm_default_init_parameters():
    """THCM::getDefaultInitParameters (THCM.C:2697-2770)."""
    pl = thcm_default_parameters()
    pl.name = "THCM Default Init Parameters"
    pl.update({
        "Problem Description": "Unnamed",
        "Global Grid-Size n": 16, "Global Grid-Size m": 16, "Global Grid-Size l": 16,
        "Global Bound xmin": 286.0, "Global Bound xmax": 350.0, "Global Bound ymin": 10.0, "Global Bound ymax": 74.0,
        "Periodic": False, "Depth hdim": 4000.0, "Grid Stretching qz": 1.0, "Topography": 1, "Flat Bottom": False,
        "Compute salinity integral": True, "Read Land Mask": False, "Land Mask": "no_mask_specified",
        "Inhomogeneous Mixing": 0, "Mixing": 1, "Rho Mixing": True, "Taper": 1,
        "Linear EOS: alpha T": 1.0e-4, "Linear EOS: alpha S": 7.6e-4,
        "Restoring Temperature Profile": 1, "Restoring Salinity Profile": 1, "Local SRES Only": False, "Salinity Integral Sign": -1,
        "Levitus T": 1, "Levitus S": 1, "Levitus Internal T/S": False,
        "Coupled Temperature": 0, "Coupled Salinity": 0, "Coupled Sea Ice Mask": 1, "Fix Pressure Points": False,
        "Coriolis Force": 1, "Forcing Type": 0,
        "Read Salinity Perturbation Mask": False, "Salinity Perturbation Mask": "no_mask_specified",
        "Wind Forcing Type": 2, "Wind Forcing Data": "wind/trtau.dat",
        "Temperature Forcing Data": "levitus/new/t00an1", "Salinity Forcing Data": "levitus/new/s00an1",
        "Integral row coordinate i": -1, "Integral row coordinate j": -1, "Scaling": "THCM",
    })
    return pl


def ocean_default_init_parameters():
    """Ocean::getDefaultInitParameters (Ocean.C:2262-2300): I/O switches (accepted, not acted on by the mirror), the Belos solver sublist
    and the THCM sublist."""
    pl = ParameterList("Ocean Default Init Parameters")
    pl.update({"Input file": "ocean_input.h5", "Output file": "ocean_output.h5", "Save mask": True, "Load mask": True,
               "Load state": False, "Save state": True, "Save frequency": 0, "Load salinity flux": False, "Save salinity flux": True,
               "Load temperature flux": False, "Save temperature flux": True, "Use legacy fort.3 output": False,
               "Use legacy fort.44 output": True, "Save column integral": False, "Max mask fixes": 5, "Analyze Jacobian": True})
    bs = pl.sublist("Belos Solver")
    bs.update({"FGMRES iterations": 500, "FGMRES tolerance": 1e-8, "FGMRES restarts": 0, "FGMRES output": 100,
               "FGMRES explicit residual test": False})
    pl["THCM"] = thcm_default_init_parameters()
    pl["THCM"].name = "THCM"
    return pl


def validate_parameters_and_set_defaults(params, defaults, where=None):
    """Teuchos::ParameterList::validateParametersAndSetDefaults: every name in `params` must exist in `defaults` with the same type
    (sublists recursively); entries of `defaults` that `params` lacks are added.  Returns `params`."""
    where = where or params.name
    for k, v in params.items():
        if k not in defaults:
            raise InvalidParameter(f'the parameter "{k}" in the list "{where}" is not a valid parameter name')
        d = defaults[k]
        if isinstance(d, ParameterList) != isinstance(v, ParameterList):
            raise InvalidParameter(f'"{k}" in "{where}": a sublist and a parameter cannot stand in for each other')
        if isinstance(v, ParameterList):
            validate_parameters_and_set_defaults(v, d, f"{where}->{k}")
        elif type(v) is not type(d):
            raise InvalidParameter(f'the parameter "{k}" in the list "{where}" has type {type(v).__name__}, expected {type(d).__name__}')
    for k, d in defaults.items():
        if k not in params:
            params[k] = validate_parameters_and_set_defaults(ParameterList(k), d, f"{where}->{k}") if isinstance(d, ParameterList) else d
    return params


def validate_parameters(params, defaults, where=None):
    """Teuchos::ParameterList::validateParameters: names and types checked as above, nothing added."""
    where = where or params.name
    for k, v in params.items():
        if k not in defaults:
            raise InvalidParameter(f'the parameter "{k}" in the list "{where}" is not a valid parameter name')
        d = defaults[k]
        if isinstance(d, ParameterList) != isinstance(v, ParameterList):
            raise InvalidParameter(f'"{k}" in "{where}": a sublist and a parameter cannot stand in for each other')
        if isinstance(v, ParameterList):
            validate_parameters(v, d, f"{where}->{k}")
        elif type(v) is not type(d):
            raise InvalidParameter(f'the parameter "{k}" in the list "{where}" has type {type(v).__name__}, expected {type(d).__name__}')
    return params


def thcm_setup(thcm_params, rank=0, nranks=1, device=0, balance=0, data_dir=None):
    """The host part of the THCM constructor (THCM.C:186-400, 640-760) for the list `thcm_params` (the "THCM" sublist of the ocean list):
    returns a dict with
      settings        -- Settings for thcmb_create (bounds in radians, THCM.C:203-206)
      landm           -- the GLOBAL land mask [l+2, m+2, n+2] of m_global::get_landm ("Read Land Mask" / "Land Mask", else "Topography")
      spert           -- the salinity perturbation mask [m, n] of m_global::get_spert, None unless "Read Salinity Perturbation Mask"
      integral_condition -- None, or (Nic, Mic, sign) when "Restoring Salinity Profile" is 0 (THCM.C:653-697; the cell must be ocean)
      fix_pressure_points, scaling, starting_parameters (the non-NaN ones, in list order), params (validated, defaults filled in).
    Uses the library's m_global symbols for the mask (host code), so the mask is the one the B1 boundary hands to THCM.C."""
    from . import thcm as _t
    import numpy as np
    p = validate_parameters_and_set_defaults(thcm_params, thcm_default_init_parameters())
    n, m, l = p["Global Grid-Size n"], p["Global Grid-Size m"], p["Global Grid-Size l"]
    sres, coupled_s = p["Restoring Salinity Profile"], p["Coupled Salinity"]
    if coupled_s == 1 and sres == 1:      # THCM.C:253-259: incompatible, SRES is switched off (with a warning)
        sres = 0
    if abs(p["Salinity Integral Sign"]) != 1:
        raise InvalidParameter("Invalid integral sign!")                                   # THCM.C:265-268
    for k, lo, hi in (("Global Bound xmin", -360.0, 360.0), ("Global Bound xmax", -360.0, 360.0),
                      ("Global Bound ymin", -90.0, 90.0), ("Global Bound ymax", -90.0, 90.0)):
        if not lo <= p[k] <= hi:                                                           # the validators of THCM.C:2714-2724
            raise InvalidParameter(f'"{k}" = {p[k]} is outside [{lo}, {hi}]')
    s = _t.Settings.from_degrees(n, m, l, p["Global Bound xmin"], p["Global Bound xmax"], p["Global Bound ymin"], p["Global Bound ymax"],
                                 periodic=p["Periodic"], hdim=p["Depth hdim"], qz=p["Grid Stretching qz"],
                                 ih=p["Inhomogeneous Mixing"], vmix=p["Mixing"], tap=p["Taper"], rho_mixing=int(p["Rho Mixing"]),
                                 coriolis_on=p["Coriolis Force"], TRES=p["Restoring Temperature Profile"], SRES=sres,
                                 iza=p["Wind Forcing Type"], ite=p["Levitus T"], its=p["Levitus S"],
                                 coupled_T=p["Coupled Temperature"], coupled_S=coupled_s, forcing_type=p["Forcing Type"],
                                 alphaT=p["Linear EOS: alpha T"], alphaS=p["Linear EOS: alpha S"],
                                 rank=rank, nranks=nranks, device=device, balance=balance)
    if p["Levitus Internal T/S"]:
        raise InvalidParameter('"Levitus Internal T/S" reads Levitus data files that do not ship with the reference; '
                               "hand the fields to set_internal_forcing instead")
    # the data-file options of m_global::get_windfield / get_temforcing / get_salforcing (global.F90:425-560): Trenberth winds and Levitus
    # surface fields are not part of the reference tree -- the same conditions under which the library's getters fail loudly
    if s.iza < 2:
        raise InvalidParameter('"Wind Forcing Type" 0 / 1 reads wind/trtau.dat, which does not ship with the reference; insert taux / tauy')
    if s.coupled_T == 0 and s.ite == 0 and s.TRES != 0:
        raise InvalidParameter('"Levitus T" = 0 with a restoring temperature profile reads Levitus data that do not ship with the reference')
    if s.coupled_S == 0 and s.its == 0 and s.SRES != 0:
        raise InvalidParameter('"Levitus S" = 0 with a restoring salinity profile reads Levitus data that do not ship with the reference')
    for flag, key in (("Read Land Mask", "Land Mask"), ("Read Salinity Perturbation Mask", "Salinity Perturbation Mask")):
        if p[flag]:   # (the Fortran symbols end the process on a missing file, like the reference: find out before calling them)
            base = data_dir if data_dir is not None else os.environ.get("THCM_DATA_DIR", ".")
            tried = [p[key], os.path.join(str(base), "mkmask", p[key])]
            if not any(os.path.isfile(t) for t in tried):
                raise FileNotFoundError(f'"{key}": none of {tried} exists')
    old = os.environ.get("THCM_DATA_DIR")
    if data_dir is not None:
        os.environ["THCM_DATA_DIR"] = str(data_dir)
    try:
        f = _t.FortranABI()
        f.global_initialize(s, maskfile=p["Land Mask"].encode() if p["Read Land Mask"] else b"", itopo=p["Topography"],
                            flat=p["Flat Bottom"],
                            spertmaskfile=p["Salinity Perturbation Mask"].encode() if p["Read Salinity Perturbation Mask"] else b"")
        landm = f.global_get_landm()
        spert = f.global_get_spert() if p["Read Salinity Perturbation Mask"] else None
    finally:
        if data_dir is not None:
            if old is None:
                del os.environ["THCM_DATA_DIR"]
            else:
                os.environ["THCM_DATA_DIR"] = old
    ic = None
    if sres == 0:
        nic = n - 1 if p["Integral row coordinate i"] == -1 else p["Integral row coordinate i"]
        mic = m - 1 if p["Integral row coordinate j"] == -1 else p["Integral row coordinate j"]
        if landm[l, mic + 1, nic + 1] != 0:                                                # THCM.C:662-690
            raise InvalidParameter(f"Integral row coordinates ({nic},{mic}) give a land point! Please give better coordinates in xml.")
        ic = (nic, mic, p["Salinity Integral Sign"])
    if p["Scaling"] not in ("THCM", "None"):
        raise InvalidParameter(f'unknown "Scaling" {p["Scaling"]!r} (the reference supports "THCM" and "None")')
    start = [(k, v) for k, v in p["Starting Parameters"].items() if not (isinstance(v, float) and math.isnan(v))]
    for k, _ in start:
        par_index(k)
    return dict(settings=s, landm=np.ascontiguousarray(landm), spert=spert, integral_condition=ic,
                fix_pressure_points=bool(p["Fix Pressure Points"]), scaling=p["Scaling"], starting_parameters=start, params=p)


def solver_parameters(belos_params):
    """The "Belos Solver" sublist (Ocean.C:985-1012) as the arguments of the library's FGMRES: "FGMRES iterations" is Belos' Num Blocks
    (the basis length = restart), the iteration limit is Num Blocks x (Maximum Restarts + 1)."""
    d = ocean_default_init_parameters()["Belos Solver"]
    p = validate_parameters_and_set_defaults(belos_params, d)
    restart = p["FGMRES iterations"]
    return dict(tol=p["FGMRES tolerance"], restart=restart, maxit=restart * (p["FGMRES restarts"] + 1), precon=1)
