"""numpy restatement of the reference's surface diagnostics src/ocean/probe.F90 (TEST INFRASTRUCTURE -- never imported by
the product).  Every function takes an OracleTHCM (grid, land mask, surface fields, coupling constants), the parameter
vector and the state `un`, and follows the reference's expressions term by term (file:line in each docstring).  Surface
values T(i,j,l), S(i,j,l) of usol are the raw unknowns of the cells 1..n x 1..m (usrc.F90:1031-1042)."""
import numpy as np

RHODIM, T0, DELTAT, S0 = 1.024e+03, 15.0, 1.0, 35.0   # usr.F90:132-160
COMB, SALT, TEMP, SUNP, BIOT = 19, 15, 17, 10, 18     # par.F90:38-67


def _surface(o, un):
    n, m, l = o.n, o.m, o.l
    u = np.asarray(un).reshape(l, m, n, 6)
    land = o.landm()[l, 1:m + 1, 1:n + 1]
    return u[l - 1, :, :, 4], u[l - 1, :, :, 5], land


def qint(o, field):
    """forcing.F90:452-464 (THCM.C:2653-2686 on one rank)."""
    n, m, l = o.n, o.m, o.l
    y = o.grid()["y"]
    land = o.landm()[l, 1:m + 1, 1:n + 1]
    lf = ls = 0.0
    for j in range(m):
        for i in range(n):
            lf = field[j, i] * np.cos(y[j + 1]) * (1 - land[j, i]) + lf
            ls = np.cos(y[j + 1]) * (1 - land[j, i]) + ls
    return lf / ls


def compute_evap(o, un, coupled):
    """probe.F90:75-113."""
    T, _, land = _surface(o, un)
    c = o.coupling_state()
    out = np.zeros((o.m, o.n))
    if not coupled:
        return out
    ev = c["eo0"] + c["eta"] * c["qdim"] * (((DELTAT / c["qdim"]) * c["dqso"] * T - o.get_field("qatm")))
    out[land == 0] = ev[land == 0]
    return out


def get_salflux(o, un, coupled_S, SRES):
    """probe.F90:177-245: (salflux, correction, qsoaflux, qsosflux)."""
    T, S, land = _surface(o, un)
    c = o.coupling_state()
    gamma = o.getpar(COMB) * o.getpar(SALT)
    pQSnd = o.getpar(COMB) * o.getpar(SALT) * c["QSnd"]
    msi, qsa, qatm, patm, emip = (o.get_field(k) for k in ("msi", "qsa", "qatm", "patm", "emip"))
    QSos = pQSnd * (c["zeta"] * (c["a0"] * (S0 + S) - (T0 + T)) - (c["Qvar"] * qsa + c["Q0"])) / (RHODIM * c["Lf"])
    QSoa = pQSnd * (c["eo0"] + c["eta"] * c["qdim"] * ((DELTAT / c["qdim"]) * c["dqso"] * T - qatm) - patm)
    qsoaflux = QSoa / c["QSnd"] * (1 - msi)
    qsosflux = QSos / c["QSnd"] * msi
    salflux = np.zeros_like(T)
    if gamma != 0:
        if coupled_S:
            salflux = (QSoa + msi * (QSos - QSoa)) * (1 - land) / gamma
        else:
            salflux = (1 - land) * (1 - SRES + SRES * o.getpar(BIOT)) * emip - SRES * o.getpar(BIOT) * S / gamma
    corr = qint(o, salflux)
    return salflux - corr, corr * gamma, qsoaflux, qsosflux


def get_temflux(o, un, coupled_T, TRES):
    """probe.F90:247-351: dict of the six n*m fields; entries of non-OCEAN surface cells are not written (left 0 here)."""
    T, S, land = _surface(o, un)
    c = o.coupling_state()
    etabi = o.getpar(COMB) * o.getpar(TEMP)
    dedt = c["eta"] * c["qdim"] * (DELTAT / c["qdim"]) * c["dqso"]
    dedq = -c["eta"] * c["qdim"]
    msi, albe, tatm, qatm = (o.get_field(k) for k in ("msi", "albe", "tatm", "qatm"))
    QSW = o.getpar(COMB) * o.getpar(SUNP) * c["suno"][:, None] * (1 - c["albe0"] - c["albed"] * albe)
    QSH = c["Ooa"] * (T - tatm)
    QLH = c["lvsc"] * (c["eo0"] + dedt * T + dedq * qatm)
    QToa = QSW - QSH - QLH
    QTos = c["QTnd"] * c["zeta"] * (c["a0"] * (S0 + S) - (T0 + T))
    ocean = land == 0
    out = {"swflux": QSW / c["QTnd"] * (1 - msi), "shflux": -QSH / c["QTnd"] * (1 - msi), "lhflux": -QLH / c["QTnd"] * (1 - msi),
           "siflux": QTos / c["QTnd"] * msi, "simask": msi.copy()}
    if not coupled_T:
        out["totflux"] = (1 - TRES + TRES * o.getpar(BIOT)) * tatm - TRES * o.getpar(BIOT) * T / etabi
    else:
        out["totflux"] = (1 - land) * QToa / c["QTnd"] + msi * (QTos - QToa) / c["QTnd"]
    return {k: np.where(ocean, v, 0.0) for k, v in out.items()}


def get_derivatives(o, un, coupled_T, coupled_S):
    """probe.F90:371-438: (dftdm, dfsdq, dfsdm, dfsdg)."""
    To, So, land = _surface(o, un)
    c = o.coupling_state()
    pQSnd = o.getpar(COMB) * o.getpar(SALT) * c["QSnd"]
    msi, albe, tatm, qatm, patm, qsa = (o.get_field(k) for k in ("msi", "albe", "tatm", "qatm", "patm", "qsa"))
    z = np.zeros_like(To)
    dftdm, dfsdq, dfsdm, dfsdg = z.copy(), z.copy(), z.copy(), z.copy()
    ocean = land == 0
    if coupled_T:
        QTos = c["QTnd"] * c["zeta"] * (c["a0"] * (So + S0) - (To + T0))
        QToa = (o.getpar(COMB) * o.getpar(SUNP) * c["suno"][:, None] * (1 - c["albe0"] - c["albed"] * albe) - c["Ooa"] * (To - tatm)
                - c["lvsc"] * c["eta"] * c["qdim"] * (DELTAT / c["qdim"] * c["dqso"] * To - qatm) - c["lvsc"] * c["eo0"])
        dftdm = np.where(ocean, QTos - QToa, 0.0)
    if coupled_S:
        dfsdq = np.where(ocean, -pQSnd * c["Qvar"] / (RHODIM * c["Lf"]) * msi, 0.0)
        QSos = (c["zeta"] * (c["a0"] * (S0 + So) - (T0 + To)) - (c["Qvar"] * qsa + c["Q0"])) / (RHODIM * c["Lf"])
        QSoa = c["eo0"] + c["eta"] * c["qdim"] * ((DELTAT / c["qdim"]) * c["dqso"] * To - qatm) - patm
        dfsdm = np.where(ocean, pQSnd * (QSos - QSoa), 0.0)
        dfsdg = np.where(ocean, -1.0, 0.0)
    return dftdm, dfsdq, dfsdm, dfsdg
