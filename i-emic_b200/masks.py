"""Land-mask I/O and synthetic masks.

File format = the reference's ``readmask`` (src/ocean/topo.F90:41-64): for every level
k = 0..l+1 one header line, then rows j = m+1 .. 0, each (n+2) digits i = 0..n+1 with values
OCEAN 0 / LAND 1 / WATER 2 / PERIO 3 (par.F90:78-81).  Arrays here are ``int32[l+2, m+2, n+2]``
(i fastest in memory), i.e. exactly the C-order array THCM.C hands to ``init_``.
"""
import numpy as np

OCEAN, LAND, WATER, PERIO = 0, 1, 2, 3


def read_mask(path, n, m, l, fix_inversion=True):
    landm = np.full((l + 2, m + 2, n + 2), LAND, dtype=np.int32)
    with open(path) as f:
        lines = f.read().split("\n")
    pos = 0
    for k in range(l + 2):
        pos += 1  # header line
        for j in range(m + 1, -1, -1):
            row = lines[pos]
            pos += 1
            vals = np.frombuffer(row.encode()[: n + 2], dtype=np.uint8) - ord("0")
            landm[k, j, : len(vals)] = vals
    if fix_inversion:  # topo.F90:94-103: no ocean below land
        for k in range(l, 1, -1):
            inv = (landm[k, 1:m + 1, 1:n + 1] == LAND) & (landm[k - 1, 1:m + 1, 1:n + 1] == OCEAN)
            landm[k - 1, 1:m + 1, 1:n + 1][inv] = LAND
    return landm


def write_mask(path, landm):
    l2, m2, n2 = landm.shape
    with open(path, "w") as f:
        for k in range(l2):
            f.write(f"level = {k:8d} _________________________\n")
            for j in range(m2 - 1, -1, -1):
                f.write("".join(str(int(v)) for v in landm[k, j]) + "\n")


def all_ocean_mask(n, m, l, periodic=False):
    landm = np.full((l + 2, m + 2, n + 2), LAND, dtype=np.int32)
    landm[1:l + 1, 1:m + 1, 1:n + 1] = OCEAN
    if periodic:
        landm[1:l + 1, 1:m + 1, 0] = PERIO
        landm[1:l + 1, 1:m + 1, n + 1] = PERIO
    return landm


def synthetic_global_mask(base, n, m, l, periodic=True):
    """Nearest-neighbour resampling of a real mask (e.g. the 4-degree 96x38x12 global mask) to n x m x l,
    followed by the reference's 'no ocean below land' fix (topo.F90:94-103) and the PERIO border rule
    (topo.F90:312-318).  Seed-free and reproducible: used for the 2, 1 and 0.5 degree benchmark grids
    (SURVEY.md section 8d)."""
    l0, m0, n0 = base.shape[0] - 2, base.shape[1] - 2, base.shape[2] - 2
    ki = np.minimum((np.arange(l) * l0) // l, l0 - 1) + 1
    ji = np.minimum((np.arange(m) * m0) // m, m0 - 1) + 1
    ii = np.minimum((np.arange(n) * n0) // n, n0 - 1) + 1
    inner = base[np.ix_(ki, ji, ii)]
    inner = np.where(inner == LAND, LAND, OCEAN).astype(np.int32)
    landm = np.full((l + 2, m + 2, n + 2), LAND, dtype=np.int32)
    landm[1:l + 1, 1:m + 1, 1:n + 1] = inner
    for k in range(l, 1, -1):
        inv = (landm[k, 1:m + 1, 1:n + 1] == LAND) & (landm[k - 1, 1:m + 1, 1:n + 1] == OCEAN)
        landm[k - 1, 1:m + 1, 1:n + 1][inv] = LAND
    if periodic:
        both = (landm[:, :, 1] == OCEAN) & (landm[:, :, n] == OCEAN)
        landm[:, :, 0][both] = PERIO
        landm[:, :, n + 1][both] = PERIO
    return landm
