"""Extracts the golden grid arrays the reference ships for its domain test (src/tests/test_domain.C:40-133, fixture
test/domain/domain_values.hdf5): for qz = 1 and qz = 2 the arrays x(16), xu(17), y(16), yv(17), z(8), zw(9) of the 16 x 16 x 8 grid on
[286, 350] x [10, 74] degrees, hdim = 4000 -- 83 contiguous little-endian doubles per group at byte offsets 2848 (qz1) and 8032 (qz2),
found by a raw scan (no HDF5 library needed).  Run once in the build container; the outputs are committed.

    python tests/golden/extract_domain_values.py [/root/reference]
"""
import os
import sys

import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
raw = open(os.path.join(ref, "test", "domain", "domain_values.hdf5"), "rb").read()
assert b"qz1" in raw and b"qz2" in raw
here = os.path.dirname(os.path.abspath(__file__))
for name, off in (("qz1", 2848), ("qz2", 8032)):
    a = np.frombuffer(raw[off:off + 8 * 83], dtype="<f8").copy()
    x, xu, y, yv, z, zw = np.split(a, np.cumsum([16, 17, 16, 17, 8]))
    assert np.all(np.diff(x) > 0) and np.all(np.diff(y) > 0) and np.all(np.diff(z) > 0) and zw[0] == -1.0 and zw[-1] == 0.0
    assert abs(xu[0] - 286 * np.pi / 180) < 1e-14 and abs(yv[-1] - 74 * np.pi / 180) < 1e-14
    a.astype("<f8").tofile(os.path.join(here, f"domain_values_{name}.f64"))
    print("wrote", name, z)
