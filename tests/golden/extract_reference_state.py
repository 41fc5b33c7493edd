"""Extracts the converged state the reference ships for its regression test (src/tests/reft_ocean.C:59-89) from
test/ocean/ocean_reference.h5: dataset `State` is a contiguous little-endian block of 16*16*16*6 = 24576 doubles at byte
offset 2144 (found by a raw scan; no HDF5 library needed).  Run once in the build container (the GPU box has no
/root/reference); the output tests/golden/ocean_reference_state.f64 is committed.

    python tests/golden/extract_reference_state.py [/root/reference]
"""
import os
import sys

import numpy as np

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
raw = open(os.path.join(ref, "test", "ocean", "ocean_reference.h5"), "rb").read()
n = 16 * 16 * 16 * 6
state = np.frombuffer(raw[2144:2144 + 8 * n], dtype="<f8").copy()
norms = [float(np.linalg.norm(state[q::6])) for q in range(6)]
# per-field 2-norms quoted in SURVEY.md / BASELINE.md (u, v, w, p, T, S)
want = [0.0979069, 0.0224030, 0.3858530, 0.0346863, 3.5162058, 0.0351621]
assert all(abs(a - b) < 1e-6 for a, b in zip(norms, want)), norms
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ocean_reference_state.f64")
state.astype("<f8").tofile(out)
print("wrote", out, norms)
