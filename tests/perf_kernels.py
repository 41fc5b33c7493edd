"""Kernel micro-benchmarks on the 1-degree grid (not a pytest; run on the GPU box):
   python tests/perf_kernels.py [n m l]   -- prints per-kernel avg ms and GB/s from the library's own event pairs."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import cases  # noqa: E402
import iemic_b200  # noqa: E402

n, m, l = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (360, 152, 24)
s, landm = cases.global_synth(n, m, l)
t = iemic_b200.THCM(s, landm)
for k, v in {"COMB": 1.0, "WIND": 1.0, "TEMP": 10.0, "SALT": 1.0}.items():
    t.setParameter(k, v)
x = torch.from_numpy(cases.consistent_state(s, landm, scale=0.05)).cuda()
F, y, dx = t.new_vector(), t.new_vector(), t.new_vector()
t.evaluate(x, F, True)
for _ in range(3):
    t.applyMatrix(x, y)
    t.evaluate(x, F, True)
t.newton_step_dev(x, dx, tol=0.0, maxit=19, restart=20, precon=1)
t.profile(True)
for _ in range(20):
    t.applyMatrix(x, y)
for _ in range(10):
    t.evaluate(x, F, True)
t.newton_step_dev(x, dx, tol=0.0, maxit=29, restart=30, precon=1)
rep = t.profile_report()
t.profile(False)
ncell, ndim, nnz = t.ndim // 6, t.ndim, t.nnz
alg = {"spmv_csr": nnz * 12 + ndim * 20, "thcm_assemble<JAC_GRAPH>": ncell * 49 + 8 * nnz, "thcm_assemble<RHS>": ncell * 145,
       "mgs_step": 32 * ndim, "dot": 16 * ndim, "axpby": 24 * ndim, "axpy_negdev": 24 * ndim, "scale_invsqrt": 16 * ndim,
       "blockdiag_apply": 48 * 8 * ncell, "blockdiag_build": ncell * 288 + 12 * nnz}
out = {}
for k, (cnt, tot) in rep.items():
    avg = tot / cnt
    out[k] = dict(n=cnt, avg_ms=round(avg, 5), gbs=round(alg[k] / avg / 1e6, 1) if k in alg else None)
print(json.dumps(dict(env={k: v for k, v in os.environ.items() if k.startswith("THCM_")}, kernels=out)))
