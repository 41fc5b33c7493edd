"""Jacobian assembly only on the 1-degree grid (for ncu): python scripts/jac_only.py [reps]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, cases, iemic_b200
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
s, landm = cases.global_synth(360, 152, 24)
t = iemic_b200.THCM(s, landm)
for k, v in {"COMB": 1.0, "WIND": 1.0, "TEMP": 10.0, "SALT": 1.0}.items():
    t.setParameter(k, v)
x = torch.from_numpy(cases.consistent_state(s, landm, scale=0.05)).cuda()
for _ in range(reps):
    t.evaluate(x, None, True)
torch.cuda.synchronize()
print("done")
