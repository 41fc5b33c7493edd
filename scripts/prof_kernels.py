"""Short workloads for ncu (GPU box): python scripts/prof_kernels.py asm|krylov"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch, cases, iemic_b200
what = sys.argv[1] if len(sys.argv) > 1 else "asm"
s, landm = cases.global_synth(360, 152, 24)
t = iemic_b200.THCM(s, landm)
for k, v in {"COMB": 1.0, "WIND": 1.0, "TEMP": 10.0, "SALT": 1.0}.items():
    t.setParameter(k, v)
t.set_ortho("dgks")
x = torch.from_numpy(cases.consistent_state(s, landm, scale=0.05)).cuda()
F, y, dx = t.new_vector(), t.new_vector(), t.new_vector()
if what == "asm":
    for _ in range(4):
        t.evaluate(x, F, True)
        t.applyMatrix(x, y)
        t.buildPreconditioner(1)
        t.applyPrecon(y, F)
else:
    t.newton_step_dev(x, dx, tol=0.0, maxit=49, restart=50, precon=1)
torch.cuda.synchronize()
print("done")
