"""Debug helper (GPU box): Jacobian of one box through the pipelined kernel vs the emulation harness."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
import cases, iemic_b200
from emu.emu import EmuTHCM
n, m, l, per = (int(v) for v in sys.argv[1:5])
s, landm = cases.box(n, m, l, bool(per), seed=6, land_frac=0.2)
t = iemic_b200.THCM(s, landm); e = EmuTHCM(s, landm)
x = cases.random_state(s, landm, scale=0.3, zero_on_land=False)
t.evaluate(torch.from_numpy(x).cuda(), None, True)
v = t.jacobian_values_host(); ve = e.jacobian(x)
print("box", n, m, l, per, "pipe", os.environ.get("THCM_ASM_PIPE"), "equal", np.array_equal(v, ve), "ndiff", int((v != ve).sum()), flush=True)
