"""Second reference-produced number for the oracle (TEST INFRASTRUCTURE): the steady state of the reference's DEFAULT run (run/ocean/*.xml: 16 x 16 x 16
all-ocean box 300-340 E x 20-60 N, Mixing = 1, Forcing Type 2, restoring T and S), which run/ocean/workflow.org:13-21 documents as
    norm state : 542.3414237   norm rhs : 0.1590434284   norm sol : 0.01717931036   parameter : 1.000006854
after a pseudo-arclength continuation in Combined Forcing with Newton tolerance 1e-2 (run/ocean/continuation_params.xml) -- i.e. an iterate
whose last Newton update was still 0.017 long.  This script follows the same branch by natural continuation (steps of 0.1, Newton to
1e-10, direct sparse solves with two pinned pressure points, updates projected off the two singular pressure modes) on the ORACLE's F and J
and prints the norm of the exact root at parameter 1.000006854:  542.3439468  (4.7e-6 relative from the reference's loosely converged
iterate; d|x|/dpar = 589, so the parameter offset itself accounts for 0.004).  ~50 minutes on one core (96 sparse LU factorisations of a
24 576-unknown 3-D Jacobian).  The converged state is kept as tests/golden/default_run_steady_state.f64 and checked by
tests/test_oracle_pins.py::test_default_run_steady_state_matches_the_reference_norm (root of F, norm).

    python scripts/default_run_steady_state.py [output directory for the intermediate states, default /tmp/steady]
"""
import os, sys, time, numpy as np, scipy.sparse as sp, scipy.sparse.linalg as spla
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
OUT = sys.argv[1] if len(sys.argv) > 1 else '/tmp/steady'
os.makedirs(OUT, exist_ok=True)
import cases, iemic_b200
from cases import PAR_INDEX as P
from oracle.oracle import OracleTHCM
n=m=l=16
s = iemic_b200.Settings.from_degrees(n,m,l,300,340,20,60,periodic=False,hdim=4000.0,qz=1.0,vmix=1,rho_mixing=0,tap=1,forcing_type=2,TRES=1,SRES=1,iza=2,ite=1,its=1)
landm = iemic_b200.all_ocean_mask(n,m,l,periodic=False)
o = OracleTHCM(s, landm)
pars={"COMB":0.0,"SUNP":0.0,"SALT":0.1,"WIND":1.0,"TEMP":10.0,"SPL1":2.0e3,"SPL2":0.01}
for k,v in pars.items(): o.setpar(P[k], v)
nd=o.ndim; rowptr,col=o.graph()
# right null vectors: constant pressure and the B-grid checkerboard
cells=np.arange(n*m*l); ci=cells%n; cj=(cells//n)%m
N1=np.zeros(nd); N1[3::6]=1.0
N2=np.zeros(nd); N2[3::6]=(-1.0)**(ci+cj)
Nr=np.stack([N1/np.linalg.norm(N1),N2/np.linalg.norm(N2)],1)
# two pressure rows made Dirichlet (different parity) for the factorisation
prow=[6*0+3, 6*1+3]
def solve(x,b):
    val,_=o.jacobian_graph(x)
    A=sp.csr_matrix((val,col,rowptr),shape=(nd,nd))
    chk=np.abs(A@Nr).max()
    A=A.tolil()
    for r in prow:
        A.rows[r]=[r]; A.data[r]=[1.0]
    bb=b.copy(); bb[prow]=0.0
    dx=spla.splu(A.tocsc()).solve(bb)
    dx-=Nr@(Nr.T@dx)
    return dx,chk
x=np.zeros(nd)
t0=time.time()
for par in list(np.linspace(0.1,1.0,10))+[1.000006854]:
    o.setpar(P["COMB"],par)
    for it in range(12):
        F=-o.rhs(x)
        dx,chk=solve(x,-F)
        x=x+dx
        nF=np.linalg.norm(-o.rhs(x))
        if np.linalg.norm(dx)<1e-10*max(1,np.linalg.norm(x)): break
    np.save(os.path.join(OUT, f"x_{par:.9f}.npy"), x); print(f"par {par:.9f}: newton its {it+1} |F| {nF:.3e} |x| {np.linalg.norm(x):.7f} null chk {chk:.1e} ({time.time()-t0:.0f}s)",flush=True)
print("reference (run/ocean/workflow.org): norm state 542.3414237 at parameter 1.000006854")
