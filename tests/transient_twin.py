"""The reference's implicit time stepper restated for the tests (TEST INFRASTRUCTURE): src/transient/AdaptiveTransient.H:88-170 (adaptive
theta stepping), Newton.H:78-123 (the Newton iteration and its convergence test), ThetaModel.H:65-165 (theta residual and Jacobian),
Transient.hpp:136-145 (time units) -- statement by statement, over any `model` that offers

    model.F(x)      -> numpy residual in the C++ sign of Ocean::computeRHS (THCM.C:1011), integral-condition row replaced (THCM.C:1013-1026)
    model.J(x)      -> scipy CSR Jacobian of F (THCM.C:1046-1180), dense integral-condition row included
    model.mass      -> diagonal of the mass matrix (assemble.F90:18-54; 0 on the integral-condition row)

The reference's regression test src/tests/trns_ocean.C runs exactly this (test/ocean/test_oceantransient.xml + timestepper_params.xml: 8 x 8 x 4
North Atlantic box, Mixing = 1, SRES = 0, theta = 0.75, ten adaptive steps from rest) and holds a GOLDEN NUMBER produced by the reference
itself: || state || = 37.03750142 (+- 1e-4) after 30 Newton steps in total (trns_ocean.C:63-64).

Linear solves: the reference uses preconditioned FGMRES whose preconditioner projects the two singular pressure modes out of every
update (TRIOS_BlockPreconditioner.C:1566-1572); here every update is the solution of the bordered system that is orthogonal to the
right null vectors of the theta Jacobian (a direct sparse solve: the solver is not what is being tested)."""
import numpy as np
import scipy.sparse as sp
import scipy.sparse.linalg as spla

GOLDEN_NORM, GOLDEN_NEWTON_STEPS, GOLDEN_TOL = 37.03750142, 30, 1e-4       # trns_ocean.C:63-64
TIMESTEPPER = dict(theta=0.75, dt=1.0e-3, nsteps=10, tmax_years=10.0, newton_tol=1e-4, max_newton=10, min_wanted=4, max_wanted=4,
                   dt_min=1.0e-6, dt_max=1.0, increase=2.0, decrease=2.0)     # test/ocean/timestepper_params.xml
# test/ocean/test_oceantransient.xml: THCM flags and "Starting Parameters"
SETTINGS = dict(vmix=1, SRES=0, TRES=1, rho_mixing=0, tap=1, iza=2, ite=1, its=1)
PARAMETERS = {"COMB": 1.0, "SUNP": 0.0, "SALT": 0.1, "WIND": 1.0, "TEMP": 10.0, "SPL1": 2.0e3, "SPL2": 0.01}


class NullSpaceSolver:
    """x = A^+ b restricted to the complement of the (state-independent) pressure null space: [A N_l; N_r^T 0] [x; mu] = [b; 0]."""

    def __init__(self):
        self.nl = self.nr = None

    def __call__(self, A, b):
        n = A.shape[0]
        if self.nl is None:
            U, S, Vt = np.linalg.svd(A.toarray())
            k = int((S < 1e-10 * S[0]).sum())
            self.nl, self.nr, self.null_dim = U[:, n - k:], Vt[n - k:].T, k
        if self.null_dim == 0:
            return spla.spsolve(A.tocsc(), b)
        K = sp.bmat([[A, sp.csr_matrix(self.nl)], [sp.csr_matrix(self.nr.T), None]]).tocsc()
        return spla.spsolve(K, np.concatenate([b, np.zeros(self.null_dim)]))[:n]


def run(model, p=TIMESTEPPER, log=None):
    """AdaptiveTransient<ThetaModel<Ocean>>::run() from the zero state; returns (state, total Newton steps, steps taken)."""
    theta = p["theta"]
    mass = model.mass
    solve = NullSpaceSolver()
    x = np.zeros(len(mass))
    dt, time, steps, total = p["dt"], 0.0, 0, 0
    tmax = p["tmax_years"] / (737.2685 / 365.0)                         # Transient.hpp:142-145
    while time < tmax and steps < p["nsteps"]:
        old, oldF = x.copy(), model.F(x)                                  # ThetaModel::initStep
        theta_rhs = lambda y: mass * (old - y) + dt * theta * model.F(y) + dt * (1.0 - theta) * oldF   # noqa: E731  ThetaModel.H:87-113
        y, Fx, converged, k = x.copy(), None, False, 0
        Fx = theta_rhs(y)
        for k in range(p["max_newton"]):                                  # Newton.H:91-122
            A = (model.J(y) - sp.diags(mass) / (theta * dt)).tocsr()      # ThetaModel.H:118-149
            dx = solve(A, Fx / (theta * dt))                              # ThetaModel.H:153-165
            normdx = np.abs(dx).max()
            y = y - dx
            Fx = theta_rhs(y)
            normF = np.linalg.norm(Fx)
            if normdx < p["newton_tol"] and normF < p["newton_tol"]:
                converged = True
                break
            if normdx > 1e2:
                break
        if not converged:                                                 # AdaptiveTransient.H:112-131
            if dt == p["dt_min"]:
                raise RuntimeError("minimum time step reached")
            dt = max(dt / p["decrease"], p["dt_min"])
            continue
        steps += 1
        time += dt
        x = y
        if log is not None:
            log.append((steps, time, dt, k, float(np.linalg.norm(x))))
        if k < p["min_wanted"]:                                           # AdaptiveTransient.H:158-162 (newton_->steps() is the 0-based index)
            dt = min(dt * p["increase"], p["dt_max"])
        elif k > p["max_wanted"]:
            dt = max(dt / p["decrease"], p["dt_min"])
        total += k
    return x, total, steps
