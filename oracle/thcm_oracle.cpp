// =============================================================================
// thcm_oracle.cpp -- CPU ORACLE (test infrastructure, NOT product code)
//
// A plain C++17 restatement of the reference's THCM Newton-step hot path
// (nlesc-smcm/i-emic, src/ocean/*.F90) with the reference's own data layout:
// dense per-cell dependency blocks Al/An(27,6,6,n,m,l), full-array atom
// updates, `boundaries`, the `fillcolA` 162-candidate threshold scan, CRS
// `matAvec`.  It exists to check the CUDA path and to serve as the timed
// "restated reference CPU path" (bench.py cpu_baseline / --impl reference).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
// arm may load this library.  The product (i-emic_b200/csrc) never links it.
//
// PARITY PIN STATUS: the reference cannot be built here (no gfortran / MPI /
// Trilinos) and ships no golden Jacobian / residual vectors.  It does hold a
// GOLDEN NUMBER of its own making for this path: src/tests/trns_ocean.C:63-64
// asserts || state || = 37.03750142 (+- 1e-4) and 30 Newton steps after ten
// adaptive implicit theta steps from rest (8 x 8 x 4 box, Mixing = 1, salinity
// integral condition).  The reference's time stepper restated over THIS
// oracle's residual and Jacobian (tests/transient_twin.py) gives 37.0375014173
// and 30 (tests/test_oracle_pins.py, 1e-7): every term of F is pinned at the
// 1e-9 level by a number the reference produced, and J well enough to repeat
// its Newton history step for step.  The same run through the CUDA path gives
// the same number (tests/test_gpu_evaluate_rows.py).  A second reference-made
// number, on another configuration (the default run of run/ocean: 16^3 basin
// without land, Forcing Type 2, restoring salinity): run/ocean/workflow.org:13-21
// prints norm state 542.3414237 at parameter 1.000006854 (Newton tolerance
// 1e-2); the exact root of this oracle's F at that parameter has norm
// 542.3439468, 4.7e-6 relative away (scripts/default_run_steady_state.py,
// tests/test_oracle_pins.py).
// It also ships ONE
// reference-produced vector: the converged state of its regression test
// (test/ocean/ocean_reference.h5, src/tests/reft_ocean.C:59-89; committed as
// tests/golden/ocean_reference_state.f64).  That state is a root of the
// restated residual to Newton accuracy (|F(x*)| = 1.8e-4 vs |F(0)| = 19.8;
// u,v,w,p rows <= 4e-8; 850x larger without the mixing term): the RESIDUAL is
// pinned by reference data (tests/test_oracle_pins.py).  The JACOBIAN has no
// reference vector of its own ("golden parity unpinned" entry by entry), but
// Newton with this F and J started at that state converges quadratically
// (1.8e-4 -> 1e-9 -> 6e-15) to a root 2e-5 from it, and the grid arrays match
// the reference's test/domain/domain_values.hdf5 to 1e-15; it is also pinned against
// that residual by finite differences and by the reference's own invariants:
// exact mass-matrix values (test_ocean.C:61-125), FD-vs-analytic Jacobian
// (TestDefinitions.H:32-87), salt conservation integrals (test_ocean.C:242-316),
// maximal-graph containment (THCM.C:2320-2549).
//
// Every routine cites the reference file:line it follows (relative to
// /root/reference/src/ocean unless stated).  Compile with
//   g++ -O2 -ffp-contract=off   (x86-64 baseline: no FMA contraction, like the
//   reference's gfortran -O3 -fdefault-real-8 build, src/CMakeLists.txt:26)
// =============================================================================
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <string>
#include "fdlibm_tanh.h"

namespace {

// tanh of the mixing tapers (mix_imp.f:675-727, 837-857): 1 = the specified fdlibm algorithm (fdlibm_tanh.h; what the device path
// implements too), 0 = the platform libm (what a build of the reference on THIS machine would call)
int g_tanh_impl = 1;
inline double mix_tanh(double x) { return g_tanh_impl ? oracle_tanh(x) : std::tanh(x); }

// par.F90:14-81
constexpr double PI = 3.14159265358979323846;
constexpr int NUN = 6, NP = 27, NPAR = 30;
enum { AL_T = 1, RAYL, EK_V, EK_H, ROSB, MIXP, RESC, SPL1, HMTP, SUNP, PE_H, PE_V,
       P_VC, LAMB, SALT, WIND, TEMP, BIOT, COMB, ARCL, NLES, IFRICB, CONT, ENER,
       ALPC, CMPR, FPER, SPER, MKAP, SPL2 };
enum { UU = 1, VV, WW, PP, TT, SS };
enum { OCEAN = 0, LAND = 1, WATER = 2, PERIO = 3 };

// usr.F90:132-160
constexpr double omegadim = 7.292e-05, r0dim = 6.37e+06, udim = 0.1e+00, gdim = 9.8e+00,
                 rhodim = 1.024e+03, t0 = 15, deltat = 1.0, deltas = 1.0, s0 = 35.0,
                 cp0 = 4.2e+03, alpt1 = 2.93, alpt2 = 8.3e-02, alpt3 = 6.6e-04,
                 ah = 2.5e+05, av = 1.0e-03, kappah = 1.0e+03, kappav = 1.0e-04;
constexpr double zmin = -1.0, zmax = 0.0;  // usr.F90:42-43
// atm.F90:5-19
constexpr double hdima = 8400., rhoa = 1.25, uatm = 0.0, ce = 1.3e-03, ch = 0.94 * ce,
                 cpa = 1000., uw = 8.5, d0 = 3.1e+06, c0 = 0.43, arad = 216.0, brad = 1.5,
                 sun0 = 1360., lv = 2.5e+06;

// atom(np,n,m,l), Fortran column-major, 1-based accessors (spf.F90:11)
struct Atom {
    int n, m, l;
    std::vector<double> a;
    Atom(int n_, int m_, int l_) : n(n_), m(m_), l(l_), a((size_t)NP * n_ * m_ * l_, 0.0) {}
    inline double& operator()(int loc, int i, int j, int k) {
        return a[(size_t)(loc - 1) + (size_t)NP * ((i - 1) + (size_t)n * ((j - 1) + (size_t)m * (k - 1)))];
    }
    void zero() { std::fill(a.begin(), a.end(), 0.0); }
};

// generic 3-D array with arbitrary lower bounds (Fortran order)
struct Arr3 {
    int l0, l1, l2, n0, n1, n2;
    std::vector<double> a;
    Arr3(int lo0, int hi0, int lo1, int hi1, int lo2, int hi2)
        : l0(lo0), l1(lo1), l2(lo2), n0(hi0 - lo0 + 1), n1(hi1 - lo1 + 1), n2(hi2 - lo2 + 1),
          a((size_t)n0 * n1 * n2, 0.0) {}
    inline double& operator()(int i, int j, int k) {
        return a[(size_t)(i - l0) + (size_t)n0 * ((j - l1) + (size_t)n1 * (k - l2))];
    }
    void zero() { std::fill(a.begin(), a.end(), 0.0); }
};

struct Oracle {
    // ---- m_usr state (usr.F90) ----
    int n = 0, m = 0, l = 0, ndim = 0;
    double xmin, xmax, ymin, ymax;
    double ymin_glob, ymax_glob;  // m_global values used by temfun/salfun (forcing.F90:424-449)
    bool periodic = false;
    double dx, dy, dz;
    std::vector<double> x, y, z, xu, yv, zw, ze, zwe, dfzT, dfzW;  // index = Fortran index
    std::vector<int> landm;                                        // (0:n+1,0:m+1,0:l+1)
    double hdim = 4000., qz = 1.0;
    int ih = 0, vmix = 1, tap = 1, rho_mixing = 0;
    int TRES = 1, SRES = 1, iza = 2, its = 1, ite = 1, coriolis_on = 1, forcing_type = 0;
    int coupled_T = 0, coupled_S = 0;
    double alphaT = 1.0e-04, alphaS = 7.6e-04;
    double par[NPAR + 1];
    std::vector<double> Frc, taux, tauy, tatm, emip, spert, adapted_emip, qatm, albe, patm, msi, gsi, qsa;
    std::vector<double> internal_temp, internal_salt;
    double QTnd, QSnd;
    // ---- m_atm (atm.F90) ----
    double qdim = 0.01, nuq = 0, nus = 0, eta = 0, dqso = 0, eo0 = 0, albe0 = 0, albed = 0, lvsc = 0;
    double Ooa = 1.0, Os = 1.0;
    std::vector<double> suno;
    // ---- m_ice (ice.F90) ----
    double zeta = 0.0, a0 = -0.0575, Lf = 3.347e+05, Qvar = 0.0, Q0 = 0.0;
    // ---- m_res ----
    double p0 = 0.0;
    // ---- m_mat ----
    std::vector<double> Al, An;  // (np,nun,nun,n,m,l)
    std::vector<int> begA, jcoA;
    std::vector<double> coA, coB;
    long bad_columns = 0;  // non-periodic columns pointing outside the domain (should stay 0)

    inline int& lm(int i, int j, int k) { return landm[(size_t)i + (size_t)(n + 2) * (j + (size_t)(m + 2) * k)]; }
    inline size_t aidx(int loc, int A, int B, int i, int j, int k) const {
        return (size_t)(loc - 1) + NP * ((size_t)(A - 1) + NUN * ((size_t)(B - 1) +
               NUN * ((size_t)(i - 1) + (size_t)n * ((size_t)(j - 1) + (size_t)m * (k - 1)))));
    }
    // matetc.F90:123-131
    inline int find_row2(int i, int j, int k, int XX) const { return NUN * ((k - 1) * n * m + n * (j - 1) + i - 1) + XX; }
    inline double& f2(std::vector<double>& f, int i, int j) { return f[(size_t)(i - 1) + (size_t)n * (j - 1)]; }
    inline double& f3(std::vector<double>& f, int i, int j, int k) { return f[(size_t)(i - 1) + (size_t)n * ((j - 1) + (size_t)m * (k - 1))]; }

    // ---------------- grid.F90:95-130 ----------------
    static double fz(double zz, double q) {
        double th = std::tanh(q * (zz + 1));
        double tth = std::tanh(q);
        if (q > 1.0) return -1 + th / tth;
        return zz + (1. - q) * zz * (1 - zz);
    }
    static double dfdz(double zz, double q) {
        double chh = std::cosh(q * (zz + 1));
        double tth = std::tanh(q);
        if (q > 1.0) return q / (tth * chh * chh);
        return 1.0 + (1. - q) * (1. - 2. * zz);
    }
    // grid.F90:2-66
    void grid() {
        dx = (xmax - xmin) / n;
        dy = (ymax - ymin) / m;
        dz = (zmax - zmin) / l;
        x.assign(n + 1, 0.0); xu.assign(n + 1, 0.0);
        y.assign(m + 2, 0.0); yv.assign(m + 1, 0.0);
        z.assign(l + 1, 0.0); zw.assign(l + 1, 0.0); ze.assign(l + 1, 0.0); zwe.assign(l + 1, 0.0);
        dfzT.assign(l + 1, 0.0); dfzW.assign(l + 1, 0.0);
        for (int i = 1; i <= n; i++) {
            x[i] = ((double)i - 0.5) * dx + xmin;
            xu[i] = ((double)i) * dx + xmin;
        }
        xu[0] = xmin;
        for (int j = 1; j <= m; j++) {
            y[j] = ((double)j - 0.5) * dy + ymin;
            yv[j] = ((double)j) * dy + ymin;
        }
        y[0] = y[1] - dy;
        y[m + 1] = y[m] + dy;
        yv[0] = ymin;
        for (int k = 1; k <= l; k++) {
            ze[k] = ((double)k - 0.5) * dz + zmin;
            zwe[k] = ((double)k) * dz + zmin;
            z[k] = fz(ze[k], qz);
            zw[k] = fz(zwe[k], qz);
            dfzT[k] = dfdz(ze[k], qz);
            dfzW[k] = dfdz(zwe[k], qz);
        }
        zw[0] = zmin;
        dfzW[0] = dfdz(zmin, qz);
    }

    // usrc.F90:1153-1197 (+ vmix_par, mix_imp.f:122-137)
    void stpnt() {
        for (int i = 0; i <= NPAR; i++) par[i] = 0.0;
        par[AL_T] = 0.1 / (2 * omegadim * rhodim * hdim * udim * dz * dfzT[l]);
        par[RAYL] = alphaT * gdim * hdim / (2 * omegadim * udim * r0dim);
        par[EK_V] = av / (2 * omegadim * hdim * hdim);
        par[EK_H] = ah / (2 * omegadim * r0dim * r0dim);
        par[ROSB] = udim / (2 * omegadim * r0dim);
        par[HMTP] = 0.0;
        par[SUNP] = 0.0;
        par[PE_H] = kappah / (udim * r0dim);
        par[PE_V] = kappav * r0dim / (udim * hdim * hdim);
        par[P_VC] = 2.5e+04 * par[PE_V];
        par[LAMB] = alphaS / alphaT;
        par[SALT] = 0.0;
        par[WIND] = 0.0;
        par[TEMP] = 0.0;
        par[BIOT] = r0dim / (75. * 3600. * 24. * udim);
        par[COMB] = 0.0;
        par[NLES] = 0.0;
        par[CMPR] = 0.0;
        par[ALPC] = 1.0;
        par[ENER] = 1.0e+02;
        par[MIXP] = 0.0;
        par[MKAP] = 0.0;
        par[SPL1] = 2.0e+03;
        par[SPL2] = 0.01;
        if (vmix == 0) {
            par[MIXP] = 0.0;
            par[P_VC] = 0.0;
            par[ALPC] = 1.0;
            par[ENER] = 1.0e+2;
            par[MKAP] = 0.0;
        }
    }

    // usrc.F90:1200-1240
    void atmos_coef() {
        double muoa = rhoa * ch * cpa * uw;
        Os = sun0 * c0 / 4 * QTnd;
        Ooa = muoa * QTnd;
        nus = 0.0;
        lvsc = 0.0;
        suno.assign(m + 1, 0.0);
        for (int j = 1; j <= m; j++) {
            double sj = std::sin(y[j]);
            suno[j] = Os * (1 - .482 * (3 * (sj * sj) - 1.) / 2.);
        }
    }

    // spf.F90:792-854
    static double amh(double yy, int ih_) { return ih_ == 0 ? 1.0 : 1. + 10.0 * std::exp(-5 * yy * yy); }
    static double bmh(double yy, int ih_) { return ih_ == 0 ? 1.0 : 1.0 + 10.0 * std::exp(-5 * yy * yy); }
    static double bmhy(double yy, int ih_) { return ih_ == 0 ? 0.0 : -10. * 10.0 * yy * std::exp(-5 * yy * yy); }

    // ---------------- spf.F90:13-74 ----------------
    void uderiv(int type, Atom& atom) {
        atom.zero();
        switch (type) {
        case 1:
            for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) atom(5, i, j, k) = 1.0;
            break;
        case 2: {
            std::vector<double> cosdx2i(m + 1);
            for (int j = 0; j <= m; j++) { double t = 1.0 / (std::cos(yv[j]) * dx); cosdx2i[j] = t * t; }
            for (int j = 1; j <= m - 1; j++) for (int i = 1; i <= n; i++) for (int k = 1; k <= l; k++) {
                atom(2, i, j, k) = amh(yv[j], ih) * cosdx2i[j];
                atom(8, i, j, k) = amh(yv[j], ih) * cosdx2i[j];
                atom(5, i, j, k) = -(atom(2, i, j, k) + atom(8, i, j, k));
            }
        } break;
        case 3: {
            double t = 1.0 / dy, rdy2i = t * t;
            for (int i = 1; i <= n; i++) for (int j = 1; j <= m - 1; j++) for (int k = 1; k <= l; k++) {
                atom(4, i, j, k) = rdy2i * bmh(y[j], ih) * std::cos(y[j]) / std::cos(yv[j]);
                atom(6, i, j, k) = rdy2i * bmh(y[j + 1], ih) * std::cos(y[j + 1]) / std::cos(yv[j]);
                atom(5, i, j, k) = -(atom(4, i, j, k) + atom(6, i, j, k));
            }
        } break;
        case 4: {
            double t = 1.0 / dz, rdz2i = t * t;
            for (int k = 1; k <= l; k++) {
                double h1 = 1. / (dfzT[k] * dfzW[k]);
                double h2 = 1. / (dfzT[k] * dfzW[k - 1]);
                for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
                    atom(14, i, j, k) = h2 * rdz2i;
                    atom(23, i, j, k) = h1 * rdz2i;
                    atom(5, i, j, k) = -(atom(14, i, j, k) + atom(23, i, j, k));
                }
            }
        } break;
        case 5: {
            std::vector<double> tand2(m + 1);
            for (int j = 0; j <= m; j++) tand2[j] = 1 - std::tan(yv[j]) * std::tan(yv[j]);
            for (int j = 1; j <= m - 1; j++) for (int i = 1; i <= n; i++) for (int k = 1; k <= l; k++)
                atom(5, i, j, k) = bmh(yv[j], ih) * tand2[j] + std::tan(yv[j]) * bmhy(yv[j], ih);
        } break;
        case 6: {
            std::vector<double> tand2(m + 1), cosd2(m + 1);
            for (int j = 0; j <= m; j++) { tand2[j] = std::tan(yv[j]); cosd2[j] = std::cos(yv[j]); }
            for (int j = 1; j <= m - 1; j++) for (int i = 1; i <= n; i++) for (int k = 1; k <= l; k++) {
                atom(2, i, j, k) = (bmhy(yv[j], ih) - (amh(yv[j], ih) + bmh(yv[j], ih)) * tand2[j]) / (dx * cosd2[j]);
                atom(8, i, j, k) = -(bmhy(yv[j], ih) - (amh(yv[j], ih) + bmh(yv[j], ih)) * tand2[j]) / (dx * cosd2[j]);
            }
        } break;
        }
    }

    // ---------------- spf.F90:76-136 ----------------
    void vderiv(int type, Atom& atom) {
        atom.zero();
        switch (type) {
        case 1:
            for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) atom(5, i, j, k) = 1.0;
            break;
        case 2: {
            std::vector<double> cosdx2i(m + 1);
            for (int j = 0; j <= m; j++) { double t = 1.0 / (std::cos(yv[j]) * dx); cosdx2i[j] = t * t; }
            for (int i = 1; i <= n; i++) for (int j = 1; j <= m - 1; j++) for (int k = 1; k <= l; k++) {
                atom(2, i, j, k) = bmh(yv[j], ih) * cosdx2i[j];
                atom(5, i, j, k) = -2 * bmh(yv[j], ih) * cosdx2i[j];
                atom(8, i, j, k) = bmh(yv[j], ih) * cosdx2i[j];
            }
        } break;
        case 3: {
            double t = 1.0 / dy, dy2i = t * t;
            for (int i = 1; i <= n; i++) for (int j = 1; j <= m - 1; j++) for (int k = 1; k <= l; k++) {
                atom(4, i, j, k) = dy2i * amh(y[j], ih) * std::cos(y[j]) / std::cos(yv[j]);
                atom(6, i, j, k) = dy2i * amh(y[j + 1], ih) * std::cos(y[j + 1]) / std::cos(yv[j]);
                atom(5, i, j, k) = -(atom(4, i, j, k) + atom(6, i, j, k));
            }
        } break;
        case 4: {
            double t = 1.0 / dz, rdz2i = t * t;
            for (int k = 1; k <= l; k++) {
                double h1 = 1. / (dfzT[k] * dfzW[k]);
                double h2 = 1. / (dfzT[k] * dfzW[k - 1]);
                for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
                    atom(14, i, j, k) = h2 * rdz2i;
                    atom(23, i, j, k) = h1 * rdz2i;
                    atom(5, i, j, k) = -(atom(14, i, j, k) + atom(23, i, j, k));
                }
            }
        } break;
        case 5:
            for (int j = 1; j <= m - 1; j++) for (int i = 1; i <= n; i++) for (int k = 1; k <= l; k++)
                atom(5, i, j, k) = bmh(yv[j], ih) - amh(yv[j], ih) * std::tan(yv[j]) * std::tan(yv[j]) +
                                   bmhy(yv[j], ih) * std::tan(yv[j]);
            break;
        case 6: {
            std::vector<double> tand2(m + 1), cosd2(m + 1);
            for (int j = 0; j <= m; j++) { tand2[j] = std::tan(yv[j]); cosd2[j] = std::cos(yv[j]); }
            for (int j = 1; j <= m - 1; j++) for (int i = 1; i <= n; i++) for (int k = 1; k <= l; k++) {
                atom(2, i, j, k) = -((amh(yv[j], ih) + bmh(yv[j], ih)) * tand2[j] - bmhy(yv[j], ih)) / (dx * cosd2[j]);
                atom(8, i, j, k) = ((amh(yv[j], ih) + bmh(yv[j], ih)) * tand2[j] - bmhy(yv[j], ih)) / (dx * cosd2[j]);
            }
        } break;
        }
    }

    // ---------------- spf.F90:138-187 ----------------
    void pderiv(int type, Atom& atom) {
        atom.zero();
        switch (type) {
        case 1: {
            std::vector<double> cos2i(m + 2);
            for (int j = 0; j <= m + 1; j++) cos2i[j] = 1.0 / (2 * std::cos(y[j]) * dx);
            for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) for (int k = 1; k <= l; k++) {
                atom(2, i, j, k) = -cos2i[j];
                atom(4, i, j, k) = cos2i[j];
                atom(1, i, j, k) = -cos2i[j];
                atom(5, i, j, k) = cos2i[j];
            }
        } break;
        case 2: {
            std::vector<double> cos2i(m + 2), cos2v(m + 1);
            for (int j = 0; j <= m; j++) cos2v[j] = std::cos(yv[j]);
            for (int j = 0; j <= m + 1; j++) cos2i[j] = 1. / (2 * std::cos(y[j]) * dy);
            for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) for (int k = 1; k <= l; k++) {
                atom(4, i, j, k) = -cos2v[j - 1] * cos2i[j];
                atom(2, i, j, k) = cos2v[j] * cos2i[j];
                atom(1, i, j, k) = -cos2v[j - 1] * cos2i[j];
                atom(5, i, j, k) = cos2v[j] * cos2i[j];
            }
        } break;
        case 3: {
            double dzi = 1.0 / dz;
            for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
                atom(5, i, j, k) = dzi / dfzT[k];
                atom(14, i, j, k) = -dzi / dfzT[k];
            }
        } break;
        }
    }

    // ---------------- spf.F90:189-268 ----------------
    void tderiv(int type, Atom& atom) {
        atom.zero();
        switch (type) {
        case 1:
        case 2:
            for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) atom(5, i, j, l) = 1.0;
            break;
        case 3: {
            std::vector<double> cosdx2i(m + 2);
            for (int j = 0; j <= m + 1; j++) { double t = 1.0 / (std::cos(y[j]) * dx); cosdx2i[j] = t * t; }
            for (int i = 1; i <= n; i++) for (int j = 1; j <= m; j++) for (int k = 1; k <= l; k++) {
                atom(2, i, j, k) = cosdx2i[j] * (1 - lm(i, j, l));
                atom(5, i, j, k) = -2 * cosdx2i[j] * (1 - lm(i, j, l));
                atom(8, i, j, k) = cosdx2i[j] * (1 - lm(i, j, l));
            }
        } break;
        case 4: {
            double t = 1.0 / dy, dy2i = t * t;
            for (int i = 1; i <= n; i++) for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) {
                atom(4, i, j, k) = (dy2i * std::cos(yv[j - 1]) / std::cos(y[j])) * (1 - lm(i, j, l));
                atom(6, i, j, k) = (dy2i * std::cos(yv[j]) / std::cos(y[j])) * (1 - lm(i, j, l));
                atom(5, i, j, k) = -(atom(4, i, j, k) + atom(6, i, j, k));
            }
        } break;
        case 5: {
            double t = 1.0 / dz, dz2i = t * t;
            for (int k = 1; k <= l - 1; k++) {
                double h1 = 1. / (dfzT[k] * dfzW[k]);
                double h2 = 1. / (dfzT[k] * dfzW[k - 1]);
                for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
                    atom(14, i, j, k) = h2 * dz2i * (1 - lm(i, j, l));
                    atom(23, i, j, k) = h1 * dz2i * (1 - lm(i, j, l));
                    atom(5, i, j, k) = -(atom(14, i, j, k) + atom(23, i, j, k));
                }
            }
            int k = l;
            double h2 = 1. / (dfzT[k] * dfzW[k - 1]);
            for (int i = 1; i <= n; i++) for (int j = 1; j <= m; j++) {
                atom(14, i, j, k) = h2 * dz2i * (1 - lm(i, j, l));
                atom(23, i, j, k) = 0.0;
                atom(5, i, j, k) = -(atom(14, i, j, k) + atom(23, i, j, k));
            }
        } break;
        case 6:
            for (int i = 1; i <= n; i++) for (int j = 1; j <= m; j++) for (int k = 1; k <= l; k++) {
                atom(23, i, j, k) = 1.0 * (1 - lm(i, j, l));
                atom(5, i, j, k) = 1.0 * (1 - lm(i, j, l));
            }
            break;
        case 7:
            for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) atom(5, i, j, 1) = 1.0;
            break;
        }
    }

    // ---------------- spf.F90:271-302 ----------------
    void coriolis(int type, Atom& atom) {
        atom.zero();
        std::vector<double> corv(m + 1);
        for (int j = 0; j <= m; j++) corv[j] = std::sin(yv[j]) * coriolis_on;
        if (type == 1 || type == 2)
            for (int i = 1; i <= n; i++) for (int j = 1; j <= m - 1; j++) for (int k = 1; k <= l; k++) atom(5, i, j, k) = corv[j];
    }

    // ---------------- spf.F90:305-345 ----------------
    void gradp(int type, Atom& atom) {
        atom.zero();
        switch (type) {
        case 1: {
            std::vector<double> cosdxi(m + 1);
            for (int j = 0; j <= m; j++) cosdxi[j] = 1. / (2 * std::cos(yv[j]) * dx);
            for (int i = 1; i <= n; i++) for (int j = 1; j <= m - 1; j++) for (int k = 1; k <= l; k++) {
                atom(5, i, j, k) = -cosdxi[j];
                atom(6, i, j, k) = -cosdxi[j];
                atom(8, i, j, k) = cosdxi[j];
                atom(9, i, j, k) = cosdxi[j];
            }
        } break;
        case 2: {
            double dyi = 1. / (2 * dy);
            for (int i = 1; i <= n; i++) for (int j = 1; j <= m - 1; j++) for (int k = 1; k <= l; k++) {
                atom(5, i, j, k) = -dyi;
                atom(8, i, j, k) = -dyi;
                atom(6, i, j, k) = dyi;
                atom(9, i, j, k) = dyi;
            }
        } break;
        case 3: {
            double dzi = 1. / dz;
            for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
                atom(5, i, j, k) = -dzi / dfzW[k];
                atom(23, i, j, k) = dzi / dfzW[k];
            }
        } break;
        }
    }

    // spf.F90:347-359
    void masksi(Atom& atom, std::vector<double>& mask) {
        atom.zero();
        for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) atom(5, i, j, l) = f2(mask, i, j);
    }

    // ---------------- spf.F90:362-484 ----------------
    void tnlin(int type, Atom& atom, Arr3& u, Arr3& v, Arr3& w, Arr3& t) {
        atom.zero();
        switch (type) {
        case 1:
            for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) atom(5, i, j, k) = 1.0;
            break;
        case 2: {
            std::vector<double> c(m + 2);
            for (int j = 0; j <= m + 1; j++) c[j] = 1.0 / (4 * std::cos(y[j]) * dx);
            for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
                atom(2, i, j, k) = -(t(i, j, k) + t(i - 1, j, k)) * c[j] * (1 - lm(i, j, l));
                atom(4, i, j, k) = (t(i + 1, j, k) + t(i, j, k)) * c[j] * (1 - lm(i, j, l));
                atom(1, i, j, k) = -(t(i, j, k) + t(i - 1, j, k)) * c[j] * (1 - lm(i, j, l));
                atom(5, i, j, k) = (t(i + 1, j, k) + t(i, j, k)) * c[j] * (1 - lm(i, j, l));
            }
        } break;
        case 3: {
            std::vector<double> c(m + 2);
            for (int j = 0; j <= m + 1; j++) c[j] = 1.0 / (4 * std::cos(y[j]) * dx);
            for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
                atom(2, i, j, k) = -(u(i - 1, j, k) + u(i - 1, j - 1, k)) * c[j] * (1 - lm(i, j, l));
                atom(8, i, j, k) = (u(i, j, k) + u(i, j - 1, k)) * c[j] * (1 - lm(i, j, l));
                atom(5, i, j, k) = atom(2, i, j, k) + atom(8, i, j, k);
            }
        } break;
        case 4: {
            std::vector<double> c(m + 2);
            for (int j = 0; j <= m + 1; j++) c[j] = 1.0 / (4 * std::cos(y[j]) * dy);
            for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
                atom(4, i, j, k) = -c[j] * (t(i, j, k) + t(i, j - 1, k)) * std::cos(yv[j - 1]) * (1 - lm(i, j, l));
                atom(1, i, j, k) = -c[j] * (t(i, j, k) + t(i, j - 1, k)) * std::cos(yv[j - 1]) * (1 - lm(i, j, l));
                atom(5, i, j, k) = c[j] * (t(i, j + 1, k) + t(i, j, k)) * std::cos(yv[j]) * (1 - lm(i, j, l));
                atom(2, i, j, k) = c[j] * (t(i, j + 1, k) + t(i, j, k)) * std::cos(yv[j]) * (1 - lm(i, j, l));
            }
        } break;
        case 5: {
            std::vector<double> c(m + 2);
            for (int j = 0; j <= m + 1; j++) c[j] = 1.0 / (4 * std::cos(y[j]) * dy);
            for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
                atom(4, i, j, k) = -(v(i, j - 1, k) + v(i - 1, j - 1, k)) * c[j] * std::cos(yv[j - 1]) * (1 - lm(i, j, l));
                atom(6, i, j, k) = (v(i, j, k) + v(i - 1, j, k)) * c[j] * std::cos(yv[j]) * (1 - lm(i, j, l));
                atom(5, i, j, k) = atom(4, i, j, k) + atom(6, i, j, k);
            }
        } break;
        case 6: {
            double tdzi = 1.0 / (2 * dz);
            for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
                for (int k = 1; k <= l - 1; k++) {
                    atom(14, i, j, k) = -tdzi * (1 - lm(i, j, l)) * (t(i, j, k) + t(i, j, k - 1)) / dfzT[k];
                    atom(5, i, j, k) = tdzi * (1 - lm(i, j, l)) * (t(i, j, k + 1) + t(i, j, k)) / dfzT[k];
                }
                int k = l;
                atom(14, i, j, k) = -tdzi * (1 - lm(i, j, l)) * (t(i, j, k) + t(i, j, k - 1)) / dfzT[k];
                atom(5, i, j, k) = 0.0;
            }
        } break;
        case 7: {
            double tdzi = 1.0 / (2 * dz);
            for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
                atom(14, i, j, k) = -w(i, j, k - 1) * (1 - lm(i, j, l)) * tdzi / dfzT[k];
                atom(23, i, j, k) = w(i, j, k) * (1 - lm(i, j, l)) * tdzi / dfzT[k];
                atom(5, i, j, k) = atom(14, i, j, k) + atom(23, i, j, k);
            }
        } break;
        }
    }

    // ---------------- spf.F90:486-542 ----------------
    void wnlin(int type, Atom& atom, Arr3& t) {
        atom.zero();
        for (int k = 1; k <= l - 1; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            switch (type) {
            case 1:
                atom(23, i, j, k) = (t(i, j, k) + t(i, j, k + 1)) / 2.;
                atom(5, i, j, k) = (t(i, j, k) + t(i, j, k + 1)) / 2.;
                break;
            case 2:
                atom(23, i, j, k) = t(i, j, k + 1) / 4.;
                atom(5, i, j, k) = (t(i, j, k) + 2 * t(i, j, k + 1)) / 4.;
                break;
            case 3: {
                double sm = t(i, j, k) + t(i, j, k + 1);
                atom(5, i, j, k) = 0.375 * (sm * sm);
                atom(23, i, j, k) = 0.375 * (sm * sm);
            } break;
            case 4:
                atom(5, i, j, k) = 0.125 * (t(i, j, k) * t(i, j, k) + 3 * t(i, j, k + 1) * t(i, j, k) +
                                            3 * t(i, j, k + 1) * t(i, j, k + 1));
                atom(23, i, j, k) = 0.125 * t(i, j, k + 1) * t(i, j, k + 1);
                break;
            }
        }
    }

    // ---------------- spf.F90:544-665 ----------------
    void unlin(int type, Atom& atom, Arr3& u, Arr3& v, Arr3& w) {
        atom.zero();
        switch (type) {
        case 1:
        case 2: {
            std::vector<double> c(m + 1);
            for (int j = 0; j <= m; j++) c[j] = 1.0 / (2 * std::cos(yv[j]) * dx);
            for (int j = 1; j <= m; j++) for (int k = 1; k <= l; k++) {
                for (int i = 1; i <= n - 1; i++)
                    atom(8, i, j, k) = (type == 1) ? u(i + 1, j, k) * c[j] : 2 * u(i + 1, j, k) * c[j];
                for (int i = 2; i <= n; i++)
                    atom(2, i, j, k) = (type == 1) ? -u(i - 1, j, k) * c[j] : -2 * u(i - 1, j, k) * c[j];
            }
        } break;
        case 3:
        case 4: {
            std::vector<double> c(m + 1);
            for (int j = 0; j <= m; j++) c[j] = 1.0 / (2 * std::cos(yv[j]) * dy);
            Arr3& q = (type == 3) ? v : u;
            for (int k = 1; k <= l; k++) for (int i = 1; i <= n; i++) {
                for (int j = 2; j <= m; j++) atom(4, i, j, k) = -q(i, j - 1, k) * std::cos(yv[j - 1]) * c[j];
                for (int j = 1; j <= m - 1; j++) atom(6, i, j, k) = q(i, j + 1, k) * std::cos(yv[j + 1]) * c[j];
            }
        } break;
        case 5: {
            for (int k = 1; k <= l; k++) {
                double tdzi = 1.0 / (8 * dfzT[k] * dz);
                for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
                    atom(23, i, j, k) = (w(i, j, k) + w(i, j + 1, k) + w(i + 1, j, k) + w(i + 1, j + 1, k)) * tdzi;
                    atom(14, i, j, k) = -(w(i, j, k - 1) + w(i, j + 1, k - 1) + w(i + 1, j, k - 1) + w(i + 1, j + 1, k - 1)) * tdzi;
                    atom(5, i, j, k) = atom(14, i, j, k) + atom(23, i, j, k);
                }
            }
        } break;
        case 6: {
            for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) for (int k = 1; k <= l; k++) {
                double tdzi = 1.0 / (8 * dfzT[k] * dz);
                double up = (u(i, j, k) + u(i, j, k + 1)) * tdzi;
                double dn = -(u(i, j, k) + u(i, j, k - 1)) * tdzi;
                atom(5, i, j, k) = up; atom(6, i, j, k) = up; atom(8, i, j, k) = up; atom(9, i, j, k) = up;
                atom(14, i, j, k) = dn; atom(15, i, j, k) = dn; atom(17, i, j, k) = dn; atom(18, i, j, k) = dn;
            }
        } break;
        case 7:
        case 8: {
            Arr3& q = (type == 7) ? v : u;
            for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++)
                atom(5, i, j, k) = q(i, j, k) * std::tan(yv[j]);
        } break;
        }
    }

    // ---------------- spf.F90:667-790 ----------------
    void vnlin(int type, Atom& atom, Arr3& u, Arr3& v, Arr3& w) {
        atom.zero();
        switch (type) {
        case 1:
        case 2: {
            std::vector<double> c(m + 1);
            for (int j = 0; j <= m; j++) c[j] = 1.0 / (2 * std::cos(yv[j]) * dx);
            Arr3& q = (type == 1) ? u : v;
            for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) {
                for (int i = 1; i <= n - 1; i++) atom(8, i, j, k) = q(i + 1, j, k) * c[j];
                for (int i = 2; i <= n; i++) atom(2, i, j, k) = -q(i - 1, j, k) * c[j];
            }
        } break;
        case 3:
        case 4: {
            std::vector<double> c(m + 1);
            for (int j = 0; j <= m; j++) c[j] = 1.0 / (2 * std::cos(yv[j]) * dy);
            for (int k = 1; k <= l; k++) for (int i = 1; i <= n; i++) {
                for (int j = 1; j <= m - 1; j++)
                    atom(6, i, j, k) = (type == 3) ? v(i, j + 1, k) * std::cos(yv[j + 1]) * c[j]
                                                   : 2 * v(i, j + 1, k) * std::cos(yv[j + 1]) * c[j];
                for (int j = 2; j <= m; j++)
                    atom(4, i, j, k) = (type == 3) ? -v(i, j - 1, k) * std::cos(yv[j - 1]) * c[j]
                                                   : -2 * v(i, j - 1, k) * std::cos(yv[j - 1]) * c[j];
            }
        } break;
        case 5: {
            for (int k = 1; k <= l; k++) {
                double tdzi = 1.0 / (8 * dfzT[k] * dz);
                for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
                    atom(23, i, j, k) = (w(i, j, k) + w(i, j + 1, k) + w(i + 1, j, k) + w(i + 1, j + 1, k)) * tdzi;
                    atom(14, i, j, k) = -(w(i, j, k - 1) + w(i, j + 1, k - 1) + w(i + 1, j, k - 1) + w(i + 1, j + 1, k - 1)) * tdzi;
                    atom(5, i, j, k) = atom(14, i, j, k) + atom(23, i, j, k);
                }
            }
        } break;
        case 6: {
            for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) for (int k = 1; k <= l; k++) {
                double tdzi = 1.0 / (8 * dfzT[k] * dz);
                double up = (v(i, j, k) + v(i, j, k + 1)) * tdzi;
                double dn = -(v(i, j, k) + v(i, j, k - 1)) * tdzi;
                atom(5, i, j, k) = up; atom(6, i, j, k) = up; atom(8, i, j, k) = up; atom(9, i, j, k) = up;
                atom(14, i, j, k) = dn; atom(15, i, j, k) = dn; atom(17, i, j, k) = dn; atom(18, i, j, k) = dn;
            }
        } break;
        case 7:
            for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++)
                atom(5, i, j, k) = u(i, j, k) * std::tan(yv[j]);
            break;
        case 8:
            for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++)
                atom(5, i, j, k) = 2 * u(i, j, k) * std::tan(yv[j]);
            break;
        }
    }

    // Al(:,A,B,:,:,1:l) = expr  /  An(:,A,B,...) += expr helpers
    template <class F> void setblk(std::vector<double>& M, int A, int B, F f) {
        for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            size_t base = aidx(1, A, B, i, j, k);
            size_t ab = (size_t)NP * ((i - 1) + (size_t)n * ((j - 1) + (size_t)m * (k - 1)));
            for (int loc = 0; loc < NP; loc++) M[base + loc] = f(ab + loc);
        }
    }
    template <class F> void addblk(std::vector<double>& M, int A, int B, F f) {
        for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            size_t base = aidx(1, A, B, i, j, k);
            size_t ab = (size_t)NP * ((i - 1) + (size_t)n * ((j - 1) + (size_t)m * (k - 1)));
            for (int loc = 0; loc < NP; loc++) M[base + loc] = M[base + loc] + f(ab + loc);
        }
    }

    // ---------------- usrc.F90:605-789 ----------------
    void lin() {
        Atom ucsi(n, m, l), uxx(n, m, l), uyy(n, m, l), uzz(n, m, l), uxs(n, m, l), fu(n, m, l), px(n, m, l);
        Atom sc(n, m, l), tcb(n, m, l), mc(n, m, l);
        double EV = par[EK_V], EH = par[EK_H], ph = (1 - par[MIXP]) * par[PE_H], pv = par[PE_V];
        double lambda = par[LAMB], xes = par[NLES], bi = par[BIOT], Ra = par[RAYL];
        std::fill(Al.begin(), Al.end(), 0.0);
        // u-equation (usrc.F90:683-693)
        uderiv(2, uxx); uderiv(3, uyy); uderiv(4, uzz); uderiv(5, ucsi); uderiv(6, uxs); coriolis(1, fu); gradp(1, px);
        setblk(Al, UU, UU, [&](size_t q) { return -EH * (uxx.a[q] + uyy.a[q] + ucsi.a[q]) - EV * uzz.a[q]; });
        setblk(Al, UU, VV, [&](size_t q) { return -fu.a[q] - EH * uxs.a[q]; });
        setblk(Al, UU, PP, [&](size_t q) { return px.a[q]; });
        // v-equation (usrc.F90:699-709)
        vderiv(2, uxx); vderiv(3, uyy); vderiv(4, uzz); vderiv(5, ucsi); vderiv(6, uxs); coriolis(2, fu); gradp(2, px);
        setblk(Al, VV, UU, [&](size_t q) { return fu.a[q] - EH * uxs.a[q]; });
        setblk(Al, VV, VV, [&](size_t q) { return -EH * (uxx.a[q] + uyy.a[q] + ucsi.a[q]) - EV * uzz.a[q]; });
        setblk(Al, VV, PP, [&](size_t q) { return px.a[q]; });
        // w-equation (usrc.F90:714-718)
        gradp(3, px); tderiv(6, uxs);
        setblk(Al, WW, PP, [&](size_t q) { return px.a[q]; });
        setblk(Al, WW, TT, [&](size_t q) { return -Ra * (1. + xes * alpt1) * uxs.a[q] / 2.; });
        setblk(Al, WW, SS, [&](size_t q) { return lambda * Ra * uxs.a[q] / 2.; });
        // p-equation (usrc.F90:723-728)
        pderiv(1, uxx); pderiv(2, uyy); pderiv(3, uzz);
        setblk(Al, PP, UU, [&](size_t q) { return uxx.a[q]; });
        setblk(Al, PP, VV, [&](size_t q) { return uyy.a[q]; });
        setblk(Al, PP, WW, [&](size_t q) { return uzz.a[q]; });
        // T-equation (usrc.F90:733-759)
        Atom& tc = fu;
        tderiv(1, tc); tderiv(2, sc); tderiv(3, uxx); tderiv(4, uyy); tderiv(5, uzz); tderiv(7, tcb);
        masksi(mc, msi);
        double dedt = lvsc * eta * qdim * (deltat / qdim) * dqso;
        if (coupled_T == 1) {
            setblk(Al, TT, TT, [&](size_t q) {
                return -ph * (uxx.a[q] + uyy.a[q]) - pv * uzz.a[q] + Ooa * tc.a[q] + dedt * sc.a[q] +
                       mc.a[q] * (QTnd * zeta * tc.a[q] - Ooa * tc.a[q] - dedt * sc.a[q]);
            });
            setblk(Al, TT, SS, [&](size_t q) { return -QTnd * zeta * a0 * mc.a[q]; });
        } else {
            setblk(Al, TT, TT, [&](size_t q) { return -ph * (uxx.a[q] + uyy.a[q]) - pv * uzz.a[q] + TRES * bi * tc.a[q]; });
        }
        // S-equation (usrc.F90:766-786)
        dedt = nus * (deltat / qdim) * dqso;
        double pQSnd = par[COMB] * par[SALT] * QSnd;
        if (coupled_S == 1) {
            setblk(Al, SS, SS, [&](size_t q) {
                return -ph * (uxx.a[q] + uyy.a[q]) - pv * uzz.a[q] - mc.a[q] * pQSnd * zeta * a0 / (rhodim * Lf);
            });
            double QSos = pQSnd * zeta / (rhodim * Lf);
            setblk(Al, SS, TT, [&](size_t q) { double QSoa = -dedt * sc.a[q]; return QSoa + mc.a[q] * (QSos - QSoa); });
        } else {
            setblk(Al, SS, SS, [&](size_t q) { return -ph * (uxx.a[q] + uyy.a[q]) - pv * uzz.a[q] + SRES * bi * sc.a[q]; });
        }
    }

    // ---------------- usrc.F90:1014-1121 ----------------
    void usol(const double* un, Arr3& u, Arr3& v, Arr3& w, Arr3& p, Arr3& t, Arr3& s) {
        u.zero(); v.zero(); w.zero(); p.zero(); t.zero(); s.zero();
        for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            u(i, j, k) = un[find_row2(i, j, k, UU) - 1];
            v(i, j, k) = un[find_row2(i, j, k, VV) - 1];
            w(i, j, k) = un[find_row2(i, j, k, WW) - 1];
            p(i, j, k) = un[find_row2(i, j, k, PP) - 1];
            t(i, j, k) = un[find_row2(i, j, k, TT) - 1];
            s(i, j, k) = un[find_row2(i, j, k, SS) - 1];
        }
        int N = n, M = m;
        for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) {
            if (periodic) {
                u(0, j, k) = u(N, j, k); v(0, j, k) = v(N, j, k);
                w(N + 1, j, k) = w(1, j, k); w(0, j, k) = w(N, j, k);
                p(N + 1, j, k) = p(1, j, k); p(0, j, k) = p(N, j, k);
                t(N + 1, j, k) = t(1, j, k); t(0, j, k) = t(N, j, k);
                s(N + 1, j, k) = s(1, j, k); s(0, j, k) = s(N, j, k);
            } else {
                u(0, j, k) = 0.0; u(N, j, k) = 0.0; v(0, j, k) = 0.0; v(N, j, k) = 0.0;
                p(0, j, k) = 0.0; p(N + 1, j, k) = 0.0;
                t(0, j, k) = t(1, j, k); t(N + 1, j, k) = t(N, j, k);
                s(0, j, k) = s(1, j, k); s(N + 1, j, k) = s(N, j, k);
            }
        }
        for (int k = 1; k <= l; k++) for (int i = 1; i <= n; i++) {
            u(i, 0, k) = 0.0; u(i, M, k) = 0.0; v(i, 0, k) = 0.0; v(i, M, k) = 0.0;
            p(i, 0, k) = 0.0; p(i, M + 1, k) = 0.0;
            t(i, 0, k) = t(i, 1, k); t(i, M + 1, k) = t(i, M, k);
            s(i, 0, k) = s(i, 1, k); s(i, M + 1, k) = s(i, M, k);
        }
        for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            u(i, j, 0) = u(i, j, 1); u(i, j, l + 1) = u(i, j, l);
            v(i, j, 0) = v(i, j, 1); v(i, j, l + 1) = v(i, j, l);
            w(i, j, l) = 0.0; w(i, j, 0) = 0.0;
            p(i, j, l + 1) = 0.0; p(i, j, 0) = 0.0;
            t(i, j, l + 1) = t(i, j, l); t(i, j, 0) = t(i, j, 1);
            s(i, j, l + 1) = s(i, j, l); s(i, j, 0) = s(i, j, 1);
        }
        for (int i = 1; i <= n; i++) for (int j = 1; j <= m; j++) for (int k = 1; k <= l; k++) {
            if (lm(i, j, k) == 1) {
                u(i, j, k) = 0.0; v(i, j, k) = 0.0;
                u(i - 1, j, k) = 0.0; v(i - 1, j, k) = 0.0;
                u(i, j - 1, k) = 0.0; v(i, j - 1, k) = 0.0;
                u(i - 1, j - 1, k) = 0.0; v(i - 1, j - 1, k) = 0.0;
            }
        }
    }

    struct Fields {
        Arr3 u, v, w, p, t, s;
        Fields(int n, int m, int l)
            : u(0, n, 0, m, 0, l + 1), v(0, n, 0, m, 0, l + 1), w(0, n + 1, 0, m + 1, 0, l),
              p(0, n + 1, 0, m + 1, 0, l + 1), t(0, n + 1, 0, m + 1, 0, l + 1), s(0, n + 1, 0, m + 1, 0, l + 1) {}
    };

    // ---------------- usrc.F90:792-887 ----------------
    void nlin_rhs(const double* un) {
        Fields F(n, m, l);
        Atom a1(n, m, l), a2(n, m, l), a3(n, m, l), a4(n, m, l);
        double epsr = par[ROSB], Ra = par[RAYL], xes = par[NLES];
        usol(un, F.u, F.v, F.w, F.p, F.t, F.s);
        // u-equation
        unlin(1, a1, F.u, F.v, F.w); unlin(3, a2, F.u, F.v, F.w); unlin(5, a3, F.u, F.v, F.w); unlin(7, a4, F.u, F.v, F.w);
        addblk(An, UU, UU, [&](size_t q) { return epsr * (a1.a[q] + a2.a[q] + a3.a[q] + a4.a[q]); });
        // v-equation: uvx, vvy, vwz, ut2
        vnlin(1, a1, F.u, F.v, F.w); vnlin(3, a2, F.u, F.v, F.w); vnlin(5, a3, F.u, F.v, F.w); vnlin(7, a4, F.u, F.v, F.w);
        addblk(An, VV, UU, [&](size_t q) { return epsr * a4.a[q]; });
        addblk(An, VV, VV, [&](size_t q) { return epsr * (a1.a[q] + a2.a[q] + a3.a[q]); });
        // w-equation
        wnlin(2, a1, F.t); wnlin(4, a2, F.t);
        for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            size_t base = aidx(1, WW, TT, i, j, k);
            size_t ab = (size_t)NP * ((i - 1) + (size_t)n * ((j - 1) + (size_t)m * (k - 1)));
            for (int loc = 0; loc < NP; loc++)
                An[base + loc] = An[base + loc] - Ra * xes * alpt2 * a1.a[ab + loc] + Ra * xes * alpt3 * a2.a[ab + loc];
        }
        // T-equation
        tnlin(3, a1, F.u, F.v, F.w, F.t); tnlin(5, a2, F.u, F.v, F.w, F.t); tnlin(7, a3, F.u, F.v, F.w, F.t);
        for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            size_t base = aidx(1, TT, TT, i, j, k);
            size_t ab = (size_t)NP * ((i - 1) + (size_t)n * ((j - 1) + (size_t)m * (k - 1)));
            for (int loc = 0; loc < NP; loc++) An[base + loc] = An[base + loc] + a1.a[ab + loc] + a2.a[ab + loc] + a3.a[ab + loc];
        }
        // S-equation
        tnlin(3, a1, F.u, F.v, F.w, F.s); tnlin(5, a2, F.u, F.v, F.w, F.s); tnlin(7, a3, F.u, F.v, F.w, F.s);
        for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            size_t base = aidx(1, SS, SS, i, j, k);
            size_t ab = (size_t)NP * ((i - 1) + (size_t)n * ((j - 1) + (size_t)m * (k - 1)));
            for (int loc = 0; loc < NP; loc++) An[base + loc] = An[base + loc] + a1.a[ab + loc] + a2.a[ab + loc] + a3.a[ab + loc];
        }
    }

    // ---------------- usrc.F90:890-1012 ----------------
    void nlin_jac(const double* un) {
        Fields F(n, m, l);
        Atom Urux(n, m, l), uvy1(n, m, l), Urvy1(n, m, l), uwz(n, m, l), Urwz(n, m, l), uvy2(n, m, l), Urvy2(n, m, l);
        double epsr = par[ROSB], Ra = par[RAYL], xes = par[NLES];
        usol(un, F.u, F.v, F.w, F.p, F.t, F.s);
        // u-equation (usrc.F90:943-952)
        unlin(2, Urux, F.u, F.v, F.w); unlin(3, uvy1, F.u, F.v, F.w); unlin(4, Urvy1, F.u, F.v, F.w);
        unlin(5, uwz, F.u, F.v, F.w); unlin(6, Urwz, F.u, F.v, F.w); unlin(7, uvy2, F.u, F.v, F.w); unlin(8, Urvy2, F.u, F.v, F.w);
        addblk(An, UU, UU, [&](size_t q) { return epsr * (Urux.a[q] + uvy1.a[q] + uwz.a[q] + uvy2.a[q]); });
        addblk(An, UU, VV, [&](size_t q) { return epsr * (Urvy1.a[q] + Urvy2.a[q]); });
        addblk(An, UU, WW, [&](size_t q) { return epsr * Urwz.a[q]; });
        // v-equation (usrc.F90:959-967)
        Atom &uvx = Urux, &uVrx = uvy1, &Vrvy = Urvy1, &vwz = uwz, &Vrwz = Urwz, &Urt2 = uvy2;
        vnlin(1, uvx, F.u, F.v, F.w); vnlin(2, uVrx, F.u, F.v, F.w); vnlin(4, Vrvy, F.u, F.v, F.w);
        vnlin(5, vwz, F.u, F.v, F.w); vnlin(6, Vrwz, F.u, F.v, F.w); vnlin(8, Urt2, F.u, F.v, F.w);
        addblk(An, VV, UU, [&](size_t q) { return epsr * (Urt2.a[q] + uVrx.a[q]); });
        addblk(An, VV, VV, [&](size_t q) { return epsr * (uvx.a[q] + Vrvy.a[q] + vwz.a[q]); });
        addblk(An, VV, WW, [&](size_t q) { return epsr * Vrwz.a[q]; });
        // w-equation (usrc.F90:973-976)
        Atom &t2r = Urux, &t3r = uvy1;
        wnlin(1, t2r, F.t); wnlin(3, t3r, F.t);
        for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            size_t base = aidx(1, WW, TT, i, j, k);
            size_t ab = (size_t)NP * ((i - 1) + (size_t)n * ((j - 1) + (size_t)m * (k - 1)));
            for (int loc = 0; loc < NP; loc++)
                An[base + loc] = An[base + loc] - Ra * xes * alpt2 * t2r.a[ab + loc] + Ra * xes * alpt3 * t3r.a[ab + loc];
        }
        // T- and S-equations (usrc.F90:982-1007)
        Atom &urTx = Urux, &Utrx = uvy1, &vrTy = Urvy1, &Vtry = uwz, &wrTz = Urwz, &Wtrz = uvy2;
        for (int pass = 0; pass < 2; pass++) {
            Arr3& tr = pass == 0 ? F.t : F.s;
            int R = pass == 0 ? TT : SS;
            tnlin(2, urTx, F.u, F.v, F.w, tr); tnlin(3, Utrx, F.u, F.v, F.w, tr); tnlin(4, vrTy, F.u, F.v, F.w, tr);
            tnlin(5, Vtry, F.u, F.v, F.w, tr); tnlin(6, wrTz, F.u, F.v, F.w, tr); tnlin(7, Wtrz, F.u, F.v, F.w, tr);
            addblk(An, R, UU, [&](size_t q) { return urTx.a[q]; });
            addblk(An, R, VV, [&](size_t q) { return vrTy.a[q]; });
            addblk(An, R, WW, [&](size_t q) { return wrTz.a[q]; });
            for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
                size_t base = aidx(1, R, R, i, j, k);
                size_t ab = (size_t)NP * ((i - 1) + (size_t)n * ((j - 1) + (size_t)m * (k - 1)));
                for (int loc = 0; loc < NP; loc++)
                    An[base + loc] = An[base + loc] + Utrx.a[ab + loc] + Vtry.a[ab + loc] + Wtrz.a[ab + loc];
            }
        }
    }

    // ---------------- boundary.F90:2-393 ----------------
    void boundaries() {
        for (int i = 1; i <= n; i++) for (int j = 1; j <= m; j++) for (int k = 1; k <= l; k++) {
            double* B = &An[aidx(1, 1, 1, i, j, k)];
            auto A = [&](int loc, int r, int c) -> double& { return B[(loc - 1) + NP * ((r - 1) + NUN * (c - 1))]; };
            auto addcol = [&](int dst, int src, int c) { for (int r = 1; r <= NUN; r++) A(dst, r, c) = A(dst, r, c) + A(src, r, c); };
            auto zcol = [&](int loc, int c) { for (int r = 1; r <= NUN; r++) A(loc, r, c) = 0.0; };
            auto zloc = [&](int loc) { for (int c = 1; c <= NUN; c++) for (int r = 1; r <= NUN; r++) A(loc, r, c) = 0.0; };
            auto zrow = [&](int r) { for (int c = 1; c <= NUN; c++) for (int loc = 1; loc <= NP; loc++) A(loc, r, c) = 0.0; };
            auto zuv = [&](int loc) { zcol(loc, UU); zcol(loc, VV); };
            int southw = lm(i - 1, j - 1, k), west = lm(i - 1, j, k), nwest = lm(i - 1, j + 1, k);
            int south = lm(i, j - 1, k), center = lm(i, j, k), north = lm(i, j + 1, k);
            int southe = lm(i + 1, j - 1, k), east = lm(i + 1, j, k), neast = lm(i + 1, j + 1, k);
            int southwb = lm(i - 1, j - 1, k - 1), westb = lm(i - 1, j, k - 1), nwestb = lm(i - 1, j + 1, k - 1);
            int southb = lm(i, j - 1, k - 1), bottom = lm(i, j, k - 1), northb = lm(i, j + 1, k - 1);
            int southeb = lm(i + 1, j - 1, k - 1), eastb = lm(i + 1, j, k - 1), neastb = lm(i + 1, j + 1, k - 1);
            int southwt = lm(i - 1, j - 1, k + 1), westt = lm(i - 1, j, k + 1), nwestt = lm(i - 1, j + 1, k + 1);
            int southt = lm(i, j - 1, k + 1), top = lm(i, j, k + 1), northt = lm(i, j + 1, k + 1);
            int southet = lm(i + 1, j - 1, k + 1), eastt = lm(i + 1, j, k + 1), neastt = lm(i + 1, j + 1, k + 1);
            int southee = 0, easteast = 0, northee = 0, nnorthee = 0, nnwest = 0, nnorth = 0, nneast = 0;
            if (i < n) {
                southee = lm(i + 2, j - 1, k); easteast = lm(i + 2, j, k); northee = lm(i + 2, j + 1, k);
                if (j < m) nnorthee = lm(i + 2, j + 2, k);
            }
            if (j < m) { nnwest = lm(i, j + 2, k); nnorth = lm(i, j + 2, k); nneast = lm(i, j + 2, k); }  // sic: boundary.F90:75-77

            if (center == OCEAN) {
                if (bottom == LAND) {
                    if (westb == LAND && southwb == LAND && southb == LAND) { addcol(1, 10, UU); addcol(1, 10, VV); }
                    zuv(10);
                    if (westb == LAND && neastb == LAND && northb == LAND) { addcol(2, 11, UU); addcol(2, 11, VV); }  // sic :91
                    zuv(11);
                    if (eastb == LAND && southeb == LAND && southb == LAND) { addcol(4, 13, UU); addcol(4, 13, VV); }
                    zuv(13);
                    if (eastb == LAND && neastb == LAND && northb == LAND) { addcol(5, 14, UU); addcol(5, 14, VV); }
                    addcol(5, 14, TT); addcol(5, 14, SS);
                    zloc(14);
                }
                if (southwb == LAND) zloc(10);
                if (westb == LAND) zloc(11);
                if (nwestb == LAND) zloc(12);
                if (southb == LAND) zloc(13);
                if (northb == LAND) zloc(15);
                if (southeb == LAND) zloc(16);
                if (eastb == LAND) zloc(17);
                if (neastb == LAND) zloc(18);
                if (top == LAND) {
                    if (westt == LAND && southwt == LAND && southt == LAND) { addcol(1, 19, UU); addcol(1, 19, VV); }
                    zuv(19);
                    if (westt == LAND && nwestt == LAND && northt == LAND) { addcol(2, 20, UU); addcol(2, 20, VV); }
                    zuv(20);
                    if (eastt == LAND && southet == LAND && southt == LAND) { addcol(4, 22, UU); addcol(4, 22, VV); }
                    zuv(22);
                    if (eastt == LAND && neastt == LAND && northt == LAND) { addcol(5, 23, UU); addcol(5, 23, VV); }
                    addcol(5, 23, TT); addcol(5, 23, SS);
                    zloc(23);
                    Frc[find_row2(i, j, k, WW) - 1] = 0.0;
                    zrow(WW);
                    for (int r = 1; r <= NUN; r++) { A(5, r, WW) = 1.0e-10; A(6, r, WW) = 1.0e-10; A(8, r, WW) = 1.0e-10; A(9, r, WW) = 1.0e-10; }
                    A(5, WW, WW) = 1.0;
                }
                if (southwt == LAND) zloc(19);
                if (westt == LAND) zloc(20);
                if (nwestt == LAND) zloc(21);
                if (southt == LAND) zloc(22);
                if (northt == LAND) zloc(24);
                if (southet == LAND) zloc(25);
                if (eastt == LAND) zloc(26);
                if (neastt == LAND) zloc(27);
                if (southw == LAND) zuv(1);
                if (west == LAND) { addcol(5, 2, TT); addcol(5, 2, SS); zloc(2); zuv(1); }
                if (nwest == LAND) { zuv(2); zuv(3); }
                else if (j < m) { if (nnwest == LAND) zuv(3); }
                if (south == LAND) { addcol(5, 4, SS); addcol(5, 4, TT); zloc(4); zuv(1); }
                if (north == LAND) {
                    zuv(2);
                    A(2, PP, UU) = 0.0; A(2, PP, VV) = 0.0; A(5, PP, UU) = 0.0; A(5, PP, VV) = 0.0;
                    Frc[find_row2(i, j, k, VV) - 1] = 0.0;
                    zrow(VV); zcol(5, VV); A(5, VV, VV) = 1.0;
                    Frc[find_row2(i, j, k, UU) - 1] = 0.0;
                    zrow(UU); zcol(5, UU); A(5, UU, UU) = 1.0;
                    addcol(5, 6, SS); addcol(5, 6, TT);
                    zloc(6);
                } else if (j < m) {
                    if (nnorth == LAND) { zuv(3); zuv(6); }
                }
                if (southe == LAND) { zuv(4); zuv(7); }
                else if (i < n) { if (southee == LAND) zuv(7); }
                if (east == LAND) {
                    zuv(4);
                    A(4, PP, UU) = 0.0; A(4, PP, VV) = 0.0; A(5, PP, UU) = 0.0; A(5, PP, VV) = 0.0;
                    Frc[find_row2(i, j, k, UU) - 1] = 0.0;
                    zrow(UU); zcol(5, UU); A(5, UU, UU) = 1.0;
                    Frc[find_row2(i, j, k, VV) - 1] = 0.0;
                    zrow(VV); zcol(5, VV); A(5, VV, VV) = 1.0;
                    addcol(5, 8, SS); addcol(5, 8, TT);
                    zloc(8);
                    zuv(7);
                } else if (i < n) {
                    if (easteast == LAND) { zuv(7); zuv(8); }
                }
                if (neast == LAND) {
                    Frc[find_row2(i, j, k, UU) - 1] = 0.0;
                    zrow(UU); zcol(5, UU); A(5, UU, UU) = 1.0;
                    Frc[find_row2(i, j, k, VV) - 1] = 0.0;
                    zrow(VV); zcol(5, VV); A(5, VV, VV) = 1.0;
                    zuv(7);
                } else if (i < n || j < m) {
                    if (i < n) {
                        if (northee == LAND) { zuv(8); zuv(9); }
                        else if (j < m) { if (nnorthee == LAND) zuv(9); }
                    }
                    if (j < m) { if (nneast == LAND) { zuv(6); zuv(9); } }
                }
            } else {
                for (int q = 0; q < NP * NUN * NUN; q++) B[q] = 0.0;
                for (int ii = 1; ii <= NUN; ii++) { Frc[find_row2(i, j, k, ii) - 1] = 0.0; A(5, ii, ii) = 1.0; }
            }
        }
    }

    // assemble.F90:142-179
    void shift(int i, int j, int k, int& i2, int& j2, int& k2, int kk) const {
        if (kk < 10) { k2 = k; j2 = j - 1 + (kk + 2) % 3; i2 = i - 1 + (kk - 1) / 3; }
        else if (kk < 19) { k2 = k - 1; j2 = j - 1 + (kk + 2) % 3; i2 = i - 1 + (kk - 10) / 3; }
        else { k2 = k + 1; j2 = j - 1 + (kk + 2) % 3; i2 = i - 1 + (kk - 19) / 3; }
        if (periodic) { if (i2 == 0) i2 = n; else if (i2 == n + 1) i2 = 1; }
    }

    // ---------------- assemble.F90:57-139 ----------------
    void fillcolA() {
        std::fill(begA.begin(), begA.end(), 0);
        int v = 1, row = 1;
        for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            const double* Alocal = &An[aidx(1, 1, 1, i, j, k)];
            for (int ii = 1; ii <= NUN; ii++) {
                begA[row - 1] = v;
                for (int kk = 1; kk <= NP; kk++) for (int jj = 1; jj <= NUN; jj++) {
                    double a = Alocal[(kk - 1) + NP * ((ii - 1) + NUN * (jj - 1))];
                    if (std::fabs(a) > 1.0e-10) {
                        coA[v - 1] = a;
                        int i2, j2, k2;
                        shift(i, j, k, i2, j2, k2, kk);
                        if (i2 < 1 || i2 > n || j2 < 1 || j2 > m || k2 < 1 || k2 > l) bad_columns++;
                        jcoA[v - 1] = find_row2(i2, j2, k2, jj);
                        v++;
                    }
                }
                row++;
            }
        }
        begA[ndim] = v;
    }

    // ---------------- assemble.F90:18-54 ----------------
    void fillcolB() {
        std::fill(coB.begin(), coB.end(), 0.0);
        for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            if (lm(i, j, k) == OCEAN) {
                if (lm(i + 1, j, k) != LAND) coB[find_row2(i, j, k, UU) - 1] = -par[ROSB];
                if (lm(i, j + 1, k) != LAND) coB[find_row2(i, j, k, VV) - 1] = -par[ROSB];
                coB[find_row2(i, j, k, TT) - 1] = -1.0;
                coB[find_row2(i, j, k, SS) - 1] = -1.0;
            }
        }
    }

    // ---------------- matetc.F90:147-166 ----------------
    void matAvec(const double* v1, double* v2) {
        for (int i = 0; i < ndim; i++) v2[i] = 0.0;
        for (int i = 1; i <= ndim; i++)
            for (int v = begA[i - 1]; v <= begA[i] - 1; v++) v2[i - 1] = coA[v - 1] * v1[jcoA[v - 1] - 1] + v2[i - 1];
    }

    // ---------------- forcing.F90:405-449 ----------------
    double wfun(double yy, int v1) {
        if (v1 == 1)
            return 0.2 - 0.8 * std::sin(6 * std::fabs(yy)) - 0.5 * (1 - std::tanh(10 * std::fabs(yy))) -
                   0.5 * (1 - std::tanh(10 * (PI / 2 - std::fabs(yy))));
        return 0.0;
    }
    double temfun(double yy) {
        if (forcing_type == 2) return std::cos(PI * (yy - ymin_glob) / (ymax_glob - ymin_glob));
        return std::cos(PI * yy / ymax_glob) + par[CMPR] * std::sin(PI * yy / ymax_glob);
    }
    double salfun(double yy) {
        if (forcing_type == 2) return std::cos(PI * (yy - ymin_glob) / (ymax_glob - ymin_glob));
        if (forcing_type == 1) return (std::cos(PI * yy / ymax_glob) + par[FPER] * yy / ymax_glob) / std::cos(yy);
        return std::cos(PI * yy / ymax_glob) + par[FPER] * yy / ymax_glob;
    }
    // forcing.F90:452-464 -> THCM.C:2653-2686 (single rank: all cells are "real")
    double qint(std::vector<double>& field) {
        double lfsint = 0.0, lsint = 0.0;
        for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            lfsint = f2(field, i, j) * std::cos(y[j]) * (1 - lm(i, j, l)) + lfsint;
            lsint = std::cos(y[j]) * (1 - lm(i, j, l)) + lsint;
        }
        return lfsint / lsint;
    }

    // ---------------- forcing.F90:4-218 ----------------
    void forcing() {
        std::fill(Frc.begin(), Frc.end(), 0.0);
        double sigma = par[COMB] * par[WIND] * par[AL_T];
        if (iza == 2)
            for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) { f2(taux, i, j) = wfun(yv[j], 1); f2(tauy, i, j) = wfun(yv[j], 2); }
        for (int j = 1; j <= m - 1; j++) for (int i = 1; i <= n; i++) {
            Frc[find_row2(i, j, l, UU) - 1] = sigma * f2(taux, i, j);
            Frc[find_row2(i, j, l, VV) - 1] = sigma * f2(tauy, i, j);
        }
        double etabi = par[COMB] * par[TEMP] * (1 - TRES + TRES * par[BIOT]);
        double temcor = 0.0;
        if (ite == 1 && coupled_T == 0) {
            for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) f2(tatm, i, j) = temfun(y[j]);
            if (TRES == 0) temcor = qint(tatm);
        }
        for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            if (coupled_T == 1) {
                double QToa = par[COMB] * par[SUNP] * suno[j] * (1 - albe0 - albed * f2(albe, i, j)) + Ooa * f2(tatm, i, j) +
                              lvsc * eta * qdim * f2(qatm, i, j) - lvsc * eo0;
                double QTos = QTnd * zeta * (a0 * s0 - t0);
                Frc[find_row2(i, j, l, TT) - 1] = (QToa + f2(msi, i, j) * (QTos - QToa)) * (1 - lm(i, j, l));
            } else {
                Frc[find_row2(i, j, l, TT) - 1] = etabi * (f2(tatm, i, j) - temcor);
            }
        }
        double gamma;
        if (coupled_S == 1) gamma = par[COMB] * par[SALT];
        else gamma = par[COMB] * par[SALT] * (1 - SRES + SRES * par[BIOT]);
        double salcor = 0.0, adapted_salcor = 0.0, spertcor = 0.0;
        if (its == 1) {
            for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) f2(emip, i, j) = salfun(y[j]) * (1 - lm(i, j, l));
            if (SRES == 0 && coupled_S == 0) salcor = qint(emip);
        }
        if (SRES == 0 && coupled_S == 0) { adapted_salcor = qint(adapted_emip); spertcor = qint(spert); }
        double pQSnd = par[COMB] * par[SALT] * QSnd;
        for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            if (coupled_S == 1) {
                double QSoa = pQSnd * (eo0 - eta * qdim * f2(qatm, i, j) - f2(patm, i, j));
                double QSos = pQSnd * (zeta * (a0 * s0 - t0) - Qvar * f2(qsa, i, j) - Q0) / (rhodim * Lf);
                Frc[find_row2(i, j, l, SS) - 1] = (QSoa + f2(msi, i, j) * (QSos - QSoa) - f2(gsi, i, j)) * (1 - lm(i, j, l));
            } else {
                Frc[find_row2(i, j, l, SS) - 1] = gamma * (1 - par[HMTP]) * (f2(emip, i, j) - salcor) +
                                                  gamma * par[HMTP] * (f2(adapted_emip, i, j) - adapted_salcor) +
                                                  par[SPER] * (1 - SRES + SRES * par[BIOT]) * (f2(spert, i, j) - spertcor);
            }
        }
        for (int k = 1; k <= l - 1; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            Frc[find_row2(i, j, k, WW) - 1] =
                -par[COMB] * (1 - lm(i, j, k)) * par[RAYL] *
                (par[LAMB] * (f3(internal_salt, i, j, k) + f3(internal_salt, i, j, k + 1)) / 2. -
                 (f3(internal_temp, i, j, k) + f3(internal_temp, i, j, k + 1)) / 2.);
        }
    }

    // ---------------- usrc.F90:6-139 ----------------
    void init(int a_n, int a_m, int a_l, double a_xmin, double a_xmax, double a_ymin, double a_ymax, const int* a_landm) {
        n = a_n; m = a_m; l = a_l; ndim = n * m * l * NUN;
        xmin = a_xmin; xmax = a_xmax; ymin = a_ymin; ymax = a_ymax;
        landm.assign((size_t)(n + 2) * (m + 2) * (l + 2), OCEAN);
        size_t nm = (size_t)n * m;
        Frc.assign(ndim, 0.0);
        for (auto* f : {&taux, &tauy, &tatm, &emip, &spert, &adapted_emip, &qatm, &albe, &patm, &msi, &gsi, &qsa}) f->assign(nm, 0.0);
        internal_temp.assign(nm * l, 0.0); internal_salt.assign(nm * l, 0.0);
        Al.assign((size_t)NP * NUN * NUN * nm * l, 0.0);
        An.assign((size_t)NP * NUN * NUN * nm * l, 0.0);
        begA.assign(ndim + 1, 0);
        size_t maxnnz = (size_t)ndim * (NUN * NP + 1);  // mat.F90:56-68
        jcoA.assign(maxnnz, 0); coA.assign(maxnnz, 0.0); coB.assign(ndim, 0.0);
        set_landmask_raw(a_landm, false);
        grid();
        double dzne = dz * dfzT[l];
        QTnd = r0dim / (udim * cp0 * rhodim * hdim * dzne);
        QSnd = s0 * r0dim / (deltas * udim * hdim * dzne);
        stpnt();
        atmos_coef();
        forcing();
        lin();
        vmix_init();   // usrc.F90:133
    }

    // usrc.F90:79-107 / 353-408
    void set_landmask_raw(const int* a_landm, bool fix_inversion) {
        size_t pos = 0;
        for (int k = 0; k <= l + 1; k++) for (int j = 0; j <= m + 1; j++) for (int i = 0; i <= n + 1; i++) {
            lm(i, j, k) = a_landm[pos];
            if (!periodic && lm(i, j, k) == PERIO) lm(i, j, k) = OCEAN;
            pos++;
        }
        if (fix_inversion)
            for (int i = 1; i <= n; i++) for (int j = 1; j <= m; j++) for (int k = l; k >= 2; k--)
                if (lm(i, j, k) == LAND && lm(i, j, k - 1) == OCEAN) lm(i, j, k - 1) = LAND;
        for (int k = 0; k <= l + 1; k++) for (int j = 0; j <= m + 1; j++) {
            if (!periodic) { lm(0, j, k) = LAND; lm(n + 1, j, k) = LAND; }
        }
        for (int k = 0; k <= l + 1; k++) for (int i = 0; i <= n + 1; i++) { lm(i, 0, k) = LAND; lm(i, m + 1, k) = LAND; }
        for (int j = 0; j <= m + 1; j++) for (int i = 0; i <= n + 1; i++) { lm(i, j, 0) = LAND; lm(i, j, l + 1) = LAND; }
    }


    // =====================================================================================
    // Tracer mixing (mix_imp.f).  vmix_flag / vmix_temp / vmix_salt / vmix_fix as in mix_imp.f:61-100, mix.F90.
    // =====================================================================================
    int vmix_flag = 0, vmix_temp = 0, vmix_salt = 0, vmix_fix = 1;
    int vmix_dim = 0;   // number of (row, column) pairs of the mixing pattern (vmix_part): mixing is applied only when > 0
    // mix_imp.f:171-228 with vmix_el_1/2 (mix_imp.f:860-1048): only the COUNT of pairs matters here
    void vmix_part() {
        long cnt = 0;
        for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            if (lm(i, j, k) != OCEAN) continue;
            int tel = 0;
            for (int dk = -1; dk <= 1; dk++) for (int dj = -1; dj <= 1; dj++) for (int di = -1; di <= 1; di++) {
                int kind = lm(i + di, j + dj, k + dk);
                if (kind == OCEAN || kind == PERIO) tel++;
            }
            if (vmix_flag == 1) cnt += 4 * tel;
            else cnt += (long)tel * ((vmix_temp ? 1 + vmix_salt : 0) + (vmix_salt ? 1 + vmix_temp : 0));
        }
        vmix_dim = cnt > 0 ? 1 : 0;
    }
    // mix_imp.f:61-100
    void vmix_init() {
        if (vmix == 0) { vmix_flag = 0; vmix_fix = 1; }
        else if (vmix == 1) { vmix_flag = 1; vmix_fix = 1; }
        else if (vmix == 2) { vmix_flag = 2; vmix_fix = 0; }
        else vmix_flag = -1;
        if (vmix_flag == 1) { vmix_temp = 1; vmix_salt = 1; vmix_part(); } else { vmix_temp = 0; vmix_salt = 0; }
    }
    // mix_imp.f:139-169 (l2nrm, mix_sup.F90: sqrt of the sum of squares over the n*m*l field values)
    void vmix_control(const double* un) {
        auto l2 = [&](int XX) { double a = 0.0; for (int q = XX - 1; q < ndim; q += NUN) a += un[q] * un[q]; return std::sqrt(a); };
        int test_temp = l2(TT) > 1.0e-12 ? 1 : 0, test_salt = l2(SS) > 1.0e-12 ? 1 : 0;
        if (vmix_temp != test_temp || vmix_salt != test_salt) {
            vmix_temp = test_temp; vmix_salt = test_salt;
            if ((vmix_temp != 0) && (vmix_temp != 0)) vmix_part();   // (sic: the reference tests vmix_temp twice, mix_imp.f:163)
        }
        vmix_fix = 1;
    }
    // mix_imp.f:817-835
    inline double isoc(int i, int j, int k) { int v = lm(i, j, k); return (v == OCEAN || v == PERIO) ? 1.0 : 0.0; }
    // mix_imp.f:837-857
    inline double tprstb(double grad, double spl) const {
        double fac = alphaT * spl;
        double a = -grad * fac;
        return std::max(mix_tanh(a * a * a), 0.0);
    }
    // mix_imp.f:675-727
    void tprslp(double drdh, double& drdz, double spl, double& slp, double& tpr) const {
        const double width = 1.0, epsln = 1.0e-20;
        if (drdz == 0.0) drdz = epsln;
        slp = -drdh / drdz;
        double absslp = std::sqrt(slp * slp);
        double delta = (r0dim / hdim) * spl;
        double sd = width * delta;
        if (tap == 1) { tpr = absslp > delta ? (delta / absslp) * (delta / absslp) : 1.0; }
        else if (tap == 2) { tpr = 0.5 * (1.0 - mix_tanh((absslp - delta) / sd)); }
        else if (tap == 3) {
            if (absslp < delta - sd && drdz < 0.0) tpr = 1.0;
            else if (absslp >= delta - sd && absslp < delta && drdz < 0.0) { double dum = (absslp - (delta - sd)) / sd; tpr = 1.0 - 3.0 * (dum * dum) + 2.0 * (dum * dum * dum); }
            else tpr = 0.0;
        } else tpr = 1.0;
    }
    // mix_imp.f:231-562
    void vmix_fun(const double* un, double* mix) {
        Fields f(n, m, l);
        usol(un, f.u, f.v, f.w, f.p, f.t, f.s);
        Arr3 &t = f.t, &s = f.s;
        Arr3 dtdxe(0, n, 0, m + 1, 0, l + 1), dsdxe(0, n, 0, m + 1, 0, l + 1), dtdyn(0, n + 1, 0, m, 0, l + 1), dsdyn(0, n + 1, 0, m, 0, l + 1),
            dtdzt(0, n + 1, 0, m + 1, 0, l), dsdzt(0, n + 1, 0, m + 1, 0, l), rho(0, n + 1, 0, m + 1, 0, l + 1),
            drhods(0, n + 1, 0, m + 1, 0, l + 1), drhodt(0, n + 1, 0, m + 1, 0, l + 1), drhodzt(0, n + 1, 0, m + 1, 0, l);
        Arr3 Ftxe(0, n, 1, m, 1, l), Fsxe(0, n, 1, m, 1, l), Ftyn(1, n, 0, m, 1, l), Fsyn(1, n, 0, m, 1, l), Ftzt(1, n, 1, m, 0, l),
            Fszt(1, n, 1, m, 0, l), Ftimp(1, n, 1, m, 0, l), Fsimp(1, n, 1, m, 0, l);
        const double xes = par[NLES], lambda = par[LAMB], piso = par[MIXP] * par[PE_H], pgm = par[MKAP] * par[PE_H],
                     eps = (1.0 - par[ALPC]) * par[ENER] * par[PE_V], kvc = par[P_VC], sp1 = par[SPL1], sp2 = par[SPL2];
        const double epsln = 1.0e-20;
        // mix_imp.f:564-641
        auto dCdxt = [&](Arr3& C, Arr3& d) { for (int k = 0; k <= l + 1; k++) for (int j = 0; j <= m + 1; j++) for (int i = 0; i <= n; i++)
            d(i, j, k) = isoc(i + 1, j, k) * isoc(i, j, k) * (C(i + 1, j, k) - C(i, j, k)) / (dx * std::cos(y[j])); };
        auto dCdyt = [&](Arr3& C, Arr3& d) { for (int k = 0; k <= l + 1; k++) for (int j = 0; j <= m; j++) for (int i = 0; i <= n + 1; i++)
            d(i, j, k) = isoc(i, j + 1, k) * isoc(i, j, k) * (C(i, j + 1, k) - C(i, j, k)) / dy; };
        auto dCdzt = [&](Arr3& C, Arr3& d) { for (int k = 0; k <= l; k++) for (int j = 0; j <= m + 1; j++) for (int i = 0; i <= n + 1; i++)
            d(i, j, k) = isoc(i, j, k + 1) * isoc(i, j, k) * (C(i, j, k + 1) - C(i, j, k)) / (dz * dfzW[k]); };
        dCdxt(t, dtdxe); dCdxt(s, dsdxe); dCdyt(t, dtdyn); dCdyt(s, dsdyn); dCdzt(t, dtdzt); dCdzt(s, dsdzt);
        for (size_t q = 0; q < rho.a.size(); q++) {
            double tt = t.a[q];
            rho.a[q] = lambda * s.a[q] - tt - xes * (alpt1 * tt + alpt2 * tt * tt - alpt3 * tt * tt * tt);
            drhodt.a[q] = -1.0 - xes * (alpt1 + 2.0 * alpt2 * tt - 3.0 * alpt3 * (tt * tt));   // drhodC, mix_imp.f:643-673
            drhods.a[q] = lambda;
        }
        dCdzt(rho, drhodzt);
        const bool npgm = (piso != 0.0) || (pgm != 0.0);
        // east-face flux of neutral physics + GM at (i,j,k) (identical statements at i = 0 and in the interior)
        auto east_face = [&](int i, int j, int k) {
            double dumt = 0.0, dums = 0.0;
            for (int kr = 0; kr <= 1; kr++) for (int ip = 0; ip <= 1; ip++) {
                double drdh = (drhodt(i + ip, j, k) * dtdxe(i, j, k) + drhods(i + ip, j, k) * dsdxe(i, j, k));
                double drdz = (drhodt(i + ip, j, k) * dtdzt(i + ip, j, k - 1 + kr) + drhods(i + ip, j, k) * dsdzt(i + ip, j, k - 1 + kr));
                double slp, tpr; tprslp(drdh, drdz, sp2, slp, tpr);
                dumt = dumt + dfzW[k - 1 + kr] * (tpr * (piso) * dtdxe(i, j, k) + tpr * (piso - pgm) * slp * dtdzt(i + ip, j, k - 1 + kr));
                dums = dums + dfzW[k - 1 + kr] * (tpr * (piso) * dsdxe(i, j, k) + tpr * (piso - pgm) * slp * dsdzt(i + ip, j, k - 1 + kr));
            }
            Ftxe(i, j, k) = Ftxe(i, j, k) - dumt / (4 * dfzT[k]);
            Fsxe(i, j, k) = Fsxe(i, j, k) - dums / (4 * dfzT[k]);
        };
        for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) {
            if (npgm) east_face(0, j, k);
            for (int i = 1; i <= n; i++) {
                if (npgm) {
                    east_face(i, j, k);
                    double dumt = 0.0, dums = 0.0;   // north face
                    for (int kr = 0; kr <= 1; kr++) for (int jq = 0; jq <= 1; jq++) {
                        double drdh = (drhodt(i, j + jq, k) * dtdyn(i, j, k) + drhods(i, j + jq, k) * dsdyn(i, j, k));
                        double drdz = (drhodt(i, j + jq, k) * dtdzt(i, j + jq, k - 1 + kr) + drhods(i, j + jq, k) * dsdzt(i, j + jq, k - 1 + kr));
                        double slp, tpr; tprslp(drdh, drdz, sp2, slp, tpr);
                        dumt = dumt + dfzW[k - 1 + kr] * std::cos(y[j + jq]) * (tpr * (piso) * dtdyn(i, j, k) + tpr * (piso - pgm) * slp * dtdzt(i, j + jq, k - 1 + kr));
                        dums = dums + dfzW[k - 1 + kr] * std::cos(y[j + jq]) * (tpr * (piso) * dsdyn(i, j, k) + tpr * (piso - pgm) * slp * dsdzt(i, j + jq, k - 1 + kr));
                    }
                    Ftyn(i, j, k) = Ftyn(i, j, k) - dumt / (4 * dfzT[k] * std::cos(yv[j]));
                    Fsyn(i, j, k) = Fsyn(i, j, k) - dums / (4 * dfzT[k] * std::cos(yv[j]));
                    dumt = 0.0; dums = 0.0;          // top face, zonal variations
                    for (int ip = 0; ip <= 1; ip++) for (int kr = 0; kr <= 1; kr++) {
                        double drdh = (drhodt(i, j, k + kr) * dtdxe(i - 1 + ip, j, k + kr) + drhods(i, j, k + kr) * dsdxe(i - 1 + ip, j, k + kr));
                        double drdz = (drhodt(i, j, k + kr) * dtdzt(i, j, k) + drhods(i, j, k + kr) * dsdzt(i, j, k));
                        double slp, tpr; tprslp(drdh, drdz, sp2, slp, tpr);
                        dumt = dumt + (tpr * (piso) * slp * slp * dtdzt(i, j, k) + tpr * (piso + pgm) * slp * dtdxe(i - 1 + ip, j, k + kr));
                        dums = dums + (tpr * (piso) * slp * slp * dsdzt(i, j, k) + tpr * (piso + pgm) * slp * dsdxe(i - 1 + ip, j, k + kr));
                    }
                    Ftzt(i, j, k) = Ftzt(i, j, k) - dumt / 4;
                    Fszt(i, j, k) = Fszt(i, j, k) - dums / 4;
                    dumt = 0.0; dums = 0.0;          // top face, meridional variations
                    for (int jq = 0; jq <= 1; jq++) for (int kr = 0; kr <= 1; kr++) {
                        double drdh = (drhodt(i, j, k + kr) * dtdyn(i, j - 1 + jq, k + kr) + drhods(i, j, k + kr) * dsdyn(i, j - 1 + jq, k + kr));
                        double drdz = (drhodt(i, j, k + kr) * dtdzt(i, j, k) + drhods(i, j, k + kr) * dsdzt(i, j, k));
                        double slp, tpr; tprslp(drdh, drdz, sp2, slp, tpr);
                        dumt = dumt + (tpr * (piso) * slp * slp * dtdzt(i, j, k) + tpr * (piso + pgm) * slp * dtdyn(i, j - 1 + jq, k + kr));
                        dums = dums + (tpr * (piso) * slp * slp * dsdzt(i, j, k) + tpr * (piso + pgm) * slp * dsdyn(i, j - 1 + jq, k + kr));
                    }
                    Ftzt(i, j, k) = Ftzt(i, j, k) - dumt / 4;
                    Fszt(i, j, k) = Fszt(i, j, k) - dums / 4;
                }
                if (eps != 0.0) {   // consistent vertical mixing
                    Ftzt(i, j, k) = Ftzt(i, j, k) + tprstb(drhodzt(i, j, k), sp1) * eps * dtdzt(i, j, k) / (drhodzt(i, j, k) - epsln);
                    Fszt(i, j, k) = Fszt(i, j, k) + tprstb(drhodzt(i, j, k), sp1) * eps * dsdzt(i, j, k) / (drhodzt(i, j, k) - epsln);
                }
                if (kvc != 0.0) {   // implicit vertical mixing / convective adjustment
                    Ftimp(i, j, k) = -tprstb(-drhodzt(i, j, k), sp1) * kvc * dtdzt(i, j, k);
                    Fsimp(i, j, k) = -tprstb(-drhodzt(i, j, k), sp1) * kvc * dsdzt(i, j, k);
                }
            }
        }
        // divergence of the fluxes (mix_imp.f:495-560)
        for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            int row = find_row2(i, j, k, TT) - 1;
            mix[row] = 0.0;
            if (vmix_temp == 1) {
                mix[row] = (Ftxe(i, j, k) - Ftxe(i - 1, j, k)) / (dx * std::cos(y[j])) + mix[row];
                mix[row] = (Ftyn(i, j, k) * std::cos(yv[j]) - Ftyn(i, j - 1, k) * std::cos(yv[j - 1])) / (dy * std::cos(y[j])) + mix[row];
                mix[row] = (Ftzt(i, j, k) - Ftzt(i, j, k - 1)) / (dz * dfzT[k]) + mix[row];
                if (rho_mixing && xes == 0.0)
                    mix[row] = ((Ftimp(i, j, k) - Ftimp(i, j, k - 1)) - (Fsimp(i, j, k) - Fsimp(i, j, k - 1)) * lambda) / (2.0 * dz * dfzT[k]) + mix[row];
                else
                    mix[row] = (Ftimp(i, j, k) - Ftimp(i, j, k - 1)) / (dz * dfzT[k]) + mix[row];
            }
            row = find_row2(i, j, k, SS) - 1;
            mix[row] = 0.0;
            if (vmix_salt == 1) {
                mix[row] = (Fsxe(i, j, k) - Fsxe(i - 1, j, k)) / (dx * std::cos(y[j])) + mix[row];
                mix[row] = (Fsyn(i, j, k) * std::cos(yv[j]) - Fsyn(i, j - 1, k) * std::cos(yv[j - 1])) / (dy * std::cos(y[j])) + mix[row];
                mix[row] = (Fszt(i, j, k) - Fszt(i, j, k - 1)) / (dz * dfzT[k]) + mix[row];
                if (rho_mixing && xes == 0.0)
                    mix[row] = ((Fsimp(i, j, k) - Fsimp(i, j, k - 1)) - (Ftimp(i, j, k) - Ftimp(i, j, k - 1)) / lambda) / (2.0 * dz * dfzT[k]) + mix[row];
                else
                    mix[row] = (Fsimp(i, j, k) - Fsimp(i, j, k - 1)) / (dz * dfzT[k]) + mix[row];
            }
        }
    }
    // mix_imp.f:729-815: forward-difference Jacobian of vmix_fun over groups of structurally orthogonal columns
    // (Coleman-More).  The reference colours the pattern of vmix_el_1/2 (mix_imp.f:860-1048: T,S rows of OCEAN cells x T,S
    // unknowns of their OCEAN/PERIO neighbours among the 27) with MINPACK's DSM; since no row meets two columns of one
    // group, fjac(row,col) = (mix(un + eps e_group)(row) - mix(un)(row)) / eps does not depend on WHICH valid colouring is
    // used -- here: colour = (i, j, k) position modulo 3 (with extra colours closing a periodic ring) x {T,S}.
    void vmix_jac(const double* un) {
        const double eps = 1.0e-08;
        std::vector<double> mix(ndim, 0.0), mixd(ndim, 0.0), und(ndim), d(ndim);
        vmix_fun(un, mix.data());
        auto ring = [&](int i, int nn, bool per) { if (!per || nn % 3 == 0) return (i - 1) % 3; int full = nn - nn % 3; return i <= full ? (i - 1) % 3 : 3 + (i - 1 - full); };
        const int ci = (periodic && n % 3 != 0) ? 3 + n % 3 : 3;
        if (periodic && n < 4 && n % 3 != 0) { fprintf(stderr, "thcm_oracle: vmix_jac needs n >= 4 on periodic grids\n"); std::abort(); }
        for (int var = 0; var < 2; var++) for (int gk = 0; gk < 3; gk++) for (int gj = 0; gj < 3; gj++) for (int gi = 0; gi < ci; gi++) {
            if ((var == 0 && !vmix_temp) || (var == 1 && !vmix_salt)) continue;
            std::fill(d.begin(), d.end(), 0.0);
            bool any = false;
            for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
                if ((k - 1) % 3 != gk || (j - 1) % 3 != gj || ring(i, n, periodic) != gi) continue;
                if (lm(i, j, k) != OCEAN) continue;            // only unknowns of OCEAN cells are columns (vmix_el)
                d[find_row2(i, j, k, var == 0 ? TT : SS) - 1] = eps; any = true;
            }
            if (!any) continue;
            for (int q = 0; q < ndim; q++) und[q] = un[q] + d[q];
            vmix_fun(und.data(), mixd.data());
            for (int q = 0; q < ndim; q++) mixd[q] = mixd[q] - mix[q];
            // fdjs: fjac(row, col) = mixd(row) / d(col) for the columns of this group; an(s,ie,je,ix,iy,iz) += fjac
            for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
                if (lm(i, j, k) != OCEAN) continue;            // rows of OCEAN cells only
                for (int h = 1; h <= 27; h++) {
                    int di = (h - 1) / 9 - 1, dj = ((h - 1) / 3) % 3 - 1, dk = (h - 1) % 3 - 1;   // any enumeration of the 27 offsets
                    int i2 = i + di, j2 = j + dj, k2 = k + dk;
                    int kind = lm(i2, j2, k2);
                    if (kind == PERIO) i2 = i - di * (n - 1);
                    else if (kind != OCEAN) continue;
                    if (j2 < 1 || j2 > m || k2 < 1 || k2 > l || i2 < 1 || i2 > n) continue;
                    if ((k2 - 1) % 3 != gk || (j2 - 1) % 3 != gj || ring(i2, n, periodic) != gi) continue;
                    int loc = 1 + (dj + 1) + 3 * (di + 1) + 9 * (dk == 0 ? 0 : (dk == -1 ? 1 : 2));
                    int je = var == 0 ? TT : SS;
                    if (vmix_temp) An[aidx(loc, TT, je, i, j, k)] += mixd[find_row2(i, j, k, TT) - 1] / eps;
                    if (vmix_salt) An[aidx(loc, SS, je, i, j, k)] += mixd[find_row2(i, j, k, SS) - 1] / eps;
                }
            }
        }
    }

    // ---------------- scaling.F90:29-64: average diagonal 6x6 block over the OCEAN cells, from the CRS Jacobian ----------------
    void average_block(double* db /* (nun,nun) column-major */) {
        for (int q = 0; q < NUN * NUN; q++) db[q] = 0.0;
        int nl = 0;
        for (int i = 1; i <= ndim; i++) {
            int ii = (i - 1) % NUN + 1, cell = (i - 1) / NUN;
            int ix = cell % n + 1, iy = (cell / n) % m + 1, iz = cell / (n * m) + 1;
            if (lm(ix, iy, iz) != OCEAN) continue;
            if (ii == PP) nl++;
            for (int j = begA[i - 1]; j <= begA[i] - 1; j++) {
                int c = jcoA[j - 1];
                if ((c - 1) / NUN == cell) { int jj = (c - 1) % NUN + 1; db[(ii - 1) + NUN * (jj - 1)] += coA[j - 1]; }
            }
        }
        if (nl > 0) for (int q = 0; q < NUN * NUN; q++) db[q] = db[q] / nl;
    }
    // scaling.F90:185-279 (the LAPACK variant that is compiled): dgetrf + dgetri = inverse by LU with partial pivoting -- here
    // plain Gauss-Jordan with partial pivoting (equal up to rounding) -- then the "special for oceanography" formulas
    static bool scal(double* mat /* (6,6) column-major, overwritten by its inverse */, double* dr, double* dc) {
        const int N6 = NUN;
        auto M = [&](int i, int j) -> double& { return mat[(i - 1) + N6 * (j - 1)]; };
        double Anorm = 0.0;
        for (int i = 1; i <= N6; i++) { double r = 0.0; for (int j = 1; j <= N6; j++) r += std::fabs(M(i, j)); Anorm = std::max(Anorm, r); }
        double inv[36];
        for (int q = 0; q < 36; q++) inv[q] = 0.0;
        for (int i = 0; i < N6; i++) inv[i + N6 * i] = 1.0;
        double a[36]; std::memcpy(a, mat, sizeof(a));
        auto A = [&](int i, int j) -> double& { return a[i + N6 * j]; };
        auto B = [&](int i, int j) -> double& { return inv[i + N6 * j]; };
        bool singular = false;
        for (int p = 0; p < N6 && !singular; p++) {
            int piv = p; double best = std::fabs(A(p, p));
            for (int r = p + 1; r < N6; r++) if (std::fabs(A(r, p)) > best) { best = std::fabs(A(r, p)); piv = r; }
            if (best == 0.0) { singular = true; break; }
            if (piv != p) for (int j = 0; j < N6; j++) { std::swap(A(p, j), A(piv, j)); std::swap(B(p, j), B(piv, j)); }
            double d = 1.0 / A(p, p);
            for (int j = 0; j < N6; j++) { A(p, j) *= d; B(p, j) *= d; }
            for (int r = 0; r < N6; r++) if (r != p) { double f = A(r, p); if (f != 0.0) for (int j = 0; j < N6; j++) { A(r, j) -= f * A(p, j); B(r, j) -= f * B(p, j); } }
        }
        double inorm = 0.0;
        for (int i = 0; i < N6; i++) { double r = 0.0; for (int j = 0; j < N6; j++) r += std::fabs(B(i, j)); inorm = std::max(inorm, r); }
        double rcond = singular ? 0.0 : 1.0 / (Anorm * inorm);
        if (1.0 + rcond == 1.0) return false;   // "diagonal block is singular up to working precision": dr, dc stay unset
        std::memcpy(mat, inv, sizeof(inv));
        dr[0] = 1.0; dc[0] = 1.0;
        double idc = std::sqrt(M(1, 1) / M(2, 2));
        dr[1] = 1 / idc; dc[1] = dr[1];
        double idr = std::sqrt(std::fabs(M(1, 1) / M(4, 4)));
        dr[3] = 1 / idr; dc[3] = dr[3];
        if (std::fabs(M(4, 3)) > std::fabs(M(3, 3))) idr = M(1, 1) / (idr * M(4, 3));
        else idr = std::sqrt(std::fabs(M(1, 1) / M(3, 3)));
        dr[2] = 2 / idr; dc[2] = dr[2];
        if (std::fabs(M(4, 5) * M(5, 4)) < .01 * std::fabs(M(4, 4) * M(5, 5))) { M(4, 5) = 1; M(5, 4) = 1; }
        idc = std::sqrt(std::fabs(M(1, 1) * M(4, 5) / (M(5, 4) * M(5, 5))));
        idr = M(1, 1) / (idc * M(5, 5));
        dr[4] = 1 / idr; dc[4] = 1 / idc;
        if (std::fabs(M(4, 6) * M(6, 4)) < .01 * std::fabs(M(4, 4) * M(6, 6))) { M(4, 6) = 1; M(6, 4) = 1; }
        idc = std::sqrt(std::fabs(M(1, 1) * M(4, 6) / (M(6, 4) * M(6, 6))));
        idr = M(1, 1) / (idc * M(6, 6));
        dr[5] = 1 / idr; dc[5] = 1 / idc;
        return true;
    }
    // scaling.F90:70-105
    bool compute_scaling(const double* db_in, double* row_scaling, double* col_scaling) {
        double db[36], rs[NUN], cs[NUN];
        std::memcpy(db, db_in, sizeof(db));
        for (int q = 0; q < NUN; q++) { rs[q] = 1.0; cs[q] = 1.0; }
        bool ok = scal(db, rs, cs);
        for (int i = 1; i <= ndim; i++) {
            int ii = (i - 1) % NUN + 1, cell = (i - 1) / NUN;
            int ix = cell % n + 1, iy = (cell / n) % m + 1, iz = cell / (n * m) + 1;
            bool oc = lm(ix, iy, iz) == OCEAN;
            row_scaling[i - 1] = oc ? rs[ii - 1] : 1.0;
            col_scaling[i - 1] = oc ? cs[ii - 1] : 1.0;
        }
        return ok;
    }
    // thcm_utils.F90:285-309
    int intcond_scaling(double* val, int* ind) {
        int v = 0;
        for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++)
            if (lm(i, j, k) == OCEAN) { val[v] = std::cos(y[j]) * dfzT[k]; ind[v] = find_row2(i, j, k, SS); v++; }
        return v;
    }

    // ---------------- usrc.F90:449-521 ----------------
    void matrix(const double* un) {
        An = Al;
        fillcolB();
        nlin_jac(un);
        if (vmix_flag >= 1) {
            if (vmix_fix == 0 && vmix_flag >= 2) vmix_control(un);
            if ((vmix_temp == 1 || vmix_salt == 1) && vmix_dim > 0) vmix_jac(un);
        }
        boundaries();
        fillcolA();
    }

    // ---------------- integrals.F90:17-88 ----------------
    void salt_advection(const double* un, double* check) {
        Fields F(n, m, l);
        usol(un, F.u, F.v, F.w, F.p, F.t, F.s);
        Arr3 &u = F.u, &v = F.v, &w = F.w, &s = F.s;
        size_t pos = 0;
        for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++, pos++) {
            if (lm(i, j, l) != OCEAN) continue;
            check[pos] = (u(i, j, k) + u(i, j - 1, k)) * (s(i + 1, j, k) + s(i, j, k)) / (4 * dx) -
                         (u(i - 1, j, k) + u(i - 1, j - 1, k)) * (s(i, j, k) + s(i - 1, j, k)) / (4 * dx) +
                         (v(i, j, k) + v(i - 1, j, k)) * (s(i, j + 1, k) + s(i, j, k)) * std::cos(yv[j]) / (4 * dy) -
                         (v(i, j - 1, k) + v(i - 1, j - 1, k)) * (s(i, j, k) + s(i, j - 1, k)) * std::cos(yv[j - 1]) / (4 * dy) +
                         w(i, j, k) * (s(i, j, k + 1) + s(i, j, k)) * std::cos(y[j]) / (2 * dz * dfzW[k]) -
                         w(i, j, k - 1) * (s(i, j, k) + s(i, j, k - 1)) * std::cos(y[j]) / (2 * dz * dfzW[k - 1]);
        }
    }
    void salt_diffusion(const double* un, double* check) {
        Fields F(n, m, l);
        usol(un, F.u, F.v, F.w, F.p, F.t, F.s);
        Arr3& s = F.s;
        size_t pos = 0;
        for (int k = 1; k <= l; k++) {
            double h1 = 1. / (dfzT[k] * dfzW[k]), h2 = 1. / (dfzT[k] * dfzW[k - 1]);
            for (int j = 1; j <= m; j++) {
                double cay = std::cos(y[j]), c1 = std::cos(yv[j]), c2 = std::cos(yv[j - 1]);
                for (int i = 1; i <= n; i++, pos++) {
                    if (lm(i, j, k) != OCEAN) continue;
                    check[pos] = std::cos(y[j]) * dfzT[k] *
                                 ((s(i + 1, j, k) + s(i - 1, j, k) - 2 * s(i, j, k)) / (dx * dx * cay * cay) +
                                  (c1 * s(i, j + 1, k) + c2 * s(i, j - 1, k) - (c1 + c2) * s(i, j, k)) / (dy * dy * cay) +
                                  (h1 * s(i, j, k + 1) + h2 * s(i, j, k - 1) - (h1 + h2) * s(i, j, k)) / (dz * dz));
                }
            }
        }
    }
    // ---------------- forcing.F90:235-280 ----------------
    void get_stochastic_forcing(int* begF, int* jcoF, double* coF) {
        double oldpar = par[SPER];
        par[SPER] = 0.0;
        forcing();
        int v = 1, prev_row = 0;
        for (int i = 0; i <= ndim; i++) begF[i] = 0;
        for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            if (coupled_S != 1) {
                int row = find_row2(i, j, l, SS);
                if (prev_row > row) { fprintf(stderr, "Forcing ordered in the wrong way\n"); abort(); }
                jcoF[v - 1] = j;
                coF[v - 1] = Frc[row - 1];
                v = v + 1;
                begF[row] = v;
                prev_row = row;
            }
        }
        v = 1;
        for (int i = 0; i <= ndim; i++) { if (begF[i] == 0) begF[i] = v; else v = begF[i]; }
        par[SPER] = oldpar;
        forcing();
    }

    // ---------------- usrc.F90:523-603 ----------------
    void rhs(const double* un, double* B) {
        std::vector<double> Au(ndim), mix(ndim, 0.0);
        An = Al;
        nlin_rhs(un);
        boundaries();
        fillcolA();
        matAvec(un, Au.data());
        if (vmix_flag >= 1) {
            if (vmix_fix == 0 && vmix_flag >= 2) vmix_control(un);
            if ((vmix_temp == 1 || vmix_salt == 1) && vmix_dim > 0) vmix_fun(un, mix.data());
        }
        const double ures = 0.0;
        for (int r = 0; r < ndim; r++) B[r] = -Au[r] - mix[r] + Frc[r] - p0 * (1 - par[RESC]) * ures;
        for (int i = 1; i <= n; i++) for (int j = 1; j <= m; j++) for (int k = 1; k <= l; k++) for (int k1 = 1; k1 <= NUN; k1++) {
            int row = find_row2(i, j, k, k1);
            B[row - 1] = B[row - 1] * (1 - lm(i, j, k));
        }
    }
};

// ---- THCM.C:2300-2598 maximal graph on one rank; rows sorted by ascending GID
//      (Epetra FillComplete order on a single rank).  0-based i,j,k here.
struct Graph {
    std::vector<int> rowptr, col;
};
inline int FIND_ROW2(int N, int M, int i, int j, int k, int XX) { return NUN * (k * N * M + N * j + i) + XX - 1; }  // THCMdefs.H:21

void insert_graph_entry(std::vector<int>& ind, int i, int j, int k, int xx, int N, int M, int L, bool perio) {
    int ii = i;
    if (perio) ii = ((i % N) + N) % N;  // THCM.C:2589-2592
    if (ii >= 0 && j >= 0 && k >= 0 && ii < N && j < M && k < L) ind.push_back(FIND_ROW2(N, M, ii, j, k, xx));
}

Graph maximal_graph(int N, int M, int L, bool perio) {
    Graph g;
    g.rowptr.assign((size_t)NUN * N * M * L + 1, 0);
    std::vector<int> ind;
    auto ins = [&](int i, int j, int k, int xx) { insert_graph_entry(ind, i, j, k, xx, N, M, L, perio); };
    auto flush = [&](int row) {
        std::sort(ind.begin(), ind.end());
        ind.erase(std::unique(ind.begin(), ind.end()), ind.end());
        g.rowptr[row + 1] = (int)ind.size();
        g.col.insert(g.col.end(), ind.begin(), ind.end());
        ind.clear();
    };
    for (int k = 0; k < L; k++) for (int j = 0; j < M; j++) for (int i = 0; i < N; i++) {
        int gid0 = FIND_ROW2(N, M, i, j, k, UU) - 1;
        auto seven = [&](int xx) { ins(i, j, k, xx); ins(i - 1, j, k, xx); ins(i + 1, j, k, xx); ins(i, j - 1, k, xx);
                                   ins(i, j + 1, k, xx); ins(i, j, k - 1, xx); ins(i, j, k + 1, xx); };
        auto wavg = [&]() { ins(i, j, k, WW); ins(i + 1, j, k, WW); ins(i + 1, j + 1, k, WW); ins(i, j + 1, k, WW);
                            ins(i, j, k - 1, WW); ins(i + 1, j, k - 1, WW); ins(i + 1, j + 1, k - 1, WW); ins(i, j + 1, k - 1, WW); };
        auto pgrad = [&]() { ins(i, j, k, PP); ins(i + 1, j, k, PP); ins(i, j + 1, k, PP); ins(i + 1, j + 1, k, PP); };
        // U (THCM.C:2357-2388)
        seven(UU); ins(i, j, k, VV); ins(i - 1, j, k, VV); ins(i + 1, j, k, VV); ins(i, j - 1, k, VV); ins(i, j + 1, k, VV);
        wavg(); pgrad(); flush(gid0 + UU);
        // V (THCM.C:2409-2438)
        seven(VV); ins(i, j, k, UU); ins(i - 1, j, k, UU); ins(i + 1, j, k, UU); wavg(); pgrad(); flush(gid0 + VV);
        // W (THCM.C:2446-2452)
        ins(i, j, k, WW); ins(i, j, k, PP); ins(i, j, k + 1, PP); ins(i, j, k, TT); ins(i, j, k + 1, TT); ins(i, j, k, SS); ins(i, j, k + 1, SS);
        flush(gid0 + WW);
        // P (THCM.C:2460-2473)
        ins(i, j, k, PP);
        ins(i, j, k, UU); ins(i - 1, j, k, UU); ins(i, j - 1, k, UU); ins(i - 1, j - 1, k, UU);
        ins(i, j, k, VV); ins(i - 1, j, k, VV); ins(i, j - 1, k, VV); ins(i - 1, j - 1, k, VV);
        ins(i, j, k, WW); ins(i, j, k - 1, WW);
        flush(gid0 + PP);
        // T, S (THCM.C:2481-2541)
        for (int pass = 0; pass < 2; pass++) {
            int R = pass == 0 ? TT : SS, O = pass == 0 ? SS : TT;
            seven(R);
            ins(i, j, k, UU); ins(i - 1, j, k, UU); ins(i - 1, j - 1, k, UU); ins(i, j - 1, k, UU);
            ins(i, j, k, VV); ins(i - 1, j, k, VV); ins(i - 1, j - 1, k, VV); ins(i, j - 1, k, VV);
            ins(i, j, k, WW); ins(i, j, k - 1, WW);
            ins(i, j, k, O); ins(i, j, k - 1, O); ins(i, j, k + 1, O);
            flush(gid0 + R);
        }
    }
    for (size_t r = 0; r + 1 < g.rowptr.size(); r++) g.rowptr[r + 1] += g.rowptr[r];
    return g;
}

}  // namespace

// =============================================================================
// C ABI for ctypes (tests / bench only)
// =============================================================================
extern "C" {

void oracle_set_tanh(int impl) { g_tanh_impl = impl; }
double oracle_tanh_value(double x) { return oracle_tanh(x); }

struct oracle_settings {
    double hdim, qz, alphaT, alphaS, ymin_glob, ymax_glob;
    int periodic, ih, vmix, tap, rho_mixing, coriolis_on, TRES, SRES, iza, ite, its, coupled_T, coupled_S, forcing_type;
};

void* oracle_create(int n, int m, int l, double xmin, double xmax, double ymin, double ymax,
                    const oracle_settings* s, const int* landm) {
    Oracle* o = new Oracle();
    o->hdim = s->hdim; o->qz = s->qz; o->alphaT = s->alphaT; o->alphaS = s->alphaS;
    o->ymin_glob = s->ymin_glob; o->ymax_glob = s->ymax_glob;
    o->periodic = s->periodic != 0; o->ih = s->ih; o->vmix = s->vmix; o->tap = s->tap; o->rho_mixing = s->rho_mixing;
    o->coriolis_on = s->coriolis_on; o->TRES = s->TRES; o->SRES = s->SRES; o->iza = s->iza; o->ite = s->ite; o->its = s->its;
    o->coupled_T = s->coupled_T; o->coupled_S = s->coupled_S; o->forcing_type = s->forcing_type;
    if (o->vmix < 0 || o->vmix > 2) { fprintf(stderr, "thcm_oracle: Mixing must be 0, 1 or 2\n"); delete o; return nullptr; }
    // init() needs n,m,l before set_landmask_raw touches lm()
    o->n = n; o->m = m; o->l = l;
    o->init(n, m, l, xmin, xmax, ymin, ymax, landm);
    return o;
}
void oracle_destroy(void* h) { delete (Oracle*)h; }
void oracle_average_block(void* h, double* db) { ((Oracle*)h)->average_block(db); }
int oracle_compute_scaling(void* h, const double* db, double* rs, double* cs) { return ((Oracle*)h)->compute_scaling(db, rs, cs) ? 1 : 0; }
int oracle_intcond_scaling(void* h, double* val, int* ind) { return ((Oracle*)h)->intcond_scaling(val, ind); }
void oracle_set_vmix_fix(void* h, int fix) { ((Oracle*)h)->vmix_fix = fix; }   // mix.F90:52-59
void oracle_vmix_fun(void* h, const double* un, double* mix) { Oracle* o = (Oracle*)h; for (int q = 0; q < o->ndim; q++) mix[q] = 0.0; o->vmix_fun(un, mix); }
void oracle_vmix_flags(void* h, int* out) { Oracle* o = (Oracle*)h; out[0] = o->vmix_flag; out[1] = o->vmix_temp; out[2] = o->vmix_salt; out[3] = o->vmix_fix; }
int oracle_ndim(void* h) { return ((Oracle*)h)->ndim; }
// usrc.F90:163-198
void oracle_setpar(void* h, int idx, double val) {
    Oracle* o = (Oracle*)h;
    if (idx >= 1 && idx <= NPAR) o->par[idx] = val;
    o->forcing();
    o->lin();
}
double oracle_getpar(void* h, int idx) { return ((Oracle*)h)->par[idx]; }
void oracle_rhs(void* h, const double* un, double* B) { ((Oracle*)h)->rhs(un, B); }
void oracle_matrix(void* h, const double* un) { ((Oracle*)h)->matrix(un); }
void oracle_fillcolb(void* h) { ((Oracle*)h)->fillcolB(); }
int oracle_nnz(void* h) { Oracle* o = (Oracle*)h; return o->begA[o->ndim] - 1; }
long oracle_bad_columns(void* h) { return ((Oracle*)h)->bad_columns; }
void oracle_get_crs(void* h, int* beg, int* jco, double* co) {
    Oracle* o = (Oracle*)h;
    int nnz = o->begA[o->ndim] - 1;
    std::memcpy(beg, o->begA.data(), sizeof(int) * (o->ndim + 1));
    std::memcpy(jco, o->jcoA.data(), sizeof(int) * nnz);
    std::memcpy(co, o->coA.data(), sizeof(double) * nnz);
}
void oracle_get_cob(void* h, double* cob) { Oracle* o = (Oracle*)h; std::memcpy(cob, o->coB.data(), sizeof(double) * o->ndim); }
void oracle_get_forcing(void* h, double* frc) { Oracle* o = (Oracle*)h; std::memcpy(frc, o->Frc.data(), sizeof(double) * o->ndim); }
void oracle_get_landm(void* h, int* out) { Oracle* o = (Oracle*)h; std::memcpy(out, o->landm.data(), sizeof(int) * o->landm.size()); }
void oracle_get_grid(void* h, double* x, double* y, double* z, double* xu, double* yv, double* zw, double* dfzT, double* dfzW) {
    Oracle* o = (Oracle*)h;
    std::memcpy(x, o->x.data(), sizeof(double) * (o->n + 1)); std::memcpy(xu, o->xu.data(), sizeof(double) * (o->n + 1));
    std::memcpy(y, o->y.data(), sizeof(double) * (o->m + 2)); std::memcpy(yv, o->yv.data(), sizeof(double) * (o->m + 1));
    std::memcpy(z, o->z.data(), sizeof(double) * (o->l + 1)); std::memcpy(zw, o->zw.data(), sizeof(double) * (o->l + 1));
    std::memcpy(dfzT, o->dfzT.data(), sizeof(double) * (o->l + 1)); std::memcpy(dfzW, o->dfzW.data(), sizeof(double) * (o->l + 1));
}
// surface fields (inserts.F90:11-281): which = 0 taux, 1 tauy, 2 tatm (atmosphere_t), 3 emip, 4 spert (emip_pert),
// 5 adapted_emip, 6 qatm (atmosphere_q), 7 albe (atmosphere_a), 8 patm (atmosphere_p), 9 qsa (seaice_q), 10 msi (seaice_m),
// 11 gsi (seaice_g).  The E-P fields are masked by the surface land mask (inserts.F90:179,198,217); q / a / p / g are only
// taken when the matching coupling flag is on (inserts.F90:45,66,87,145).  No recompute until the next setpar.
void oracle_set_field(void* h, int which, const double* f) {
    Oracle* o = (Oracle*)h;
    std::vector<double>* dst[] = {&o->taux, &o->tauy, &o->tatm, &o->emip, &o->spert, &o->adapted_emip,
                                  &o->qatm, &o->albe, &o->patm, &o->qsa, &o->msi, &o->gsi};
    if (which == 6 && !(o->coupled_T == 1 || o->coupled_S == 1)) return;
    if (which == 7 && o->coupled_T != 1) return;
    if ((which == 8 || which == 11) && o->coupled_S != 1) return;
    const bool masked = which >= 3 && which <= 5;
    size_t pos = 0;
    for (int j = 1; j <= o->m; j++) for (int i = 1; i <= o->n; i++, pos++)
        o->f2(*dst[which], i, j) = masked ? f[pos] * (1 - o->lm(i, j, o->l)) : f[pos];
}
// usrc.F90:353-418: new land mask (land-inversion fix, dummy frame), optionally re-running vmix_init + forcing + lin
void oracle_set_landmask(void* h, const int* landm, int periodic, int reinit) {
    Oracle* o = (Oracle*)h;
    o->periodic = periodic != 0;
    o->set_landmask_raw(landm, true);
    if (reinit == 1) { o->vmix_init(); o->forcing(); o->lin(); }
}
// usrc.F90:434-446
void oracle_setsres(void* h, int sres) { Oracle* o = (Oracle*)h; o->SRES = sres; o->forcing(); o->lin(); }
void oracle_salt_advection(void* h, const double* un, double* check) { ((Oracle*)h)->salt_advection(un, check); }
void oracle_salt_diffusion(void* h, const double* un, double* check) { ((Oracle*)h)->salt_diffusion(un, check); }
void oracle_stochastic_forcing(void* h, int* begF, int* jcoF, double* coF) { ((Oracle*)h)->get_stochastic_forcing(begF, jcoF, coF); }
// usr.F90:267-300: n*m*l internal temperature / salinity fields (w-row forcing, forcing.F90:199-209)
void oracle_set_internal_forcing(void* h, const double* temp, const double* salt) {
    Oracle* o = (Oracle*)h;
    std::memcpy(o->internal_temp.data(), temp, sizeof(double) * o->internal_temp.size());
    std::memcpy(o->internal_salt.data(), salt, sizeof(double) * o->internal_salt.size());
}
// model constants the numpy restatement of probe.F90 (oracle/probe_oracle.py) needs: QTnd QSnd Ooa Os nus lvsc qdim eta dqso
// eo0 albe0 albed zeta a0 Lf Qvar Q0, then suno(1..m)
void oracle_get_coupling_state(void* h, double* out17, double* suno) {
    Oracle* o = (Oracle*)h;
    double v[17] = {o->QTnd, o->QSnd, o->Ooa, o->Os, o->nus, o->lvsc, o->qdim, o->eta, o->dqso, o->eo0, o->albe0, o->albed, o->zeta,
                    o->a0, o->Lf, o->Qvar, o->Q0};
    std::memcpy(out17, v, sizeof(v));
    for (int j = 1; j <= o->m; j++) suno[j - 1] = o->suno[j];
}
void oracle_get_field(void* h, int which, double* f) {
    Oracle* o = (Oracle*)h;
    std::vector<double>* src[] = {&o->taux, &o->tauy, &o->tatm, &o->emip, &o->spert, &o->adapted_emip,
                                  &o->qatm, &o->albe, &o->patm, &o->qsa, &o->msi, &o->gsi};
    std::memcpy(f, src[which]->data(), sizeof(double) * o->n * o->m);
}
// usrc.F90:254-310: pars = the 18 doubles of Atmosphere::CommPars (tdim qdim nuq eta dqso dqsi dqdt Eo0 Ei0 Cs t0o t0i a0 da
// tauf tauc comb albf); nus and lvsc are frozen at the COMB / SALT / TEMP values of the moment of the call
void oracle_set_atmos_parameters(void* h, const double* pars) {
    Oracle* o = (Oracle*)h;
    o->qdim = pars[1]; o->nuq = pars[2]; o->eta = pars[3]; o->dqso = pars[4]; o->eo0 = pars[7];
    o->albe0 = pars[12]; o->albed = pars[13];
    o->nus = o->par[COMB] * o->par[SALT] * o->eta * o->qdim * o->QSnd;
    o->lvsc = o->par[COMB] * o->par[TEMP] * rhodim * lv * o->QTnd;
    o->forcing();
    o->lin();
}
// usrc.F90:313-350: pars = the 7 doubles of SeaIce::CommPars (zeta a0 Lf s0 rhoo Qvar Q0)
void oracle_set_seaice_parameters(void* h, const double* pars) {
    Oracle* o = (Oracle*)h;
    o->zeta = pars[0]; o->a0 = pars[1]; o->Lf = pars[2]; o->Qvar = pars[5]; o->Q0 = pars[6];
    o->forcing();
    o->lin();
}

// maximal graph (THCM.C:2300-2580): call with col == nullptr to get nnz
int oracle_graph(int N, int M, int L, int periodic, int* rowptr, int* col) {
    Graph g = maximal_graph(N, M, L, periodic != 0);
    if (rowptr) std::memcpy(rowptr, g.rowptr.data(), sizeof(int) * g.rowptr.size());
    if (col) std::memcpy(col, g.col.data(), sizeof(int) * g.col.size());
    return (int)g.col.size();
}

// THCM.C:1082-1162 on one rank: zero the graph values, then ReplaceGlobalValues
// row by row from the 1-based Fortran CRS.  Returns the number of CRS entries
// that are NOT in the graph (Epetra would return ierr=2 and the reference aborts).
long oracle_scatter_to_graph(int nrow, const int* beg, const int* jco, const double* co,
                             const int* rowptr, const int* col, double* val) {
    long missing = 0;
    for (int q = 0; q < rowptr[nrow]; q++) val[q] = 0.0;
    for (int i = 0; i < nrow; i++) {
        for (int v = beg[i]; v < beg[i + 1]; v++) {
            int c = jco[v - 1] - 1;
            const int* lo = std::lower_bound(col + rowptr[i], col + rowptr[i + 1], c);
            if (lo == col + rowptr[i + 1] || *lo != c) { missing++; continue; }
            val[lo - col] = co[v - 1];
        }
    }
    return missing;
}

// Epetra_CrsMatrix::Apply equivalent (Ocean.C:1369-1374): row-wise FP64 dot in stored order, 0-based CSR
void oracle_spmv(int nrow, const int* rowptr, const int* col, const double* val, const double* x, double* y) {
    for (int i = 0; i < nrow; i++) {
        double s = 0.0;
        for (int q = rowptr[i]; q < rowptr[i + 1]; q++) s += val[q] * x[col[q]];
        y[i] = s;
    }
}
// matetc.F90:147-166 on caller-provided 1-based CRS
void oracle_matavec(int nrow, const int* beg, const int* jco, const double* co, const double* v1, double* v2) {
    for (int i = 0; i < nrow; i++) {
        double s = 0.0;
        for (int v = beg[i]; v < beg[i + 1]; v++) s = co[v - 1] * v1[jco[v - 1] - 1] + s;
        v2[i] = s;
    }
}

}  // extern "C"
