// Host-side setup of the B200 THCM path: domain decomposition, grid, parameters, forcing vector,
// 1-D coefficient tables, land-mask derived per-cell data, static maximal graph and halo plan.
// Everything here runs at init / parameter-change / land-mask-change frequency (never per Newton
// step); the per-step work is in thcm_assembly.cu / krylov.cu.  Transcendentals are evaluated here
// with glibc libm -- like the reference's Fortran -- so the kernels need none.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include "thcm_cell.cuh"

extern "C" void thcm_throw_error_(char* msg);
extern "C" void thcm_forcing_integral_(double* field, double* y, int* landm, double* out);

namespace thcm {

static std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
const std::string& last_error() { return g_last_error; }
// errors: no return codes at the Fortran boundary; failure = thcm_throw_error_ (THCM.C:2690) which aborts
void fatal(const std::string& msg) {
    g_last_error = msg;
    std::string m = "thcm_b200: " + msg;
    thcm_throw_error_(const_cast<char*>(m.c_str()));
    fprintf(stderr, "%s\n", m.c_str());
    abort();
}

// usr.F90:132-160
static constexpr double PI = 3.14159265358979323846;
static constexpr double omegadim = 7.292e-05, r0dim = 6.37e+06, udim = 0.1e+00, gdim = 9.8e+00, rhodim = 1.024e+03,
                        t0 = 15, deltat = 1.0, deltas = 1.0, s0 = 35.0, cp0 = 4.2e+03, alpt1 = 2.93, alpt2 = 8.3e-02, alpt3 = 6.6e-04,
                        ah = 2.5e+05, av = 1.0e-03, kappah = 1.0e+03, kappav = 1.0e-04;
static constexpr double zmin = -1.0, zmax = 0.0;
// atm.F90:5-19 (the values the ocean uses in coupled mode)
static constexpr double rhoa = 1.25, ce = 1.3e-03, ch = 0.94 * ce, cpa = 1000., uw = 8.5, c0 = 0.43, sun0 = 1360., lv = 2.5e+06;

// ---------------------------------------------------------------------------------------------
// TRIOS::Domain::Decomp2D (src/trios/TRIOS_Domain.C:201-315): factor nprocs = npN x npM minimising
// |m/npM - n/npN| (integer division, r_min starts at 100 -- quirk kept), row-major rank grid,
// remainders to the first ranks.  Ghost layers are NOT added here: kernels index globally and use a
// width-1 halo buffer instead of the reference's 2 overlapping layers (multi-GPU).
// ---------------------------------------------------------------------------------------------
bool decomp2d(int nprocs, int pid, int N, int M, int L, int periodic, Block& b, const Cuts* cuts) {
    int t1 = nprocs, t2 = 1, npM = t1, npN = t2;
    double r, r_min = 100;
    while (t1 > 0) {
        t2 = nprocs / t1;
        r = std::abs(M / t1 - N / t2);
        if (t1 * t2 == nprocs && r <= r_min) { r_min = r; npM = t1; npN = t2; }
        t1--;
    }
    b.N = N; b.M = M; b.L = L; b.npN = npN; b.npM = npM; b.rank = pid; b.nranks = nprocs; b.periodic = periodic;
    b.pidN = pid % npN;
    b.pidM = (pid - b.pidN) / npN;
    b.m0 = M / npM; b.n0 = N / npN;
    b.j0 = b.pidM * (M / npM); b.i0 = b.pidN * (N / npN);
    int remM = M % npM, remN = N % npN;
    if (b.pidM < remM) b.m0++;
    if (b.pidN < remN) b.n0++;
    b.j0 += std::min(remM, b.pidM);
    b.i0 += std::min(remN, b.pidN);
    b.cuts = nullptr;
    if (cuts && cuts->npM == npM && cuts->npN == npN) {   // ocean-weighted cut lines (same rank grid, same rank -> (pidN, pidM) map)
        b.cuts = cuts;
        b.j0 = cuts->jc[b.pidM]; b.m0 = cuts->jc[b.pidM + 1] - b.j0;
        const int* ic = cuts->ic.data() + (size_t)b.pidM * (npN + 1);
        b.i0 = ic[b.pidN]; b.n0 = ic[b.pidN + 1] - b.i0;
    }
    b.wrap_x = (periodic && npN == 1) ? 1 : 0;
    b.halo_w = (b.pidN > 0 || (periodic && npN > 1)) ? 1 : 0;
    b.halo_e = (b.pidN < npN - 1 || (periodic && npN > 1)) ? 1 : 0;
    b.halo_s = b.pidM > 0 ? 1 : 0;
    b.halo_n = b.pidM < npM - 1 ? 1 : 0;
    b.hk = (b.halo_s + b.halo_n) * (b.n0 + b.halo_w + b.halo_e) + (b.halo_w + b.halo_e) * b.m0;
    return b.n0 > 0 && b.m0 > 0;
}

// prefix-balanced cuts of a weight sequence into `parts` pieces of at least `minw` entries each
static std::vector<int> balanced_cuts(const std::vector<double>& w, int parts, int minw) {
    const int len = (int)w.size();
    std::vector<double> cum(len + 1, 0.0);
    for (int q = 0; q < len; q++) cum[q + 1] = cum[q] + w[q];
    std::vector<int> cut(parts + 1, 0);
    cut[parts] = len;
    for (int p = 1; p < parts; p++) {
        const double target = cum[len] * p / parts;
        const int lo = cut[p - 1] + minw, hi = len - (parts - p) * minw;
        int best = lo;
        for (int k = lo; k <= hi; k++) if (std::fabs(cum[k] - target) < std::fabs(cum[best] - target)) best = k;
        cut[p] = best;
    }
    return cut;
}
// Ocean-weighted cut lines on the reference's npN x npM rank grid: an OCEAN cell weighs 1 (assembly, SpMV and every Krylov vector
// operation), a LAND cell 0.05 (its tile is skipped by the assembly kernels, its rows are never streamed).  Deterministic in the global
// mask, so every rank computes the same cuts.
void compute_cuts(const int* landm, int N, int M, int L, int nprocs, Cuts& cuts) {
    Block tmp;
    decomp2d(nprocs, 0, N, M, L, 0, tmp);
    const int npN = tmp.npN, npM = tmp.npM;
    cuts = Cuts();
    if (N < 2 * npN || M < 2 * npM) return;   // too small to move a cut: uniform
    std::vector<double> w((size_t)N * M, 0.0);
    for (int k = 1; k <= L; k++) for (int j = 1; j <= M; j++) for (int i = 1; i <= N; i++)
        w[(size_t)(j - 1) * N + (i - 1)] += landm[(size_t)i + (size_t)(N + 2) * (j + (size_t)(M + 2) * k)] == OCEAN ? 1.0 : 0.05;
    std::vector<double> rowsum(M, 0.0);
    for (int j = 0; j < M; j++) for (int i = 0; i < N; i++) rowsum[j] += w[(size_t)j * N + i];
    cuts.npN = npN; cuts.npM = npM;
    cuts.jc = balanced_cuts(rowsum, npM, 2);
    cuts.ic.assign((size_t)npM * (npN + 1), 0);
    for (int pm = 0; pm < npM; pm++) {
        std::vector<double> colsum(N, 0.0);
        for (int j = cuts.jc[pm]; j < cuts.jc[pm + 1]; j++) for (int i = 0; i < N; i++) colsum[i] += w[(size_t)j * N + i];
        const std::vector<int> ic = balanced_cuts(colsum, npN, 2);
        for (int pn = 0; pn <= npN; pn++) cuts.ic[(size_t)pm * (npN + 1) + pn] = ic[pn];
    }
}
bool setup_block(thcmb_ctx* c, const int* landm_global) {
    const thcmb_settings& s = c->s;
    c->cuts = Cuts();
    if (s.balance && s.nranks > 1) compute_cuts(landm_global, s.N, s.M, s.L, s.nranks, c->cuts);
    return decomp2d(s.nranks, s.rank, s.N, s.M, s.L, s.periodic, c->blk, c->cuts.npM ? &c->cuts : nullptr);
}

int halo_slot(const Block& b, int ie, int je, int k) {
    int wrow = b.n0 + b.halo_w + b.halo_e;
    if (je == -1) {
        if (!b.halo_s || ie < -b.halo_w || ie > b.n0 - 1 + b.halo_e) return -1;
        return k * b.hk + (ie + b.halo_w);
    }
    if (je == b.m0) {
        if (!b.halo_n || ie < -b.halo_w || ie > b.n0 - 1 + b.halo_e) return -1;
        return k * b.hk + b.halo_s * wrow + (ie + b.halo_w);
    }
    if (je < 0 || je >= b.m0) return -1;
    int off = (b.halo_s + b.halo_n) * wrow;
    if (ie == -1) { if (!b.halo_w) return -1; return k * b.hk + off + je; }
    if (ie == b.n0) { if (!b.halo_e) return -1; return k * b.hk + off + b.halo_w * b.m0 + je; }
    return -1;
}

// grid.F90:95-130
static double fz(double zz, double q) {
    double th = std::tanh(q * (zz + 1)), tth = std::tanh(q);
    return q > 1.0 ? -1 + th / tth : zz + (1. - q) * zz * (1 - zz);
}
static double dfdz(double zz, double q) {
    double chh = std::cosh(q * (zz + 1)), tth = std::tanh(q);
    return q > 1.0 ? q / (tth * chh * chh) : 1.0 + (1. - q) * (1. - 2. * zz);
}

// grid.F90:2-66 on the GLOBAL domain (decomposition-invariant tables; compare with the 1-rank reference)
void build_grid(thcmb_ctx* c) {
    const thcmb_settings& s = c->s;
    int n = s.N, m = s.M, l = s.L;
    c->dx = (s.xmax - s.xmin) / n; c->dy = (s.ymax - s.ymin) / m; c->dz = (zmax - zmin) / l;
    c->x.assign(n + 1, 0.0); c->xu.assign(n + 1, 0.0); c->y.assign(m + 2, 0.0); c->yv.assign(m + 1, 0.0);
    c->z.assign(l + 1, 0.0); c->zw.assign(l + 1, 0.0); c->ze.assign(l + 1, 0.0); c->zwe.assign(l + 1, 0.0);
    c->dfzT.assign(l + 1, 0.0); c->dfzW.assign(l + 1, 0.0);
    for (int i = 1; i <= n; i++) { c->x[i] = ((double)i - 0.5) * c->dx + s.xmin; c->xu[i] = ((double)i) * c->dx + s.xmin; }
    c->xu[0] = s.xmin;
    for (int j = 1; j <= m; j++) { c->y[j] = ((double)j - 0.5) * c->dy + s.ymin; c->yv[j] = ((double)j) * c->dy + s.ymin; }
    c->y[0] = c->y[1] - c->dy; c->y[m + 1] = c->y[m] + c->dy; c->yv[0] = s.ymin;
    for (int k = 1; k <= l; k++) {
        c->ze[k] = ((double)k - 0.5) * c->dz + zmin; c->zwe[k] = ((double)k) * c->dz + zmin;
        c->z[k] = fz(c->ze[k], s.qz); c->zw[k] = fz(c->zwe[k], s.qz);
        c->dfzT[k] = dfdz(c->ze[k], s.qz); c->dfzW[k] = dfdz(c->zwe[k], s.qz);
    }
    c->zw[0] = zmin; c->dfzW[0] = dfdz(zmin, s.qz);
    double dzne = c->dz * c->dfzT[l];  // usrc.F90:125-127
    c->QTnd = r0dim / (udim * cp0 * rhodim * s.hdim * dzne);
    c->QSnd = s0 * r0dim / (deltas * udim * s.hdim * dzne);
}

// usrc.F90:1153-1197 + vmix_par (mix_imp.f:122-137)
void stpnt(thcmb_ctx* c) {
    const thcmb_settings& s = c->s;
    double* par = c->par;
    for (int i = 0; i <= NPAR; i++) par[i] = 0.0;
    par[AL_T] = 0.1 / (2 * omegadim * rhodim * s.hdim * udim * c->dz * c->dfzT[s.L]);
    par[RAYL] = s.alphaT * gdim * s.hdim / (2 * omegadim * udim * r0dim);
    par[EK_V] = av / (2 * omegadim * s.hdim * s.hdim);
    par[EK_H] = ah / (2 * omegadim * r0dim * r0dim);
    par[ROSB] = udim / (2 * omegadim * r0dim);
    par[PE_H] = kappah / (udim * r0dim);
    par[PE_V] = kappav * r0dim / (udim * s.hdim * s.hdim);
    par[P_VC] = 2.5e+04 * par[PE_V];
    par[LAMB] = s.alphaS / s.alphaT;
    par[BIOT] = r0dim / (75. * 3600. * 24. * udim);
    par[ALPC] = 1.0; par[ENER] = 1.0e+02; par[SPL1] = 2.0e+03; par[SPL2] = 0.01;
    if (s.vmix == 0) { par[MIXP] = 0.0; par[P_VC] = 0.0; par[ALPC] = 1.0; par[ENER] = 1.0e+2; par[MKAP] = 0.0; }
}

static inline int& LM(thcmb_ctx* c, int i, int j, int k) {
    return c->landm[(size_t)i + (size_t)(c->s.N + 2) * (j + (size_t)(c->s.M + 2) * k)];
}

// usrc.F90:79-107 (init) and :375-408 (set_landmask) applied to the GLOBAL mask
void apply_landmask_rules(thcmb_ctx* c, const int* in, bool fix_inversion) {
    int n = c->s.N, m = c->s.M, l = c->s.L;
    bool periodic = c->s.periodic != 0;
    c->landm.assign((size_t)(n + 2) * (m + 2) * (l + 2), OCEAN);
    size_t pos = 0;
    for (int k = 0; k <= l + 1; k++) for (int j = 0; j <= m + 1; j++) for (int i = 0; i <= n + 1; i++) {
        int v = in ? in[pos] : OCEAN;
        if (!periodic && v == PERIO) v = OCEAN;
        LM(c, i, j, k) = v;
        pos++;
    }
    if (fix_inversion)
        for (int i = 1; i <= n; i++) for (int j = 1; j <= m; j++) for (int k = l; k >= 2; k--)
            if (LM(c, i, j, k) == LAND && LM(c, i, j, k - 1) == OCEAN) LM(c, i, j, k - 1) = LAND;
    for (int k = 0; k <= l + 1; k++) for (int j = 0; j <= m + 1; j++)
        if (!periodic) { LM(c, 0, j, k) = LAND; LM(c, n + 1, j, k) = LAND; }
    for (int k = 0; k <= l + 1; k++) for (int i = 0; i <= n + 1; i++) { LM(c, i, 0, k) = LAND; LM(c, i, m + 1, k) = LAND; }
    for (int j = 0; j <= m + 1; j++) for (int i = 0; i <= n + 1; i++) { LM(c, i, j, 0) = LAND; LM(c, i, j, l + 1) = LAND; }
}

// forcing.F90:405-449
static double wfun(double yy, int v1) {
    if (v1 == 1)
        return 0.2 - 0.8 * std::sin(6 * std::fabs(yy)) - 0.5 * (1 - std::tanh(10 * std::fabs(yy))) -
               0.5 * (1 - std::tanh(10 * (PI / 2 - std::fabs(yy))));
    return 0.0;
}
// temfun / salfun use the GLOBAL latitude bounds of m_global, also on a sub-domain (forcing.F90:418-449)
static inline double glob_ymin(const thcmb_ctx* c) { return (c->s.ymin_glob == 0.0 && c->s.ymax_glob == 0.0) ? c->s.ymin : c->s.ymin_glob; }
static inline double glob_ymax(const thcmb_ctx* c) { return (c->s.ymin_glob == 0.0 && c->s.ymax_glob == 0.0) ? c->s.ymax : c->s.ymax_glob; }
static double temfun(const thcmb_ctx* c, double yy) {
    const thcmb_settings& s = c->s;
    const double ymin = glob_ymin(c), ymax = glob_ymax(c);
    if (s.forcing_type == 2) return std::cos(PI * (yy - ymin) / (ymax - ymin));
    return std::cos(PI * yy / ymax) + c->par[CMPR] * std::sin(PI * yy / ymax);
}
static double salfun(const thcmb_ctx* c, double yy) {
    const thcmb_settings& s = c->s;
    const double ymin = glob_ymin(c), ymax = glob_ymax(c);
    if (s.forcing_type == 2) return std::cos(PI * (yy - ymin) / (ymax - ymin));
    if (s.forcing_type == 1) return (std::cos(PI * yy / ymax) + c->par[FPER] * yy / ymax) / std::cos(yy);
    return std::cos(PI * yy / ymax) + c->par[FPER] * yy / ymax;
}
// forcing.F90:452-464 -> THCM.C:2653-2686.  Every rank holds the global surface fields and mask, so the
// global integral is evaluated redundantly in the reference's 1-rank summation order (no collective).
static double qint(thcmb_ctx* c, const std::vector<double>& f) {
    int n = c->s.N, m = c->s.M, l = c->s.L;
    if (c->use_integral_callback) {   // created through init_: like the Fortran, let the C++ side form the (MPI-summed) integral
        double cor = 0.0;
        thcm_forcing_integral_(const_cast<double*>(f.data()), c->y.data() + 1, c->landm.data(), &cor);
        return cor;
    }
    double lf = 0.0, ls = 0.0;
    for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
        lf = f[(size_t)(i - 1) + (size_t)n * (j - 1)] * std::cos(c->y[j]) * (1 - LM(c, i, j, l)) + lf;
        ls = std::cos(c->y[j]) * (1 - LM(c, i, j, l)) + ls;
    }
    return lf / ls;
}

// Rows that `boundaries` turns into identity rows get Frc = 0 (boundary.F90:167,256,263,309,316,344,351,384).
static bool frc_row_zeroed(thcmb_ctx* c, int i, int j, int k, int XX) {
    if (LM(c, i, j, k) != OCEAN) return true;
    if (XX == WW) return LM(c, i, j, k + 1) == LAND;
    if (XX == UU || XX == VV) return LM(c, i, j + 1, k) == LAND || LM(c, i + 1, j, k) == LAND || LM(c, i + 1, j + 1, k) == LAND;
    return false;
}

// forcing.F90:4-218: Frc on the owned rows (idealised or inserted surface fields)
void compute_forcing(thcmb_ctx* c) {
    const thcmb_settings& s = c->s;
    const Block& b = c->blk;
    const double* par = c->par;
    int n = s.N, m = s.M, l = s.L;
    auto F2 = [&](std::vector<double>& f, int i, int j) -> double& { return f[(size_t)(i - 1) + (size_t)n * (j - 1)]; };
    if (s.iza == 2)
        for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) { F2(c->taux, i, j) = wfun(c->yv[j], 1); F2(c->tauy, i, j) = wfun(c->yv[j], 2); }
    double sigma = par[COMB] * par[WIND] * par[AL_T];
    double etabi = par[COMB] * par[TEMP] * (1 - s.TRES + s.TRES * par[BIOT]);
    double temcor = 0.0;
    if (s.ite == 1 && s.coupled_T == 0) {
        for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) F2(c->tatm, i, j) = temfun(c, c->y[j]);
        if (s.TRES == 0) temcor = qint(c, c->tatm);
    }
    double gamma;
    if (s.coupled_S == 1) gamma = par[COMB] * par[SALT];
    else gamma = par[COMB] * par[SALT] * (1 - s.SRES + s.SRES * par[BIOT]);
    double salcor = 0.0, adapted_salcor = 0.0, spertcor = 0.0;
    if (s.its == 1) {
        for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) F2(c->emip, i, j) = salfun(c, c->y[j]) * (1 - LM(c, i, j, l));
        if (s.SRES == 0 && s.coupled_S == 0) salcor = qint(c, c->emip);
    }
    if (s.SRES == 0 && s.coupled_S == 0) { adapted_salcor = qint(c, c->adapted_emip); spertcor = qint(c, c->spert); }
    const double pQSnd = par[COMB] * par[SALT] * c->QSnd;

    c->frc_raw.assign(b.ndim(), 0.0);
    auto row = [&](int gi, int gj, int k, int XX) {  // 1-based global cell -> local owned row (0-based)
        return (size_t)NUN * (((size_t)(k - 1) * b.m0 + (gj - 1 - b.j0)) * b.n0 + (gi - 1 - b.i0)) + XX - 1;
    };
    for (int gj = b.j0 + 1; gj <= b.j0 + b.m0; gj++) for (int gi = b.i0 + 1; gi <= b.i0 + b.n0; gi++) {
        if (gj <= m - 1) {
            c->frc_raw[row(gi, gj, l, UU)] = sigma * F2(c->taux, gi, gj);
            c->frc_raw[row(gi, gj, l, VV)] = sigma * F2(c->tauy, gi, gj);
        }
        if (s.coupled_T == 1) {   // externally coupled atmosphere / sea ice (forcing.F90:66-80)
            double QToa = par[COMB] * par[SUNP] * c->suno[gj] * (1 - c->atm_albe0 - c->atm_albed * F2(c->albe, gi, gj)) +
                          c->atm_Ooa * F2(c->tatm, gi, gj) + c->atm_lvsc * c->atm_eta * c->atm_qdim * F2(c->qatm, gi, gj) -
                          c->atm_lvsc * c->atm_eo0;
            double QTos = c->QTnd * c->ice_zeta * (c->ice_a0 * s0 - t0);
            c->frc_raw[row(gi, gj, l, TT)] = (QToa + F2(c->msi, gi, gj) * (QTos - QToa)) * (1 - LM(c, gi, gj, l));
        } else {
            c->frc_raw[row(gi, gj, l, TT)] = etabi * (F2(c->tatm, gi, gj) - temcor);
        }
        if (s.coupled_S == 1) {   // forcing.F90:152-164
            double QSoa = pQSnd * (c->atm_eo0 - c->atm_eta * c->atm_qdim * F2(c->qatm, gi, gj) - F2(c->patm, gi, gj));
            double QSos = pQSnd * (c->ice_zeta * (c->ice_a0 * s0 - t0) - c->ice_Qvar * F2(c->qsa, gi, gj) - c->ice_Q0) / (rhodim * c->ice_Lf);
            c->frc_raw[row(gi, gj, l, SS)] = (QSoa + F2(c->msi, gi, gj) * (QSos - QSoa) - F2(c->gsi, gi, gj)) * (1 - LM(c, gi, gj, l));
        } else {
            c->frc_raw[row(gi, gj, l, SS)] = gamma * (1 - par[HMTP]) * (F2(c->emip, gi, gj) - salcor) +
                                             gamma * par[HMTP] * (F2(c->adapted_emip, gi, gj) - adapted_salcor) +
                                             par[SPER] * (1 - s.SRES + s.SRES * par[BIOT]) * (F2(c->spert, gi, gj) - spertcor);
        }
    }
    // forcing.F90:199-209: the w-row forcing from internal_temp / internal_salt -- identically zero unless the caller
    // provided them (m_usr::set_internal_forcing; the Levitus files the reference would read do not ship with it)
    if (c->internal_set) {
        auto F3 = [&](const std::vector<double>& f, int i, int j, int k) { return f[(size_t)(i - 1) + (size_t)n * ((j - 1) + (size_t)m * (k - 1))]; };
        for (int k = 1; k <= l - 1; k++) for (int gj = b.j0 + 1; gj <= b.j0 + b.m0; gj++) for (int gi = b.i0 + 1; gi <= b.i0 + b.n0; gi++)
            c->frc_raw[row(gi, gj, k, WW)] = -par[COMB] * (1 - LM(c, gi, gj, k)) * par[RAYL] *
                                             (par[LAMB] * (F3(c->internal_salt, gi, gj, k) + F3(c->internal_salt, gi, gj, k + 1)) / 2. -
                                              (F3(c->internal_temp, gi, gj, k) + F3(c->internal_temp, gi, gj, k + 1)) / 2.);
    }
    c->frc_local = c->frc_raw;
    mask_forcing_rows(c);
    c->frc_masked = false;
}

// boundary.F90 writes Frc(row) = 0 for every row it turns into an identity row, inside each rhs / matrix call and on top of what
// earlier calls zeroed (the zeros are cumulative until `forcing` refills Frc): applied to frc_local as it stands
void mask_forcing_rows(thcmb_ctx* c) {
    const Block& b = c->blk;
    const int l = c->s.L;
    auto row = [&](int gi, int gj, int k, int XX) { return (size_t)NUN * (((size_t)(k - 1) * b.m0 + (gj - 1 - b.j0)) * b.n0 + (gi - 1 - b.i0)) + XX - 1; };
    for (int k = 1; k <= l; k++) for (int gj = b.j0 + 1; gj <= b.j0 + b.m0; gj++) for (int gi = b.i0 + 1; gi <= b.i0 + b.n0; gi++)
        for (int XX = 1; XX <= NUN; XX++)
            if (frc_row_zeroed(c, gi, gj, k, XX)) c->frc_local[row(gi, gj, k, XX)] = 0.0;
}

// usrc.F90:1200-1240: the members of m_atm the ocean reads (Ooa, Os, suno); nus and lvsc wait for set_atmos_parameters
void atmos_coef(thcmb_ctx* c) {
    const int m = c->s.M;
    const double muoa = rhoa * ch * cpa * uw;
    c->atm_Os = sun0 * c0 / 4 * c->QTnd;
    c->atm_Ooa = muoa * c->QTnd;
    c->atm_nus = 0.0;
    c->atm_lvsc = 0.0;
    c->suno.assign(m + 1, 0.0);
    for (int j = 1; j <= m; j++) {
        const double sj = std::sin(c->y[j]);
        c->suno[j] = c->atm_Os * (1 - .482 * (3 * (sj * sj) - 1.) / 2.);
    }
}

// m_inserts (inserts.F90:11-281): GLOBAL N*M surface field, i fastest.  The E-P fields are masked by the surface land
// mask (:179,198,217); q / a / p / g are only taken when the matching coupling flag is on (:45,66,87,145).  Like the
// reference, nothing is recomputed until the next setparcs / set_*_parameters call.
void insert_surface_field(thcmb_ctx* c, int which, const double* f) {
    std::vector<double>* dst[SF_COUNT] = {&c->taux, &c->tauy, &c->tatm, &c->emip, &c->spert, &c->adapted_emip,
                                          &c->qatm, &c->albe, &c->patm, &c->qsa, &c->msi, &c->gsi};
    if (which < 0 || which >= SF_COUNT) fatal("insert_surface_field: unknown field");
    const thcmb_settings& s = c->s;
    if (which == SF_QATM && !(s.coupled_T == 1 || s.coupled_S == 1)) return;
    if (which == SF_ALBE && s.coupled_T != 1) return;
    if ((which == SF_PATM || which == SF_GSI) && s.coupled_S != 1) return;
    const bool masked = which == SF_EMIP || which == SF_SPERT || which == SF_ADAPTED_EMIP;
    size_t pos = 0;
    for (int j = 1; j <= s.M; j++) for (int i = 1; i <= s.N; i++, pos++)
        (*dst[which])[pos] = masked ? f[pos] * (1 - LM(c, i, j, s.L)) : f[pos];
}
// usrc.F90:254-310: the 18 doubles of Atmosphere::CommPars (tdim qdim nuq eta dqso dqsi dqdt Eo0 Ei0 Cs t0o t0i a0 da tauf
// tauc comb albf); nus and lvsc are frozen at the COMB / SALT / TEMP values of the moment of the call.  The caller re-runs
// forcing + lin (refresh_params).
void set_atmos_parameters(thcmb_ctx* c, const double* p) {
    c->atm_qdim = p[1]; c->atm_nuq = p[2]; c->atm_eta = p[3]; c->atm_dqso = p[4]; c->atm_eo0 = p[7];
    c->atm_albe0 = p[12]; c->atm_albed = p[13];
    c->atm_nus = c->par[COMB] * c->par[SALT] * c->atm_eta * c->atm_qdim * c->QSnd;
    c->atm_lvsc = c->par[COMB] * c->par[TEMP] * rhodim * lv * c->QTnd;
}
// usrc.F90:313-350: the 7 doubles of SeaIce::CommPars (zeta a0 Lf s0 rhoo Qvar Q0)
void set_seaice_parameters(thcmb_ctx* c, const double* p) {
    c->ice_zeta = p[0]; c->ice_a0 = p[1]; c->ice_Lf = p[2]; c->ice_Qvar = p[5]; c->ice_Q0 = p[6];
    if (p[3] != s0) fprintf(stderr, "thcm_b200: WARNING conflicting reference salinity s0\n");
    if (p[4] != rhodim) fprintf(stderr, "thcm_b200: WARNING conflicting sea water density rhodim\n");
}
// m_usr::set_internal_forcing (usr.F90:267-300): N*M*L fields, i fastest; takes effect at the next forcing
void set_internal_forcing(thcmb_ctx* c, const double* temp, const double* salt) {
    const size_t tot = (size_t)c->s.N * c->s.M * c->s.L;
    c->internal_temp.assign(temp, temp + tot);
    c->internal_salt.assign(salt, salt + tot);
    c->internal_set = true;
}
// allocation of the surface fields of m_usr (usr.F90 allocate_usr) + atmos_coef, between stpnt and forcing (usrc.F90:118-131)
void init_surface_fields(thcmb_ctx* c) {
    const size_t nm = (size_t)c->s.N * c->s.M;
    for (auto* f : {&c->taux, &c->tauy, &c->tatm, &c->emip, &c->spert, &c->adapted_emip, &c->qatm, &c->albe, &c->patm, &c->msi,
                    &c->gsi, &c->qsa})
        f->assign(nm, 0.0);
    atmos_coef(c);
}

// assemble.F90:18-54 on the owned rows
void compute_cob(thcmb_ctx* c) {
    const Block& b = c->blk;
    c->cob_local.assign(b.ndim(), 0.0);
    for (int k = 1; k <= b.L; k++) for (int gj = b.j0 + 1; gj <= b.j0 + b.m0; gj++) for (int gi = b.i0 + 1; gi <= b.i0 + b.n0; gi++) {
        size_t r0 = (size_t)NUN * (((size_t)(k - 1) * b.m0 + (gj - 1 - b.j0)) * b.n0 + (gi - 1 - b.i0));
        if (LM(c, gi, gj, k) == OCEAN) {
            if (LM(c, gi + 1, gj, k) != LAND) c->cob_local[r0 + UU - 1] = -c->par[ROSB];
            if (LM(c, gi, gj + 1, k) != LAND) c->cob_local[r0 + VV - 1] = -c->par[ROSB];
            c->cob_local[r0 + TT - 1] = -1.0;
            c->cob_local[r0 + SS - 1] = -1.0;
        }
    }
}

// spf.F90:792-854
static double amh(double yy, int ih) { return ih == 0 ? 1.0 : 1. + 10.0 * std::exp(-5 * yy * yy); }
static double bmh(double yy, int ih) { return ih == 0 ? 1.0 : 1.0 + 10.0 * std::exp(-5 * yy * yy); }
static double bmhy(double yy, int ih) { return ih == 0 ? 0.0 : -10. * 10.0 * yy * std::exp(-5 * yy * yy); }

// The state-independent part of the dependency blocks (`lin`, usrc.F90:605-789 with the atoms of
// spf.F90:13-359) reduced to 1-D tables: every atom depends on j or on k only (apart from the 0/1
// surface-mask factor), so Al(loc,A,B,i,j,k) = f(table_j[j], table_k[k]) evaluated in the kernel with
// the reference's operation order.  Each table entry is the exact sub-expression the reference forms.
void compute_tables(thcmb_ctx* c) {
    const thcmb_settings& s = c->s;
    const double* par = c->par;
    int m = s.M, l = s.L, ih = s.ih;
    double dx = c->dx, dy = c->dy, dz = c->dz;
    double EV = par[EK_V], EH = par[EK_H], ph = (1 - par[MIXP]) * par[PE_H], pv = par[PE_V];
    double lambda = par[LAMB], xes = par[NLES], bi = par[BIOT], Ra = par[RAYL];
    int js = m + 2, ks = l + 2;
    c->jt_host.assign((size_t)J_COUNT * js, 0.0);
    c->kt_host.assign((size_t)K_COUNT * ks, 0.0);
    auto JTB = [&](int t, int j) -> double& { return c->jt_host[(size_t)t * js + j]; };
    auto KTB = [&](int t, int k) -> double& { return c->kt_host[(size_t)t * ks + k]; };
    const std::vector<double>&y = c->y, &yv = c->yv, &dfzT = c->dfzT, &dfzW = c->dfzW;
    double rdy = 1.0 / dy, rdy2i = rdy * rdy;
    for (int j = 0; j <= m; j++) {
        double t = 1.0 / (std::cos(yv[j]) * dx), cosdx2i = t * t;       // uderiv(2)/vderiv(2)
        JTB(J_CORV, j) = std::sin(yv[j]) * s.coriolis_on;              // coriolis
        JTB(J_C2X, j) = 1.0 / (2 * std::cos(yv[j]) * dx);              // gradp(1), unlin(1,2), vnlin(1,2)
        JTB(J_C2Y, j) = 1.0 / (2 * std::cos(yv[j]) * dy);              // unlin(3,4), vnlin(3,4)
        JTB(J_COSYV, j) = std::cos(yv[j]);
        JTB(J_TANYV, j) = std::tan(yv[j]);
        if (j >= 1 && j <= m - 1) {
            // ---- u rows: uderiv 2,3,5,6 (spf.F90:31-71) ----
            double uxx2 = amh(yv[j], ih) * cosdx2i, uxx8 = amh(yv[j], ih) * cosdx2i, uxx5 = -(uxx2 + uxx8);
            double uyy4 = rdy2i * bmh(y[j], ih) * std::cos(y[j]) / std::cos(yv[j]);
            double uyy6 = rdy2i * bmh(y[j + 1], ih) * std::cos(y[j + 1]) / std::cos(yv[j]);
            double uyy5 = -(uyy4 + uyy6);
            double tand2 = 1 - std::tan(yv[j]) * std::tan(yv[j]);
            double ucsi5 = bmh(yv[j], ih) * tand2 + std::tan(yv[j]) * bmhy(yv[j], ih);
            JTB(J_LU2, j) = -EH * (uxx2 + 0.0 + 0.0) - EV * 0.0;         // usrc.F90:690 at loc 2 (= loc 8)
            JTB(J_LU4, j) = -EH * (0.0 + uyy4 + 0.0) - EV * 0.0;
            JTB(J_LU6, j) = -EH * (0.0 + uyy6 + 0.0) - EV * 0.0;
            JTB(J_LU5, j) = -EH * (uxx5 + uyy5 + ucsi5);                  // minus EV*uzz5(k) in the kernel
            double tanv = std::tan(yv[j]), cosv = std::cos(yv[j]);
            double vxs2 = (bmhy(yv[j], ih) - (amh(yv[j], ih) + bmh(yv[j], ih)) * tanv) / (dx * cosv);
            double vxs8 = -(bmhy(yv[j], ih) - (amh(yv[j], ih) + bmh(yv[j], ih)) * tanv) / (dx * cosv);
            JTB(J_LUV2, j) = -0.0 - EH * vxs2;                            // usrc.F90:691 at loc 2
            JTB(J_LUV8, j) = -0.0 - EH * vxs8;
            // ---- v rows: vderiv 2,3,5,6 (spf.F90:94-133) ----
            double vxx2 = bmh(yv[j], ih) * cosdx2i, vxx5 = -2 * bmh(yv[j], ih) * cosdx2i;
            double vyy4 = rdy2i * amh(y[j], ih) * std::cos(y[j]) / std::cos(yv[j]);
            double vyy6 = rdy2i * amh(y[j + 1], ih) * std::cos(y[j + 1]) / std::cos(yv[j]);
            double vyy5 = -(vyy4 + vyy6);
            double vcsi5 = bmh(yv[j], ih) - amh(yv[j], ih) * std::tan(yv[j]) * std::tan(yv[j]) + bmhy(yv[j], ih) * std::tan(yv[j]);
            JTB(J_LV2, j) = -EH * (vxx2 + 0.0 + 0.0) - EV * 0.0;
            JTB(J_LV4, j) = -EH * (0.0 + vyy4 + 0.0) - EV * 0.0;
            JTB(J_LV6, j) = -EH * (0.0 + vyy6 + 0.0) - EV * 0.0;
            JTB(J_LV5, j) = -EH * (vxx5 + vyy5 + vcsi5);
            double uxs2 = -((amh(yv[j], ih) + bmh(yv[j], ih)) * tanv - bmhy(yv[j], ih)) / (dx * cosv);
            double uxs8 = ((amh(yv[j], ih) + bmh(yv[j], ih)) * tanv - bmhy(yv[j], ih)) / (dx * cosv);
            JTB(J_LVU2, j) = 0.0 - EH * uxs2;                             // usrc.F90:706 at loc 2
            JTB(J_LVU8, j) = 0.0 - EH * uxs8;
        }
    }
    for (int j = 0; j <= m + 1; j++) {
        JTB(J_CP, j) = 1.0 / (2 * std::cos(y[j]) * dx);                 // pderiv(1)
        double c2iy = 1. / (2 * std::cos(y[j]) * dy);                   // pderiv(2)
        if (j >= 1 && j <= m) { JTB(J_PVA, j) = std::cos(yv[j - 1]) * c2iy; JTB(J_PVB, j) = std::cos(yv[j]) * c2iy; }
        JTB(J_C4X, j) = 1.0 / (4 * std::cos(y[j]) * dx);                // tnlin(2,3)
        JTB(J_C4Y, j) = 1.0 / (4 * std::cos(y[j]) * dy);                // tnlin(4,5)
        if (j >= 1 && j <= m) {
            // tderiv 3,4 (spf.F90:210-231) at surface-mask factor 1, combined as in usrc.F90:758/785
            double t = 1.0 / (std::cos(y[j]) * dx), cosdx2i = t * t;
            double txx2 = cosdx2i, txx5 = -2 * cosdx2i;
            double tyy4 = rdy2i * std::cos(yv[j - 1]) / std::cos(y[j]);
            double tyy6 = rdy2i * std::cos(yv[j]) / std::cos(y[j]);
            double tyy5 = -(tyy4 + tyy6);
            JTB(J_TT2, j) = -ph * (txx2 + 0.0);
            JTB(J_TT4, j) = -ph * (0.0 + tyy4);
            JTB(J_TT6, j) = -ph * (0.0 + tyy6);
            JTB(J_TT5, j) = -ph * (txx5 + tyy5);
        }
    }
    double rdz = 1.0 / dz, rdz2i = rdz * rdz, dzi = 1.0 / dz;
    for (int k = 1; k <= l; k++) {
        double h1 = 1. / (dfzT[k] * dfzW[k]), h2 = 1. / (dfzT[k] * dfzW[k - 1]);
        double uzz14 = h2 * rdz2i, uzz23 = h1 * rdz2i, uzz5 = -(uzz14 + uzz23);  // uderiv(4) = vderiv(4)
        KTB(K_ZU5, k) = EV * uzz5; KTB(K_ZU14, k) = EV * uzz14; KTB(K_ZU23, k) = EV * uzz23;
        KTB(K_TDZI8, k) = 1.0 / (8 * dfzT[k] * dz);                     // unlin(5,6), vnlin(5,6)
        KTB(K_WP5, k) = -dzi / dfzW[k]; KTB(K_WP23, k) = dzi / dfzW[k]; // gradp(3)
        KTB(K_PW5, k) = dzi / dfzT[k]; KTB(K_PW14, k) = -dzi / dfzT[k]; // pderiv(3)
        double tzz14 = h2 * rdz2i, tzz23 = (k < l) ? h1 * rdz2i : 0.0, tzz5 = -(tzz14 + tzz23);  // tderiv(5)
        KTB(K_ZT5, k) = pv * tzz5; KTB(K_ZT14, k) = pv * tzz14; KTB(K_ZT23, k) = pv * tzz23;
        KTB(K_DFZT, k) = dfzT[k];
        // restoring term, absent from the coupled branches (usrc.F90:745-758, 771-785)
        KTB(K_RT, k) = s.coupled_T == 1 ? 0.0 : s.TRES * bi * (k == l ? 1.0 : 0.0);   // usrc.F90:758, tderiv(1)
        KTB(K_RS, k) = s.coupled_S == 1 ? 0.0 : s.SRES * bi * (k == l ? 1.0 : 0.0);   // usrc.F90:785, tderiv(2)
        KTB(K_DFZW, k) = dfzW[k]; KTB(K_DFZWM, k) = dfzW[k - 1];        // dCdzt, mix_imp.f:613-641
    }
    // per-j / per-k records of the pipelined kernel: the three j-neighbour values of every j-table, all k-tables of level k
    c->jrec_host.assign((size_t)(m + 2) * J_COUNT * JREC, 0.0);
    for (int j = 1; j <= m; j++) for (int tb = 0; tb < J_COUNT; tb++) for (int d = 0; d < JREC; d++)
        c->jrec_host[((size_t)j * J_COUNT + tb) * JREC + d] = JTB(tb, j + d - 1);
    c->krec_host.assign((size_t)(l + 2) * K_COUNT, 0.0);
    for (int k = 0; k <= l + 1; k++) for (int tb = 0; tb < K_COUNT; tb++) c->krec_host[(size_t)k * K_COUNT + tb] = KTB(tb, k);
    DevTables& t = c->tab;
    t.jstride = js; t.kstride = ks;
    t.epsr = par[ROSB];
    t.dyi = 1. / (2 * dy);                                              // gradp(2)
    t.cWT = -Ra * (1. + xes * alpt1);                                   // usrc.F90:717
    t.cWS = lambda * Ra;                                                // usrc.F90:718
    t.c2 = Ra * xes * alpt2; t.c3 = Ra * xes * alpt3;                   // usrc.F90:862-863, 975-976
    t.tdzi2 = 1.0 / (2 * dz);                                           // tnlin(6,7)
    // coupled mode (usrc.F90:742-783): constants of the surface-level sensible / latent / sea-ice terms, each the
    // sub-expression the reference forms; msi restricted to the owned columns for the kernels
    t.coupled_T = s.coupled_T == 1; t.coupled_S = s.coupled_S == 1;
    t.cpl_ooa = c->atm_Ooa;
    t.cpl_dedt_t = c->atm_lvsc * c->atm_eta * c->atm_qdim * (deltat / c->atm_qdim) * c->atm_dqso;   // usrc.F90:743
    t.cpl_qtz = c->QTnd * c->ice_zeta;
    t.cpl_ts = -c->QTnd * c->ice_zeta * c->ice_a0;                                                   // usrc.F90:755
    t.cpl_pq = par[COMB] * par[SALT] * c->QSnd;                                                       // usrc.F90:767
    t.cpl_zeta = c->ice_zeta; t.cpl_a0 = c->ice_a0; t.cpl_rl = rhodim * c->ice_Lf;
    t.cpl_dedt_s = c->atm_nus * (deltat / c->atm_qdim) * c->atm_dqso;                               // usrc.F90:766
    t.msi = nullptr;
    c->msi_local.assign((size_t)c->blk.n0 * c->blk.m0, 0.0);
    if (t.coupled_T || t.coupled_S)
        for (int lj = 0; lj < c->blk.m0; lj++) for (int li = 0; li < c->blk.n0; li++)
            c->msi_local[(size_t)lj * c->blk.n0 + li] = c->msi[(size_t)(c->blk.i0 + li) + (size_t)s.N * (c->blk.j0 + lj)];
    // tracer mixing (vmix_fun, mix_imp.f:231-562): the vertical schemes (implicit mixing / convective adjustment, consistent vertical
    // mixing).  Neutral physics and GM put T,S entries on all 27 stencil positions; the reference's own C++ layer cannot take them
    // (they are outside the maximal graph of THCM.C:2320-2549, ReplaceGlobalValues fails with "value excluded", THCM.C:1095-1104).
    const bool mix_on = c->vmix_flag >= 1 && c->vmix_dim > 0;
    if (c->vmix_flag >= 1 && (par[MIXP] != 0.0 || par[MKAP] != 0.0))
        fatal("Mixing >= 1 with MIXP != 0 (neutral physics) or MKAP != 0 (Gent-McWilliams): their Jacobian entries fall outside the "
              "maximal matrix graph of THCM.C:2320-2549, so the reference's THCM::evaluate cannot assemble them either; not built");
    t.mix_lambda = lambda; t.mix_xes = xes; t.mix_kvc = par[P_VC]; t.mix_eps = (1.0 - par[ALPC]) * par[ENER] * par[PE_V]; t.mix_fac = s.alphaT * par[SPL1]; t.mix_dz = dz;
    t.mix_temp = mix_on ? c->vmix_temp : 0; t.mix_salt = mix_on ? c->vmix_salt : 0;
    t.mix_rho = (s.rho_mixing && xes == 0.0) ? 1 : 0;
}

// mix_imp.f:61-100 (vmix_init) with the pair count of vmix_part (mix_imp.f:171-228): mixing is applied only when the
// pattern is non-empty, i.e. when the block owns at least one OCEAN cell
void vmix_init(thcmb_ctx* c) {
    const int vm = c->s.vmix;
    if (vm < 0 || vm > 2) fatal("Mixing must be 0, 1 or 2");
    c->vmix_flag = vm; c->vmix_fix = vm == 2 ? 0 : 1;
    c->vmix_temp = c->vmix_salt = vm == 1 ? 1 : 0;
    bool ocean = false;
    for (int k = 1; k <= c->s.L && !ocean; k++) for (int j = 1; j <= c->s.M && !ocean; j++) for (int i = 1; i <= c->s.N; i++)
        if (LM(c, i, j, k) == OCEAN) { ocean = true; break; }
    c->vmix_dim = (vm == 1 && ocean) ? 1 : 0;
    c->vmix_has_ocean = ocean;
}
// vmix_control (mix_imp.f:139-169) once the field norms are known
void vmix_set_flags(thcmb_ctx* c, int temp, int salt) {
    if (c->vmix_temp != temp || c->vmix_salt != salt) {
        c->vmix_temp = temp; c->vmix_salt = salt;
        if (c->vmix_temp != 0) c->vmix_dim = c->vmix_has_ocean ? 1 : 0;   // vmix_part (the reference tests vmix_temp twice, :163)
    }
    c->vmix_fix = 1;
}

// ---------------------------------------------------------------------------------------------
// class tables: sorted position of every canonical slot inside its maximal-graph row
// (THCM.C:2354-2549 inserts, Epetra FillComplete sorts by column id)
// ---------------------------------------------------------------------------------------------
const ClassTables& class_tables(int periodic) {
    static ClassTables T[2];
    static bool done[2] = {false, false};
    int p = periodic ? 1 : 0;
    if (done[p]) return T[p];
    for (int cls = 0; cls < NCLASS; cls++) {
        // representative cell of this class on a virtual grid
        auto rep = [&](int first, int last, int& dim, int& idx) {
            if (first && last) { dim = 1; idx = 0; }
            else if (first) { dim = 4; idx = 0; }
            else if (last) { dim = 4; idx = 3; }
            else { dim = 4; idx = 1; }
        };
        int Nn, Mm, Ll, i, j, k;
        rep(cls & 1, cls & 2, Nn, i); rep(cls & 4, cls & 8, Mm, j); rep(cls & 16, cls & 32, Ll, k);
        for (int R = 1; R <= NUN; R++) {
            const SlotDef* sl = row_slots(R);
            int ns = ROW_NSLOT[R - 1];
            long long key[24]; int order[24], cnt = 0;
            for (int q = 0; q < ns; q++) {
                int i2 = i + loc_di(sl[q].loc), j2 = j + loc_dj(sl[q].loc), k2 = k + loc_dk(sl[q].loc);
                if (p && Nn >= 3) i2 = ((i2 % Nn) + Nn) % Nn;
                T[p].pos[cls][ROW_OFF[R - 1] + q] = -1;
                if (i2 < 0 || i2 >= Nn || j2 < 0 || j2 >= Mm || k2 < 0 || k2 >= Ll) continue;
                key[cnt] = (((long long)k2 * Mm + j2) * Nn + i2) * NUN + sl[q].col - 1;
                order[cnt++] = q;
            }
            for (int a = 0; a < cnt; a++) {
                int rank = 0;
                for (int bq = 0; bq < cnt; bq++) if (key[bq] < key[a]) rank++;
                T[p].pos[cls][ROW_OFF[R - 1] + order[a]] = (int8_t)rank;
            }
            T[p].rowlen[cls][R - 1] = (uint8_t)cnt;
        }
    }
    done[p] = true;
    return T[p];
}

static inline int cell_class(const Block& b, int gi, int gj, int k) {  // 0-based global
    return (gi == 0 ? 1 : 0) | (gi == b.N - 1 ? 2 : 0) | (gj == 0 ? 4 : 0) | (gj == b.M - 1 ? 8 : 0) | (k == 0 ? 16 : 0) |
           (k == b.L - 1 ? 32 : 0);
}

// owner rank of a global cell column (gi, gj) under decomp2d
static int owner_of(const Block& me, int gi, int gj) {
    auto find = [](int g, int tot, int np) {
        int base = tot / np, rem = tot % np;
        // first `rem` parts have base+1 cells
        int cut = rem * (base + 1);
        return g < cut ? g / (base + 1) : rem + (g - cut) / base;
    };
    if (me.cuts) {
        const Cuts& cu = *me.cuts;
        int pm = 0; while (pm + 1 < cu.npM && gj >= cu.jc[pm + 1]) pm++;
        const int* ic = cu.ic.data() + (size_t)pm * (cu.npN + 1);
        int pn = 0; while (pn + 1 < cu.npN && gi >= ic[pn + 1]) pn++;
        return pm * cu.npN + pn;
    }
    int pn = find(gi, me.N, me.npN), pm = find(gj, me.M, me.npM);
    return pm * me.npN + pn;
}

// ---------------------------------------------------------------------------------------------
// THCM row / column scaling (scaling.F90:70-279, the LAPACK variant that is compiled): inverse of the average diagonal
// block (dgetrf + dgetri there; Gauss-Jordan with partial pivoting here, equal up to rounding), then the "special for
// oceanography" formulas.  Setup frequency (once per preconditioner build), plain host code.
// ---------------------------------------------------------------------------------------------
static bool scaling_scal(double* mat /* (6,6) column-major: in the block, out its inverse */, double* dr, double* dc) {
    auto M = [&](int i, int j) -> double& { return mat[(i - 1) + NUN * (j - 1)]; };
    double Anorm = 0.0;
    for (int i = 1; i <= NUN; i++) { double r = 0.0; for (int j = 1; j <= NUN; j++) r += std::fabs(M(i, j)); Anorm = std::max(Anorm, r); }
    double a[36], inv[36];
    memcpy(a, mat, sizeof(a));
    for (int q = 0; q < 36; q++) inv[q] = 0.0;
    for (int i = 0; i < NUN; i++) inv[i + NUN * i] = 1.0;
    auto A = [&](int i, int j) -> double& { return a[i + NUN * j]; };
    auto B = [&](int i, int j) -> double& { return inv[i + NUN * j]; };
    bool singular = false;
    for (int p = 0; p < NUN && !singular; p++) {
        int piv = p; double best = std::fabs(A(p, p));
        for (int r = p + 1; r < NUN; r++) if (std::fabs(A(r, p)) > best) { best = std::fabs(A(r, p)); piv = r; }
        if (best == 0.0) { singular = true; break; }
        if (piv != p) for (int j = 0; j < NUN; j++) { std::swap(A(p, j), A(piv, j)); std::swap(B(p, j), B(piv, j)); }
        double d = 1.0 / A(p, p);
        for (int j = 0; j < NUN; j++) { A(p, j) *= d; B(p, j) *= d; }
        for (int r = 0; r < NUN; r++) if (r != p) { double f = A(r, p); if (f != 0.0) for (int j = 0; j < NUN; j++) { A(r, j) -= f * A(p, j); B(r, j) -= f * B(p, j); } }
    }
    double inorm = 0.0;
    for (int i = 0; i < NUN; i++) { double r = 0.0; for (int j = 0; j < NUN; j++) r += std::fabs(B(i, j)); inorm = std::max(inorm, r); }
    const double rcond = singular ? 0.0 : 1.0 / (Anorm * inorm);
    if (1.0 + rcond == 1.0) return false;    // "diagonal block is singular up to working precision" (scaling.F90:222-225)
    memcpy(mat, inv, sizeof(inv));
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
// m_scaling::compute (scaling.F90:70-105) for the owned rows of this block
bool scaling_compute(const thcmb_ctx* c, const double* db36, double* row_scaling, double* col_scaling) {
    double db[36], rs[NUN], cs[NUN];
    memcpy(db, db36, sizeof(db));
    for (int q = 0; q < NUN; q++) { rs[q] = 1.0; cs[q] = 1.0; }
    const bool ok = scaling_scal(db, rs, cs);
    const Block& b = c->blk;
    thcmb_ctx* cc = const_cast<thcmb_ctx*>(c);
    for (int k = 0; k < b.L; k++) for (int lj = 0; lj < b.m0; lj++) for (int li = 0; li < b.n0; li++) {
        const bool oc = LM(cc, b.i0 + li + 1, b.j0 + lj + 1, k + 1) == OCEAN;
        const size_t cell = ((size_t)k * b.m0 + lj) * b.n0 + li;
        for (int q = 0; q < NUN; q++) { row_scaling[cell * NUN + q] = oc ? rs[q] : 1.0; col_scaling[cell * NUN + q] = oc ? cs[q] : 1.0; }
    }
    return ok;
}
// m_thcm_utils::intcond_scaling (thcm_utils.F90:285-309): cos(y_j) dfzT_k on the S rows of the owned OCEAN cells, 1-based LOCAL rows
int intcond_scaling(const thcmb_ctx* c, double* val, int* ind) {
    const Block& b = c->blk;
    thcmb_ctx* cc = const_cast<thcmb_ctx*>(c);
    int v = 0;
    for (int k = 0; k < b.L; k++) for (int lj = 0; lj < b.m0; lj++) for (int li = 0; li < b.n0; li++) {
        if (LM(cc, b.i0 + li + 1, b.j0 + lj + 1, k + 1) != OCEAN) continue;
        val[v] = std::cos(c->y[b.j0 + lj + 1]) * c->dfzT[k + 1];
        ind[v] = NUN * (((k * b.m0) + lj) * b.n0 + li) + SS;
        v++;
    }
    return v;
}

DevBlock dev_block(const Block& b) {
    return DevBlock{b.N, b.M, b.L, b.i0, b.j0, b.n0, b.m0, b.periodic, b.wrap_x, b.halo_w, b.halo_e, b.halo_s, b.halo_n, b.hk, b.ncell()};
}

// Per-tile descriptors of the pipelined assembly kernel (needs the static graph: call after build_static_host)
void build_tile_descs(const thcmb_ctx* c, const std::vector<uint32_t>& nbmask, const std::vector<uint8_t>& surf,
                      const std::vector<uint8_t>& uvlive, std::vector<TileDesc>& out) {
    const DevBlock b = dev_block(c->blk);
    const int ntile = tiles_per_row(b) * b.m0 * b.L;
    out.assign(ntile, TileDesc{});
    for (int t = 0; t < ntile; t++) {
        const TileGeom g = tile_geom_of(b, t);
        TileDesc& d = out[t];
        for (int x = 0; x < g.ncell; x++) {
            d.nbmask[x] = nbmask[g.cell0 + x];
            if (surf[(size_t)g.lj * b.n0 + (g.cell0 + x) % b.n0]) d.surfbits |= 1u << x;
        }
        for (int r = 0; r < 9; r++) for (int x = 0; x < g.ncell + 2; x++) {
            unsigned fl = position_flags(b, uvlive.data(), g.gi0 - 1 + x, g.gj + r % 3 - 1, g.k + r / 3 - 1);
            if (fl & POS_UV) d.uvbits[r] |= 1ull << x;
            if (fl & POS_W) d.wbits[r] |= 1ull << x;
        }
        d.g0 = c->rowptr_host[(size_t)NUN * g.cell0];
        d.tot = c->rowptr_host[(size_t)NUN * (g.cell0 + g.ncell)] - d.g0;
        if (d.tot == g.ncell * NSLOT_TOTAL && (d.g0 & 1) == 0) d.flags |= 1u;
        bool open_ocean = true;   // no LAND among the 27 (+5) neighbours of any cell: every statement of `boundaries` is a no-op
        for (int x = 0; x < g.ncell; x++) open_ocean = open_ocean && d.nbmask[x] == 0u;
        if (open_ocean) d.flags |= 2u;
        bool all_land = true;      // identity rows only (boundary.F90:381-386): a state-independent pattern
        for (int x = 0; x < g.ncell; x++) all_land = all_land && ((d.nbmask[x] >> 4) & 1u);
        if (all_land) d.flags |= 4u;
    }
}

// Static per-cell data, graph and halo plan.  Host arrays are returned to the caller (thcm_api.cu uploads them).
void build_static_host(thcmb_ctx* c, std::vector<uint32_t>& nbmask, std::vector<uint8_t>& surf, std::vector<uint8_t>& uvlive,
                       std::vector<int>& send_idx, std::vector<int>& recv_slot) {
    const Block& b = c->blk;
    int N = b.N, M = b.M, L = b.L;
    if (b.periodic && N < 3) fatal("periodic grids need N >= 3");
    // ---- nbmask / surf ----
    nbmask.assign(b.ncell(), 0); surf.assign((size_t)b.n0 * b.m0, 0);
    for (int k = 1; k <= L; k++) for (int gj = b.j0 + 1; gj <= b.j0 + b.m0; gj++) for (int gi = b.i0 + 1; gi <= b.i0 + b.n0; gi++) {
        uint32_t mk = 0;
        for (int loc = 1; loc <= NP; loc++) {
            int v = LM(c, gi + loc_di(loc), gj + loc_dj(loc), k + loc_dk(loc));
            bool bit = (loc == 5) ? (v != OCEAN) : (v == LAND);
            if (bit) mk |= 1u << (loc - 1);
        }
        // boundary.F90:64-78 extra neighbours (only read where the reference reads them)
        if (gi < N) {
            if (LM(c, gi + 2, gj - 1, k) == LAND) mk |= 1u << 27;                 // southee
            if (LM(c, gi + 2, gj, k) == LAND) mk |= 1u << 28;                     // easteast
            if (LM(c, gi + 2, gj + 1, k) == LAND) mk |= 1u << 29;                 // northee
            if (gj < M && LM(c, gi + 2, gj + 2, k) == LAND) mk |= 1u << 30;       // nnorthee
        }
        if (gj < M && LM(c, gi, gj + 2, k) == LAND) mk |= 1u << 31;               // nnwest = nnorth = nneast (sic)
        nbmask[((size_t)(k - 1) * b.m0 + (gj - 1 - b.j0)) * b.n0 + (gi - 1 - b.i0)] = mk;
        if (k == L) surf[(size_t)(gj - 1 - b.j0) * b.n0 + (gi - 1 - b.i0)] = (uint8_t)(1 - LM(c, gi, gj, L));
    }
    // ---- uvlive: which u,v corner values survive usol (usrc.F90:1028-1119), GLOBAL rules ----
    // corner (ic, jc) in Fortran indexing 0..N x 0..M; box stored for ic in [i0, i0+n0+1], jc in [j0, j0+m0+1]
    int bn = b.n0 + 2, bm = b.m0 + 2;
    uvlive.assign((size_t)bn * bm * L, 0);
    for (int k = 1; k <= L; k++) for (int jc = b.j0; jc <= b.j0 + b.m0 + 1; jc++) for (int ic = b.i0; ic <= b.i0 + b.n0 + 1; ic++) {
        bool live;
        if (ic > N || jc > M) live = false;
        else if (jc == 0) live = false;                       // u(i,0,k) = 0 / never set
        else if (ic == 0) live = b.periodic != 0;             // u(0,j,k) = u(N,j,k) (raw copy), j = 1..M
        else {
            live = true;
            if (jc == M) live = false;                        // u(i,M,k) = 0 for i = 1..N
            if (!b.periodic && ic == N) live = false;         // u(N,j,k) = 0
        }
        if (live) {
            // zeroed by every LAND cell (landm == 1) among the four cells around the corner, cells in 1..N x 1..M
            for (int di = 0; di <= 1 && live; di++) for (int dj = 0; dj <= 1; dj++) {
                int ci = ic + di, cj = jc + dj;
                if (ci < 1 || ci > N || cj < 1 || cj > M) continue;
                if (LM(c, ci, cj, k) == 1) { live = false; break; }
            }
        }
        uvlive[((size_t)(k - 1) * bm + (jc - b.j0)) * bn + (ic - b.i0)] = live ? 1 : 0;
    }
    // ---- static graph with local column ids ----
    const ClassTables& ct = class_tables(b.periodic);
    int ndim = b.ndim();
    c->rowptr_host.assign(ndim + 1, 0);
    for (int k = 0; k < L; k++) for (int lj = 0; lj < b.m0; lj++) for (int li = 0; li < b.n0; li++) {
        int cls = cell_class(b, b.i0 + li, b.j0 + lj, k);
        size_t cell = ((size_t)k * b.m0 + lj) * b.n0 + li;
        for (int R = 0; R < NUN; R++) c->rowptr_host[cell * NUN + R + 1] = ct.rowlen[cls][R];
    }
    for (int r = 0; r < ndim; r++) c->rowptr_host[r + 1] += c->rowptr_host[r];
    c->gnnz = c->rowptr_host[ndim];
    c->col_host.assign(c->gnnz, -1);
    c->halo_gid.assign((size_t)NUN * b.nhalo_cells(), -1);
    c->local_gid.assign(ndim, 0);
    for (int k = 0; k < L; k++) for (int lj = 0; lj < b.m0; lj++) for (int li = 0; li < b.n0; li++) {
        int gi = b.i0 + li, gj = b.j0 + lj;
        int cls = cell_class(b, gi, gj, k);
        size_t cell = ((size_t)k * b.m0 + lj) * b.n0 + li;
        for (int R = 1; R <= NUN; R++) {
            c->local_gid[cell * NUN + R - 1] = NUN * ((k * M + gj) * N + gi) + R - 1;
            const SlotDef* sl = row_slots(R);
            int base = c->rowptr_host[cell * NUN + R - 1];
            for (int q = 0; q < ROW_NSLOT[R - 1]; q++) {
                int p = ct.pos[cls][ROW_OFF[R - 1] + q];
                if (p < 0) continue;
                int ie = li + loc_di(sl[q].loc), je = lj + loc_dj(sl[q].loc), k2 = k + loc_dk(sl[q].loc);
                if (b.wrap_x) ie = ((ie % b.n0) + b.n0) % b.n0;
                int id;
                if (ie >= 0 && ie < b.n0 && je >= 0 && je < b.m0) id = NUN * ((k2 * b.m0 + je) * b.n0 + ie) + sl[q].col - 1;
                else {
                    int hs = halo_slot(b, ie, je, k2);
                    if (hs < 0) fatal("internal: graph entry without halo slot");
                    id = ndim + NUN * hs + sl[q].col - 1;
                    int g2 = ((b.i0 + ie) % N + N) % N;
                    c->halo_gid[(size_t)NUN * hs + sl[q].col - 1] = NUN * ((k2 * M + (b.j0 + je)) * N + g2) + sl[q].col - 1;
                }
                c->col_host[base + p] = id;
            }
        }
    }
    // ---- halo plan: recv lists in my slot order, send lists in the peer's slot order ----
    c->peers.clear(); send_idx.clear(); recv_slot.clear(); c->send_dst_host.clear(); c->send_peer_host.clear();
    if (b.nranks > 1) {
        auto halo_cells = [&](const Block& bb, std::vector<int>& gi_, std::vector<int>& gj_, std::vector<int>& k_, std::vector<int>& slot_) {
            for (int k = 0; k < bb.L; k++) for (int je = -1; je <= bb.m0; je++) for (int ie = -1; ie <= bb.n0; ie++) {
                if (ie >= 0 && ie < bb.n0 && je >= 0 && je < bb.m0) continue;
                int hs = halo_slot(bb, ie, je, k);
                if (hs < 0) continue;
                int gj = bb.j0 + je;
                if (gj < 0 || gj >= bb.M) continue;
                int gi = bb.i0 + ie;
                if (bb.periodic) gi = ((gi % bb.N) + bb.N) % bb.N;
                if (gi < 0 || gi >= bb.N) continue;
                gi_.push_back(gi); gj_.push_back(gj); k_.push_back(k); slot_.push_back(hs);
            }
        };
        std::vector<int> hgi, hgj, hk_, hslot;
        halo_cells(b, hgi, hgj, hk_, hslot);
        for (int p = 0; p < b.nranks; p++) {
            if (p == b.rank) continue;
            thcmb_ctx::Peer peer{p, (int)send_idx.size(), 0, (int)recv_slot.size(), 0};
            for (size_t q = 0; q < hslot.size(); q++)
                if (owner_of(b, hgi[q], hgj[q]) == p) recv_slot.push_back(hslot[q]);
            Block pb;
            decomp2d(b.nranks, p, N, M, L, b.periodic, pb, b.cuts);
            std::vector<int> pgi, pgj, pk, pslot;
            halo_cells(pb, pgi, pgj, pk, pslot);
            for (size_t q = 0; q < pslot.size(); q++)
                if (owner_of(b, pgi[q], pgj[q]) == b.rank) {
                    send_idx.push_back((pk[q] * b.m0 + (pgj[q] - b.j0)) * b.n0 + (pgi[q] - b.i0));
                    c->send_dst_host.push_back(pslot[q]);                 // where the cell lives in the peer's halo buffer
                    c->send_peer_host.push_back((int)c->peers.size());    // index into c->peers (this peer is appended below)
                }
            peer.send_cnt = (int)send_idx.size() - peer.send_off;
            peer.recv_cnt = (int)recv_slot.size() - peer.recv_off;
            if (peer.send_cnt || peer.recv_cnt) c->peers.push_back(peer);
        }
    }
    c->nsend_cells = (int)send_idx.size(); c->nrecv_cells = (int)recv_slot.size();
    // cell compaction maps (groundwork for the ocean-only Krylov space): ocean cells of the block in cell order,
    // and the inverse map (compact index or -1 for LAND)
    c->ocell_host.clear();
    c->ccell_host.assign((size_t)b.ncell(), -1);
    for (int cell = 0; cell < b.ncell(); cell++)
        if (!((nbmask[(size_t)cell] >> 4) & 1u)) { c->ccell_host[(size_t)cell] = (int)c->ocell_host.size(); c->ocell_host.push_back(cell); }
    // column ids of the ocean-only (cell-compacted) space, stored at the SAME offsets as the graph's column array: compact position of
    // an owned ocean column, -1 for a column on LAND (its value is an exact zero and x = 0 there), nc6 + halo offset for a halo column
    // (halo cells keep their slots; a neighbour pushes zeros for its LAND cells).  Entries of LAND rows are never read.
    const int nc6 = NUN * (int)c->ocell_host.size(), nd = b.ndim();
    c->colc_host.assign(c->col_host.size(), -1);
    for (size_t ci = 0; ci < c->ocell_host.size(); ci++)
        for (int r = 0; r < NUN; r++) {
            const int row = NUN * c->ocell_host[ci] + r;
            for (int q = c->rowptr_host[row]; q < c->rowptr_host[row + 1]; q++) {
                const int id = c->col_host[q];
                if (id >= nd) c->colc_host[q] = nc6 + (id - nd);
                else { const int cc = c->ccell_host[(size_t)(id / NUN)]; c->colc_host[q] = cc < 0 ? -1 : NUN * cc + id % NUN; }
            }
        }
    // compact source of every cell of the send lists (-1 = LAND: a zero record is pushed)
    c->send_cidx_host.resize(send_idx.size());
    for (size_t q = 0; q < send_idx.size(); q++) c->send_cidx_host[q] = c->ccell_host[(size_t)send_idx[q]];
}


}  // namespace thcm

extern "C" const char* thcmb_last_error(void) { return thcm::last_error().c_str(); }
