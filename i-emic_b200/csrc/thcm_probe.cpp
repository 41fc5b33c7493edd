// Host-side diagnostics and setup symbols of the B1 boundary that are NOT on the Newton-step hot path but that
// src/ocean/THCM.C / Ocean.C bind (THCM.C:49-176, Ocean.C:42-49): m_probe (probe.F90), m_integrals (integrals.F90),
// get_stochastic_forcing (forcing.F90:235-280), getdeps / get_parameters (usrc.F90:201-251), writeparams, write_data,
// m_thcm_utils::get_landm / loadbal_weights, m_usr::set_internal_forcing.  They run at parameter-change / output
// frequency on n*m surface fields, are plain host C++ in the reference's statement order, and exist so that THCM.C links
// against libthcm_b200.so without a single unresolved symbol.  State-dependent ones take the HOST state vector the
// reference passes (single-rank sub-domain view, like the Fortran they replace).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include "thcm_internal.h"

namespace thcm {

namespace {
constexpr double rhodim = 1.024e+03, t0 = 15, deltat = 1.0, s0 = 35.0, r0dim = 6.37e+06, udim = 0.1e+00;   // usr.F90:132-160

inline int LMc(const thcmb_ctx* c, int i, int j, int k) {
    return c->landm[(size_t)i + (size_t)(c->s.N + 2) * (j + (size_t)(c->s.M + 2) * k)];
}
inline size_t frow(const thcmb_ctx* c, int i, int j, int k, int XX) {   // find_row2 - 1 (matetc.F90:123-131)
    return (size_t)NUN * (((size_t)(k - 1) * c->s.M + (j - 1)) * c->s.N + (i - 1)) + XX - 1;
}
inline double F2c(const std::vector<double>& f, const thcmb_ctx* c, int i, int j) { return f[(size_t)(i - 1) + (size_t)c->s.N * (j - 1)]; }
void need_single_rank(const thcmb_ctx* c, const char* what) {
    if (c->blk.nranks != 1) fatal(std::string(what) + ": the Fortran-symbol diagnostics see one sub-domain = the whole domain (nranks = 1)");
}

// usol (usrc.F90:1014-1121) on the host, all ghost layers, in the reference's statement order
struct Usol {
    int n, m, l;
    std::vector<double> u, v, w, t, s;   // all dimensioned (0:n+1, 0:m+1, 0:l+1) for simplicity
    inline size_t ix(int i, int j, int k) const { return (size_t)i + (size_t)(n + 2) * (j + (size_t)(m + 2) * k); }
    Usol(const thcmb_ctx* c, const double* un) : n(c->s.N), m(c->s.M), l(c->s.L) {
        const size_t tot = (size_t)(n + 2) * (m + 2) * (l + 2);
        u.assign(tot, 0.0); v.assign(tot, 0.0); w.assign(tot, 0.0); t.assign(tot, 0.0); s.assign(tot, 0.0);
        const bool periodic = c->s.periodic != 0;
        for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            const double* r = un + frow(c, i, j, k, 1);
            u[ix(i, j, k)] = r[0]; v[ix(i, j, k)] = r[1]; w[ix(i, j, k)] = r[2]; t[ix(i, j, k)] = r[4]; s[ix(i, j, k)] = r[5];
        }
        for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) {
            if (periodic) {
                u[ix(0, j, k)] = u[ix(n, j, k)]; v[ix(0, j, k)] = v[ix(n, j, k)];
                w[ix(n + 1, j, k)] = w[ix(1, j, k)]; w[ix(0, j, k)] = w[ix(n, j, k)];
                t[ix(n + 1, j, k)] = t[ix(1, j, k)]; t[ix(0, j, k)] = t[ix(n, j, k)];
                s[ix(n + 1, j, k)] = s[ix(1, j, k)]; s[ix(0, j, k)] = s[ix(n, j, k)];
            } else {
                u[ix(0, j, k)] = 0.0; u[ix(n, j, k)] = 0.0; v[ix(0, j, k)] = 0.0; v[ix(n, j, k)] = 0.0;
                t[ix(0, j, k)] = t[ix(1, j, k)]; t[ix(n + 1, j, k)] = t[ix(n, j, k)];
                s[ix(0, j, k)] = s[ix(1, j, k)]; s[ix(n + 1, j, k)] = s[ix(n, j, k)];
            }
        }
        for (int k = 1; k <= l; k++) for (int i = 1; i <= n; i++) {
            u[ix(i, 0, k)] = 0.0; u[ix(i, m, k)] = 0.0; v[ix(i, 0, k)] = 0.0; v[ix(i, m, k)] = 0.0;
            t[ix(i, 0, k)] = t[ix(i, 1, k)]; t[ix(i, m + 1, k)] = t[ix(i, m, k)];
            s[ix(i, 0, k)] = s[ix(i, 1, k)]; s[ix(i, m + 1, k)] = s[ix(i, m, k)];
        }
        for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
            u[ix(i, j, 0)] = u[ix(i, j, 1)]; u[ix(i, j, l + 1)] = u[ix(i, j, l)];
            v[ix(i, j, 0)] = v[ix(i, j, 1)]; v[ix(i, j, l + 1)] = v[ix(i, j, l)];
            w[ix(i, j, l)] = 0.0; w[ix(i, j, 0)] = 0.0;
            t[ix(i, j, l + 1)] = t[ix(i, j, l)]; t[ix(i, j, 0)] = t[ix(i, j, 1)];
            s[ix(i, j, l + 1)] = s[ix(i, j, l)]; s[ix(i, j, 0)] = s[ix(i, j, 1)];
        }
        for (int i = 1; i <= n; i++) for (int j = 1; j <= m; j++) for (int k = 1; k <= l; k++)
            if (LMc(c, i, j, k) == LAND) {
                u[ix(i, j, k)] = 0.0; v[ix(i, j, k)] = 0.0; u[ix(i - 1, j, k)] = 0.0; v[ix(i - 1, j, k)] = 0.0;
                u[ix(i, j - 1, k)] = 0.0; v[ix(i, j - 1, k)] = 0.0; u[ix(i - 1, j - 1, k)] = 0.0; v[ix(i - 1, j - 1, k)] = 0.0;
            }
    }
};

// forcing.F90:452-464 (the reference calls back into C++ for the MPI sum; one rank: the plain loop)
double qint_host(const thcmb_ctx* c, const double* f) {
    const int n = c->s.N, m = c->s.M, l = c->s.L;
    double lf = 0.0, ls = 0.0;
    for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
        lf = f[(size_t)(i - 1) + (size_t)n * (j - 1)] * std::cos(c->y[j]) * (1 - LMc(c, i, j, l)) + lf;
        ls = std::cos(c->y[j]) * (1 - LMc(c, i, j, l)) + ls;
    }
    return lf / ls;
}
}  // namespace

// m_probe::get_atmosphere_t/q/p, get_emip, get_adapted_emip, get_emip_pert, get_taux, get_tauy (probe.F90:11-175, 440-490):
// `which` as SurfaceField.  Returns false when the reference would leave the output untouched (no coupling).
bool probe_get_field(const thcmb_ctx* c, int which, double* out) {
    const thcmb_settings& s = c->s;
    const std::vector<double>* src[SF_COUNT] = {&c->taux, &c->tauy, &c->tatm, &c->emip, &c->spert, &c->adapted_emip,
                                                &c->qatm, &c->albe, &c->patm, &c->qsa, &c->msi, &c->gsi};
    if (which < 0 || which >= SF_COUNT) fatal("probe_get_field: unknown field");
    if (which == SF_QATM && !(s.coupled_T == 1 || s.coupled_S == 1)) return false;
    if (which == SF_PATM && s.coupled_S != 1) return false;
    size_t pos = 0;
    for (int j = 1; j <= s.M; j++) for (int i = 1; i <= s.N; i++, pos++)
        out[pos] = which == SF_EMIP ? (*src[which])[pos] * (1 - LMc(c, i, j, s.L)) : (*src[which])[pos];   // probe.F90:124
    return true;
}
// m_probe::get_suno (probe.F90:353-369)
void probe_get_suno(const thcmb_ctx* c, double* out) {
    size_t pos = 0;
    for (int j = 1; j <= c->s.M; j++) for (int i = 1; i <= c->s.N; i++, pos++) out[pos] = c->suno[j];
}
// m_probe::compute_evap (probe.F90:75-113)
void probe_compute_evap(const thcmb_ctx* c, const double* un, double* evap) {
    need_single_rank(c, "compute_evap");
    const thcmb_settings& s = c->s;
    std::fill(evap, evap + (size_t)s.N * s.M, 0.0);
    if (!(s.coupled_T == 1 || s.coupled_S == 1)) return;
    size_t pos = 0;
    for (int j = 1; j <= s.M; j++) for (int i = 1; i <= s.N; i++, pos++)
        if (LMc(c, i, j, s.L) == 0)
            evap[pos] = c->atm_eo0 + c->atm_eta * c->atm_qdim * (((deltat / c->atm_qdim) * c->atm_dqso * un[frow(c, i, j, s.L, TT)] - F2c(c->qatm, c, i, j)));
}
// m_probe::get_salflux (probe.F90:177-245)
void probe_get_salflux(thcmb_ctx* c, const double* un, double* salflux, double* correction, double* qsoaflux, double* qsosflux) {
    need_single_rank(c, "get_salflux");
    const thcmb_settings& s = c->s;
    const double* par = c->par;
    const int n = s.N, m = s.M, l = s.L;
    const double gamma = par[COMB] * par[SALT];
    std::fill(salflux, salflux + (size_t)n * m, 0.0);
    const double pQSnd = par[COMB] * par[SALT] * c->QSnd;
    // side effect of the reference kept: the module variable nus is overwritten WITHOUT the eta*qdim factor (probe.F90:224), which
    // the next `lin` then uses until set_atmos_parameters / get_derivatives restore it
    c->atm_nus = par[COMB] * par[SALT] * c->QSnd;
    size_t pos = 0;
    for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++, pos++) {
        const double T = un[frow(c, i, j, l, TT)], S = un[frow(c, i, j, l, SS)];
        const double QSos = pQSnd * (c->ice_zeta * (c->ice_a0 * (s0 + S) - (t0 + T)) - (c->ice_Qvar * F2c(c->qsa, c, i, j) + c->ice_Q0)) /
                            (rhodim * c->ice_Lf);
        const double QSoa = pQSnd * (c->atm_eo0 + c->atm_eta * c->atm_qdim * ((deltat / c->atm_qdim) * c->atm_dqso * T - F2c(c->qatm, c, i, j)) -
                                     F2c(c->patm, c, i, j));
        qsoaflux[pos] = QSoa / c->QSnd * (1 - F2c(c->msi, c, i, j));
        qsosflux[pos] = QSos / c->QSnd * F2c(c->msi, c, i, j);
        if (gamma != 0) {
            if (s.coupled_S == 1)
                salflux[pos] = (QSoa + F2c(c->msi, c, i, j) * (QSos - QSoa)) * (1 - LMc(c, i, j, l)) / gamma;
            else
                salflux[pos] = (1 - LMc(c, i, j, l)) * (1 - s.SRES + s.SRES * par[BIOT]) * F2c(c->emip, c, i, j) - s.SRES * par[BIOT] * S / gamma;
        }
    }
    const double corr = qint_host(c, salflux);
    for (size_t q = 0; q < (size_t)n * m; q++) salflux[q] = salflux[q] - corr;
    *correction = corr * gamma;
}
// m_probe::get_temflux (probe.F90:247-351); entries of non-OCEAN surface cells are left untouched, like the reference
void probe_get_temflux(const thcmb_ctx* c, const double* un, double* totflux, double* swflux, double* shflux, double* lhflux, double* siflux,
                       double* simask) {
    need_single_rank(c, "get_temflux");
    const thcmb_settings& s = c->s;
    const double* par = c->par;
    const int n = s.N, m = s.M, l = s.L;
    const double etabi = par[COMB] * par[TEMP];
    const double dedt = c->atm_eta * c->atm_qdim * (deltat / c->atm_qdim) * c->atm_dqso;
    const double dedq = -c->atm_eta * c->atm_qdim;
    size_t pos = 0;
    for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++, pos++) {
        if (LMc(c, i, j, l) != OCEAN) continue;
        const double T = un[frow(c, i, j, l, TT)], S = un[frow(c, i, j, l, SS)], msi = F2c(c->msi, c, i, j);
        const double QSW = par[COMB] * par[SUNP] * c->suno[j] * (1 - c->atm_albe0 - c->atm_albed * F2c(c->albe, c, i, j));
        const double QSH = c->atm_Ooa * (T - F2c(c->tatm, c, i, j));
        const double QLH = c->atm_lvsc * (c->atm_eo0 + dedt * T + dedq * F2c(c->qatm, c, i, j));
        const double QToa = QSW - QSH - QLH;
        const double QTos = c->QTnd * c->ice_zeta * (c->ice_a0 * (s0 + S) - (t0 + T));
        swflux[pos] = QSW / c->QTnd * (1 - msi);
        shflux[pos] = -QSH / c->QTnd * (1 - msi);
        lhflux[pos] = -QLH / c->QTnd * (1 - msi);
        siflux[pos] = QTos / c->QTnd * msi;
        simask[pos] = msi;
        if (s.coupled_T == 0)
            totflux[pos] = (1 - s.TRES + s.TRES * par[BIOT]) * F2c(c->tatm, c, i, j) - s.TRES * par[BIOT] * T / etabi;
        else
            totflux[pos] = (1 - LMc(c, i, j, l)) * QToa / c->QTnd + msi * (QTos - QToa) / c->QTnd;
    }
}
// m_probe::get_derivatives (probe.F90:371-438); like the reference this refreshes nus from the current COMB / SALT
void probe_get_derivatives(thcmb_ctx* c, const double* un, double* dftdm, double* dfsdq, double* dfsdm, double* dfsdg) {
    need_single_rank(c, "get_derivatives");
    const thcmb_settings& s = c->s;
    const double* par = c->par;
    const int n = s.N, m = s.M, l = s.L;
    const size_t nm = (size_t)n * m;
    std::fill(dftdm, dftdm + nm, 0.0); std::fill(dfsdq, dfsdq + nm, 0.0); std::fill(dfsdm, dfsdm + nm, 0.0); std::fill(dfsdg, dfsdg + nm, 0.0);
    c->atm_nus = par[COMB] * par[SALT] * c->atm_eta * c->atm_qdim * c->QSnd;
    const double pQSnd = par[COMB] * par[SALT] * c->QSnd;
    size_t pos = 0;
    for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++, pos++) {
        if (LMc(c, i, j, l) != OCEAN) continue;
        const double To = un[frow(c, i, j, l, TT)], So = un[frow(c, i, j, l, SS)];
        const double Ta = F2c(c->tatm, c, i, j), Ab = F2c(c->albe, c, i, j), qa = F2c(c->qatm, c, i, j), pa = F2c(c->patm, c, i, j),
                     Ms = F2c(c->msi, c, i, j), qs = F2c(c->qsa, c, i, j);
        if (s.coupled_T == 1) {
            const double QTos = c->QTnd * c->ice_zeta * (c->ice_a0 * (So + s0) - (To + t0));
            const double QToa = par[COMB] * par[SUNP] * c->suno[j] * (1 - c->atm_albe0 - c->atm_albed * Ab) - c->atm_Ooa * (To - Ta) -
                                c->atm_lvsc * c->atm_eta * c->atm_qdim * (deltat / c->atm_qdim * c->atm_dqso * To - qa) - c->atm_lvsc * c->atm_eo0;
            dftdm[pos] = QTos - QToa;
        }
        if (s.coupled_S == 1) {
            dfsdq[pos] = -pQSnd * c->ice_Qvar / (rhodim * c->ice_Lf) * Ms;
            const double QSos = (c->ice_zeta * (c->ice_a0 * (s0 + So) - (t0 + To)) - (c->ice_Qvar * qs + c->ice_Q0)) / (rhodim * c->ice_Lf);
            const double QSoa = c->atm_eo0 + c->atm_eta * c->atm_qdim * ((deltat / c->atm_qdim) * c->atm_dqso * To - qa) - pa;
            dfsdm[pos] = pQSnd * (QSos - QSoa);
            dfsdg[pos] = -1.0;
        }
    }
}

// m_integrals::salt_advection / salt_diffusion (integrals.F90:17-88): per-cell integrands, entries of skipped cells untouched
void integrals_salt_advection(const thcmb_ctx* c, const double* un, double* check) {
    need_single_rank(c, "salt_advection");
    const Usol f(c, un);
    const int n = f.n, m = f.m, l = f.l;
    const double dx = c->dx, dy = c->dy, dz = c->dz;
    size_t pos = 0;
    for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++, pos++) {
        if (LMc(c, i, j, l) != OCEAN) continue;   // sic: the surface mask (integrals.F90:35)
        auto U = [&](int a, int b, int d) { return f.u[f.ix(a, b, d)]; };
        auto V = [&](int a, int b, int d) { return f.v[f.ix(a, b, d)]; };
        auto W = [&](int a, int b, int d) { return f.w[f.ix(a, b, d)]; };
        auto S = [&](int a, int b, int d) { return f.s[f.ix(a, b, d)]; };
        check[pos] = (U(i, j, k) + U(i, j - 1, k)) * (S(i + 1, j, k) + S(i, j, k)) / (4 * dx) -
                     (U(i - 1, j, k) + U(i - 1, j - 1, k)) * (S(i, j, k) + S(i - 1, j, k)) / (4 * dx) +
                     (V(i, j, k) + V(i - 1, j, k)) * (S(i, j + 1, k) + S(i, j, k)) * std::cos(c->yv[j]) / (4 * dy) -
                     (V(i, j - 1, k) + V(i - 1, j - 1, k)) * (S(i, j, k) + S(i, j - 1, k)) * std::cos(c->yv[j - 1]) / (4 * dy) +
                     W(i, j, k) * (S(i, j, k + 1) + S(i, j, k)) * std::cos(c->y[j]) / (2 * dz * c->dfzW[k]) -
                     W(i, j, k - 1) * (S(i, j, k) + S(i, j, k - 1)) * std::cos(c->y[j]) / (2 * dz * c->dfzW[k - 1]);
    }
}
void integrals_salt_diffusion(const thcmb_ctx* c, const double* un, double* check) {
    need_single_rank(c, "salt_diffusion");
    const Usol f(c, un);
    const int n = f.n, m = f.m, l = f.l;
    const double dx = c->dx, dy = c->dy, dz = c->dz;
    size_t pos = 0;
    for (int k = 1; k <= l; k++) {
        const double h1 = 1. / (c->dfzT[k] * c->dfzW[k]), h2 = 1. / (c->dfzT[k] * c->dfzW[k - 1]);
        for (int j = 1; j <= m; j++) {
            const double cay = std::cos(c->y[j]), c1 = std::cos(c->yv[j]), c2 = std::cos(c->yv[j - 1]);
            for (int i = 1; i <= n; i++, pos++) {
                if (LMc(c, i, j, k) != OCEAN) continue;
                auto S = [&](int a, int b, int d) { return f.s[f.ix(a, b, d)]; };
                check[pos] = std::cos(c->y[j]) * c->dfzT[k] *
                             ((S(i + 1, j, k) + S(i - 1, j, k) - 2 * S(i, j, k)) / (dx * dx * cay * cay) +
                              (c1 * S(i, j + 1, k) + c2 * S(i, j - 1, k) - (c1 + c2) * S(i, j, k)) / (dy * dy * cay) +
                              (h1 * S(i, j, k + 1) + h2 * S(i, j, k - 1) - (h1 + h2) * S(i, j, k)) / (dz * dz));
            }
        }
    }
}

// get_stochastic_forcing (forcing.F90:235-280): the S-row surface forcing at SPER = 0 as a 1-based CRS "matrix" F with
// one entry per surface cell (column = j), into the begF / jcoF / coF buffers of m_mat::set_pointers
void stochastic_forcing(thcmb_ctx* c, int* begF, int* jcoF, double* coF) {
    need_single_rank(c, "get_stochastic_forcing");
    if (!begF || !jcoF || !coF) fatal("get_stochastic_forcing: set_pointers was not called with begF / jcoF / coF");
    const thcmb_settings& s = c->s;
    const int n = s.N, m = s.M, l = s.L, ndim = c->blk.ndim();
    const double oldpar = c->par[SPER];
    c->par[SPER] = 0.0;
    compute_forcing(c);
    int v = 1;
    std::fill(begF, begF + ndim + 1, 0);
    for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++)
        if (s.coupled_S != 1) {
            const int row = (int)frow(c, i, j, l, SS) + 1;
            jcoF[v - 1] = j;
            coF[v - 1] = c->frc_raw[row - 1];
            v = v + 1;
            begF[row] = v;   // begF(row+1), 1-based
        }
    v = 1;
    for (int i = 0; i <= ndim; i++) { if (begF[i] == 0) begF[i] = v; else v = begF[i]; }
    c->par[SPER] = oldpar;
    compute_forcing(c);
}

// getdeps / get_parameters (usrc.F90:201-234)
void get_deps(const thcmb_ctx* c, double* out7) {
    out7[0] = c->atm_Ooa; out7[1] = c->atm_Os; out7[2] = c->atm_nus; out7[3] = c->atm_eta; out7[4] = c->atm_lvsc; out7[5] = c->atm_qdim;
    out7[6] = c->par[COMB] * c->par[SALT] * c->QSnd;
}
// Ocean::getBlock(Atmosphere) (Ocean.C:1603-1730): the dependence of the ocean rows on the atmosphere's unknowns as a CRS block over
// ALL 6 N M L ocean rows in FIND_ROW2 order (0-based beg[ndim + 1]); only the surface T rows (coupled_T: columns T, A, Q in that order)
// and the surface S rows (coupled_S, the integral-condition row excepted: columns Q and, when the atmosphere carries the auxiliary
// precipitation unknown, P) of OCEAN surface points hold entries.  "Negating as the Jacobian is taken negative": the entries are
// d F_ocean / d x_atmos for the C++-sign residual F = A u - Frc.  col*: the atmosphere's interface_row(i, j, XX) per surface point
// (n*m, i fastest; colP < 0 = no auxiliary row), pdist = Atmosphere::getPdist (NULL = 1), albed = Atmosphere::CommPars::da.
int ocean_block_atmosphere(thcmb_ctx* c, double albed, const double* pdist, const int* colT, const int* colQ, const int* colA,
                           const int* colP, int* beg, int* jco, double* co) {
    need_single_rank(c, "Ocean::getBlock(Atmosphere)");
    const thcmb_settings& s = c->s;
    const int n = s.N, m = s.M, l = s.L;
    double dep[7]; get_deps(c, dep);
    const double Ooa = dep[0], nus = dep[2], eta = dep[3], lvsc = dep[4], qdim = dep[5];
    const double comb = c->par[COMB], sunp = c->par[SUNP];
    const int rowIntCon = c->ic_on ? c->ic_grow : -1;
    int el = 0;
    size_t row = 0;
    for (int k = 0; k < l; k++) for (int j = 0; j < m; j++) for (int i = 0; i < n; i++) {
        const size_t sr = (size_t)j * n + i;
        const double M = c->msi[sr], S = c->suno[j + 1], Pd = pdist ? pdist[sr] : 1.0;
        for (int xx = UU; xx <= SS; xx++, row++) {
            beg[row] = el;
            if (k != l - 1 || LMc(c, i + 1, j + 1, l) != OCEAN) continue;
            if (xx == TT && s.coupled_T) {
                co[el] = -(Ooa * (1.0 - M)); jco[el++] = colT[sr];
                co[el] = -(-comb * sunp * S * albed * (1.0 - M)); jco[el++] = colA[sr];
                co[el] = -(lvsc * eta * qdim * (1.0 - M)); jco[el++] = colQ[sr];
            } else if (xx == SS && s.coupled_S && (int)row != rowIntCon) {
                co[el] = -(-nus * (1.0 - M)); jco[el++] = colQ[sr];
                if (colP && colP[sr] >= 0) { co[el] = -(-nus * Pd * (1.0 - M)); jco[el++] = colP[sr]; }
            }
        }
    }
    beg[row] = el;
    return el;
}
// Ocean::getBlock(SeaIce) (Ocean.C:1733-1810): likewise for the sea-ice unknowns Q (heat flux), M (mask), G (integral correction), from
// m_probe::get_derivatives at the HOST state un (THCM::getDerivatives).  Surface T rows: column M; surface S rows: columns Q, M, G.
int ocean_block_seaice(thcmb_ctx* c, const double* un, const int* colQ, const int* colM, const int* colG, int* beg, int* jco, double* co) {
    need_single_rank(c, "Ocean::getBlock(SeaIce)");
    const thcmb_settings& s = c->s;
    const int n = s.N, m = s.M, l = s.L;
    const size_t nm = (size_t)n * m;
    std::vector<double> dftdm(nm), dfsdq(nm), dfsdm(nm), dfsdg(nm);
    probe_get_derivatives(c, un, dftdm.data(), dfsdq.data(), dfsdm.data(), dfsdg.data());
    const int rowIntCon = c->ic_on ? c->ic_grow : -1;
    int el = 0;
    size_t row = 0;
    for (int k = 0; k < l; k++) for (int j = 0; j < m; j++) for (int i = 0; i < n; i++) {
        const size_t sr = (size_t)j * n + i;
        for (int xx = UU; xx <= SS; xx++, row++) {
            beg[row] = el;
            if (k != l - 1 || LMc(c, i + 1, j + 1, l) != OCEAN) continue;
            if (xx == TT && s.coupled_T) { co[el] = -dftdm[sr]; jco[el++] = colM[sr]; }
            else if (xx == SS && s.coupled_S && (int)row != rowIntCon) {
                co[el] = -dfsdq[sr]; jco[el++] = colQ[sr];
                co[el] = -dfsdm[sr]; jco[el++] = colM[sr];
                co[el] = -dfsdg[sr]; jco[el++] = colG[sr];
            }
        }
    }
    beg[row] = el;
    return el;
}
void get_dim_parameters(const thcmb_ctx* c, double* r0, double* u0, double* h0) { *r0 = r0dim; *u0 = udim; *h0 = c->s.hdim; }

// m_thcm_utils::loadbal_weights (thcm_utils.F90:325-353): OCEAN cells per water column / l.  The vmix_counts terms are the
// cell counts of neutral physics / consistent mixing / convective adjustment of mix_imp.f, all zero for the mixing schemes
// this library implements (implicit vertical mixing only), whatever the factors.
void loadbal_weights(const thcmb_ctx* c, double* array) {
    const int n = c->s.N, m = c->s.M, l = c->s.L;
    for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
        int cnt = 0;
        for (int k = 0; k <= l + 1; k++) cnt += LMc(c, i, j, k) == OCEAN ? 1 : 0;
        array[(size_t)(i - 1) + (size_t)n * (j - 1)] = (double)cnt / l;
    }
}

// writeparams (usrc.F90:421-431): the 30 parameters to fort.7, '(5e15.5)'
void write_params(const thcmb_ctx* c) {
    FILE* f = fopen("fort.7", "a");
    if (!f) { fprintf(stderr, "thcm_b200: cannot open fort.7\n"); return; }
    fprintf(f, " ---------------------------\n");
    for (int i = 1; i <= NPAR; i++) fprintf(f, "%15.5E%s", c->par[i], (i % 5 == 0 || i == NPAR) ? "\n" : "");
    fprintf(f, " ---------------------------\n");
    fclose(f);
}
// write_data (inout.F90:20-93): the legacy fort.3 solution file (header, parameters, one unknown per line in i,j,k,XX
// order).  The geometry file fort.44 of write_geometry is not produced (it needs m_global's file-backed arrays).
void write_data(const thcmb_ctx* c, const double* u, int ofile, int* lab) {
    if (ofile == 0) return;
    *lab = *lab + 1;
    FILE* f = fopen("fort.3", "w");
    if (!f) { fprintf(stderr, "thcm_b200: cannot open fort.3\n"); return; }
    const int n = c->s.N, m = c->s.M, l = c->s.L, ndim = NUN * n * m * l, nf = 0, icp = 0;
    const int nskip = (NPAR - 1) / 5 + 1 + 1 + nf + ndim * ((nf + 1) / 10 + 1);
    fprintf(f, "Version   0%4d%4d%4d%4d%4d%4d%4d%4d%12d%12d\n", *lab, icp, NPAR, nf, n, m, l, NUN, ndim, nskip);
    for (int i = 1; i <= NPAR; i++) fprintf(f, "%18.10E %s", c->par[i], (i % 5 == 0 || i == NPAR) ? "\n" : "");
    fprintf(f, "%18.10E %16.8E %16.8E\n", 0.0, 0.0, 0.0);
    for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) for (int XX = 1; XX <= NUN; XX++)
        fprintf(f, "%18.10E\n", u[frow(c, i, j, k, XX)]);
    fclose(f);
}

}  // namespace thcm
