// =============================================================================
// test_ocean_mirror.cpp -- the reference's ocean unit tests (src/tests/test_ocean.C) re-stated over the C++ mirror
// include/thcm_model.hpp (thcm_b200::Ocean), same test names and checks, a minimal EXPECT_* in place of GoogleTest:
//   Ocean.Initialization    test_ocean.C:28-57    construction, ||F(0)|| < 1e-6 and ||Frc|| = 0 at Combined Forcing = 0
//   Ocean.MassMat           test_ocean.C:60-125   applyMassMat(1) = (-Ro | 0, -Ro | 0, 0, 0, -1, -1) on ocean cells, 0 at the integral row
//   Ocean.ComputeJacobian   test_ocean.C:127-182  computeJacobian at Combined Forcing = 0.1; J against centred differences of F
//   Ocean.IntegralCondition THCM.C:1013-1026      F_row = sign (c.x - correction), (J v)_row = sign c.v  (SRES = 0, the test's own setting)
// Configuration = test/ocean/ocean_params.xml (8 x 8 x 4 North Atlantic box, mask_natl8, Restoring Salinity Profile = 0).
// Built by tests/cpp/Makefile into tests/cpp/_bin/; run by tests/test_zz_unverified.py on a B200.
// =============================================================================
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "thcm_model.hpp"

static int g_failed = 0, g_checks = 0;
#define EXPECT_TRUE(c) do { g_checks++; if (!(c)) { g_failed++; std::printf("  FAILED %s:%d: %s\n", __FILE__, __LINE__, #c); } } while (0)
#define EXPECT_EQ(a, b) EXPECT_TRUE((a) == (b))
#define EXPECT_LT(a, b) EXPECT_TRUE((a) < (b))
#define EXPECT_NEAR(a, b, tol) EXPECT_TRUE(std::fabs((a) - (b)) <= (tol))
#define TEST(suite, name) static void suite##_##name(); static void suite##_##name()
#define RUN(suite, name) do { int f0 = g_failed; std::printf("[ RUN      ] %s.%s\n", #suite, #name); suite##_##name(); \
                              std::printf("[ %s ] %s.%s\n", g_failed == f0 ? "      OK" : " FAILED ", #suite, #name); } while (0)

using thcm_b200::Ocean;
using thcm_b200::Vector;

static std::shared_ptr<Ocean> ocean;
static std::vector<int> landm;
static int N = 8, M = 8, L = 4;

static double rnd(unsigned long long& s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return (double)((s >> 11) & ((1ull << 53) - 1)) / (double)(1ull << 52) - 1.0; }
static bool is_ocean(int i, int j, int k) { return landm[(size_t)(i + 1) + (size_t)(N + 2) * ((j + 1) + (size_t)(M + 2) * (k + 1))] == 0; }
static int lm(int i, int j, int k) { return landm[(size_t)i + (size_t)(N + 2) * (j + (size_t)(M + 2) * k)]; }   // Fortran indices 0..n+1
// rows `boundaries` turns into identity rows (boundary.F90): Dirichlet unknowns, zero in every state the solver produces
static bool dirichlet(int i, int j, int k, int xx) {   // 0-based cell, xx 0..5
    const int I = i + 1, J = j + 1, K = k + 1;
    if (lm(I, J, K) != 0) return true;
    if (xx <= 1) return lm(I, J + 1, K) == 1 || lm(I + 1, J, K) == 1 || lm(I + 1, J + 1, K) == 1;
    if (xx == 2) return K == L || lm(I, J, K + 1) == 1;
    return false;
}
static std::vector<double> manifold_state(unsigned long long seed, double scale) {
    std::vector<double> x((size_t)6 * N * M * L);
    for (int k = 0; k < L; k++) for (int j = 0; j < M; j++) for (int i = 0; i < N; i++) for (int xx = 0; xx < 6; xx++) {
        const double r = rnd(seed);
        x[(size_t)6 * ((k * M + j) * N + i) + xx] = dirichlet(i, j, k, xx) ? 0.0 : scale * r;
    }
    return x;
}

TEST(Ocean, Initialization) {
    ocean->setPar("Combined Forcing", 0.0);
    ocean->getState('V')->zero();
    ocean->computeRHS();
    EXPECT_LT(ocean->getRHS('V')->norm(), 1e-6);                 // test_ocean.C:42-49
    ocean->setPar("Combined Forcing", 0.1);
    ocean->computeRHS();
    EXPECT_TRUE(ocean->getRHS('V')->norm() > 1e-6);              // the forcing is what drives the residual
}

TEST(Ocean, MassMat) {
    Vector v(ocean->context()), out(ocean->context());
    v.fill(1.0);
    ocean->computeMassMat();
    ocean->applyMassMat(v, out);
    const std::vector<double> o = out.toHost();
    EXPECT_EQ((int)o.size(), 6 * N * M * L);
    const double rosb = ocean->getPar("Rossby-Number");
    for (int k = 0; k < L; k++) for (int j = 0; j < M; j++) for (int i = 0; i < N; i++) {
        if (!is_ocean(i, j, k)) continue;
        const double* r = &o[(size_t)6 * ((k * M + j) * N + i)];
        if (std::fabs(r[0]) > 0) EXPECT_EQ(r[0], -rosb);
        if (std::fabs(r[1]) > 0) EXPECT_EQ(r[1], -rosb);
        EXPECT_EQ(r[2], 0.0);
        EXPECT_EQ(r[3], 0.0);
        EXPECT_EQ(r[4], -1.0);
        if (std::fabs(r[5]) > 0) EXPECT_EQ(r[5], -1.0);
    }
    const int rowIntCon = 6 * (((L - 1) * M + (M - 1)) * N + (N - 1)) + 5;   // FIND_ROW2(6,N,M,L,N-1,M-1,L-1,SS), 0-based
    EXPECT_EQ(o[(size_t)rowIntCon], 0.0);                         // test_ocean.C:109-124
}

TEST(Ocean, ComputeJacobian) {
    ocean->setPar("Combined Forcing", 0.1);
    const std::vector<double> x = manifold_state(12345ull, 0.05);
    ocean->getState('V')->fromHost(x.data());
    ocean->computeJacobian();
    Vector d(ocean->context()), Jd(ocean->context()), Fp(ocean->context()), Fm(ocean->context());
    const double h = 1e-6;
    for (int dir = 0; dir < 4; dir++) {
        const std::vector<double> dv = manifold_state(777ull + dir, 1.0);
        d.fromHost(dv.data());
        ocean->applyMatrix(d, Jd);
        std::vector<double> xp(x), xm(x);
        for (size_t q = 0; q < x.size(); q++) { xp[q] += h * dv[q]; xm[q] -= h * dv[q]; }
        ocean->getState('V')->fromHost(xp.data()); ocean->computeRHS(); Fp = *ocean->getRHS('V');
        ocean->getState('V')->fromHost(xm.data()); ocean->computeRHS(); Fm = *ocean->getRHS('V');
        Fp.update(-1.0, Fm, 1.0);
        Fp.scale(1.0 / (2 * h));
        // the integral-condition row is linear in x: included; compare in the 2-norm
        Fp.update(-1.0, Jd, 1.0);
        EXPECT_LT(Fp.norm(), 1e-6 * Jd.norm());
    }
    ocean->getState('V')->fromHost(x.data());
}

TEST(Ocean, IntegralCondition) {
    const int rowIntCon = 6 * (((L - 1) * M + (M - 1)) * N + (N - 1)) + 5;
    std::vector<double> coeff;
    const double volume = ocean->getTHCM().getIntCondCoeff(coeff);
    EXPECT_TRUE(volume > 0.0);
    const std::vector<double> x = manifold_state(99ull, 0.05);
    EXPECT_EQ(coeff.size(), x.size());
    double cx = 0.0;
    for (size_t q = 0; q < coeff.size(); q++) cx += coeff[q] * x[q];
    ocean->getState('V')->fromHost(x.data());
    ocean->computeRHS();
    const std::vector<double> F = ocean->getRHS('V')->toHost();
    EXPECT_NEAR(F[(size_t)rowIntCon], -1.0 * (cx - 0.0), 1e-12 * (1.0 + std::fabs(cx)));   // intSign_ = -1, no correction
    ocean->computeJacobian();
    Vector v(ocean->context()), Jv(ocean->context());
    v.fromHost(x.data());
    ocean->applyMatrix(v, Jv);
    EXPECT_NEAR(Jv.toHost()[(size_t)rowIntCon], -cx, 1e-12 * (1.0 + std::fabs(cx)));
}

int main(int argc, char** argv) {
    if (argc < 2) { std::fprintf(stderr, "usage: test_ocean_mirror <mask_natl8 file>\n"); return 2; }
    const double PI = 3.14159265358979323846;
    int periodic = 0, zero = 0, one = 1, iza = 2;
    double xmin = 286 * PI / 180.0, xmax = 350 * PI / 180.0, ymin = 10 * PI / 180.0, ymax = 74 * PI / 180.0, hdim = 4000.0, qz = 1.0;
    __m_global_MOD_initialize(&N, &M, &L, &xmin, &xmax, &ymin, &ymax, &hdim, &qz, &periodic, &zero, &zero, &one, &one, &zero, &iza, &one, &one,
                              &zero, &zero, &zero, &zero, argv[1], "", "", "", "");
    landm.resize((size_t)(N + 2) * (M + 2) * (L + 2));
    __m_global_MOD_get_landm(landm.data());
    // the frame of dummy cells is LAND for the model (usrc.F90:100-107): mirror it for the Dirichlet bookkeeping of this test
    for (int k = 0; k <= L + 1; k++) for (int j = 0; j <= M + 1; j++) for (int i = 0; i <= N + 1; i++)
        if (i == 0 || i == N + 1 || j == 0 || j == M + 1 || k == 0 || k == L + 1) landm[(size_t)i + (size_t)(N + 2) * (j + (size_t)(M + 2) * k)] = 1;
    thcmb_settings s;
    thcmb_default_settings(&s);
    s.N = N; s.M = M; s.L = L; s.xmin = xmin; s.xmax = xmax; s.ymin = ymin; s.ymax = ymax; s.hdim = hdim; s.qz = qz; s.SRES = 0;
    ocean = std::make_shared<Ocean>(s, landm.data());
    ocean->setPar("Wind Forcing", 1.0); ocean->setPar("Temperature Forcing", 10.0); ocean->setPar("Salinity Forcing", 1.0);
    thcmb_enable_intcond(ocean->context(), -1, -1, -1);      // what THCM's constructor does for Restoring Salinity Profile = 0
    RUN(Ocean, Initialization);
    RUN(Ocean, MassMat);
    RUN(Ocean, ComputeJacobian);
    RUN(Ocean, IntegralCondition);
    std::printf("[==========] %d checks, %d failed\n", g_checks, g_failed);
    return g_failed ? 1 : 0;
}
