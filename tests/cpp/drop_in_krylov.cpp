// =============================================================================
// drop_in_krylov.cpp -- TEST PROGRAM (test infrastructure): the reference's own, UNMODIFIED Krylov drivers
//   /root/reference/src/gmressolver/GMRESSolver.H   GMRESSolver<Model, VectorPointer>
//   /root/reference/src/idrsolver/IDRSolver.H       IDRSolver<Model, VectorPointer>
// instantiated over the C++ mirror of this repo (include/thcm_model.hpp: thcm_b200::Ocean as the Model,
// std::shared_ptr<thcm_b200::Vector> as the VectorPointer), i.e. the drop-in claim of BASELINE.json for src/gmressolver and
// src/idrsolver: the reference's solver loops run unchanged while every vector operation, the SpMV and the preconditioner
// are kernels of libthcm_b200.so on device-resident data.
//
// Flow = the reference's own (THCM.C:328-390, 582-611; Ocean.C:1277-1309): m_global::initialize reads the mask file,
// get_landm hands it back, the context is created, parameters are set by XML name, computeRHS + computeJacobian, then
// J dx = F is solved three ways -- reference GMRES template, reference IDR(s) template, in-library thcmb_gmres -- and the
// residual histories are printed as one JSON line for tests/test_zz_cpp_mirror.py.
// Built by tests/cpp/Makefile into tests/cpp/_bin/ (git-ignored, travels to the GPU box prebuilt like oracle/_ref).
// =============================================================================
#include <cstdio>
#include <cstring>
#include <deque>
#include <iostream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "GlobalDefinitions.H"   // oracle/stubs: TIMER_* / PRINT / WARNING as the reference's headers expect them

namespace kref {
std::vector<std::string>& log() { static std::vector<std::string> l; return l; }
std::deque<std::vector<double>>& random_queue() { static std::deque<std::vector<double>> q; return q; }
}
extern "C" void dgels_(char*, int*, int*, int*, double*, int*, double*, int*, double*, int*, int* info) { *info = -1; }
extern "C" void dgesv_(int*, int*, double*, int*, int*, double*, int*, int* info) { *info = -1; }

#include "GMRESSolver.H"
#include "IDRSolver.H"
#include "thcm_model.hpp"

using thcm_b200::Ocean;
using thcm_b200::Vector;
using VecPtr = std::shared_ptr<Vector>;

struct Pars {
    std::map<std::string, double> v;
    template <typename T> T get(const char* name, T def) { auto it = v.find(name); return it == v.end() ? def : (T)it->second; }
};

static std::vector<double> parse_hist(const char* key) {
    std::vector<double> h;
    for (auto& s : kref::log()) {
        size_t p = s.find(key);
        if (p != std::string::npos) h.push_back(std::strtod(s.c_str() + p + std::strlen(key), nullptr));
    }
    return h;
}
static void print_vec(const char* name, const std::vector<double>& v, bool last = false) {
    printf("\"%s\": [", name);
    for (size_t i = 0; i < v.size(); i++) printf("%s%.17g", i ? ", " : "", v[i]);
    printf("]%s", last ? "" : ", ");
}

int main(int argc, char** argv) {
    if (argc < 2) { fprintf(stderr, "usage: drop_in_krylov <mask_natl8 file>\n"); return 2; }
    const double PI = 3.14159265358979323846;   // THCMdefs.H:19; degrees are converted as value * PI_ / 180.0 (THCM.C:203-206)
    // test/ocean/ocean_params.xml: 8 x 8 x 4 North Atlantic box, non-periodic
    int N = 8, M = 8, L = 4, periodic = 0, itopo = 0, flat = 0, rd_mask = 1, TRES = 1, SRES = 1, iza = 2, ite = 1, its = 1, rd_spertm = 0,
        cT = 0, cS = 0, ftype = 0;
    double xmin = 286 * PI / 180.0, xmax = 350 * PI / 180.0, ymin = 10 * PI / 180.0, ymax = 74 * PI / 180.0, hdim = 4000.0, qz = 1.0;
    __m_global_MOD_initialize(&N, &M, &L, &xmin, &xmax, &ymin, &ymax, &hdim, &qz, &periodic, &itopo, &flat, &rd_mask, &TRES, &SRES, &iza,
                              &ite, &its, &rd_spertm, &cT, &cS, &ftype, argv[1], "", "", "", "");
    std::vector<int> landm((size_t)(N + 2) * (M + 2) * (L + 2));
    __m_global_MOD_get_landm(landm.data());

    thcmb_settings s;
    thcmb_default_settings(&s);
    s.N = N; s.M = M; s.L = L; s.xmin = xmin; s.xmax = xmax; s.ymin = ymin; s.ymax = ymax; s.hdim = hdim; s.qz = qz; s.periodic = periodic;
    Ocean::SolverParameters sp;
    sp.tol = 1e-8; sp.maxit = 60; sp.restart = 30; sp.precon = 1;
    Ocean ocean(s, landm.data(), sp);
    ocean.setPar("Combined Forcing", 1.0); ocean.setPar("Wind Forcing", 1.0); ocean.setPar("Temperature Forcing", 10.0);
    ocean.setPar("Salinity Forcing", 1.0); ocean.setPar("NLES", 1.0);
    const int n = ocean.getTHCM().ndim();
    // deterministic state (the pytest side rebuilds it with the same formula); zero on land like any state of the solver
    std::vector<double> x((size_t)n);
    for (int i = 0; i < n; i++) {
        const int cell = i / 6, ci = cell % N, cj = (cell / N) % M, ck = cell / (N * M);
        const bool land = landm[(size_t)(ci + 1) + (size_t)(N + 2) * ((cj + 1) + (size_t)(M + 2) * (ck + 1))] != 0;
        // exactly representable arithmetic only (numpy's SIMD sin and glibc's differ in the last bit)
        x[(size_t)i] = land ? 0.0 : 0.05 * ((double)((i * 37) % 101) / 101.0 - 0.5);
    }
    ocean.getState('V')->fromHost(x.data());
    ocean.computeRHS();
    ocean.computeJacobian();
    ocean.buildPreconditioner();
    VecPtr b = ocean.getRHS('C');
    const double normb = b->norm();

    // ---- the reference's GMRES template over the device model ----
    VecPtr xg = std::make_shared<Vector>(ocean.context());
    GMRESSolver<Ocean, VecPtr> gmres(ocean);
    gmres.setSolution(xg); gmres.setRHS(b);
    auto gp = std::make_shared<Pars>();
    gp->v["GMRES tolerance"] = sp.tol; gp->v["GMRES iterations"] = sp.maxit; gp->v["GMRES restart"] = sp.restart;
    gp->v["GMRES verbosity"] = 8; gp->v["GMRES preconditioning"] = 1; gp->v["GMRES left prec"] = 0; gp->v["GMRES flexible"] = 1;
    gmres.setParameters(gp);
    kref::log().clear();
    gmres.solve();
    std::vector<double> hist_ref = parse_hist("impl res: ");
    hist_ref.push_back(gmres.residual());
    const int iters_ref = gmres.getNumIters();
    Vector r(*b), Ax(*b);
    ocean.applyMatrix(*gmres.getSolution(), Ax);
    r.update(-1.0, Ax, 1.0);
    const double true_res_ref = r.norm() / normb;

    // ---- the in-library GMRES (thcmb_gmres) on the same system ----
    VecPtr xl = std::make_shared<Vector>(ocean.context());
    std::vector<double> hist_lib(4096);
    thcmb_krylov_result res;
    thcmb_gmres(ocean.context(), b->data(), xl->data(), sp.tol, sp.maxit, sp.restart, 1 | 4, hist_lib.data(), 4096, &res);
    hist_lib.resize((size_t)res.nhist);
    Vector dsol(*xl);
    dsol.update(-1.0, *gmres.getSolution(), 1.0);
    const double sol_diff = dsol.norm() / (xl->norm() + 1e-300);

    // ---- the reference's IDR(s) template over the device model ----
    VecPtr xi = std::make_shared<Vector>(ocean.context());
    IDRSolver<Ocean, VecPtr> idr(ocean, xi, b);
    auto ip = std::make_shared<Pars>();
    ip->v["IDR s"] = 4; ip->v["IDR tolerance"] = 1e-6; ip->v["IDR iterations"] = 40; ip->v["IDR save search space"] = 0; ip->v["IDR verbosity"] = 0;
    idr.setParameters(ip);
    idr.solve();
    Vector ri(*b);
    ocean.applyMatrix(*idr.getSolution(), Ax);
    ri.update(-1.0, Ax, 1.0);
    const double true_res_idr = ri.norm() / normb;

    printf("{\"n\": %d, \"normb\": %.17g, \"iters_ref\": %d, \"iters_lib\": %d, \"true_res_ref\": %.17g, \"sol_diff\": %.17g, "
           "\"true_res_idr\": %.17g, ", n, normb, iters_ref, res.iters, true_res_ref, sol_diff, true_res_idr);
    std::vector<double> bh = b->toHost();
    print_vec("rhs", bh);
    print_vec("hist_ref", hist_ref);
    print_vec("hist_lib", hist_lib, true);
    printf("}\n");
    return 0;
}
