// =============================================================================
// krylov_ref.cpp -- ORACLE HARNESS (test infrastructure, NOT product code)
//
// Instantiates the reference's own, UNMODIFIED Krylov templates
//   /root/reference/src/gmressolver/GMRESSolver.H   (GMRESSolver<Model,VectorPointer>)
//   /root/reference/src/idrsolver/IDRSolver.H       (IDRSolver<Model,VectorPointer>)
// against a plain host vector and a CSR operator, and exposes them through a
// small C ABI so pytest can compare the CUDA GMRES / IDR(s) residual histories
// with the reference's.  Built by oracle/Makefile into oracle/_ref/ (git-ignored,
// travels to the GPU box prebuilt).  No reference source is copied: the headers
// are included from where they lie (-I/root/reference/src/...).
// =============================================================================
#include <cstdio>
#include <cstring>
#include <memory>
#include <map>
#include <deque>
#include <string>
#include <vector>
#include <cmath>
#include <iostream>
#include <sstream>

#include "GlobalDefinitions.H"  // oracle/stubs

namespace kref {
std::vector<std::string>& log() { static std::vector<std::string> l; return l; }
std::deque<std::vector<double>>& random_queue() { static std::deque<std::vector<double>> q; return q; }
}

// LAPACK symbols the GMRES header declares; only reached with minimiser scheme 'Q'
// (GMRESSolver.H:71,450) which this harness never selects.
extern "C" void dgels_(char*, int*, int*, int*, double*, int*, double*, int*, double*, int*, int* info) { *info = -1; }
extern "C" void dgesv_(int*, int*, double*, int*, int*, double*, int*, int* info) { *info = -1; }

#include "GMRESSolver.H"
#include "IDRSolver.H"

namespace {

// Host vector with the interface the templates need (GMRESSolverDecl.H:12-16, IDRSolverDecl.H:12-16)
struct Vec {
    std::vector<double> d;
    Vec() {}
    explicit Vec(size_t n) : d(n, 0.0) {}
    double dot(Vec const& o) const { double s = 0.0; for (size_t i = 0; i < d.size(); i++) s += d[i] * o.d[i]; return s; }
    double norm() const { return std::sqrt(dot(*this)); }
    void update(double a, Vec const& A, double b) { for (size_t i = 0; i < d.size(); i++) d[i] = a * A.d[i] + b * d[i]; }
    void scale(double a) { for (size_t i = 0; i < d.size(); i++) d[i] = a * d[i]; }
    void zero() { std::fill(d.begin(), d.end(), 0.0); }
    void random() {  // shadow-space vectors are injected by the caller (IDRSolver.H:84-104)
        auto& q = kref::random_queue();
        if (q.empty()) { for (size_t i = 0; i < d.size(); i++) d[i] = std::sin(1.0 + 0.37 * (double)i); return; }
        d = q.front(); q.pop_front();
    }
    void print() const {}
};

struct CsrModel {
    int n; const int* rowptr; const int* col; const double* val;
    int prec_kind;            // 0 identity, 1 block-diagonal (nb x nb dense inverse blocks, row-major)
    int nb; const double* minv;
    long n_matvec = 0, n_prec = 0;
    void applyMatrix(Vec const& v, Vec& out) {
        n_matvec++;
        if (out.d.size() != (size_t)n) out.d.resize(n);
        for (int i = 0; i < n; i++) { double s = 0.0; for (int q = rowptr[i]; q < rowptr[i + 1]; q++) s += val[q] * v.d[col[q]]; out.d[i] = s; }
    }
    void applyPrecon(Vec const& v, Vec& out) {
        n_prec++;
        if (out.d.size() != (size_t)n) out.d.resize(n);
        if (prec_kind == 0) { out.d = v.d; return; }
        for (int c = 0; c < n / nb; c++) {
            const double* M = minv + (size_t)c * nb * nb;
            for (int r = 0; r < nb; r++) { double s = 0.0; for (int q = 0; q < nb; q++) s += M[r * nb + q] * v.d[c * nb + q]; out.d[c * nb + r] = s; }
        }
    }
};

struct Pars {
    std::map<std::string, double> v;
    template <typename T> T get(const char* name, T def) { auto it = v.find(name); return it == v.end() ? def : (T)it->second; }
};

}  // namespace

extern "C" {

// Runs GMRESSolver::solve() (GMRESSolver.H:81-255).  hist receives resid_ after every inner
// iteration (parsed from the solver's own status prints at verbosity 8), nhist its length.
// flags: bit0 = use preconditioner, bit1 = left prec, bit2 = flexible.
int kref_gmres(int n, const int* rowptr, const int* col, const double* val, int prec_kind, int nb, const double* minv,
               const double* b, double* x, double tol, int maxit, int restart, int flags,
               double* hist, int hist_cap, int* nhist, int* iters, double* final_resid, long* n_matvec) {
    CsrModel model{n, rowptr, col, val, prec_kind, nb, minv};
    auto xs = std::make_shared<Vec>(n); auto bs = std::make_shared<Vec>(n);
    std::memcpy(xs->d.data(), x, sizeof(double) * n); std::memcpy(bs->d.data(), b, sizeof(double) * n);
    GMRESSolver<CsrModel, std::shared_ptr<Vec>> solver(model);
    solver.setSolution(xs); solver.setRHS(bs);
    auto pars = std::make_shared<Pars>();
    pars->v["GMRES tolerance"] = tol; pars->v["GMRES iterations"] = maxit; pars->v["GMRES restart"] = restart;
    pars->v["GMRES verbosity"] = 8; pars->v["GMRES preconditioning"] = (flags & 1) ? 1 : 0;
    pars->v["GMRES left prec"] = (flags & 2) ? 1 : 0; pars->v["GMRES flexible"] = (flags & 4) ? 1 : 0;
    solver.setParameters(pars);
    kref::log().clear();
    int rc = solver.solve();
    // "iteration: K impl res: R expl res: E" is printed at the START of inner iteration K with the residual
    // reached by iteration K-1 (GMRESSolver.H:148-149); the last one comes from solver.residual().
    int nh = 0;
    for (auto& s : kref::log()) {
        size_t p = s.find("impl res: ");
        if (p == std::string::npos) continue;
        double r = std::strtod(s.c_str() + p + 10, nullptr);
        if (nh < hist_cap) hist[nh] = r;
        nh++;
    }
    if (nh < hist_cap) hist[nh] = solver.residual();
    nh++;
    *nhist = nh < hist_cap ? nh : hist_cap;
    *iters = solver.getNumIters(); *final_resid = solver.residual(); *n_matvec = model.n_matvec;
    std::memcpy(x, solver.getSolution()->d.data(), sizeof(double) * n);
    return rc;
}

// Runs IDRSolver::solve() (IDRSolver.H:109-340).  P_raw: s vectors (row-major s x n) handed to
// Vector::random() in createP (IDRSolver.H:84-104).  hist = normr_ sequence (resvec_), parsed from
// the solver's own per-iteration status line (verbosity 5) with cout at 17 digits.
int kref_idrs(int n, const int* rowptr, const int* col, const double* val, int prec_kind, int nb, const double* minv,
              const double* b, double* x, double tol, int maxit, int s, const double* P_raw,
              double* hist, int hist_cap, int* nhist, int* iters, double* final_resid, long* n_matvec) {
    CsrModel model{n, rowptr, col, val, prec_kind, nb, minv};
    auto xs = std::make_shared<Vec>(n); auto bs = std::make_shared<Vec>(n);
    std::memcpy(xs->d.data(), x, sizeof(double) * n); std::memcpy(bs->d.data(), b, sizeof(double) * n);
    kref::random_queue().clear();
    for (int j = 0; j < s; j++) kref::random_queue().push_back(std::vector<double>(P_raw + (size_t)j * n, P_raw + (size_t)(j + 1) * n));
    IDRSolver<CsrModel, std::shared_ptr<Vec>> solver(model, xs, bs);
    auto pars = std::make_shared<Pars>();
    pars->v["IDR s"] = s; pars->v["IDR tolerance"] = tol; pars->v["IDR iterations"] = maxit;
    pars->v["IDR save search space"] = 0; pars->v["IDR verbosity"] = 5;
    solver.setParameters(pars);
    std::ostringstream cap;
    std::streambuf* old = std::cout.rdbuf(cap.rdbuf());
    auto oldprec = std::cout.precision(17);
    int rc = solver.solve();
    std::cout.rdbuf(old); std::cout.precision(oldprec);
    std::istringstream in(cap.str());
    std::string line; int nh = 0;
    while (std::getline(in, line)) {
        size_t p = line.find("residual: ");
        if (p == std::string::npos || line.find("iteration: ") == std::string::npos) continue;
        double r = std::strtod(line.c_str() + p + 10, nullptr);
        if (nh < hist_cap) hist[nh] = r;
        nh++;
    }
    *nhist = nh < hist_cap ? nh : hist_cap;
    *iters = solver.getNumIters(); *final_resid = solver.implicitResNorm(); *n_matvec = model.n_matvec;
    std::memcpy(x, solver.getSolution()->d.data(), sizeof(double) * n);
    return rc;
}

}  // extern "C"
