// =============================================================================
// thcm_model.hpp -- C++ host-side mirror of the reference's interface for the Newton-step path, over the C ABI of
// libthcm_b200.so (include/thcm_b200.h).  Header-only, C++14, no CUDA or Trilinos headers needed.
//
// The reference's host code for this path is C++ (src/ocean/THCM.C, src/ocean/Ocean.C, src/utils/Model.H,
// src/transient/ThetaModel.H) on top of Epetra vectors.  Trilinos is not available here, so the mirror keeps the
// reference's class and method names, argument meaning and error behaviour and replaces Epetra_Vector by a
// device-resident Vector:
//
//   thcm_b200::Vector        the Vector concept of src/gmressolver/GMRESSolverDecl.H:12-16 and
//                            src/idrsolver/IDRSolverDecl.H:12-16 (dot, norm, update, scale, zero, random, copy) --
//                            every operation is one kernel of the library, data never leaves HBM.  The reference's
//                            UNMODIFIED GMRESSolver<Model, VectorPointer> / IDRSolver<Model, VectorPointer> templates
//                            instantiate over it (tests/cpp/drop_in_krylov.cpp does exactly that).
//   thcm_b200::THCM          src/ocean/THCM.H:76-330: evaluate(), evaluateB(), setParameter / getParameter by XML name,
//                            RecomputeScaling, getIntCondCoeff, fixMixing, the coupling setters of THCM.C:1395-1560.
//   thcm_b200::Ocean         the Model API (src/utils/Model.H:54-117 as implemented by Ocean.C:1070-1391): getState /
//                            getRHS / getSolution ('V' view, 'C' copy), computeRHS, computeJacobian, computeMassMat,
//                            applyMatrix, applyMassMat, buildPreconditioner, applyPrecon, solve, setPar / getPar.
//   thcm_b200::ThetaModel<M> src/transient/ThetaModel.H:18-165.
//
// Errors: the reference aborts through ERROR(...) (GlobalDefinitions.H:79-96); the mirror throws std::runtime_error with
// the same message style, and the library itself calls thcm_throw_error_ for anything below the ABI.
// =============================================================================
#ifndef THCM_MODEL_HPP
#define THCM_MODEL_HPP

#include <cmath>
#include <cstddef>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "thcm_b200.h"

namespace thcm_b200 {

// ---------------------------------------------------------------------------------------------------------------------
// Device vector of the owned unknowns of one rank (standard-map ordering, 6 interleaved unknowns per cell)
// ---------------------------------------------------------------------------------------------------------------------
class Vector {
    thcmb_ctx* c_ = nullptr;
    double* d_ = nullptr;
    int n_ = 0;
    void alloc(thcmb_ctx* c, int n) {
        release();
        c_ = c; n_ = n;
        if (c && n > 0) {
            d_ = static_cast<double*>(thcmb_device_alloc(c, (long long)n * (long long)sizeof(double)));
            if (!d_) throw std::runtime_error("thcm_b200::Vector: device allocation failed");
        }
    }
    void release() { if (d_) thcmb_device_free(c_, d_); d_ = nullptr; n_ = 0; }

public:
    Vector() {}
    explicit Vector(thcmb_ctx* c) { alloc(c, thcmb_ndim_local(c)); zero(); }
    Vector(const Vector& o) { *this = o; }
    Vector(Vector&& o) noexcept : c_(o.c_), d_(o.d_), n_(o.n_) { o.d_ = nullptr; o.n_ = 0; }
    Vector& operator=(const Vector& o) {
        if (this == &o) return *this;
        if (!o.d_) { release(); c_ = o.c_; return *this; }
        if (n_ != o.n_ || c_ != o.c_) alloc(o.c_, o.n_);
        thcmb_axpby(c_, n_, 1.0, o.d_, 0.0, d_);          // this = 1*o + 0*this
        return *this;
    }
    Vector& operator=(Vector&& o) noexcept {
        if (this != &o) { release(); c_ = o.c_; d_ = o.d_; n_ = o.n_; o.d_ = nullptr; o.n_ = 0; }
        return *this;
    }
    ~Vector() { release(); }

    int length() const { return n_; }
    // the reference's templates hand default-constructed vectors to applyMatrix / applyPrecon as outputs (GMRESSolver.H:131-132)
    void ensure(const Vector& like) { if (n_ != like.n_ || c_ != like.c_) alloc(like.c_, like.n_); }
    double* data() { return d_; }
    const double* data() const { return d_; }
    thcmb_ctx* context() const { return c_; }

    // Vector concept of the reference's Krylov templates (GMRESSolverDecl.H:12-16)
    double dot(const Vector& o) const { return thcmb_dot(c_, n_, d_, o.d_); }      // all-reduced over the ranks
    double norm() const { return thcmb_nrm2(c_, n_, d_); }
    void update(double scalarA, const Vector& A, double scalarThis) { thcmb_axpby(c_, n_, scalarA, A.d_, scalarThis, d_); }
    void scale(double a) { thcmb_scale(c_, n_, a, d_); }
    void zero() { fill(0.0); }   // (a scale by 0 would keep NaNs of uninitialised memory)
    void fill(double a) {
        std::vector<double> h((size_t)n_, a);
        fromHost(h.data());
    }
    // Vector::random() of the reference seeds Epetra's generator; here a reproducible host sequence (the shadow space of
    // IDR(s) only has to be generic).  setRandomSource lets a caller inject the vectors (parity tests).
    void random() {
        std::vector<double> h((size_t)n_);
        auto& q = randomSource();
        if (!q.empty()) { h = q.front(); q.erase(q.begin()); h.resize((size_t)n_, 0.0); }
        else {
            static unsigned long long state = 0x9E3779B97F4A7C15ull;
            for (int i = 0; i < n_; i++) {
                state = state * 6364136223846793005ull + 1442695040888963407ull;
                h[(size_t)i] = (double)((state >> 11) & ((1ull << 53) - 1)) / (double)(1ull << 52) - 1.0;
            }
        }
        fromHost(h.data());
    }
    static std::vector<std::vector<double>>& randomSource() { static std::vector<std::vector<double>> q; return q; }
    void print() const {}

    void fromHost(const double* h) { thcmb_h2d(c_, d_, h, (long long)n_ * 8); thcmb_sync(c_); }
    void toHost(double* h) const { thcmb_d2h(c_, h, d_, (long long)n_ * 8); thcmb_sync(c_); }
    std::vector<double> toHost() const { std::vector<double> h((size_t)n_); toHost(h.data()); return h; }
};
using VectorPtr = std::shared_ptr<Vector>;

// ---------------------------------------------------------------------------------------------------------------------
// THCM (src/ocean/THCM.H): owner of the library context of this rank
// ---------------------------------------------------------------------------------------------------------------------
class THCM {
    thcmb_ctx* c_ = nullptr;

public:
    // THCM::par2int (THCM.C:1841-1890): XML parameter names -> par.F90 indices
    static int par2int(const std::string& label) {
        static const std::map<std::string, int> tbl = {
            {"Time", 0}, {"AL_T", 1}, {"Rayleigh-Number", 2}, {"Vertical Ekman-Number", 3}, {"Horizontal Ekman-Number", 4},
            {"Rossby-Number", 5}, {"MIXP", 6}, {"RESC", 7}, {"SPL1", 8}, {"Salinity Homotopy", 9}, {"Solar Forcing", 10},
            {"Horizontal Peclet-Number", 11}, {"Vertical Peclet-Number", 12}, {"P_VC", 13}, {"LAMB", 14}, {"Salinity Forcing", 15},
            {"Wind Forcing", 16}, {"Temperature Forcing", 17}, {"Nonlinear Factor", 18}, {"Combined Forcing", 19}, {"ARCL", 20},
            {"NLES", 21}, {"CMPR", 26}, {"ALPC", 25}, {"Energy", 24}, {"Flux Perturbation", 27}, {"MKAP", 29}, {"SPL2", 30},
            {"IFRICB", 22}, {"CONT", 23}, {"Salinity Perturbation", 28}};
        auto it = tbl.find(label);
        return it == tbl.end() ? -1 : it->second;     // -1 like the reference
    }

    THCM(const thcmb_settings& s, const int* landm_global) {
        c_ = thcmb_create(&s, landm_global);
        if (!c_) throw std::runtime_error(std::string("THCM: thcmb_create failed: ") + thcmb_last_error());
    }
    THCM(const THCM&) = delete;
    THCM& operator=(const THCM&) = delete;
    ~THCM() { if (c_) thcmb_destroy(c_); }

    thcmb_ctx* context() const { return c_; }
    int ndim() const { return thcmb_ndim_local(c_); }

    // THCM::evaluate (THCM.C:957-1199): rhs <- F(soln) (C++ sign), Jacobian values into the static graph
    bool evaluate(const Vector& soln, Vector* rhs, bool computeJac) {
        if (rhs) thcmb_residual_dev(c_, soln.data(), rhs->data());
        if (computeJac) thcmb_jacobian_dev(c_, soln.data());
        thcmb_sync(c_);
        return true;
    }
    // THCM::evaluateB (THCM.C:1202-1230): the diagonal of the mass matrix (host copy, coB of assemble.F90:18-54)
    std::vector<double> evaluateB() { std::vector<double> b((size_t)ndim()); thcmb_get_cob(c_, b.data()); return b; }
    bool setParameter(const std::string& label, double value) {                    // THCM.C:1945-1957
        const int param = par2int(label);
        if (param < 0 || param > 30) throw std::runtime_error("THCM::setParameter: invalid parameter '" + label + "'");
        thcmb_set_par(c_, param, value);
        return true;
    }
    bool getParameter(const std::string& label, double& value) {                   // THCM.C:1960-1971
        const int param = par2int(label);
        if (param < 0 || param > 30) throw std::runtime_error("THCM::getParameter: invalid parameter '" + label + "'");
        value = thcmb_get_par(c_, param);
        return true;
    }
    void applyMatrix(const Vector& v, Vector& out) { out.ensure(v); thcmb_spmv_dev(c_, v.data(), out.data()); }
    // THCM::RecomputeScaling (THCM.C:1781-1834), getIntCondCoeff (THCM.C:2608-2637), fixMixing (THCM.C:2639-2647)
    int RecomputeScaling(std::vector<double>& rowScaling, std::vector<double>& colScaling) {
        rowScaling.resize((size_t)ndim()); colScaling.resize((size_t)ndim());
        return thcmb_recompute_scaling(c_, rowScaling.data(), colScaling.data(), nullptr);
    }
    // coefficients on ALL owned unknowns (non-zero on the S rows of ocean cells), like the Epetra vector intcondCoeff_; returns the volume
    double getIntCondCoeff(std::vector<double>& coeff) { coeff.resize((size_t)ndim()); return thcmb_intcond_coeff(c_, coeff.data()); }
    void fixMixing(int value) { thcmb_set_vmix_fix(c_, value); }
    // THCM::setLandMask(global mask, init) (THCM.C:1362-1392): with init the instance follows the new GLOBAL mask (set_landmask_, reinit = 1)
    void setLandMask(const std::vector<int>& landmGlobal, bool init = true) {
        if (thcmb_set_landmask(c_, landmGlobal.data(), init ? 1 : 0) != 0) throw std::runtime_error(thcmb_last_error());
    }
    // coupling setters (THCM.C:1395-1560 -> m_inserts); `which` as in thcmb_insert_field
    void setSurfaceField(int which, const std::vector<double>& fieldGlobal) { thcmb_insert_field(c_, which, fieldGlobal.data()); }
    void setAtmosphereParameters(const double* commPars18) { thcmb_set_atmos_parameters(c_, commPars18); }
    void setSeaIceParameters(const double* commPars7) { thcmb_set_seaice_parameters(c_, commPars7); }
};

// ---------------------------------------------------------------------------------------------------------------------
// Ocean: the Model API (src/utils/Model.H:54-117) as Ocean.C implements it
// ---------------------------------------------------------------------------------------------------------------------
struct SolverParameters {   // run/ocean/solver_params.xml: FGMRES tolerance 1e-4, 500 iterations, no restarts
    double tol = 1e-4; int maxit = 500; int restart = 400; int precon = 1; bool dgks = false;
};

class Ocean {
public:
    using VectorPtr = thcm_b200::VectorPtr;
    using ConstVectorPtr = std::shared_ptr<const Vector>;
    using SolverParameters = thcm_b200::SolverParameters;

protected:
    std::shared_ptr<THCM> thcm_;
    VectorPtr state_, rhs_, sol_;
    std::vector<double> massMat_;
    SolverParameters sp_;
    bool precInitialized_ = false;
    thcmb_krylov_result lastSolve_{};

    VectorPtr getVector(char mode, const VectorPtr& v) const {   // Utils::getVector: 'V' view, 'C' copy
        if (mode == 'V') return v;
        if (mode == 'C') return std::make_shared<Vector>(*v);
        throw std::runtime_error("Ocean::getVector: invalid mode");
    }

public:
    Ocean(const thcmb_settings& s, const int* landm_global, SolverParameters sp = SolverParameters())
        : thcm_(std::make_shared<THCM>(s, landm_global)), sp_(sp) {
        state_ = std::make_shared<Vector>(thcm_->context());
        rhs_ = std::make_shared<Vector>(thcm_->context());
        sol_ = std::make_shared<Vector>(thcm_->context());
    }
    // over a THCM that exists already (e.g. thcm_b200::makeTHCM of thcm_paramlist.hpp)
    explicit Ocean(std::shared_ptr<THCM> thcm, SolverParameters sp = SolverParameters()) : thcm_(std::move(thcm)), sp_(sp) {
        state_ = std::make_shared<Vector>(thcm_->context());
        rhs_ = std::make_shared<Vector>(thcm_->context());
        sol_ = std::make_shared<Vector>(thcm_->context());
    }
    virtual ~Ocean() {}

    THCM& getTHCM() { return *thcm_; }
    thcmb_ctx* context() const { return thcm_->context(); }
    VectorPtr getState(char mode = 'C') { return getVector(mode, state_); }
    VectorPtr getRHS(char mode = 'C') { return getVector(mode, rhs_); }
    VectorPtr getSolution(char mode = 'C') { return getVector(mode, sol_); }
    const std::vector<double>& getMassMat() const { return massMat_; }

    virtual void setPar(const std::string& name, double value) { thcm_->setParameter(name, value); precInitialized_ = false; }
    virtual double getPar(const std::string& name) { double v = 0.0; thcm_->getParameter(name, v); return v; }

    virtual void preProcess() {}
    virtual void computeRHS() { thcm_->evaluate(*state_, rhs_.get(), false); }                           // Ocean.C:1277-1295
    virtual void computeJacobian() { thcm_->evaluate(*state_, nullptr, true); precInitialized_ = false; } // Ocean.C:1297-1309
    virtual void computeMassMat() { massMat_ = thcm_->evaluateB(); }
    virtual void applyMatrix(const Vector& v, Vector& out) { thcm_->applyMatrix(v, out); }               // Ocean.C:1369-1374
    virtual void applyMassMat(const Vector& v, Vector& out) {                                            // diagonal B
        out.ensure(v);
        if (thcmb_apply_mass_dev(context(), v.data(), out.data()) != 0) throw std::runtime_error(thcmb_last_error());
    }
    virtual void buildPreconditioner() {                                                                 // Ocean.C:1377-1391
        if (!precInitialized_) { thcmb_build_precon(context(), sp_.precon); precInitialized_ = true; }
    }
    virtual void applyPrecon(const Vector& v, Vector& out) {
        buildPreconditioner();
        out.ensure(v);
        thcmb_apply_precon_dev(context(), v.data(), out.data());
    }
    // Ocean::solve (Ocean.C:1070-1147): sol_ = J^{-1} rhs with right-preconditioned flexible GMRES, zero initial guess
    virtual int solve(ConstVectorPtr rhs = ConstVectorPtr()) {
        const Vector& b = rhs ? *rhs : *rhs_;
        buildPreconditioner();
        sol_->zero();
        const int flags = (sp_.precon != 0 ? 1 : 0) | 4 | (sp_.dgks ? 8 : 0);
        thcmb_gmres(context(), b.data(), sol_->data(), sp_.tol, sp_.maxit, sp_.restart, flags, nullptr, 0, &lastSolve_);
        return lastSolve_.status;
    }
    const thcmb_krylov_result& lastSolve() const { return lastSolve_; }

    // Utils::CRSMat of the reference (src/utils/Utils.H): beg / jco / co, 0-based
    struct CRSMat { std::vector<int> beg, jco; std::vector<double> co; };
    // Ocean::getBlock(std::shared_ptr<Atmosphere>) (Ocean.C:1603-1730): d F_ocean / d x_atmosphere.  AtmosLike: interface_row(i, j, XX)
    // (Atmosphere.C:1746-1763), commParsDa() (Atmosphere::CommPars::da), pdist() (n*m doubles or nullptr).  N, M: the global surface grid
    template <class AtmosLike> CRSMat getBlockAtmosphere(AtmosLike& atmos, int N, int M) {
        std::vector<int> cT((size_t)N * M), cQ(cT.size()), cA(cT.size()), cP(cT.size());
        for (int j = 0; j < M; j++) for (int i = 0; i < N; i++) {
            const size_t sr = (size_t)j * N + i;
            cT[sr] = atmos.interface_row(i, j, 1); cQ[sr] = atmos.interface_row(i, j, 2);     // ATMOS_TT_, ATMOS_QQ_
            cA[sr] = atmos.interface_row(i, j, 3); cP[sr] = atmos.interface_row(i, j, 4);     // ATMOS_AA_, ATMOS_PP_
        }
        CRSMat b;
        b.beg.resize((size_t)thcmb_ndim_local(context()) + 1); b.jco.resize(6 * cT.size()); b.co.resize(6 * cT.size());
        const int nnz = thcmb_ocean_block_atmosphere(context(), atmos.commParsDa(), atmos.pdist(), cT.data(), cQ.data(), cA.data(), cP.data(),
                                                     b.beg.data(), b.jco.data(), b.co.data());
        b.jco.resize((size_t)nnz); b.co.resize((size_t)nnz);
        return b;
    }
    // Ocean::getBlock(std::shared_ptr<SeaIce>) (Ocean.C:1733-1810): d F_ocean / d x_seaice at the current state
    template <class SeaIceLike> CRSMat getBlockSeaIce(SeaIceLike& seaice, int N, int M) {
        std::vector<int> cQ((size_t)N * M), cM(cQ.size()), cG(cQ.size());
        for (int j = 0; j < M; j++) for (int i = 0; i < N; i++) {
            const size_t sr = (size_t)j * N + i;
            cQ[sr] = seaice.interface_row(i, j, 2); cM[sr] = seaice.interface_row(i, j, 3); cG[sr] = seaice.interface_row(i, j, 5);   // SEAICE_QQ_, _MM_, _GG_
        }
        const std::vector<double> un = state_->toHost();
        CRSMat b;
        b.beg.resize(un.size() + 1); b.jco.resize(6 * cQ.size()); b.co.resize(6 * cQ.size());
        const int nnz = thcmb_ocean_block_seaice(context(), un.data(), cQ.data(), cM.data(), cG.data(), b.beg.data(), b.jco.data(), b.co.data());
        b.jco.resize((size_t)nnz); b.co.resize((size_t)nnz);
        return b;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// ThetaModel (src/transient/ThetaModel.H:18-165): M u_n + dt theta F(u_{n+1}) + dt (1-theta) F(u_n) - M u_{n+1} = 0
// ---------------------------------------------------------------------------------------------------------------------
template <typename Model>
class ThetaModel : public Model {
    double theta_, timestep_ = 1.0e-3;
    VectorPtr oldState_, oldRhs_;

public:
    template <typename... Args>
    explicit ThetaModel(double theta, Args&&... args) : Model(std::forward<Args>(args)...), theta_(theta) {
        oldState_ = Model::getState('C');
        oldRhs_ = Model::getRHS('C');
    }
    void initStep(double timestep) {                    // ThetaModel.H:65-75
        timestep_ = timestep;
        *oldState_ = *Model::getState('V');
        Model::preProcess();
        Model::computeRHS();
        *oldRhs_ = *Model::getRHS('V');
    }
    void setState(const VectorPtr& state) { if (Model::getState('V') != state) *Model::getState('V') = *state; }
    void computeRHS() override {                        // ThetaModel.H:87-113
        if (theta_ < 0 || theta_ > 1) throw std::runtime_error("ThetaModel: Incorrect theta");
        Model::computeRHS();
        thcmb_theta_rhs_dev(Model::context(), theta_, timestep_, Model::getState('V')->data(), oldState_->data(), oldRhs_->data(),
                            Model::getRHS('V')->data());
        thcmb_sync(Model::context());
    }
    void computeJacobian() override {                   // ThetaModel.H:118-149
        Model::computeJacobian();
        thcmb_theta_jacobian_dev(Model::context(), theta_, timestep_);
    }
    int solve(typename Model::ConstVectorPtr rhs = typename Model::ConstVectorPtr()) override {   // ThetaModel.H:153-165
        if (theta_ == 0.0) throw std::runtime_error("ThetaModel: theta = 0 divides by the singular mass matrix of THCM");
        auto b = std::make_shared<Vector>(rhs ? *rhs : *Model::getRHS('V'));
        b->scale(1.0 / timestep_ / theta_);
        return Model::solve(b);
    }
};

}  // namespace thcm_b200

#endif  // THCM_MODEL_HPP
