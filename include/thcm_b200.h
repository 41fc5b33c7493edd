/* =============================================================================
 * thcm_b200.h -- C ABI of libthcm_b200.so, the B200-native THCM Newton-step path.
 *
 * Two layers:
 *
 *  B1  The gfortran-mangled symbols that the reference's src/ocean/THCM.C binds
 *      (extern "C" list THCM.C:49-176, name mangling src/utils/my_f2c.H:15-17).
 *      Same names, same by-pointer arguments, same semantics (caller-owned CRS
 *      buffers, no return codes, one global instance per process).  They take
 *      HOST pointers; host<->device copies happen inside.
 *
 *  B1' A handle-based device API (thcmb_*) so that residual, Jacobian, SpMV and
 *      the Krylov vectors never leave HBM.  All pointers named d_* are DEVICE
 *      pointers; everything else is host memory.
 *
 * All reals are IEEE FP64, all integers 32-bit.  Unknown ordering (matetc.F90:123):
 *   row = 6*((k-1)*n*m + (j-1)*n + (i-1)) + XX,  XX in {u=1,v,w,p,T,S=6}.
 * No CPU fallback exists: every compute entry point fails loudly (thcm_throw_error_
 * callback if installed, else abort()) when no CUDA device is usable.
 * ========================================================================== */
#ifndef THCM_B200_H
#define THCM_B200_H
#ifdef __cplusplus
extern "C" {
#endif

/* --------------------------------------------------------------------------
 * B1: reference Fortran symbols (host pointers)
 * ------------------------------------------------------------------------ */

/* replaces m_global::initialize, src/ocean/global.F90:65-157 (THCM.C:331-338).
 * Angles in radians.  Stores the settings and clears the mask; of the file names the land mask and the salinity perturbation mask
 * are used (data/ in the reference holds nothing else).  The mask itself is made by get_landm, like the reference's topofit
 * (topo.F90:6-38): rd_mask != 0 reads mkmask/<maskfile> (a path, or a name below $THCM_DATA_DIR/mkmask; readmask, topo.F90:41-127),
 * rd_mask == 0 builds the idealised continents of itopo = 1..4 (depth3land, topo.F90:129-330; itopo = 0 needs bathymetry data
 * the reference does not ship and fails loudly); flat copies the surface level down. */
void __m_global_MOD_initialize(int* N, int* M, int* L, double* xmin, double* xmax, double* ymin, double* ymax,
                               double* hdim, double* qz, int* periodic, int* itopo, int* flat, int* rd_mask,
                               int* TRES, int* SRES, int* iza, int* ite, int* its, int* rd_spertm,
                               int* coupled_T, int* coupled_S, int* forcing_type,
                               const char* maskfile, const char* spertmaskfile, const char* windfile,
                               const char* sstfile, const char* sssfile);
/* replaces m_global::get_landm (global.F90:299-318, THCM.C:389): runs topofit, then (N+2)(M+2)(L+2) ints, i fastest */
void __m_global_MOD_get_landm(int* landm);
void __m_global_MOD_finalize(void);

/* replaces SUBROUTINE init, src/ocean/usrc.F90:6-139 (called THCM.C:603-611) */
void init_(int* n, int* m, int* l, int* nmlglob, double* xmin, double* xmax, double* ymin, double* ymax,
           double* alphaT, double* alphaS, int* ih, int* vmix, int* tap, int* rho_mixing, int* coriolis_on,
           int* periodic, int* landm, double* taux, double* tauy, double* tatm, double* emip, double* spert);
/* replaces SUBROUTINE finalize, usrc.F90:143-160 */
void finalize_(void);
/* replaces m_mat::get_array_sizes / set_pointers, src/ocean/mat.F90:56-103 (THCM.C:619-638) */
void __m_mat_MOD_get_array_sizes(int* nrows, int* nnz);
void __m_mat_MOD_set_pointers(int* nrows, int* nnz, int* begA, int* jcoA, double* coA, double* coB,
                              int* begF, int* jcoF, double* coF);
/* replaces SUBROUTINE rhs(un,B), usrc.F90:523-603 (THCM.C:1001): B has the THCM sign, C++ negates */
void rhs_(double* un, double* B);
/* replaces SUBROUTINE matrix(un), usrc.F90:449-521 (THCM.C:1066): fills begA/jcoA/coA (1-based,
 * Fortran entry order, |a|>1e-10 thresholded) and coB through the borrowed pointers */
void matrix_(double* un);
/* replaces SUBROUTINE fillcolB, src/ocean/assemble.F90:18-54 (THCM.C:1209) */
void fillcolb_(void);
/* replaces setparcs/getparcs, usrc.F90:163-198: idx 1..30 (par.F90:38-67); set re-runs forcing+lin */
void setparcs_(int* idx, double* val);
void getparcs_(int* idx, double* val);
/* replaces setsres, usrc.F90:434-446 */
void setsres_(int* sres);
/* replaces set_landmask, usrc.F90:353-418 */
void set_landmask_(int* landm, int* periodic, int* reinit);
/* replaces get_forcing, src/ocean/forcing.F90:220-233 */
void get_forcing_(double* frc);
/* replace m_inserts::insert_*, src/ocean/inserts.F90:11-281 (THCM.C:85-98): n*m surface fields, i fastest; no recompute
 * until the next setparcs / set_*_parameters.  emip / adapted_emip / emip_pert are masked by the surface land mask
 * (inserts.F90:179,198,217); atmosphere_q / _a / _p and seaice_g are ignored unless the matching coupling flag is on
 * (inserts.F90:45,66,87,145) */
void __m_inserts_MOD_insert_taux(double* f);
void __m_inserts_MOD_insert_tauy(double* f);
void __m_inserts_MOD_insert_atmosphere_t(double* f);
void __m_inserts_MOD_insert_atmosphere_q(double* f);
void __m_inserts_MOD_insert_atmosphere_a(double* f);
void __m_inserts_MOD_insert_atmosphere_p(double* f);
void __m_inserts_MOD_insert_seaice_q(double* f);
void __m_inserts_MOD_insert_seaice_m(double* f);
void __m_inserts_MOD_insert_seaice_g(double* f);
void __m_inserts_MOD_insert_emip(double* f);
void __m_inserts_MOD_insert_adapted_emip(double* f);
void __m_inserts_MOD_insert_emip_pert(double* f);
/* replace set_atmos_parameters / set_seaice_parameters, src/ocean/usrc.F90:254-350 (Ocean.C:48-49, 1486, 1507): pars points
 * to Atmosphere::CommPars (18 doubles: tdim qdim nuq eta dqso dqsi dqdt Eo0 Ei0 Cs t0o t0i a0 da tauf tauc comb albf) /
 * SeaIce::CommPars (7 doubles: zeta a0 Lf s0 rhoo Qvar Q0); both re-run forcing + lin */
void set_atmos_parameters_(void* pars);
void set_seaice_parameters_(void* pars);
/* replaces m_mix::set_vmix_fix, src/ocean/mix.F90:52-59 */
void __m_mix_MOD_set_vmix_fix(int* fix);

/* ---- setup / diagnostics symbols THCM.C and Ocean.C bind besides the hot path (host code, parameter-change or output
 * frequency).  They exist so that the reference links against this library without an unresolved symbol. ---- */
/* m_global (src/ocean/global.F90:215-608; THCM.C:86-102): global-domain mask and forcing arrays read on the root.  Options
 * that read a Levitus / Trenberth data file (iza < 2, ite = 0 or its = 0 with restoring, internal T/S) fail loudly: those files
 * do not ship with the reference -- feed the fields through m_inserts instead */
void __m_global_MOD_set_maskfile(const char* maskfile);
void __m_global_MOD_get_current_landm(int* landm);
void __m_global_MOD_set_landm(int* landm);
void __m_global_MOD_get_windfield(double* taux, double* tauy);
void __m_global_MOD_get_temforcing(double* tatm);
void __m_global_MOD_get_salforcing(double* emip);
void __m_global_MOD_get_internal_temforcing(double* temp);
void __m_global_MOD_get_internal_salforcing(double* salt);
void __m_global_MOD_get_spert(double* spert);   /* SRES everywhere, or (1 - digit)(1 - landm(i,j,l)) from mkmask/<spertmaskfile> (read_spertm, forcing.F90:372-402) */
void __m_global_MOD_get_land_temp(double* land);   /* THCM.C:170 declares it; the reference's Fortran never defines it: fails loudly */
/* global.F90:241-293 (THCM.C:51-56): the caller's global grid arrays; checked against the library's own grid.F90 arrays */
void set_global_x(int* n, double* a);
void set_global_y(int* n, double* a);
void set_global_z(int* n, double* a);
void set_global_xu(int* n, double* a);
void set_global_yv(int* n, double* a);
void set_global_zw(int* n, double* a);
/* m_usr::set_internal_forcing (src/ocean/usr.F90:267-300; THCM.C:594): n*m*l internal T / S fields for the w-row forcing */
void __m_usr_MOD_set_internal_forcing(double* temp, double* salt);
/* m_thcm_utils::get_landm / loadbal_weights (src/ocean/thcm_utils.F90:259-277, 325-353) */
void __m_thcm_utils_MOD_get_landm(int* landm);
void __m_thcm_utils_MOD_loadbal_weights(double* weights, double* fac_ntrphys, double* fac_consmix, double* fac_convadj);
/* m_probe (src/ocean/probe.F90; THCM.C:136-170, 1568-1763): n*m surface diagnostics of the coupled model; `un` / `sol` is the
 * HOST state vector */
void __m_probe_MOD_get_atmosphere_t(double* f);
void __m_probe_MOD_get_atmosphere_q(double* f);
void __m_probe_MOD_get_atmosphere_p(double* f);
void __m_probe_MOD_get_emip(double* f);
void __m_probe_MOD_get_adapted_emip(double* f);
void __m_probe_MOD_get_emip_pert(double* f);
void __m_probe_MOD_get_taux(double* f);
void __m_probe_MOD_get_tauy(double* f);
void __m_probe_MOD_get_suno(double* f);
void __m_probe_MOD_compute_evap(double* evap, double* un);
void __m_probe_MOD_get_salflux(double* un, double* salflux, double* scorr, double* qsoaflux, double* qsosflux);
void __m_probe_MOD_get_temflux(double* un, double* totflux, double* swflux, double* shflux, double* lhflux, double* siflux,
                               double* simask);
void __m_probe_MOD_get_derivatives(double* un, double* dftdm, double* dfsdq, double* dfsdm, double* dfsdg);
/* m_integrals (src/ocean/integrals.F90:17-88; THCM.C:2133, 2155): per-cell salt advection / diffusion integrands (n*m*l) */
void __m_integrals_MOD_salt_advection(double* un, double* check);
void __m_integrals_MOD_salt_diffusion(double* un, double* check);
/* get_stochastic_forcing (src/ocean/forcing.F90:235-280; THCM.C:848): fills begF / jcoF / coF of set_pointers */
void get_stochastic_forcing_(void);
/* getdeps / get_parameters / get_nondimensionalization_parameters (src/ocean/usrc.F90:201-251; Ocean.C:889, 1610) */
void getdeps_(double* Ooa, double* Os, double* nus, double* eta, double* lvsc, double* qdim, double* pqsnd);
void get_parameters_(double* r0dim, double* udim, double* hdim);
void get_nondimensionalization_parameters_(double* out27);
/* writeparams (usrc.F90:421-431 -> fort.7), write_data (src/ocean/inout.F90:20-93 -> fort.3; Ocean.C:1877), write_levitus */
void writeparams_(void);
void write_data_(double* u, int* ofile, int* lab);
void write_levitus_(const char* filename);
/* replace m_scaling::average_block / compute (src/ocean/scaling.F90:29-105; THCM.C:106-107, 1798-1807): the local average
 * 6x6 diagonal block of the Jacobian of the last matrix_ call over the OCEAN cells, db(nun,nun) column-major; and the THCM
 * row / column scaling vectors built from the (globally averaged) block */
void __m_scaling_MOD_average_block(double* db);
void __m_scaling_MOD_compute(double* db, double* rowscales, double* colscales);
/* replaces m_thcm_utils::intcond_scaling (src/ocean/thcm_utils.F90:285-309; THCM.C:119, 2619): cos(y_j) dfzT_k on the S rows
 * of the OCEAN cells (1-based local row ids) */
void __m_thcm_utils_MOD_intcond_scaling(double* values, int* indices, int* len);

/* Callbacks the reference's Fortran calls back into C++ (THCM.C:2653,2690; GlobalDefinitions.C:145,154).
 * When the library is linked against the reference these resolve to its definitions; standalone,
 * weak defaults inside the library are used. */
void thcm_forcing_integral_(double* field, double* y, int* landm, double* out);
void thcm_throw_error_(char* msg);
void timer_start_(const char* label);
void timer_stop_(const char* label);

/* --------------------------------------------------------------------------
 * B1': handle-based device API
 * ------------------------------------------------------------------------ */
typedef struct thcmb_ctx thcmb_ctx;

typedef struct thcmb_settings {
    /* global domain (global.F90:65-157); angles in radians */
    int N, M, L;
    double xmin, xmax, ymin, ymax, hdim, qz;
    int periodic;
    /* model flags (usr.F90:50-85) */
    int ih, vmix, tap, rho_mixing, coriolis_on, TRES, SRES, iza, ite, its, coupled_T, coupled_S, forcing_type;
    double alphaT, alphaS;
    /* domain decomposition (TRIOS_Domain.C:201-315): this process is rank `rank` of `nranks` */
    int rank, nranks;
    /* CUDA device ordinal to use */
    int device;
    /* latitude bounds of the GLOBAL domain when this context is one sub-domain of an MPI-decomposed run that calls the
     * Fortran symbols per rank (temfun / salfun use m_global's ymin, ymax, forcing.F90:418-449); both 0 = same as ymin, ymax */
    double ymin_glob, ymax_glob;
    /* 0: the reference's uniform cut lines (TRIOS_Domain.C:258-273); 1: same rank grid and rectangular blocks, cut lines placed by
     * OCEAN-cell count (the reference's "load balancing: not implemented", TRIOS_Domain.C:384-392) -- LAND cells are identity rows and
     * cost almost nothing, so uniform cuts leave the slowest of 8 ranks with 1.6x the mean work on a global mask */
    int balance;
} thcmb_settings;

void thcmb_default_settings(thcmb_settings* s);

/* Ocean::getBlock(std::shared_ptr<Atmosphere>) / getBlock(std::shared_ptr<SeaIce>) (Ocean.C:1603-1730, 1733-1810): the coupling blocks
 * d F_ocean / d x_atmosphere and d F_ocean / d x_seaice of the coupled model's Jacobian, CRS over all 6 N M L ocean rows in FIND_ROW2
 * order (0-based beg[ndim + 1]; at most 3 entries per surface T / S row: allocate 3 * 2 * N * M).  The column ids are the other model's
 * interface_row(i, j, XX) per surface point (N*M ints, i fastest; colP: -1 where the atmosphere has no auxiliary precipitation row).
 * One rank (the reference all-gathers the surface fields for these blocks too).  Return the number of entries. */
int thcmb_ocean_block_atmosphere(thcmb_ctx* c, double albed, const double* pdist, const int* colT, const int* colQ, const int* colA,
                                 const int* colP, int* beg, int* jco, double* co);
int thcmb_ocean_block_seaice(thcmb_ctx* c, const double* un_host, const int* colQ, const int* colM, const int* colG, int* beg, int* jco,
                             double* co);

/* landm_global: (N+2)(M+2)(L+2) ints, i fastest, values OCEAN 0 / LAND 1 / WATER 2 / PERIO 3 (par.F90:78-81).
 * Every rank passes the same global mask (replaces the bcast+import of THCM.C:378-565). Returns NULL on error. */
thcmb_ctx* thcmb_create(const thcmb_settings* s, const int* landm_global);
void thcmb_destroy(thcmb_ctx* c);
const char* thcmb_last_error(void);
/* The ONE instance behind the Fortran symbols of B1 (created by init_, replaced by the next init_, gone after finalize_; NULL outside):
 * lets code that keeps calling the Fortran symbols use the device API on the same model (include/thcm_epetra_bridge.hpp) */
thcmb_ctx* thcmb_fortran_context(void);

/* decomposition queries (TRIOS_Domain.H:114-247 subset) */
void thcmb_local_block(const thcmb_ctx* c, int* i0, int* j0, int* n0, int* m0, int* npN, int* npM);
int thcmb_ndim_local(const thcmb_ctx* c);  /* 6*n0*m0*L owned unknowns */
long long thcmb_graph_nnz(const thcmb_ctx* c);
/* tiles of 32 cells of this rank's block, and how many of them hold at least one non-LAND cell (the Jacobian kernels write the identity
 * rows of the all-LAND tiles once and revisit only the others) */
void thcmb_tile_counts(const thcmb_ctx* c, int* ntiles, int* nactive);
/* static maximal graph of the owned rows (THCM.C:2300-2580): 0-based CSR, columns sorted ascending by global id;
 * col[] holds LOCAL column ids: < ndim_local owned, >= ndim_local halo slot (see thcmb_halo_gids) */
void thcmb_get_graph(const thcmb_ctx* c, int* rowptr, int* col);
int thcmb_halo_size(const thcmb_ctx* c);              /* number of halo unknowns (6 per halo cell) */
void thcmb_halo_gids(const thcmb_ctx* c, int* gids);  /* global id of each halo unknown */
void thcmb_local_gids(const thcmb_ctx* c, int* gids); /* global id of each owned unknown (standard map) */

/* parameters (setparcs/getparcs); set recomputes forcing and the linear tables */
void thcmb_set_par(thcmb_ctx* c, int idx, double val);
double thcmb_get_par(const thcmb_ctx* c, int idx);
/* coupled mode (coupled_T / coupled_S = 1; BASELINE configs[2]: the ocean block of the coupled model): surface fields on
 * the GLOBAL N*M grid (every rank passes the same field), `which` = 0 taux, 1 tauy, 2 atmosphere_t, 3 emip, 4 emip_pert,
 * 5 adapted_emip, 6 atmosphere_q, 7 atmosphere_a, 8 atmosphere_p, 9 seaice_q, 10 seaice_m, 11 seaice_g (inserts.F90);
 * the two parameter setters follow usrc.F90:254-350 and recompute forcing and the linear tables */
void thcmb_insert_field(thcmb_ctx* c, int which, const double* field_global);
void thcmb_set_atmos_parameters(thcmb_ctx* c, const double* pars18);
void thcmb_set_seaice_parameters(thcmb_ctx* c, const double* pars7);
void thcmb_get_forcing(thcmb_ctx* c, double* frc_host);
void thcmb_get_cob(thcmb_ctx* c, double* cob_host);

/* NCCL plumbing for nranks>1: the caller creates a 128-byte ncclUniqueId on rank 0 (thcmb_nccl_unique_id),
 * broadcasts it (torch.distributed) and every rank calls thcmb_nccl_init. */
int thcmb_nccl_unique_id(void* id128);
int thcmb_nccl_init(thcmb_ctx* c, const void* id128);
/* Optional (same node, NVLink): fused reduction + all-reduce over peer memory.  Every rank exports the CUDA IPC handle
 * (64 bytes) of its mailbox, the caller all-gathers them (nranks x 64 bytes, rank order) and hands them back; from then
 * on dot products / MGS steps finish their cross-GPU sum inside the reduction kernel instead of calling ncclAllReduce. */
int thcmb_p2p_local_handle(thcmb_ctx* c, void* handle64);
int thcmb_p2p_open(thcmb_ctx* c, const void* handles_all);

/* ---- hot path, device pointers, asynchronous on the context's stream ---- */
/* halo exchange of a state/Krylov vector (width 1, edges+corners, periodic wrap): fills the context's halo buffer */
int thcmb_halo_exchange(thcmb_ctx* c, const double* d_x);
/* F(x): d_F = +A(u)u + mix - Frc  (the sign Ocean::computeRHS returns, THCM.C:1011); does its own halo exchange */
int thcmb_residual_dev(thcmb_ctx* c, const double* d_un, double* d_F);
/* THCM-sign residual exactly as rhs_ returns it (B = -Au - mix + Frc, masked) */
int thcmb_rhs_dev(thcmb_ctx* c, const double* d_un, double* d_B);
/* Jacobian values into the static graph (explicit zeros kept), stored inside the context */
int thcmb_jacobian_dev(thcmb_ctx* c, const double* d_un);
const double* thcmb_jacobian_values(const thcmb_ctx* c);   /* device pointer, graph order */
const int* thcmb_graph_rowptr_dev(const thcmb_ctx* c);
const int* thcmb_graph_col_dev(const thcmb_ctx* c);
/* out[slot[e]] = in[e], e < n, on the context's stream: the stored values (graph order) into the value layout of the caller's
 * matrix object -- what replaces the ReplaceGlobalValues loop of THCM.C:1082-1104 (include/thcm_epetra_bridge.hpp) */
int thcmb_scatter_values_dev(thcmb_ctx* c, long long n, const int* d_slot, const double* d_in, double* d_out);
/* Fortran-order thresholded CRS on the device (count -> scan -> fill); returns nnz, arrays 1-based like matrix_ */
long long thcmb_jacobian_crs_dev(thcmb_ctx* c, const double* d_un, int* d_begA, int* d_jcoA, double* d_coA);
/* y = J x on the stored Jacobian (Ocean::applyMatrix, Ocean.C:1369-1374), with halo exchange when nranks>1 */
int thcmb_spmv_dev(thcmb_ctx* c, const double* d_x, double* d_y);
/* generic CSR SpMV on caller-provided device arrays (0-based) */
int thcmb_csr_spmv_dev(thcmb_ctx* c, int nrow, const int* d_rowptr, const int* d_col, const double* d_val,
                       const double* d_x, double* d_y);

/* Theta time stepping, src/transient/ThetaModel.H:87-165.  rhs: d_F holds F(u_{n+1}) on entry and
 * M (u_n - u_{n+1}) + dt (1-theta) F(u_n) + dt theta F(u_{n+1}) on return (M = the mass diagonal coB); jacobian: adds
 * -M / (theta dt) to the diagonal of the stored Jacobian (no-op for theta = 0, like the reference) */
int thcmb_theta_rhs_dev(thcmb_ctx* c, double theta, double dt, const double* d_state, const double* d_old_state,
                        const double* d_old_rhs, double* d_F);
int thcmb_theta_jacobian_dev(thcmb_ctx* c, double theta, double dt);
/* Ocean::applyMassMat (Ocean.C:1448-1457): d_out = M d_v with the diagonal mass matrix (coB of assemble.F90:18-54; 0 on w, p, LAND and
 * the replaced rows of thcmb_enable_intcond / thcmb_fix_pressure_points).  d_out is not read; returns -1 when it aliases d_v */
int thcmb_apply_mass_dev(thcmb_ctx* c, const double* d_v, double* d_out);

/* vector kernels (Epetra_MultiVector::Dot/Norm2/Update/Scale as used by GMRESSolver.H / IDRSolver.H);
 * dot/nrm2 all-reduce over ranks and return on the host */
double thcmb_dot(thcmb_ctx* c, int n, const double* d_x, const double* d_y);
double thcmb_nrm2(thcmb_ctx* c, int n, const double* d_x);
int thcmb_axpby(thcmb_ctx* c, int n, double a, const double* d_x, double b, double* d_y); /* y = a x + b y */
int thcmb_scale(thcmb_ctx* c, int n, double a, double* d_x);

/* preconditioner for the in-library Krylov solvers: 0 identity, 1 6x6 block-diagonal of the stored Jacobian */
int thcmb_build_precon(thcmb_ctx* c, int kind);
int thcmb_apply_precon_dev(thcmb_ctx* c, const double* d_x, double* d_y);

typedef struct thcmb_krylov_result {
    int status;          /* 0 converged, 1 not converged, <0 breakdown */
    int iters;
    double resid;        /* final (relative for GMRES, absolute for IDR) implicit residual */
    int nhist;           /* entries written to hist */
    long long n_matvec;
} thcmb_krylov_result;

/* Restarted right-preconditioned (F)GMRES with modified Gram-Schmidt and Givens rotations: the algorithm of
 * src/gmressolver/GMRESSolver.H:81-255 (minimiser scheme 'B').  d_b, d_x on device; hist (host, may be NULL)
 * receives the scaled residual after every inner iteration.  flags: bit0 precondition, bit2 flexible, bit4 (16) full-length Krylov vectors (no ocean-only compaction),
 * bit3 batched Gram-Schmidt with the DGKS criterion (the orthogonalisation Belos uses on the reference's production path,
 * Ocean.C:977-1024) instead of the template's modified Gram-Schmidt: 2-4 global reductions per iteration instead of i+2. */
int thcmb_gmres(thcmb_ctx* c, const double* d_b, double* d_x, double tol, int maxit, int restart, int flags,
                double* hist, int hist_cap, thcmb_krylov_result* res);
/* IDR(s) with bi-orthogonalisation: src/idrsolver/IDRSolver.H:109-340.  d_P_raw: s vectors (s x ndim) that
 * replace Vector::random() in createP (IDRSolver.H:84-104); orthonormalised on the device. */
int thcmb_idrs(thcmb_ctx* c, const double* d_b, double* d_x, double tol, int maxit, int s, const double* d_P_raw,
               double* hist, int hist_cap, thcmb_krylov_result* res);

/* One Newton step from HOST buffers (the end-to-end call): H2D un, F(un), J(un), solve J dx = -F with
 * preconditioned GMRES, D2H dx.  Returns the Krylov result; fnorm receives ||F||_2. */
int thcmb_newton_step(thcmb_ctx* c, const double* un_host, double* dx_host, double tol, int maxit, int restart,
                      int precon_kind, double* fnorm, thcmb_krylov_result* res);

/* orthogonalisation used by thcmb_newton_step*: 0 = modified Gram-Schmidt (GMRESSolver.H:177-181), 1 = batched DGKS */
void thcmb_set_ortho(thcmb_ctx* c, int mode);
/* tracer mixing (Mixing = 1, 2; mix_imp.f): m_mix::set_vmix_fix (mix.F90:52-59, THCM.C:2639-2647) and the current
 * {vmix_flag, vmix_temp, vmix_salt, vmix_fix} */
void thcmb_set_vmix_fix(thcmb_ctx* c, int fix);
/* THCM::RecomputeScaling (THCM.C:1781-1834) on the stored Jacobian: host vectors of ndim_local entries in Trilinos' convention
 * (the inverse of THCM's), T and S scaled alike; db36_out (optional) = the averaged block, column-major.  Returns 1 when the
 * block is singular to working precision (the reference then leaves the scaling unset; here it is the identity). */
int thcmb_recompute_scaling(thcmb_ctx* c, double* row_scaling, double* col_scaling, double* db36_out);
/* THCM::getIntCondCoeff (THCM.C:2608-2637): integral-condition coefficients on the owned S rows; returns the local volume */
double thcmb_intcond_coeff(const thcmb_ctx* c, double* coeff);
void thcmb_get_vmix_flags(const thcmb_ctx* c, int* out4);
/* The row replacements THCM::evaluate makes above the Fortran core (THCM.C:653-697, 1013-1041, 1164-1172, 2180-2296).
 * thcmb_enable_intcond: with SRES = 0 the S row of cell (Nic, Mic, L-1) (0-based, -1 = N-1 / M-1; "Integral row coordinate
 * i / j") becomes the salinity integral condition: F_row = sign (c.x - correction), J_row = sign c^T (c = thcmb_intcond_coeff,
 * sign = "Salinity Integral Sign").  The row is dense: it is not stored in the graph, thcmb_spmv_dev adds it as one fused
 * dot product + all-reduce.  thcmb_set_intcond_correction = THCM::setIntCondCorrection (returns the correction);
 * thcmb_fix_pressure_points = "Fix Pressure Points" (identity rows for p of the cells (N-1, M-1, L-1), (N-2, M-1, L-1)).
 * All three also zero the mass diagonal of the replaced rows.  thcmb_intcond_row: global row id (0-based) or -1 */
void thcmb_enable_intcond(thcmb_ctx* c, int Nic, int Mic, int sign);
/* THCM::setLandMask (THCM.C:1362-1392) / set_landmask_ with reinit = 1 (usrc.F90:353-418) for a handle: the instance is rebuilt on
 * the new GLOBAL mask ((N+2)(M+2)(L+2) ints, i fastest), same decomposition.  init = 0: no-op (the reference then only updates m_global's
 * copy).  Rebuild the preconditioner afterwards.  Returns -1 (thcmb_last_error) when an enabled integral condition's cell became land */
int thcmb_set_landmask(thcmb_ctx* c, const int* landm_global, int init);
double thcmb_set_intcond_correction(thcmb_ctx* c, const double* d_vec);
void thcmb_fix_pressure_points(thcmb_ctx* c, int on);
int thcmb_intcond_row(const thcmb_ctx* c);
/* the same step with the state already in HBM (bench.py "value") */
int thcmb_newton_step_dev(thcmb_ctx* c, const double* d_un, double* d_dx, double tol, int maxit, int restart,
                          int precon_kind, double* fnorm, thcmb_krylov_result* res);

/* per-kernel device timing: while on, every kernel launch of the library is bracketed by a CUDA event pair on the
 * context's stream; thcmb_profile_report sums them per kernel id (0 .. thcmb_kernel_count()-1) */
void thcmb_profile(thcmb_ctx* c, int on);
int thcmb_kernel_count(void);
const char* thcmb_kernel_name(int kid);
int thcmb_profile_report(thcmb_ctx* c, int kid, int* count, double* total_ms);

/* utilities */
void* thcmb_device_alloc(thcmb_ctx* c, long long bytes);
void thcmb_device_free(thcmb_ctx* c, void* p);
int thcmb_h2d(thcmb_ctx* c, void* d_dst, const void* h_src, long long bytes);
int thcmb_d2h(thcmb_ctx* c, void* h_dst, const void* d_src, long long bytes);
int thcmb_sync(thcmb_ctx* c);
void* thcmb_stream(thcmb_ctx* c);               /* cudaStream_t the hot path runs on */
long long thcmb_launch_count(const thcmb_ctx* c); /* kernels launched by this library so far */
/* per-stage device time in ms of the last call, by the reference's profile labels (GlobalDefinitions.C:145-172):
 * "nlin_rhs+boundaries+matAvec", "nlin_jac+boundaries+fillcolA", "matAvec", ... ; returns -1 if unknown */
double thcmb_last_stage_ms(const thcmb_ctx* c, const char* label);

#ifdef __cplusplus
}
#endif
#endif /* THCM_B200_H */
