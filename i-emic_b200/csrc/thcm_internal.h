// Internal declarations of libthcm_b200 (not part of the public ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <string>
#include <vector>
#include <map>
#include "../../include/thcm_b200.h"

namespace thcm {

// par.F90:17-81
constexpr int NUN = 6, NP = 27, NPAR = 30;
enum { AL_T = 1, RAYL, EK_V, EK_H, ROSB, MIXP, RESC, SPL1, HMTP, SUNP, PE_H, PE_V, P_VC, LAMB, SALT, WIND, TEMP,
       BIOT, COMB, ARCL, NLES, IFRICB, CONT, ENER, ALPC, CMPR, FPER, SPER, MKAP, SPL2 };
enum { UU = 1, VV, WW, PP, TT, SS };
enum { OCEAN = 0, LAND = 1, WATER = 2, PERIO = 3 };

// ---- 1-D coefficient tables (host-computed with glibc libm, uploaded on every parameter change) ----
// j-tables are indexed by GLOBAL j in 0..M+1, k-tables by k in 0..L+1.
enum JT {
    J_LU2, J_LU4, J_LU6, J_LU5,      // Al(UU,UU): -(EH*uxx2), -(EH*uyy4), -(EH*uyy6), -(EH*((uxx5+uyy5)+ucsi5))
    J_LUV2, J_LUV8,                  // Al(UU,VV) loc 2/8: -(EH*vxs)
    J_CORV,                          // sin(yv)*coriolis_on
    J_C2X, J_C2Y,                    // 1/(2 cos(yv) dx), 1/(2 cos(yv) dy)
    J_COSYV, J_TANYV,                // cos(yv_j), tan(yv_j)
    J_LV2, J_LV4, J_LV6, J_LV5,      // Al(VV,VV)
    J_LVU2, J_LVU8,                  // Al(VV,UU) loc 2/8: -(EH*uxs)
    J_CP,                            // 1/(2 cos(y) dx)            (pderiv 1)
    J_PVA, J_PVB,                    // cos(yv_{j-1})*c, cos(yv_j)*c with c = 1/(2 cos(y) dy)  (pderiv 2)
    J_TT2, J_TT4, J_TT6, J_TT5,      // Al(TT,TT)=Al(SS,SS) horizontal part at surface-mask 1
    J_C4X, J_C4Y,                    // 1/(4 cos(y) dx), 1/(4 cos(y) dy)
    J_COUNT
};
enum KT {
    K_ZU5, K_ZU14, K_ZU23,           // EV*uzz(5|14|23)
    K_TDZI8,                         // 1/(8 dfzT dz)
    K_WP5, K_WP23,                   // gradp(3)
    K_PW5, K_PW14,                   // pderiv(3)
    K_ZT5, K_ZT14, K_ZT23,           // pv*tzz(5|14|23) at surface-mask 1
    K_DFZT,
    K_RT, K_RS,                      // (TRES*bi)*tc5, (SRES*bi)*sc5
    K_DFZW, K_DFZWM,                 // dfzW(k), dfzW(k-1)  (tracer mixing, mix_imp.f:613-641)
    K_COUNT
};

struct DevTables {
    const double* jt;  // [J_COUNT][jstride]
    const double* kt;  // [K_COUNT][kstride]
    int jstride, kstride;
    double epsr, dyi, cWT, cWS, c2, c3, tdzi2;
    // tracer mixing (mix_imp.f): implicit vertical mixing / convective adjustment.  mix_temp / mix_salt = vmix_temp / vmix_salt
    // (0 when mixing is off), mix_fac = alphaT * SPL1 (tprstb), mix_rho = rho_mixing .and. xes == 0
    double mix_lambda, mix_xes, mix_kvc, mix_eps, mix_fac, mix_dz;   // mix_eps = (1 - ALPC) * ENER * PE_V (consistent vertical mixing)
    int mix_temp, mix_salt, mix_rho;
    // coupled mode (coupled_T / coupled_S = 1, usrc.F90:733-786): surface-level terms of the T and S rows that replace the
    // restoring term.  msi = sea-ice mask of the owned columns (n0*m0, i fastest); every constant is the sub-expression the
    // reference forms: cpl_ooa = Ooa, cpl_dedt_t = lvsc*eta*qdim*(deltat/qdim)*dqso, cpl_qtz = QTnd*zeta,
    // cpl_ts = -QTnd*zeta*a0, cpl_pq = COMB*SALT*QSnd, cpl_rl = rhodim*Lf, cpl_dedt_s = nus*(deltat/qdim)*dqso
    const double* msi;
    int coupled_T, coupled_S;
    double cpl_ooa, cpl_dedt_t, cpl_qtz, cpl_ts, cpl_pq, cpl_zeta, cpl_a0, cpl_rl, cpl_dedt_s;
};

// Cut lines of the lon x lat block partition.  The reference cuts the index ranges uniformly (TRIOS_Domain.C:258-273) and notes load
// balancing as "not implemented" (TRIOS_Domain.C:384-392); with LAND = identity rows that costs up to 1.6x at 8 ranks on a global mask
// (the slowest rank owns 1.6x the mean number of OCEAN cells).  thcmb_settings.balance = 1 keeps the reference's npN x npM rank grid and
// rectangular blocks but places the cuts by ocean-cell count: latitude cuts jc[0..npM] first, then longitude cuts per latitude band,
// ic[pm * (npN + 1) + 0..npN].  Empty (npM == 0) = the reference's uniform cuts.
struct Cuts { int npN = 0, npM = 0; std::vector<int> jc, ic; };

// ---- local block of the global grid (TRIOS_Domain.C:201-315 decomposition, global indexing kept) ----
struct Block {
    const Cuts* cuts = nullptr;   // non-uniform cut lines (owned by the context), nullptr = uniform
    int N, M, L;        // global sizes
    int i0, j0;         // 0-based global offset of the first owned cell
    int n0, m0;         // owned cells in x, y (full depth L)
    int npN, npM, pidN, pidM, rank, nranks;
    int periodic;       // global periodicity in x
    int wrap_x;         // periodic and a single rank in x: wrap by index, no x-halo
    int halo_w, halo_e, halo_s, halo_n;  // 1 if a halo strip exists on that side
    // halo cells per level: south row, north row (both n0 + halo_w + halo_e wide incl. corners), west col, east col
    int hk;             // halo cells per k level
    int ncell() const { return n0 * m0 * L; }
    int ndim() const { return NUN * ncell(); }
    int nhalo_cells() const { return hk * L; }
};

// device view used by the assembly kernels
struct DevBlock {
    int N, M, L, i0, j0, n0, m0, periodic, wrap_x, halo_w, halo_e, halo_s, halo_n, hk, ncell;
};

// class tables: position of every canonical slot inside the sorted graph row, per boundary class
// class bits: 1 i==first, 2 i==last, 4 j==first, 8 j==last, 16 k==first, 32 k==last (GLOBAL indices)
constexpr int NCLASS = 64;
constexpr int NSLOT_TOTAL = 104;
struct ClassTables {
    int8_t pos[NCLASS][NSLOT_TOTAL];  // position within its row, -1 = not in graph (clipped)
    uint8_t rowlen[NCLASS][NUN];
};

// Static per-tile descriptor of the pipelined assembly kernel (host-built from the land mask, one TMA bulk load per tile):
// the per-cell neighbour masks and, for each of the 9 staged grid lines, which positions keep u,v / w after usol.
struct alignas(16) TileDesc {
    uint32_t nbmask[32];   // per cell of the tile (0 beyond ncell)
    uint64_t uvbits[9];    // bit x (0..33) of line r: u,v survive usol's no-slip rule at staged column x
    uint64_t wbits[9];     // likewise for w (lid / bottom / ghost-column rule)
    uint32_t surfbits;     // bit lane: 1 - landm(i,j,L) of the cell's column
    uint32_t flags;        // bit 0: nothing clipped and the tile's first entry is 16-byte aligned (TMA bulk store);
                           // bit 1: open ocean (all neighbour masks zero: `boundaries` is a no-op for the whole tile);
                           // bit 2: every cell of the tile is LAND
    int32_t g0;            // graph offset of the tile's first entry
    int32_t tot;           // number of graph entries of the tile
};
static_assert(sizeof(TileDesc) == 288, "TileDesc is copied with cp.async.bulk (16-byte multiples)");
constexpr int JREC = 3;   // j-table record: [J_COUNT][3] values at gj-1, gj, gj+1

struct Timer {
    cudaEvent_t a = nullptr, b = nullptr;
};

}  // namespace thcm

struct thcmb_ctx {
    thcmb_settings s;
    thcm::Cuts cuts;
    thcm::Block blk;
    int device = 0;
    cudaStream_t stream = nullptr;
    // ---- host model state ----
    double par[thcm::NPAR + 1];
    double dx, dy, dz, QTnd, QSnd;
    std::vector<double> x, y, z, xu, yv, zw, ze, zwe, dfzT, dfzW;  // GLOBAL grid, Fortran index = vector index
    std::vector<int> landm;  // GLOBAL (0:N+1,0:M+1,0:L+1) after init's frame rules
    std::vector<double> taux, tauy, tatm, emip, spert, adapted_emip;  // GLOBAL N*M surface fields
    std::vector<double> qatm, albe, patm, msi, gsi, qsa;              // coupled mode (m_usr, usr.F90; inserts.F90:33-157)
    // m_atm (atm.F90) / m_ice (ice.F90) values the ocean needs in coupled mode; set_atmos/seaice_parameters (usrc.F90:254-350)
    double atm_qdim = 0.01, atm_nuq = 0.0, atm_nus = 0.0, atm_eta = 0.0, atm_dqso = 0.0, atm_eo0 = 0.0, atm_albe0 = 0.0,
           atm_albed = 0.0, atm_lvsc = 0.0, atm_Ooa = 1.0, atm_Os = 1.0;
    std::vector<double> suno;                                         // shortwave profile (usrc.F90:1228), index j
    double ice_zeta = 0.0, ice_a0 = -0.0575, ice_Lf = 3.347e+05, ice_Qvar = 0.0, ice_Q0 = 0.0;
    std::vector<double> internal_temp, internal_salt;                 // m_usr::set_internal_forcing (usr.F90:267-300), N*M*L
    bool internal_set = false;
    std::vector<double> msi_local;                                    // msi on the owned columns (n0*m0, i fastest)
    double* d_msi = nullptr;
    std::vector<double> frc_local;   // owned rows, masked by the rows `boundaries` turns into identity rows
    std::vector<double> frc_raw;     // owned rows, as `forcing` leaves it
    bool use_integral_callback = false;   // qint goes through thcm_forcing_integral_ (THCM.C:2653): contexts created by init_
    bool frc_masked = false;         // get_forcing_ semantics (boundary.F90 zeroes Frc lazily inside rhs/matrix)
    std::vector<double> cob_local;
    std::vector<double> jt_host, kt_host;
    thcm::DevTables tab;
    // ---- static device data ----
    double *d_jt = nullptr, *d_kt = nullptr;
    uint32_t* d_nbmask = nullptr;   // per owned cell
    uint8_t* d_surf = nullptr;      // per owned column: 1 - landm(i,j,L)
    uint8_t* d_uvlive = nullptr;    // (n0+2)(m0+2)L corner box: 1 if usol keeps u,v there
    uint8_t* d_cls = nullptr;       // per owned cell: boundary class (0..63)
    thcm::TileDesc* d_tdesc = nullptr;  // per tile (pipelined assembly kernel)
    double *d_jrec = nullptr, *d_krec = nullptr;   // per-j / per-k table records (one bulk load each)
    std::vector<double> jrec_host, krec_host;
    unsigned int* d_tilectr = nullptr;
    int asm_pipe = 1;               // Jacobian kernel: 1 block per tile with TMA staging (default), 0 block per tile with
                                    // per-position loads (THCM_ASM_PIPE=0; the kernel family of the residual and of the Fortran-order CRS)
    int* d_active_tiles = nullptr; int n_active_tiles = 0;   // tiles with at least one non-LAND cell
    bool land_tiles_written = false;                          // the identity rows of the all-LAND tiles are in d_val
    double* d_frc = nullptr;        // owned rows (masked)
    double* d_cob = nullptr;        // mass diagonal coB of the owned rows (theta stepping)
    int *d_rowptr = nullptr, *d_col = nullptr;  // static graph, local column ids
    double* d_val = nullptr;        // Jacobian values in graph order
    long long gnnz = 0;
    std::vector<int> rowptr_host, col_host, halo_gid, local_gid;
    std::vector<int> ocell_host, ccell_host;   // ocean cells of the block (cell order) and the inverse map (-1 = LAND)
    int *d_ocell = nullptr, *d_ccell = nullptr; int n_ocell = 0;
    std::vector<int> colc_host, send_cidx_host;   // compact column ids (graph offsets), compact source cell of the send lists
    int *d_colc = nullptr, *d_send_cidx = nullptr;
    double* d_iccoeff_c = nullptr;   // integral-condition coefficients gathered to the ocean cells
    void* d_peer_ll = nullptr;       // device array [2][npeers]: the neighbours' LL halo buffers (compact SpMV), per parity
    void* d_halo_ll[2] = {nullptr, nullptr};   // my LL halo buffers (inside the IPC-shared allocation)
    double* d_halo_plain_c = nullptr;            // the LL halo of one exchange landed as plain doubles (compact SpMV after the fused head kernel)
    unsigned long long halo_landed_seq = 0;      // the exchange d_halo_plain_c holds (0 = none)
    unsigned long long halo_ll_seq = 0;
    int krylov_compact = 1;          // GMRES on the ocean cells only (THCM_KRYLOV_COMPACT=0 switches it off)
    signed char* d_cpos = nullptr;   // [64 classes][6 rows][6 cols]: position of the in-cell entry inside its sorted graph row, -1 = none
    uint8_t* d_landcell = nullptr;   // per owned cell: 1 = LAND (identity rows; the SpMV answers y = x for them without streaming the row)
    // ---- halo exchange ----
    double *d_halo = nullptr;       // 6*nhalo doubles, laid out per Block::hk
    double *d_sendbuf = nullptr, *d_recvbuf = nullptr;
    int* d_send_idx = nullptr;      // owned-cell ids to pack, grouped by neighbour
    int* d_recv_slot = nullptr;     // halo slot for every received cell, grouped by neighbour
    // direct halo push over NVLink peer memory (thcm_linalg.cu): destination slot in the PEER's halo buffer and peer index
    // of every packed cell; the halo buffers (two, alternating) live inside the IPC-shared mailbox allocation
    std::vector<int> send_dst_host, send_peer_host;
    int *d_send_dst = nullptr, *d_send_peer = nullptr;
    bool halo_p2p = false;
    double* d_halo_p2p[2] = {nullptr, nullptr};
    double** d_peer_halo = nullptr;  // device array [2][npeers]: peer halo buffer base per parity
    unsigned long long halo_seq = 0;
    unsigned int* d_halo_counter = nullptr;
    struct Peer { int rank; int send_off, send_cnt, recv_off, recv_cnt; };
    std::vector<Peer> peers;
    int nsend_cells = 0, nrecv_cells = 0;
    void* nccl_comm = nullptr;
    // fused reduction + all-reduce over NVLink peer memory (thcm_linalg.cu)
    bool p2p_on = false;
    void* d_mailbox = nullptr;
    void* d_peer_mailboxes = nullptr;
    std::vector<void*> p2p_peer_ptrs;
    unsigned long long* d_p2p_seq = nullptr;   // device counters of completed exchanges: [0] scalar slots, [1] vector slots
    double* d_mdpartial = nullptr;   // multi_dot partials
    int* d_flags = nullptr;          // device flags (DGKS second-pass decision)
    // ---- workspaces ----
    double* d_un = nullptr;         // staging for host-pointer entry points
    double* d_tmp = nullptr;
    double* d_partial = nullptr;    // reduction partials
    double* d_scalars = nullptr;    // device scalars (dot results, H column)
    double* h_scalars = nullptr;    // pinned
    unsigned int* d_counter = nullptr;
    int* d_blockcnt = nullptr;      // CRS count per assembly block
    int n_asm_blocks = 0;
    double* d_minv = nullptr;       // block-diagonal inverse, 36 per cell
    int precon_kind = 0;
    std::vector<double*> krylov_pool;  // device vectors reused across solves (slots allocated on first use)
    double* d_work[4] = {nullptr, nullptr, nullptr, nullptr};   // dx of thcmb_newton_step, compact b / x and the Arnoldi work vector of thcmb_gmres: never pool slots
    long long launches = 0;
    std::map<std::string, double> stage_ms;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t ev_slot[2] = {nullptr, nullptr};   // pipelined GMRES: scalars of step i have reached the host
    // optional per-kernel device timing (bench.py roofline): event pairs recorded around every launch
    bool prof_on = false;
    std::vector<cudaEvent_t> prof_ev;
    std::vector<int> prof_kid;
    // borrowed host CRS pointers (m_mat::set_pointers)
    int *begA = nullptr, *jcoA = nullptr; double *coA = nullptr, *coB = nullptr;
    int *begF = nullptr, *jcoF = nullptr; double *coF = nullptr;   // get_stochastic_forcing (forcing.F90:235-280)
    // row replacements of THCM::evaluate (THCM.C:1013-1041, 1164-1172): salinity integral condition (SRES = 0), pressure Dirichlet rows
    bool ic_on = false, pfix_on = false;
    int ic_sign = -1, ic_grow = -1, ic_lrow = -1;   // "Salinity Integral Sign", global row (rowintcon_), local row or -1 when not owned
    double ic_correction = 0.0;                      // THCM::setIntCondCorrection (THCM.C:2078-2097)
    double* d_iccoeff = nullptr;                     // intcondCoeff_ on the owned rows
    int pfix_grow[2] = {-1, -1}, pfix_lrow[2] = {-1, -1};
    int vmix_fix = 1, vmix_flag = 0, vmix_temp = 0, vmix_salt = 0, vmix_dim = 0;   // mix_imp.f:61-169
    long long crs_cap = 0;   // entries the caller's jcoA / coA hold (set_pointers' nnz argument)
    bool vmix_has_ocean = false;
    int fused_cgs2 = 2;             // DGKS: first update + second projection in one sweep over the basis (three reads of the basis per
                                    // iteration instead of four).  0 separate kernels, 1 basis values parked in shared memory, 2 parked
                                    // for nv <= 16 and L2-tiled above (default on any rank count).  THCM_FUSED_CGS2 overrides
    int gmres_ortho = 0;            // thcmb_newton_step: 0 modified Gram-Schmidt (GMRESSolver.H), 1 batched DGKS (Belos)
};

namespace thcm {
// host side (thcm_host.cpp)
void set_error(const std::string& msg);
void fatal(const std::string& msg);
bool decomp2d(int nprocs, int pid, int N, int M, int L, int periodic, Block& b, const Cuts* cuts = nullptr);
void compute_cuts(const int* landm_global, int N, int M, int L, int nprocs, Cuts& cuts);
bool setup_block(thcmb_ctx* c, const int* landm_global);   // cuts (when settings.balance) + decomp2d for the context's rank
void build_grid(thcmb_ctx* c);
void stpnt(thcmb_ctx* c);
void apply_landmask_rules(thcmb_ctx* c, const int* landm_in, bool fix_inversion);
void compute_forcing(thcmb_ctx* c);
void mask_forcing_rows(thcmb_ctx* c);
void atmos_coef(thcmb_ctx* c);
void init_surface_fields(thcmb_ctx* c);
enum SurfaceField { SF_TAUX = 0, SF_TAUY, SF_TATM, SF_EMIP, SF_SPERT, SF_ADAPTED_EMIP, SF_QATM, SF_ALBE, SF_PATM, SF_QSA, SF_MSI,
                    SF_GSI, SF_COUNT };
void insert_surface_field(thcmb_ctx* c, int which, const double* f);
void set_atmos_parameters(thcmb_ctx* c, const double* pars18);
void set_seaice_parameters(thcmb_ctx* c, const double* pars7);
void set_internal_forcing(thcmb_ctx* c, const double* temp, const double* salt);
// diagnostics / output symbols of the B1 boundary (thcm_probe.cpp; probe.F90, integrals.F90, forcing.F90:235, usrc.F90:201-251)
bool probe_get_field(const thcmb_ctx* c, int which, double* out);
void probe_get_suno(const thcmb_ctx* c, double* out);
void probe_compute_evap(const thcmb_ctx* c, const double* un, double* evap);
void probe_get_salflux(thcmb_ctx* c, const double* un, double* salflux, double* correction, double* qsoaflux, double* qsosflux);
void probe_get_temflux(const thcmb_ctx* c, const double* un, double* totflux, double* swflux, double* shflux, double* lhflux,
                       double* siflux, double* simask);
void probe_get_derivatives(thcmb_ctx* c, const double* un, double* dftdm, double* dfsdq, double* dfsdm, double* dfsdg);
void integrals_salt_advection(const thcmb_ctx* c, const double* un, double* check);
void integrals_salt_diffusion(const thcmb_ctx* c, const double* un, double* check);
void stochastic_forcing(thcmb_ctx* c, int* begF, int* jcoF, double* coF);
void get_deps(const thcmb_ctx* c, double* out7);
int ocean_block_atmosphere(thcmb_ctx* c, double albed, const double* pdist, const int* colT, const int* colQ, const int* colA,
                           const int* colP, int* beg, int* jco, double* co);
int ocean_block_seaice(thcmb_ctx* c, const double* un, const int* colQ, const int* colM, const int* colG, int* beg, int* jco, double* co);
void get_dim_parameters(const thcmb_ctx* c, double* r0, double* u0, double* h0);
void loadbal_weights(const thcmb_ctx* c, double* array);
void write_params(const thcmb_ctx* c);
void write_data(const thcmb_ctx* c, const double* u, int ofile, int* lab);
void compute_tables(thcmb_ctx* c);
void vmix_init(thcmb_ctx* c);
void vmix_set_flags(thcmb_ctx* c, int temp, int salt);
void compute_cob(thcmb_ctx* c);
const ClassTables& class_tables(int periodic);
int halo_slot(const Block& b, int ie, int je, int k);  // extended local coords (-1..n0, -1..m0); -1 if not a halo cell
// device side
void upload_class_tables(const ClassTables& t);
int scatter_slots(thcmb_ctx* c, long long n, const int* d_slot, const double* d_in, double* d_out);
int launch_assembly(thcmb_ctx* c, int mode, const double* d_un, double* d_out, int* d_begA, int* d_jcoA, double* d_coA);
enum { MODE_RHS = 0, MODE_JAC_GRAPH = 1, MODE_JAC_COUNT = 2, MODE_JAC_CRS = 3 };
int scan_block_counts(thcmb_ctx* c);
int asm_block_count(const Block& b);
int spmv(thcmb_ctx* c, int nrow, const int* rp, const int* col, const double* val, const double* x, const double* halo,
         int nlocal, double* y);
int halo_exchange(thcmb_ctx* c, const double* d_x, bool wait = true);
int field_sumsq(thcmb_ctx* c, const double* d_un, double* h_out2);   // global sum of squares of the T and S fields
// vector kernels (device-scalar flavours keep the Krylov inner loops free of host syncs)
int dot_dev(thcmb_ctx* c, int n, const double* x, const double* y, double* d_out);
int allreduce_dev(thcmb_ctx* c, double* d_buf, int count);
int multi_dot_dev(thcmb_ctx* c, int n, int nv, double* const* vecs, const double* w, const int* d_skip, double* d_out);
int multi_axpy_dev(thcmb_ctx* c, int n, int nv, double* const* vecs, const double* d_h, const int* d_skip, double* w);
int multi_axpy_dot_dev(thcmb_ctx* c, int n, int nv, double* const* vecs, const double* d_h, const int* d_skip, double* w, double* d_ww,
                       const double* d_ww_old, int* d_flag_out, double* d_final_out, int kid = -1, bool rare_path = false);
int fused_axpy_dot_dev(thcmb_ctx* c, int n, int nv, double* const* vecs, const double* d_h1, double* w, double* d_out,
                       const double* d_ww_old, int* d_flag_out, double* d_final_out, int* d_flag2_out = nullptr);
int dgks_flag_dev(thcmb_ctx* c, const double* ww_old, const double* ww_new, int* d_flag);
int mgs_step_dev(thcmb_ctx* c, int n, const double* d_hk, const double* vk, const double* vnext, double* w, double* d_out);
int nccl_unique_id(void* id128);
int nccl_init(thcmb_ctx* c, const void* id128);
void nccl_destroy(thcmb_ctx* c);
int p2p_local_handle(thcmb_ctx* c, void* handle64);
int p2p_open(thcmb_ctx* c, const void* handles_all);
void p2p_close(thcmb_ctx* c);
DevBlock dev_block(const Block& b);
void build_tile_descs(const thcmb_ctx* c, const std::vector<uint32_t>& nbmask, const std::vector<uint8_t>& surf,
                      const std::vector<uint8_t>& uvlive, std::vector<TileDesc>& out);
void build_static_host(thcmb_ctx* c, std::vector<uint32_t>& nbmask, std::vector<uint8_t>& surf, std::vector<uint8_t>& uvlive,
                       std::vector<int>& send_idx, std::vector<int>& recv_slot);
const std::string& last_error();
int axpby(thcmb_ctx* c, int n, double a, const double* x, double b, double* y);
int axpy_negdev(thcmb_ctx* c, int n, const double* d_h, const double* x, double* y);          // y -= (*d_h) x
int scale_invsqrt_dev(thcmb_ctx* c, int n, const double* d_nrm2, double* x, double* d_nrm);   // x /= sqrt(*d_nrm2)
int copy(thcmb_ctx* c, int n, const double* x, double* y);
int fill(thcmb_ctx* c, int n, double a, double* x);
int mass_apply(thcmb_ctx* c, int n, const double* d_cob, const double* v, double* out);
int theta_rhs(thcmb_ctx* c, int n, double theta, double dt, const double* state, const double* old_state, const double* old_rhs,
              const double* d_cob, double* F);
int theta_jacobian(thcmb_ctx* c, double theta, double dt, const double* d_cob);
int fix_residual_rows(thcmb_ctx* c, const double* d_un, double* d_F);
int fix_spmv_rows(thcmb_ctx* c, const double* d_x, double* d_y);
int fix_jacobian_rows(thcmb_ctx* c);
int build_blockdiag(thcmb_ctx* c);
int average_block(thcmb_ctx* c, double* db36);
bool scaling_compute(const thcmb_ctx* c, const double* db36, double* row_scaling, double* col_scaling);
int intcond_scaling(const thcmb_ctx* c, double* val, int* ind);
int apply_blockdiag(thcmb_ctx* c, const double* x, double* y);
int gather_cells(thcmb_ctx* c, const double* in, double* out);
int scatter_cells(thcmb_ctx* c, const double* in, double* out);
int spmv_compact(thcmb_ctx* c, const double* xc, double* yc);   // incl. the LL halo push on more than one rank and the integral row
int spmv_compact_rows(thcmb_ctx* c, const double* xc, double* yc, unsigned long long seq);   // the SpMV alone: the halo of exchange `seq` was pushed by the caller
unsigned long long scale_precon_push(thcmb_ctx* c, const double* w, const double* d_nrm2, double* v, double* z, double* d_nrm_out,
                                     int nv, double* const* vecs, const double* d_h2, const int* d_flag, const int* d_flag2);
bool compact_possible(const thcmb_ctx* c);
double land_nonzero_global(thcmb_ctx* c, const double* x);
int apply_blockdiag_compact(thcmb_ctx* c, const double* x, double* y);
double* pool_vec(thcmb_ctx* c, size_t idx);
double* work_vec(thcmb_ctx* c, int which);
}  // namespace thcm

namespace thcm {
enum KernelId { KID_ASM_RHS = 0, KID_ASM_JAC, KID_ASM_COUNT, KID_ASM_CRS, KID_SCAN, KID_SPMV, KID_DOT, KID_MGS, KID_AXPBY,
                KID_AXPY_DEV, KID_SCALE, KID_COPY, KID_FILL, KID_PRECON_BUILD, KID_PRECON_APPLY, KID_HALO_PACK, KID_HALO_UNPACK,
                KID_MULTIDOT, KID_MULTIAXPY, KID_SECOND_UPDATE, KID_COUNT };
struct ProfScope {   // records an event pair around one kernel launch when profiling is on
    thcmb_ctx* c; bool on;
    ProfScope(thcmb_ctx* c_, int kid) : c(c_), on(c_->prof_on && c_->prof_kid.size() < 60000) {
        if (!on) return;
        size_t i = c->prof_kid.size();
        while (c->prof_ev.size() < 2 * (i + 1)) { cudaEvent_t e; cudaEventCreate(&e); c->prof_ev.push_back(e); }
        c->prof_kid.push_back(kid);
        cudaEventRecord(c->prof_ev[2 * i], c->stream);
    }
    ~ProfScope() { if (on) cudaEventRecord(c->prof_ev[2 * (c->prof_kid.size() - 1) + 1], c->stream); }
};
}  // namespace thcm

#define THCM_CUDA(call)                                                                           \
    do {                                                                                          \
        cudaError_t e__ = (call);                                                                 \
        if (e__ != cudaSuccess)                                                                   \
            thcm::fatal(std::string("CUDA error ") + cudaGetErrorString(e__) + " at " + __FILE__ + ":" + \
                        std::to_string(__LINE__));                                                \
    } while (0)
