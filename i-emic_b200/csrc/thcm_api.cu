// =============================================================================
// thcm_api.cu -- the C ABI of libthcm_b200.so (include/thcm_b200.h):
//   * handle-based device API thcmb_* (context, residual, Jacobian, SpMV, vector kernels)
//   * host drivers of the Krylov solvers: GMRES (GMRESSolver.H:81-255) and IDR(s) (IDRSolver.H:109-340)
//   * the gfortran-mangled B1 symbols THCM.C binds (rhs_, matrix_, init_, setparcs_, ...)
// Host code only orchestrates; all arithmetic on the timed path runs in CUDA kernels.  There is
// no CPU fallback: without a usable device every entry point fails loudly.
// =============================================================================
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <memory>
#include "thcm_internal.h"
#include "thcm_slots.h"

using namespace thcm;

// ---- weak defaults of the callbacks the reference defines in C++ (THCM.C:2653,2690; GlobalDefinitions.C:145,154)
extern "C" {
__attribute__((weak)) void thcm_throw_error_(char* msg) { fprintf(stderr, "%s\n", msg); abort(); }
__attribute__((weak)) void timer_start_(const char*) {}
__attribute__((weak)) void timer_stop_(const char*) {}
// standalone default of the callback of forcing.F90:452-464: the one-rank integral over the sub-domain of the global instance
// (the reference's definition, THCM.C:2653-2686, sums the ranks' real cells with MPI instead)
void thcm_forcing_integral_default(double* field, double* y, int* landm, double* out);
__attribute__((weak)) void thcm_forcing_integral_(double* field, double* y, int* landm, double* out) {
    thcm_forcing_integral_default(field, y, landm, out);
}
}

namespace {

void require_device(int device) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0)
        fatal("no CUDA device available: the THCM B200 path has no CPU fallback (" + std::string(cudaGetErrorString(e)) + ")");
    if (device >= n) fatal("CUDA device ordinal out of range");
    THCM_CUDA(cudaSetDevice(device));
}

template <class T> void upload(T*& d, const std::vector<T>& h) {
    if (d) { cudaFree(d); d = nullptr; }
    if (h.empty()) return;
    THCM_CUDA(cudaMalloc(&d, sizeof(T) * h.size()));
    THCM_CUDA(cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
}

void upload_params(thcmb_ctx* c) {
    upload(c->d_jt, c->jt_host);
    upload(c->d_kt, c->kt_host);
    upload(c->d_frc, c->frc_local);
    upload(c->d_jrec, c->jrec_host);
    upload(c->d_krec, c->krec_host);
    upload(c->d_msi, c->msi_local);
    upload(c->d_cob, c->cob_local);
}

// mass diagonal of the rows THCM::evaluate replaces (integral condition, pressure Dirichlet rows) = 0 (THCM.C:1220-1240)
void zero_cob_of_fixed_rows(thcmb_ctx* c) {
    if (c->ic_on && c->ic_lrow >= 0) c->cob_local[(size_t)c->ic_lrow] = 0.0;
    if (c->pfix_on) for (int q = 0; q < 2; q++) if (c->pfix_lrow[q] >= 0) c->cob_local[(size_t)c->pfix_lrow[q]] = 0.0;
}

void refresh_params(thcmb_ctx* c) {  // forcing + lin (usrc.F90:178-179)
    compute_forcing(c);
    compute_tables(c);
    compute_cob(c);
    zero_cob_of_fixed_rows(c);
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    upload_params(c);
}

void build_static(thcmb_ctx* c) {
    std::vector<uint32_t> nbmask; std::vector<uint8_t> surf, uvlive; std::vector<int> send_idx, recv_slot;
    build_static_host(c, nbmask, surf, uvlive, send_idx, recv_slot);
    upload(c->d_nbmask, nbmask); upload(c->d_surf, surf); upload(c->d_uvlive, uvlive);
    std::vector<TileDesc> tdesc;
    build_tile_descs(c, nbmask, surf, uvlive, tdesc);
    upload(c->d_tdesc, tdesc);
    {   // tiles the Jacobian kernels revisit on every assembly: all but the all-LAND ones (state-independent identity rows)
        std::vector<int> active;
        for (size_t q = 0; q < tdesc.size(); q++) if (!(tdesc[q].flags & 4u)) active.push_back((int)q);
        upload(c->d_active_tiles, active); c->n_active_tiles = (int)active.size();
        c->land_tiles_written = false;   // d_val is rebuilt below
    }
    upload(c->d_rowptr, c->rowptr_host); upload(c->d_col, c->col_host);
    upload(c->d_ocell, c->ocell_host); upload(c->d_ccell, c->ccell_host); c->n_ocell = (int)c->ocell_host.size();
    upload(c->d_colc, c->colc_host); upload(c->d_send_cidx, c->send_cidx_host);
    std::vector<int>().swap(c->colc_host);   // as large as the graph's column array: the device copy is the one that is used
    {   // LAND cells: identity rows of the Jacobian whatever the state (bit 4 of the neighbour mask = centre not OCEAN)
        std::vector<uint8_t> landcell(nbmask.size());
        for (size_t q = 0; q < nbmask.size(); q++) landcell[q] = (uint8_t)((nbmask[q] >> 4) & 1u);
        upload(c->d_landcell, landcell);
    }
    upload(c->d_send_idx, send_idx); upload(c->d_recv_slot, recv_slot);
    upload(c->d_send_dst, c->send_dst_host); upload(c->d_send_peer, c->send_peer_host);
    if (c->d_val) cudaFree(c->d_val);
    THCM_CUDA(cudaMalloc(&c->d_val, sizeof(double) * (size_t)std::max<long long>(c->gnnz, 1)));
    THCM_CUDA(cudaMemset(c->d_val, 0, sizeof(double) * (size_t)std::max<long long>(c->gnnz, 1)));
    size_t nh = (size_t)NUN * std::max(c->blk.nhalo_cells(), 1);
    for (double** p : {&c->d_sendbuf, &c->d_recvbuf}) { if (*p) cudaFree(*p); *p = nullptr; }
    if (!c->halo_p2p) {   // (with the P2P halo push the buffers live in the IPC-shared allocation and keep their size)
        if (c->d_halo) cudaFree(c->d_halo);
        THCM_CUDA(cudaMalloc(&c->d_halo, sizeof(double) * nh));
        THCM_CUDA(cudaMemset(c->d_halo, 0, sizeof(double) * nh));
    }
    THCM_CUDA(cudaMalloc(&c->d_sendbuf, sizeof(double) * NUN * (size_t)std::max(c->nsend_cells, 1)));
    THCM_CUDA(cudaMalloc(&c->d_recvbuf, sizeof(double) * NUN * (size_t)std::max(c->nrecv_cells, 1)));
    upload_class_tables(class_tables(c->blk.periodic));
    {   // where the 6x6 in-cell block sits inside the sorted graph rows, per boundary class (block-diagonal preconditioner)
        const ClassTables& ct = class_tables(c->blk.periodic);
        std::vector<signed char> cpos((size_t)NCLASS * NUN * NUN, (signed char)-1);
        for (int cls = 0; cls < NCLASS; cls++) for (int R = 1; R <= NUN; R++) for (int C = 1; C <= NUN; C++) {
            const int sl = slot_of(R, 5, C);
            if (sl >= 0) cpos[((size_t)cls * NUN + (R - 1)) * NUN + (C - 1)] = (signed char)ct.pos[cls][ROW_OFF[R - 1] + sl];
        }
        upload(c->d_cpos, cpos);
    }
}

// usrc.F90:353-418 on one instance: the per-cell data, tile descriptors, compact maps and graph-dependent tables follow the new mask
void set_landmask_ctx(thcmb_ctx* c, const int* landm, bool reinit) {
    apply_landmask_rules(c, landm, true);
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    build_static(c);
    if (reinit) { vmix_init(c); refresh_params(c); }   // usrc.F90:410-415
    else {   // no re-initialisation of forcing / lin asked: the mass diagonal and the rows `boundaries` masks in Frc still follow the mask
        compute_cob(c); zero_cob_of_fixed_rows(c);
        mask_forcing_rows(c);
        upload(c->d_cob, c->cob_local); upload(c->d_frc, c->frc_local);
    }
}
void stage_begin(thcmb_ctx* c) { THCM_CUDA(cudaEventRecord(c->ev0, c->stream)); }
void stage_end(thcmb_ctx* c, const char* label) {  // device time under the reference's profile labels
    THCM_CUDA(cudaEventRecord(c->ev1, c->stream));
    THCM_CUDA(cudaEventSynchronize(c->ev1));
    float ms = 0;
    THCM_CUDA(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->stage_ms[label] = ms;
}

}  // namespace

namespace thcm {
// slot `idx` of the Krylov vector pool, allocated on first use (a solve with restart = 400 that converges in 30 iterations
// touches 30 slots, not 800)
double* pool_vec(thcmb_ctx* c, size_t idx) {
    if (c->krylov_pool.size() <= idx) c->krylov_pool.resize(idx + 1, nullptr);
    if (!c->krylov_pool[idx]) THCM_CUDA(cudaMalloc(&c->krylov_pool[idx], sizeof(double) * (size_t)c->blk.ndim()));
    return c->krylov_pool[idx];
}
// dedicated work vectors that must never alias a pool slot: 0 = dx of thcmb_newton_step, 1 / 2 = compact b / x of thcmb_gmres
double* work_vec(thcmb_ctx* c, int which) {
    if (!c->d_work[which]) THCM_CUDA(cudaMalloc(&c->d_work[which], sizeof(double) * (size_t)c->blk.ndim()));
    return c->d_work[which];
}
}  // namespace thcm

extern "C" {

void thcmb_default_settings(thcmb_settings* s) {
    memset(s, 0, sizeof(*s));
    s->N = s->M = s->L = 0;
    s->hdim = 4000.0; s->qz = 1.0;                         // usr.F90:45-46, THCM.C defaults
    s->ih = 0; s->vmix = 0; s->tap = 1; s->rho_mixing = 0; s->coriolis_on = 1;
    s->TRES = 1; s->SRES = 1; s->iza = 2; s->ite = 1; s->its = 1; s->coupled_T = 0; s->coupled_S = 0; s->forcing_type = 0;
    s->alphaT = 1.0e-04; s->alphaS = 7.6e-04;               // usr.F90:143-144
    s->rank = 0; s->nranks = 1; s->device = 0;
    s->balance = 0;                                         // the reference's uniform cut lines
}

thcmb_ctx* thcmb_create(const thcmb_settings* s, const int* landm_global) {
    require_device(s->device);
    thcmb_ctx* c = new thcmb_ctx();
    c->s = *s; c->device = s->device;
    if (!setup_block(c, landm_global)) fatal("domain decomposition produced an empty block");
    THCM_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    THCM_CUDA(cudaEventCreate(&c->ev0)); THCM_CUDA(cudaEventCreate(&c->ev1));
    for (auto& e : c->ev_slot) THCM_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    build_grid(c);
    stpnt(c);
    apply_landmask_rules(c, landm_global, false);
    vmix_init(c);      // usrc.F90:133
    init_surface_fields(c);   // allocate_usr + atmos_coef (usrc.F90:118-134)
    c->n_asm_blocks = asm_block_count(c->blk);
    if (const char* e = getenv("THCM_ASM_PIPE")) c->asm_pipe = atoi(e);
    if (const char* e = getenv("THCM_KRYLOV_COMPACT")) c->krylov_compact = atoi(e);
    c->fused_cgs2 = 2;   // 2 = L2-tiled kernel (76.2 ms per Newton step at 1 degree), 1 = shared-memory parking (78.7), 0 = unfused (82.5)
    if (const char* e = getenv("THCM_FUSED_CGS2")) c->fused_cgs2 = atoi(e);
    THCM_CUDA(cudaMalloc(&c->d_partial, sizeof(double) * 4096));
    THCM_CUDA(cudaMalloc(&c->d_scalars, sizeof(double) * 4096));
    THCM_CUDA(cudaMemset(c->d_scalars, 0, sizeof(double) * 4096));
    THCM_CUDA(cudaMalloc(&c->d_counter, sizeof(unsigned int)));
    THCM_CUDA(cudaMemset(c->d_counter, 0, sizeof(unsigned int)));
    THCM_CUDA(cudaMallocHost(&c->h_scalars, sizeof(double) * 4096));
    THCM_CUDA(cudaMalloc(&c->d_blockcnt, sizeof(int) * (size_t)(c->n_asm_blocks + 1)));
    THCM_CUDA(cudaMalloc(&c->d_un, sizeof(double) * (size_t)c->blk.ndim()));
    THCM_CUDA(cudaMalloc(&c->d_tmp, sizeof(double) * (size_t)c->blk.ndim()));
    build_static(c);
    refresh_params(c);
    return c;
}

void thcmb_destroy(thcmb_ctx* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    p2p_close(c);
    nccl_destroy(c);
    for (void* p : {(void*)c->d_jt, (void*)c->d_kt, (void*)c->d_nbmask, (void*)c->d_surf, (void*)c->d_uvlive, (void*)c->d_frc,
                    (void*)c->d_rowptr, (void*)c->d_col, (void*)c->d_val, (void*)(c->halo_p2p ? nullptr : c->d_halo), (void*)c->d_sendbuf,
                    (void*)c->d_recvbuf, (void*)c->d_send_dst, (void*)c->d_send_peer, (void*)c->d_halo_counter,
                    (void*)c->d_send_idx, (void*)c->d_recv_slot, (void*)c->d_un, (void*)c->d_tmp, (void*)c->d_partial,
                    (void*)c->d_scalars, (void*)c->d_counter, (void*)c->d_blockcnt, (void*)c->d_minv, (void*)c->d_tdesc, (void*)c->d_jrec,
                    (void*)c->d_krec, (void*)c->d_msi, (void*)c->d_cob, (void*)c->d_iccoeff,
                    (void*)c->d_flags, (void*)c->d_mdpartial, (void*)c->d_tilectr, (void*)c->d_cls, (void*)c->d_landcell, (void*)c->d_ocell, (void*)c->d_ccell, (void*)c->d_colc, (void*)c->d_send_cidx,
                    (void*)c->d_iccoeff_c, (void*)c->d_active_tiles, (void*)c->d_cpos, (void*)c->d_halo_plain_c})
        if (p) cudaFree(p);
    for (double* p : c->krylov_pool) if (p) cudaFree(p);
    for (double* p : c->d_work) if (p) cudaFree(p);
    if (c->h_scalars) cudaFreeHost(c->h_scalars);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    for (auto e : c->ev_slot) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c;
}

void thcmb_local_block(const thcmb_ctx* c, int* i0, int* j0, int* n0, int* m0, int* npN, int* npM) {
    *i0 = c->blk.i0; *j0 = c->blk.j0; *n0 = c->blk.n0; *m0 = c->blk.m0; *npN = c->blk.npN; *npM = c->blk.npM;
}
int thcmb_ndim_local(const thcmb_ctx* c) { return c->blk.ndim(); }
long long thcmb_graph_nnz(const thcmb_ctx* c) { return c->gnnz; }
void thcmb_get_graph(const thcmb_ctx* c, int* rowptr, int* col) {
    memcpy(rowptr, c->rowptr_host.data(), sizeof(int) * c->rowptr_host.size());
    memcpy(col, c->col_host.data(), sizeof(int) * c->col_host.size());
}
void thcmb_tile_counts(const thcmb_ctx* c, int* ntiles, int* nactive) { *ntiles = c->n_asm_blocks; *nactive = c->n_active_tiles; }
int thcmb_halo_size(const thcmb_ctx* c) { return NUN * c->blk.nhalo_cells(); }
void thcmb_halo_gids(const thcmb_ctx* c, int* gids) { memcpy(gids, c->halo_gid.data(), sizeof(int) * c->halo_gid.size()); }
void thcmb_local_gids(const thcmb_ctx* c, int* gids) { memcpy(gids, c->local_gid.data(), sizeof(int) * c->local_gid.size()); }

void thcmb_set_par(thcmb_ctx* c, int idx, double val) {
    if (idx >= 1 && idx <= NPAR) c->par[idx] = val;
    refresh_params(c);
}
/* coupled mode: m_inserts::insert_* on GLOBAL N*M fields, set_atmos_parameters / set_seaice_parameters (usrc.F90:254-350) */
void thcmb_insert_field(thcmb_ctx* c, int which, const double* field_global) { insert_surface_field(c, which, field_global); }
void thcmb_set_atmos_parameters(thcmb_ctx* c, const double* pars18) { set_atmos_parameters(c, pars18); refresh_params(c); }
void thcmb_set_seaice_parameters(thcmb_ctx* c, const double* pars7) { set_seaice_parameters(c, pars7); refresh_params(c); }
double thcmb_get_par(const thcmb_ctx* c, int idx) { return (idx >= 1 && idx <= NPAR) ? c->par[idx] : 0.0; }
void thcmb_get_forcing(thcmb_ctx* c, double* frc) {
    const std::vector<double>& f = c->frc_masked ? c->frc_local : c->frc_raw;
    memcpy(frc, f.data(), sizeof(double) * f.size());
}
void thcmb_get_cob(thcmb_ctx* c, double* cob) { memcpy(cob, c->cob_local.data(), sizeof(double) * c->cob_local.size()); }

/* theta time stepping (ThetaModel.H:87-165): d_F holds F(u_{n+1}) on entry and the theta-method residual on return; the Jacobian
 * variant adds -M / (theta dt) to the diagonal of the stored Jacobian (call after thcmb_jacobian_dev, before the solve) */
int thcmb_theta_rhs_dev(thcmb_ctx* c, double theta, double dt, const double* d_state, const double* d_old_state, const double* d_old_rhs,
                        double* d_F) {
    if (!(dt > 0.0)) fatal("thcmb_theta_rhs_dev: the time step must be positive");
    return theta_rhs(c, c->blk.ndim(), theta, dt, d_state, d_old_state, d_old_rhs, c->d_cob, d_F);
}
int thcmb_apply_mass_dev(thcmb_ctx* c, const double* d_v, double* d_out) {
    if (d_v == d_out) { set_error("thcmb_apply_mass_dev: v and out must be different vectors"); return -1; }
    return mass_apply(c, c->blk.ndim(), c->d_cob, d_v, d_out);
}
int thcmb_theta_jacobian_dev(thcmb_ctx* c, double theta, double dt) {
    if (theta == 0.0) return 0;                                     // ThetaModel.H:126-127
    if (!(dt > 0.0)) fatal("thcmb_theta_jacobian_dev: the time step must be positive");
    return theta_jacobian(c, theta, dt, c->d_cob);
}
int thcmb_nccl_unique_id(void* id128) { return nccl_unique_id(id128); }
int thcmb_nccl_init(thcmb_ctx* c, const void* id128) { return nccl_init(c, id128); }
int thcmb_p2p_local_handle(thcmb_ctx* c, void* handle64) { return p2p_local_handle(c, handle64); }
int thcmb_p2p_open(thcmb_ctx* c, const void* handles_all) { return p2p_open(c, handles_all); }
void thcmb_set_ortho(thcmb_ctx* c, int mode) { c->gmres_ortho = mode; }
/* THCM::RecomputeScaling (THCM.C:1781-1834): average diagonal block of the stored Jacobian (device reduction), summed over the
 * ranks and divided by their number, m_scaling::compute, inverted into Trilinos' convention, T and S scaled alike */
int thcmb_recompute_scaling(thcmb_ctx* c, double* row_scaling, double* col_scaling, double* db36_out) {
    double ldb[36], gdb[36];
    average_block(c, ldb);
    memcpy(gdb, ldb, sizeof(gdb));
    if (c->blk.nranks > 1) {
        double* d = c->d_scalars + 3900;
        THCM_CUDA(cudaMemcpyAsync(d, ldb, sizeof(ldb), cudaMemcpyHostToDevice, c->stream));
        allreduce_dev(c, d, 36);
        THCM_CUDA(cudaMemcpyAsync(gdb, d, sizeof(gdb), cudaMemcpyDeviceToHost, c->stream));
        THCM_CUDA(cudaStreamSynchronize(c->stream));
    }
    for (int q = 0; q < 36; q++) gdb[q] /= c->blk.nranks;
    if (db36_out) memcpy(db36_out, gdb, sizeof(gdb));
    const int n = c->blk.ndim();
    const bool ok = scaling_compute(c, gdb, row_scaling, col_scaling);
    for (int i = 0; i < n; i++) { row_scaling[i] = 1.0 / row_scaling[i]; col_scaling[i] = 1.0 / col_scaling[i]; }
    for (int i = TT - 1; i < n; i += NUN) {
        double mean = 0.5 * (row_scaling[i] + row_scaling[i + 1]);
        row_scaling[i] = mean; row_scaling[i + 1] = mean;
        mean = 0.5 * (col_scaling[i] + col_scaling[i + 1]);
        col_scaling[i] = mean; col_scaling[i + 1] = mean;
    }
    return ok ? 0 : 1;
}
/* THCM::getIntCondCoeff (THCM.C:2608-2637): coefficients of the salinity integral condition on the owned S rows (0 elsewhere);
 * returns their 1-norm = the local part of the total volume */
double thcmb_intcond_coeff(const thcmb_ctx* c, double* coeff) {
    const int n = c->blk.ndim();
    for (int i = 0; i < n; i++) coeff[i] = 0.0;
    std::vector<double> val((size_t)c->blk.ncell()); std::vector<int> ind((size_t)c->blk.ncell());
    const int len = intcond_scaling(c, val.data(), ind.data());
    double vol = 0.0;
    for (int q = 0; q < len; q++) { coeff[ind[q] - 1] = val[q]; vol += std::fabs(val[q]); }
    return vol;
}
/* The row replacements THCM::evaluate makes above the Fortran core.  Local row of a global cell (0-based i, j, k) or -1. */
static int local_row_of(const thcmb_ctx* c, int gi, int gj, int k, int XX) {
    const Block& b = c->blk;
    const int li = gi - b.i0, lj = gj - b.j0;
    if (li < 0 || li >= b.n0 || lj < 0 || lj >= b.m0) return -1;
    return NUN * ((k * b.m0 + lj) * b.n0 + li) + XX - 1;
}
static void refresh_fix_rows(thcmb_ctx* c) {
    compute_cob(c);
    zero_cob_of_fixed_rows(c);
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    upload(c->d_cob, c->cob_local);
}
/* THCM.C:653-697: the salinity integral condition replaces the S row of cell (Nic, Mic, L-1) (0-based; -1 = N-1 / M-1); only
 * with SRES = 0, and the cell's surface point must be ocean.  sign = "Salinity Integral Sign" (+-1, default -1) */
void thcmb_enable_intcond(thcmb_ctx* c, int Nic, int Mic, int sign) {
    const thcmb_settings& s = c->s;
    if (s.SRES != 0) fatal("the salinity integral condition needs SRES = 0 (Restoring Salinity Profile = 0)");
    if (sign != 1 && sign != -1) fatal("Salinity Integral Sign must be +1 or -1");     // THCM.C:261
    if (Nic < 0) Nic = s.N - 1;
    if (Mic < 0) Mic = s.M - 1;
    if (Nic >= s.N || Mic >= s.M) fatal("Integral row coordinates outside the domain");
    if (c->landm[(size_t)(Nic + 1) + (size_t)(s.N + 2) * ((Mic + 1) + (size_t)(s.M + 2) * s.L)] != 0)
        fatal("Integral row coordinates (" + std::to_string(Nic) + "," + std::to_string(Mic) + ") give a land point! Please give better coordinates");
    c->ic_on = true; c->ic_sign = sign;
    c->ic_grow = NUN * (((s.L - 1) * s.M + Mic) * s.N + Nic) + SS - 1;
    c->ic_lrow = local_row_of(c, Nic, Mic, s.L - 1, SS);
    std::vector<double> coeff((size_t)c->blk.ndim());
    thcmb_intcond_coeff(c, coeff.data());
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    upload(c->d_iccoeff, coeff);
    if (c->d_iccoeff_c) { cudaFree(c->d_iccoeff_c); c->d_iccoeff_c = nullptr; }
    THCM_CUDA(cudaMalloc(&c->d_iccoeff_c, sizeof(double) * (size_t)std::max(NUN * c->n_ocell, 1)));
    if (c->n_ocell > 0) gather_cells(c, c->d_iccoeff, c->d_iccoeff_c);
    refresh_fix_rows(c);
}
/* THCM::setLandMask(global mask, init) (THCM.C:1362-1392) for a handle: with init the instance is rebuilt on the new GLOBAL mask the way
 * set_landmask_ does with reinit = 1 (same decomposition and cut lines; a preconditioner built before must be rebuilt by the caller);
 * without init only m_global's copy changes in the reference, i.e. nothing here.  An enabled integral condition keeps its cell and gets
 * fresh coefficients; -1 (thcmb_last_error) when that cell is land in the new mask. */
int thcmb_set_landmask(thcmb_ctx* c, const int* landm_global, int init) {
    if (!init) return 0;
    const thcmb_settings& s = c->s;
    int Nic = -1, Mic = -1;
    if (c->ic_on) {
        const int cell = c->ic_grow / NUN;
        Nic = cell % s.N; Mic = (cell / s.N) % s.M;
        if (landm_global[(size_t)(Nic + 1) + (size_t)(s.N + 2) * ((Mic + 1) + (size_t)(s.M + 2) * s.L)] != 0) {
            set_error("thcmb_set_landmask: the integral-condition cell (" + std::to_string(Nic) + "," + std::to_string(Mic) + ") is land in the new mask");
            return -1;
        }
    }
    set_landmask_ctx(c, landm_global, true);
    if (c->ic_on) thcmb_enable_intcond(c, Nic, Mic, c->ic_sign);
    return 0;
}
/* THCM::setIntCondCorrection (THCM.C:2078-2097): the salinity integral of d_vec becomes the target of the condition */
double thcmb_set_intcond_correction(thcmb_ctx* c, const double* d_vec) {
    if (!c->ic_on) fatal("setIntCondCorrection: the integral condition is not enabled");
    c->ic_correction = thcmb_dot(c, c->blk.ndim(), c->d_iccoeff, d_vec);
    return c->ic_correction;
}
/* "Fix Pressure Points" (THCM.C:749-757, 2258-2296): Dirichlet rows for the pressure of the cells (N-1, M-1, L-1) and (N-2, M-1, L-1) */
void thcmb_fix_pressure_points(thcmb_ctx* c, int on) {
    const thcmb_settings& s = c->s;
    c->pfix_on = on != 0;
    for (int q = 0; q < 2; q++) {
        c->pfix_grow[q] = NUN * (((s.L - 1) * s.M + (s.M - 1)) * s.N + (s.N - 1 - q)) + PP - 1;
        c->pfix_lrow[q] = local_row_of(c, s.N - 1 - q, s.M - 1, s.L - 1, PP);
    }
    refresh_fix_rows(c);
}
int thcmb_intcond_row(const thcmb_ctx* c) { return c->ic_on ? c->ic_grow : -1; }
void thcmb_set_vmix_fix(thcmb_ctx* c, int fix) { c->vmix_fix = fix; }   /* m_mix::set_vmix_fix, mix.F90:52-59 */
void thcmb_get_vmix_flags(const thcmb_ctx* c, int* out4) { out4[0] = c->vmix_flag; out4[1] = c->vmix_temp; out4[2] = c->vmix_salt; out4[3] = c->vmix_fix; }

int thcmb_halo_exchange(thcmb_ctx* c, const double* d_x) { return halo_exchange(c, d_x); }
/* Ocean::getBlock(Atmosphere / SeaIce) (Ocean.C:1603-1810): the ocean block's coupling to the other models (host code at coupling
 * frequency: surface points only) */
int thcmb_ocean_block_atmosphere(thcmb_ctx* c, double albed, const double* pdist, const int* colT, const int* colQ, const int* colA,
                                 const int* colP, int* beg, int* jco, double* co) {
    return ocean_block_atmosphere(c, albed, pdist, colT, colQ, colA, colP, beg, jco, co);
}
int thcmb_ocean_block_seaice(thcmb_ctx* c, const double* un_host, const int* colQ, const int* colM, const int* colG, int* beg, int* jco,
                             double* co) {
    return ocean_block_seaice(c, un_host, colQ, colM, colG, beg, jco, co);
}

// vmix_control (mix_imp.f:139-169), called from rhs / matrix when Mixing = 2 and the partition is not fixed
// (usrc.F90:496, 558): which of the T and S fields are non-zero decides whether their mixing terms are on
static void vmix_control(thcmb_ctx* c, const double* d_un) {
    if (c->vmix_flag < 2 || c->vmix_fix != 0) return;
    double nrm2[2];
    field_sumsq(c, d_un, nrm2);
    vmix_set_flags(c, std::sqrt(nrm2[0]) > 1.0e-12 ? 1 : 0, std::sqrt(nrm2[1]) > 1.0e-12 ? 1 : 0);
    compute_tables(c);   // only the mixing switches of c->tab change (kernel arguments, no upload needed)
}

int thcmb_rhs_dev(thcmb_ctx* c, const double* d_un, double* d_B) {
    c->frc_masked = true;
    vmix_control(c, d_un);
    halo_exchange(c, d_un);
    return launch_assembly(c, MODE_RHS, d_un, d_B, nullptr, nullptr, nullptr);
}
int thcmb_residual_dev(thcmb_ctx* c, const double* d_un, double* d_F) {
    c->frc_masked = true;
    vmix_control(c, d_un);
    halo_exchange(c, d_un);
    launch_assembly(c, MODE_RHS | 0x100, d_un, d_F, nullptr, nullptr, nullptr);
    return fix_residual_rows(c, d_un, d_F);   // integral condition / pressure Dirichlet rows (THCM.C:1013-1041), no-op unless enabled
}
int thcmb_jacobian_dev(thcmb_ctx* c, const double* d_un) {
    c->frc_masked = true;
    vmix_control(c, d_un);
    halo_exchange(c, d_un);
    launch_assembly(c, MODE_JAC_GRAPH, d_un, nullptr, nullptr, nullptr, nullptr);
    return fix_jacobian_rows(c);              // THCM.C:1164-1172, no-op unless enabled
}
const double* thcmb_jacobian_values(const thcmb_ctx* c) { return c->d_val; }
const int* thcmb_graph_rowptr_dev(const thcmb_ctx* c) { return c->d_rowptr; }
int thcmb_scatter_values_dev(thcmb_ctx* c, long long n, const int* d_slot, const double* d_in, double* d_out) {
    if (n < 0 || (n > 0 && (!d_slot || !d_in || !d_out))) { set_error("thcmb_scatter_values_dev: null argument"); return -1; }
    return scatter_slots(c, n, d_slot, d_in, d_out);
}
const int* thcmb_graph_col_dev(const thcmb_ctx* c) { return c->d_col; }

long long thcmb_jacobian_crs_dev(thcmb_ctx* c, const double* d_un, int* d_begA, int* d_jcoA, double* d_coA) {
    c->frc_masked = true;
    vmix_control(c, d_un);
    halo_exchange(c, d_un);
    launch_assembly(c, MODE_JAC_COUNT, d_un, nullptr, nullptr, nullptr, nullptr);
    scan_block_counts(c);
    launch_assembly(c, MODE_JAC_CRS, d_un, nullptr, d_begA, d_jcoA, d_coA);
    int last = 0;
    THCM_CUDA(cudaMemcpyAsync(&last, d_begA + c->blk.ndim(), sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    return (long long)last - 1;
}

int thcmb_spmv_dev(thcmb_ctx* c, const double* d_x, double* d_y) {
    halo_exchange(c, d_x);
    spmv(c, c->blk.ndim(), c->d_rowptr, c->d_col, c->d_val, d_x, c->d_halo, c->blk.ndim(), d_y);
    return fix_spmv_rows(c, d_x, d_y);        // the dense integral-condition row: one more dot product, only when enabled
}
int thcmb_csr_spmv_dev(thcmb_ctx* c, int nrow, const int* d_rowptr, const int* d_col, const double* d_val, const double* d_x, double* d_y) {
    return spmv(c, nrow, d_rowptr, d_col, d_val, d_x, d_x, 0x7fffffff, d_y);
}

double thcmb_dot(thcmb_ctx* c, int n, const double* d_x, const double* d_y) {
    dot_dev(c, n, d_x, d_y, c->d_scalars);
    THCM_CUDA(cudaMemcpyAsync(c->h_scalars, c->d_scalars, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    return c->h_scalars[0];
}
double thcmb_nrm2(thcmb_ctx* c, int n, const double* d_x) { return std::sqrt(thcmb_dot(c, n, d_x, d_x)); }
int thcmb_axpby(thcmb_ctx* c, int n, double a, const double* d_x, double b, double* d_y) { return axpby(c, n, a, d_x, b, d_y); }
int thcmb_scale(thcmb_ctx* c, int n, double a, double* d_x) { return axpby(c, n, 0.0, d_x, a, d_x); }

int thcmb_build_precon(thcmb_ctx* c, int kind) {
    c->precon_kind = kind;
    if (kind == 1) return build_blockdiag(c);
    return kind == 0 ? 0 : -1;
}
int thcmb_apply_precon_dev(thcmb_ctx* c, const double* d_x, double* d_y) {
    if (c->precon_kind == 1) return apply_blockdiag(c, d_x, d_y);
    return copy(c, c->blk.ndim(), d_x, d_y);
}

void* thcmb_device_alloc(thcmb_ctx* c, long long bytes) {
    void* p = nullptr;
    THCM_CUDA(cudaSetDevice(c->device));
    THCM_CUDA(cudaMalloc(&p, (size_t)std::max<long long>(bytes, 8)));
    return p;
}
void thcmb_device_free(thcmb_ctx*, void* p) { if (p) cudaFree(p); }
int thcmb_h2d(thcmb_ctx* c, void* d, const void* h, long long bytes) {
    THCM_CUDA(cudaMemcpyAsync(d, h, (size_t)bytes, cudaMemcpyHostToDevice, c->stream));
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int thcmb_d2h(thcmb_ctx* c, void* h, const void* d, long long bytes) {
    THCM_CUDA(cudaMemcpyAsync(h, d, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream));
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}
int thcmb_sync(thcmb_ctx* c) { THCM_CUDA(cudaStreamSynchronize(c->stream)); return 0; }
void* thcmb_stream(thcmb_ctx* c) { return (void*)c->stream; }
long long thcmb_launch_count(const thcmb_ctx* c) { return c->launches; }
double thcmb_last_stage_ms(const thcmb_ctx* c, const char* label) {
    auto it = c->stage_ms.find(label);
    return it == c->stage_ms.end() ? -1.0 : it->second;
}

// =============================================================================
// GMRES -- the algorithm of src/gmressolver/GMRESSolver.H:81-255 (right / flexible preconditioning,
// modified Gram-Schmidt, Givens rotations, back substitution = minimiser scheme 'B').
// The MGS chain runs without host syncs: every projection coefficient stays in device memory and is
// consumed by the next fused (axpy + dot) kernel; one small D2H per iteration brings column i of H.
// =============================================================================
static void gen_rot(double& dx, double& dy, double& cs, double& sn) {  // GMRESSolver.H:258-279
    if (dy == 0.0) { cs = 1.0; sn = 0.0; }
    else if (std::abs(dy) > std::abs(dx)) { double t = dx / dy; sn = 1.0 / sqrt(1.0 + t * t); cs = t * sn; }
    else { double t = dy / dx; cs = 1.0 / sqrt(1.0 + t * t); sn = t * cs; }
}
static void app_rot(double& dx, double& dy, double& cs, double& sn) {  // GMRESSolver.H:282-290
    double t = cs * dx + sn * dy;
    dy = -sn * dx + cs * dy;
    dx = t;
}

int thcmb_gmres(thcmb_ctx* c, const double* d_b, double* d_x, double tol, int maxit, int restart, int flags, double* hist,
                int hist_cap, thcmb_krylov_result* res) {
    int n = c->blk.ndim();
    const bool prec = flags & 1, flexible = flags & 4, batched = flags & 8;
    const int m = restart;
    long long n_reorth = 0;
    // argument errors of the handle API are reported, not fatal (thcmb_last_error); the Fortran symbols keep the reference's abort semantics.
    // Batched orthogonalisation works through the basis in chunks of 64 vectors, so the reference's 500 Krylov vectors
    // (run/ocean/solver_params.xml) are fine; the limit is the pinned scalar buffer (two pipelined slots of 3 (m + 2) + 8 doubles)
    if (m < 1 || (batched ? 2 * (3 * (m + 18) + 8) > 3800 : m + 3 > 3800)) {
        set_error("thcmb_gmres: restart length " + std::to_string(m) + " outside the supported range (1 .. " + (batched ? "600" : "3796") + ")");
        if (res) { res->status = -1; res->iters = 0; res->resid = 0.0; res->nhist = 0; res->n_matvec = 0; }
        return -1;
    }
    // Ocean-only Krylov space (default; THCM_KRYLOV_COMPACT=0 or flag 16 switch it off): rows of LAND cells are identity rows, so when b
    // and the initial guess vanish on LAND -- checked, over all ranks -- every Krylov vector does; the caller's vectors are gathered once
    // and the solution is scattered back at the end
    double* const d_x_full = d_x;
    // (every rank takes the same branch: the test is global; a rank without LAND cells simply keeps all of its cells)
    bool compact = compact_possible(c) && !(flags & 16);
    if (compact && (land_nonzero_global(c, d_b) != 0.0 || land_nonzero_global(c, d_x) != 0.0)) compact = false;
    if (compact) {
        double* bc = work_vec(c, 1);
        double* xc = work_vec(c, 2);
        gather_cells(c, d_b, bc);
        gather_cells(c, d_x, xc);
        d_b = bc; d_x = xc;
        n = c->n_ocell * NUN;
    }
    auto applyA = [&](const double* v, double* out) { if (compact) spmv_compact(c, v, out); else thcmb_spmv_dev(c, v, out); };
    auto applyM = [&](const double* v, double* out) {
        if (!compact) thcmb_apply_precon_dev(c, v, out);
        else if (c->precon_kind == 1) apply_blockdiag_compact(c, v, out);
        else copy(c, n, v, out);
    };
    long long n_matvec = 0;
    int nh = 0;
    auto push = [&](double r) { if (hist && nh < hist_cap) hist[nh] = r; nh++; };
    // pool: 0 = r/tmp, 1 = tmp2, 2.. = V[0..m], then Z[0..m-1] when flexible+prec
    double* r = pool_vec(c, 0);
    double* tmp = pool_vec(c, 1);
    auto V = [&](int i) { return pool_vec(c, 2 + i); };
    auto Z = [&](int i) { return pool_vec(c, 2 + (m + 1) + i); };
    std::vector<std::vector<double>> H(m + 1, std::vector<double>(m, 0.0));
    std::vector<double> s(m + 1, 0.0), cs(m + 1, 0.0), sn(m + 1, 0.0), y;
    double normb = thcmb_nrm2(c, n, d_b);
    applyA(d_x, r); n_matvec++;
    axpby(c, n, 1.0, d_b, -1.0, r);  // r = b - A x
    double beta = thcmb_nrm2(c, n, r);
    if (normb == 0.0) normb = 1;
    double resid = beta / normb;
    int iter = 0, status = 1;
    if (resid <= tol) { status = 0; }
    else {
        while (iter <= maxit) {
            beta = thcmb_nrm2(c, n, r);
            copy(c, n, r, V(0));
            axpby(c, n, 0.0, V(0), 1.0 / beta, V(0));  // r.scale(1/beta); V[0] = r
            std::fill(s.begin(), s.end(), 0.0);
            s[0] = beta;
            int i, space = -1;
            bool converged = false;
            // ---- one Arnoldi step on the device: w = A M^-1 v_i, orthogonalised against V[0..i], V[i+1] = w / ||w||, and the
            //      column of H copied to pinned host slot `slot` (async).  MGS follows GMRESSolver.H:177-187 statement by
            //      statement; batched = classical Gram-Schmidt with the DGKS criterion (Belos "DGKS", Ocean.C:977-1024).
            const int S = std::max(80, (m + 2 + 15) / 16 * 16), HSLOT = 3 * S + 8;   // batched dh layout: [0,S) h1 + ww_old, [S,2S) ww_new, [2S,3S) h2, [3S] ||w||^2, [3S+1] ||w||
            // fused head (compact space, block-diagonal preconditioner, flexible, batched): the orthogonalisation works in a buffer of
            // its own (wbuf); the next step's first kernel turns it into V[i+1] = w / ||w|| AND Z[i+1] = M^-1 V[i+1] AND pushes the halo of
            // Z[i+1] -- scale_invsqrt + blockdiag_apply + halo push in one launch.  V[i+1] of the LAST step of a cycle is never formed:
            // nothing reads it (the update uses Z[0..i], GMRESSolver.H:293-313).
            const bool fuse_head = batched && compact && prec && flexible && c->precon_kind == 1 && c->d_minv && !getenv("THCM_NO_FUSED_HEAD");
            double* const wbuf = fuse_head ? work_vec(c, 3) : nullptr;
            // ... and the second Gram-Schmidt update moves into that kernel too, its norm taken from Pythagoras (RedEpilogue::flag2_out):
            // two all-reduces and four kernels per iteration instead of three and five
            const bool cgs2_kernel = (c->p2p_on || c->blk.nranks == 1) && c->fused_cgs2 && (n & 1) == 0;
            const bool pyth_norm = fuse_head && cgs2_kernel && !getenv("THCM_EXPLICIT_NORM");
            auto enqueue = [&](int i, int slot) {
                double* w = fuse_head ? wbuf : V(i + 1);
                if (fuse_head && i > 0) {
                    // the previous step's second update (against V[0..i-1], coefficients h2 in dh[2S..], DGKS flag d_flags[0]) rides
                    // in this kernel unless the explicit path took it (d_flags[1])
                    std::vector<double*> vprev(i);
                    for (int k = 0; k < i; k++) vprev[k] = V(k);
                    const bool pyth = pyth_norm && i <= 64;
                    const unsigned long long seq = scale_precon_push(c, wbuf, c->d_scalars + 3 * S, V(i), Z(i), nullptr, i, vprev.data(),
                                                                     c->d_scalars + 2 * S, pyth ? c->d_flags : nullptr, c->d_flags + 1);
                    spmv_compact_rows(c, Z(i), w, seq);
                } else if (prec) {
                    double* z = flexible ? Z(i) : tmp;
                    applyM(V(i), z);
                    applyA(z, w);
                } else applyA(V(i), w);
                n_matvec++;
                double* dh = c->d_scalars;
                double* hs = c->h_scalars + (size_t)slot * HSLOT;
                if (!batched) {
                    // dh[k] = H[k][i], dh[i+1] = ||w||^2, dh[i+2] = ||w||
                    dot_dev(c, n, w, V(0), dh + 0);
                    for (int k = 0; k < i; k++) mgs_step_dev(c, n, dh + k, V(k), V(k + 1), w, dh + k + 1);
                    axpy_negdev(c, n, dh + i, V(i), w);
                    dot_dev(c, n, w, w, dh + i + 1);
                    scale_invsqrt_dev(c, n, dh + i + 1, w, dh + i + 2);  // V[i+1] = w / ||w||
                    THCM_CUDA(cudaMemcpyAsync(c->h_scalars, dh, sizeof(double) * (i + 3), cudaMemcpyDeviceToHost, c->stream));
                } else {
                    const int nv = i + 1;
                    std::vector<double*> vp(nv);
                    for (int k = 0; k < nv; k++) vp[k] = V(k);
                    if (!c->d_flags) { THCM_CUDA(cudaMalloc(&c->d_flags, sizeof(int) * 8)); THCM_CUDA(cudaMemsetAsync(c->d_flags, 0, sizeof(int) * 8, c->stream)); }
                    const bool fused_cgs2 = cgs2_kernel && nv <= 64;   // (one kernel holds up to 64 basis pointers)
                    if (!fused_cgs2) THCM_CUDA(cudaMemsetAsync(dh + 2 * S, 0, sizeof(double) * S, c->stream));   // h2 of a skipped second pass
                    multi_dot_dev(c, n, nv, vp.data(), w, nullptr, dh);
                    if (fused_cgs2) {
                        // CGS2 with the basis read three times instead of four: w' = w - V h1 and h2 = V^T w', w'.w' in ONE sweep
                        // (+ all-reduce + DGKS decision); the second update only when the decision asks for it.  h2 lands in
                        // dh[2S..] and is used by the host only when the flag is set.
                        const bool pyth = pyth_norm && nv <= 64;
                        fused_axpy_dot_dev(c, n, nv, vp.data(), dh, w, dh + 2 * S, dh + nv, c->d_flags, dh + 3 * S, pyth ? c->d_flags + 1 : nullptr);
                        // explicit second update + norm: always without the Pythagorean shortcut, else only when its guard tripped
                        multi_axpy_dot_dev(c, n, nv, vp.data(), dh + 2 * S, pyth ? c->d_flags + 1 : c->d_flags, w, dh + 3 * S, nullptr, nullptr, nullptr,
                                           KID_SECOND_UPDATE, pyth);
                    } else if (c->p2p_on || c->blk.nranks == 1) {
                        // fused update + norm (+ all-reduce + DGKS decision): two reductions per iteration when no second pass
                        multi_axpy_dot_dev(c, n, nv, vp.data(), dh, nullptr, w, dh + S, dh + nv, c->d_flags, dh + 3 * S);
                        multi_dot_dev(c, n, nv, vp.data(), w, c->d_flags, dh + 2 * S);
                        multi_axpy_dot_dev(c, n, nv, vp.data(), dh + 2 * S, c->d_flags, w, dh + 3 * S, nullptr, nullptr, nullptr);
                    } else {   // plain NCCL all-reduces (THCM_P2P=0)
                        multi_axpy_dev(c, n, nv, vp.data(), dh, nullptr, w);
                        dot_dev(c, n, w, w, dh + S);
                        dgks_flag_dev(c, dh + nv, dh + S, c->d_flags);
                        multi_dot_dev(c, n, nv, vp.data(), w, c->d_flags, dh + 2 * S);
                        multi_axpy_dev(c, n, nv, vp.data(), dh + 2 * S, c->d_flags, w);
                        dot_dev(c, n, w, w, dh + 3 * S);
                    }
                    if (!fuse_head) scale_invsqrt_dev(c, n, dh + 3 * S, w, dh + 3 * S + 1);
                    THCM_CUDA(cudaMemcpyAsync(hs, dh, sizeof(double) * (3 * S + 2), cudaMemcpyDeviceToHost, c->stream));
                    THCM_CUDA(cudaMemcpyAsync(hs + 3 * S + 2, c->d_flags, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
                    THCM_CUDA(cudaEventRecord(c->ev_slot[slot], c->stream));
                }
            };
            // ---- host part of step i: column i of H, Givens rotations, residual estimate (GMRESSolver.H:190-199)
            auto process = [&](int i, int slot) -> bool {
                if (!batched) {
                    THCM_CUDA(cudaStreamSynchronize(c->stream));
                    for (int k = 0; k <= i; k++) H[k][i] = c->h_scalars[k];
                    H[i + 1][i] = c->h_scalars[i + 2];
                } else {
                    const double* hs = c->h_scalars + (size_t)slot * HSLOT;
                    THCM_CUDA(cudaEventSynchronize(c->ev_slot[slot]));
                    const bool second = *reinterpret_cast<const int*>(hs + 3 * S + 2) != 0;
                    for (int k = 0; k <= i; k++) H[k][i] = hs[k] + (second ? hs[2 * S + k] : 0.0);
                    H[i + 1][i] = fuse_head ? std::sqrt(hs[3 * S]) : hs[3 * S + 1];   // (IEEE sqrt: the same bits on host and device)
                    if (second) n_reorth++;
                }
                space = i;
                for (int k = 0; k < i; k++) app_rot(H[k][i], H[k + 1][i], cs[k], sn[k]);
                gen_rot(H[i][i], H[i + 1][i], cs[i], sn[i]);
                app_rot(H[i][i], H[i + 1][i], cs[i], sn[i]);
                app_rot(s[i], s[i + 1], cs[i], sn[i]);
                resid = std::abs(s[i + 1]) / normb;
                push(resid);
                return resid < tol;
            };
            if (!batched) {
                for (i = 0; i < m && iter <= maxit; i++, iter++) {
                    enqueue(i, 0);
                    if (process(i, 0)) { converged = true; break; }
                }
            } else {
                // software-pipelined by one step: step i+1 is queued before the host waits for the scalars of step i, so the
                // device never idles on the host's Givens update (a step queued past convergence is simply not used)
                for (i = 0; i < m && iter <= maxit; i++, iter++) {
                    enqueue(i, i & 1);
                    if (i > 0 && process(i - 1, (i - 1) & 1)) { converged = true; break; }
                }
                if (converged) iter--;
                else if (i > 0 && process(i - 1, (i - 1) & 1)) { converged = true; iter--; }
            }
            // Update (GMRESSolver.H:293-313, 402-410): back substitution, x += sum y_j Z_j | V_j | M^-1 V y
            y = s;
            for (int a = space; a >= 0; a--) {
                y[a] /= H[a][a];
                for (int b2 = a - 1; b2 >= 0; b2--) y[b2] -= H[b2][a] * y[a];
            }
            if (!prec || flexible) {
                for (int j = 0; j <= space; j++) axpby(c, n, y[j], (prec && flexible) ? Z(j) : V(j), 1.0, d_x);
            } else {
                fill(c, n, 0.0, tmp);
                for (int j = 0; j <= space; j++) axpby(c, n, y[j], V(j), 1.0, tmp);
                applyM(tmp, r);
                axpby(c, n, 1.0, r, 1.0, d_x);
            }
            applyA(d_x, r); n_matvec++;
            axpby(c, n, 1.0, d_b, -1.0, r);
            if (converged || resid < tol) { status = 0; break; }
        }
    }
    if (compact) scatter_cells(c, d_x, d_x_full);
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    if (res) { res->status = status; res->iters = iter; res->resid = resid; res->nhist = std::min(nh, hist_cap); res->n_matvec = n_matvec; }
    return status;
}

// =============================================================================
// IDR(s) -- src/idrsolver/IDRSolver.H:109-340 (bi-orthogonalisation variant, right preconditioning,
// no smoothing / residual replacement, fresh search space).  Shadow vectors come from the caller.
// =============================================================================
int thcmb_idrs(thcmb_ctx* c, const double* d_b, double* d_x, double tol, int maxit, int s, const double* d_P_raw, double* hist,
               int hist_cap, thcmb_krylov_result* res) {
    const int n = c->blk.ndim();
    long long n_matvec = 0;
    int nh = 0;
    auto push = [&](double r) { if (hist && nh < hist_cap) hist[nh] = r; nh++; };
    auto applyA = [&](const double* v, double* out) { thcmb_spmv_dev(c, v, out); n_matvec++; };
    auto applyM = [&](const double* v, double* out) { thcmb_apply_precon_dev(c, v, out); };
    auto dot = [&](const double* a, const double* b) { return thcmb_dot(c, n, a, b); };
    auto update = [&](double* self, double a, const double* A, double b) { axpby(c, n, a, A, b, self); };  // self = a*A + b*self
    // pool layout: r, v, t, P[s], G[s], U[s]
    double* r = pool_vec(c, 0); double* v = pool_vec(c, 1); double* t = pool_vec(c, 2);
    auto P = [&](int i) { return pool_vec(c, 3 + i); };
    auto G = [&](int i) { return pool_vec(c, 3 + s + i); };
    auto U = [&](int i) { return pool_vec(c, 3 + 2 * s + i); };
    // createP (IDRSolver.H:84-104): Gram-Schmidt on the injected "random" vectors
    for (int j = 0; j < s; j++) {
        copy(c, n, d_P_raw + (size_t)j * n, P(j));
        for (int k = 0; k < j; k++) { double alpha = dot(P(k), P(j)); update(P(j), -alpha, P(k), 1.0); }
        double nr = std::sqrt(dot(P(j), P(j)));
        update(P(j), 0.0, P(j), 1.0 / nr);
    }
    double normb = std::sqrt(dot(d_b, d_b));
    double tolb = tol * normb;
    applyA(d_x, r);
    update(r, 1.0, d_b, -1.0);
    double normr = std::sqrt(dot(r, r));
    push(normr);
    int flag = 0, jj = 0, iter = 0;
    double om = 1.0;
    std::vector<double> f(s, 0.0), gamma(s, 0.0);
    std::vector<std::vector<double>> M(s, std::vector<double>(s, 0.0));
    for (int i = 0; i < s; i++) copy(c, n, d_x, G(i));  // G(dim, Vector(*x_))
    copy(c, n, r, t);
    while (normr > tolb && iter < maxit) {
        for (int i = 0; i < s; i++) f[i] = dot(r, P(i));
        for (int k = 0; k < s; k++) {
            copy(c, n, r, v);
            if (jj > 0) {
                for (int i = k; i < s; i++) {
                    gamma[i] = f[i];
                    for (int j = k; j < i; j++) gamma[i] = gamma[i] - M[i][j] * gamma[j];
                    gamma[i] = gamma[i] / M[i][i];
                    update(v, -gamma[i], G(i), 1.0);
                }
                applyM(v, t);
                update(t, 0.0, t, om);  // t.scale(om)
                for (int i = k; i < s; i++) update(t, gamma[i], U(i), 1.0);
                copy(c, n, t, U(k));
            } else {
                applyM(v, U(k));
            }
            applyA(U(k), G(k));
            for (int i = 0; i < k; i++) {
                double alpha = dot(P(i), G(k)) / M[i][i];
                update(G(k), -alpha, G(i), 1.0);
                update(U(k), -alpha, U(i), 1.0);
            }
            for (int i = k; i < s; i++) M[i][k] = dot(G(k), P(i));
            if (M[k][k] == 0) { flag = 3; goto done; }
            {
                double beta = f[k] / M[k][k];
                update(r, -beta, G(k), 1.0);
                update(d_x, beta, U(k), 1.0);
                if (k < s - 1) for (int i = k + 1; i < s; i++) f[i] = f[i] - beta * M[i][k];
            }
            normr = std::sqrt(dot(r, r));
            push(normr);
            iter++;
            if (normr < tolb || iter == maxit) break;
        }
        if (normr < tolb || iter == maxit) break;
        jj++;
        applyM(r, v);
        applyA(v, t);
        {   // calc_omega (IDRSolver.H:366-380)
            double ns = std::sqrt(dot(r, r)), nt = std::sqrt(dot(t, t)), ts = dot(t, r);
            double rho = std::abs(ts / (nt * ns));
            om = ts / (nt * nt);
            if (rho < 0.7) om = om * 0.7 / rho;
        }
        update(r, -om, t, 1.0);
        update(d_x, om, v, 1.0);
        normr = std::sqrt(dot(r, r));
        push(normr);
        iter++;
    }
done:
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    int status = flag ? -flag : (normr < tolb ? 0 : 1);
    if (res) { res->status = status; res->iters = iter; res->resid = normr; res->nhist = std::min(nh, hist_cap); res->n_matvec = n_matvec; }
    return status;
}

// One Newton step with device-resident state: F(x), J(x), preconditioner, solve J dx = -F (Ocean.C:1052-1055 pattern).
int thcmb_newton_step_dev(thcmb_ctx* c, const double* d_un, double* d_dx, double tol, int maxit, int restart, int precon_kind,
                          double* fnorm, thcmb_krylov_result* res) {
    const int n = c->blk.ndim();
    double* d_F = c->d_tmp;
    thcmb_residual_dev(c, d_un, d_F);             // F(x)
    thcmb_jacobian_dev(c, d_un);                  // J(x)
    thcmb_build_precon(c, precon_kind);
    axpby(c, n, 0.0, d_F, -1.0, d_F);             // b = -F
    double fn = thcmb_nrm2(c, n, d_F);
    if (fnorm) *fnorm = fn;
    fill(c, n, 0.0, d_dx);
    return thcmb_gmres(c, d_F, d_dx, tol, maxit, restart, 1 | 4 | (c->gmres_ortho ? 8 : 0), nullptr, 0, res);   // precon 0 = identity
}

// One Newton step from HOST buffers: the end-to-end call (bench.py "e2e"): H2D state, step, D2H update.
int thcmb_newton_step(thcmb_ctx* c, const double* un_host, double* dx_host, double tol, int maxit, int restart, int precon_kind,
                      double* fnorm, thcmb_krylov_result* res) {
    const int n = c->blk.ndim();
    double* d_dx = work_vec(c, 0);
    THCM_CUDA(cudaMemcpyAsync(c->d_un, un_host, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    int rc = thcmb_newton_step_dev(c, c->d_un, d_dx, tol, maxit, restart, precon_kind, fnorm, res);
    THCM_CUDA(cudaMemcpyAsync(dx_host, d_dx, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    return rc;
}

// ---- per-kernel device timing (event pairs around every launch while profiling is on) ----
void thcmb_profile(thcmb_ctx* c, int on) {
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    c->prof_on = on != 0;
    c->prof_kid.clear();
}
static const char* kKernelNames[KID_COUNT] = {"thcm_assemble<RHS>", "thcm_assemble<JAC_GRAPH>", "thcm_assemble<JAC_COUNT>",
    "thcm_assemble<JAC_CRS>", "scan_counts", "spmv_csr", "dot", "mgs_step", "axpby", "axpy_negdev", "scale_invsqrt", "copy", "fill",
    "blockdiag_build", "blockdiag_apply", "halo_pack", "halo_unpack", "multi_dot", "multi_axpy", "second_update"};
int thcmb_kernel_count(void) { return KID_COUNT; }
const char* thcmb_kernel_name(int kid) { return (kid >= 0 && kid < KID_COUNT) ? kKernelNames[kid] : ""; }
int thcmb_profile_report(thcmb_ctx* c, int kid, int* count, double* total_ms) {
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    int n = 0; double ms = 0.0;
    for (size_t i = 0; i < c->prof_kid.size(); i++) {
        if (c->prof_kid[i] != kid) continue;
        float t = 0;
        if (cudaEventElapsedTime(&t, c->prof_ev[2 * i], c->prof_ev[2 * i + 1]) == cudaSuccess) { ms += t; n++; }
    }
    *count = n; *total_ms = ms;
    return 0;
}

// =============================================================================
// B1: the reference's Fortran symbols (one global instance per process, host pointers)
// =============================================================================
static thcmb_settings g_set;
static bool g_have_global = false;
static std::vector<int> g_landm_global;
static thcmb_ctx* g_ctx = nullptr;
static int *g_dbeg = nullptr, *g_djco = nullptr; static double* g_dco = nullptr;
// The caller's CRS arrays live from set_pointers to finalize_ (THCM.C:619-638 allocates them once, THCM::~THCM calls finalize_ before it
// deletes them, THCM.C:798-812): the prefix that actually travels (begA, and jcoA / coA up to 1.25 x the entry count) is page-locked so
// that matrix_'s device-to-host copies run at PCIe speed instead of through the driver's pageable staging (0.95 GB per Jacobian at 1
// degree).  Released by set_pointers, init_ and finalize_; THCM_PIN_CRS=0 keeps the buffers pageable.
static struct PinnedPrefix { void* p = nullptr; size_t bytes = 0, cap = 0; } g_pin[3];
static void unpin_crs() {
    for (auto& e : g_pin) { if (e.p) { cudaHostUnregister(e.p); cudaGetLastError(); } e = PinnedPrefix{}; }
}
static void pin_prefix(int which, void* p, size_t bytes, size_t cap_bytes) {
    static const bool on = !(getenv("THCM_PIN_CRS") && atoi(getenv("THCM_PIN_CRS")) == 0);
    PinnedPrefix& e = g_pin[which];
    if (!on || !p || (e.p == p && e.bytes >= bytes)) return;
    if (e.p) { cudaHostUnregister(e.p); cudaGetLastError(); e = PinnedPrefix{}; }
    const size_t want = std::min(cap_bytes, bytes + bytes / 4);
    if (cudaHostRegister(p, want, cudaHostRegisterDefault) == cudaSuccess) { e.p = p; e.bytes = want; e.cap = cap_bytes; }
    else cudaGetLastError();   // (locked-memory limit, exotic allocator ...): the copy simply stays pageable
}

static thcmb_ctx* G() { if (!g_ctx) fatal("THCM not initialised: call init_ first"); return g_ctx; }
static int g_dims[3] = {0, 0, 0};   // n, m, l of the instance init_ is creating / has created
void thcm_forcing_integral_default(double* field, double* y, int* landm, double* out) {
    const int n = g_dims[0], m = g_dims[1], l = g_dims[2];
    double lf = 0.0, ls = 0.0;
    for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
        const int land = landm[(size_t)i + (size_t)(n + 2) * (j + (size_t)(m + 2) * l)];
        lf = field[(size_t)(i - 1) + (size_t)n * (j - 1)] * std::cos(y[j - 1]) * (1 - land) + lf;
        ls = std::cos(y[j - 1]) * (1 - land) + ls;
    }
    *out = lf / ls;
}

// reads a land mask in the format of topo.F90:41-64 (per level k=0..l+1: one header line, rows j=m+1..0 of n+2 digits)
static bool read_mask_file(const char* path, int n, int m, int l, std::vector<int>& out) {
    std::ifstream f(path);
    if (!f) return false;
    out.assign((size_t)(n + 2) * (m + 2) * (l + 2), LAND);
    std::string line;
    for (int k = 0; k <= l + 1; k++) {
        if (!std::getline(f, line)) return false;
        for (int j = m + 1; j >= 0; j--) {
            if (!std::getline(f, line)) return false;
            for (int i = 0; i <= n + 1 && i < (int)line.size(); i++)
                out[(size_t)i + (size_t)(n + 2) * (j + (size_t)(m + 2) * k)] = line[i] - '0';
        }
    }
    return true;
}

// m_global keeps what topofit (topo.F90:6-38) needs later: initialize only stores it and clears landm (global.F90:100-160); every
// get_landm runs topofit -- readmask when "Read Land Mask" is set, else depth3land with the "Topography" case
static int g_itopo = 1, g_flat = 0, g_rd_mask = 0, g_rd_spertm = 0;
static std::string g_maskfile, g_spertmaskfile;
static std::string locate_mkmask(const std::string& file) {   // global.F90 locate_file: the name as given, then mkmask/<name> below the data dir
    if (std::ifstream(file)) return file;
    const char* dd = getenv("THCM_DATA_DIR");
    return std::string(dd ? dd : ".") + "/mkmask/" + file;
}
static inline int& LMglob(int i, int j, int k) {
    return g_landm_global[(size_t)i + (size_t)(g_set.N + 2) * (j + (size_t)(g_set.M + 2) * k)];
}
static void fix_land_inversion_and_flatten() {   // topo.F90:94-109 / 287-291
    const int n = g_set.N, m = g_set.M, l = g_set.L;
    for (int i = 1; i <= n; i++) for (int j = 1; j <= m; j++) for (int k = l; k >= 2; k--)
        if (LMglob(i, j, k) == LAND && LMglob(i, j, k - 1) == OCEAN) LMglob(i, j, k - 1) = LAND;
    if (g_flat) for (int k = 1; k <= l - 1; k++) for (int j = 0; j <= m + 1; j++) for (int i = 0; i <= n + 1; i++) LMglob(i, j, k) = LMglob(i, j, l);
}
// readmask (topo.F90:41-127)
static void readmask() {
    const int n = g_set.N, m = g_set.M, l = g_set.L;
    const std::string p = locate_mkmask(g_maskfile);
    if (!read_mask_file(p.c_str(), n, m, l, g_landm_global)) fatal("failed to read land mask " + g_maskfile + " (tried " + p + ")");
    fix_land_inversion_and_flatten();
}
// depth3land (topo.F90:129-330) without bathymetry data (depth = 0: every cell starts as LAND): the idealised continents of
// "Topography" 1..4; case 0 fits ETOPO data that does not ship with the reference and stops there too ("cannot find ocean point")
static void depth3land() {
    const int n = g_set.N, m = g_set.M, l = g_set.L;
    g_landm_global.assign((size_t)(n + 2) * (m + 2) * (l + 2), LAND);
    auto ocean_interior = [&]() { for (int k = 1; k <= l; k++) for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) LMglob(i, j, k) = OCEAN; };
    auto land_box = [&](int i0, int i1, int j0, int j1) {   // landm(i0:i1, j0:j1, 1:l) = LAND with Fortran bounds checking left to the caller
        if (i0 < 1 || i1 > n || j0 < 1 || j1 > m) fatal("depth3land: this Topography case needs a larger grid (topo.F90:252-270)");
        for (int k = 1; k <= l; k++) for (int j = j0; j <= j1; j++) for (int i = i0; i <= i1; i++) LMglob(i, j, k) = LAND;
    };
    switch (g_itopo) {
    case 1: ocean_interior(); break;                                   // no continents
    case 2: {                                                          // Miocene: four rectangular continents in longitude / latitude
        ocean_interior();
        constexpr double PI = 3.14159265358979323846;                  // par.F90:14
        const double dx = (g_set.xmax - g_set.xmin) / n, dy = (g_set.ymax - g_set.ymin) / m;
        const double ph1 = 250 * PI / 180., ph2 = 315 * PI / 180., ph3 = 10 * PI / 180., ph4 = 65. * PI / 180.;
        const double thd = -60 * PI / 180., thsa = -35 * PI / 180., thn = 10. * PI / 180., tha = 30 * PI / 180.;
        for (int i = 1; i <= n; i++) {
            const double x = (i - 0.5) * dx + g_set.xmin;              // grid.F90:28
            const bool am = x < ph2 && x > ph1, af = x < ph4 && x > ph3;
            for (int j = 1; j <= m; j++) {
                const double y = (j - 0.5) * dy + g_set.ymin;          // grid.F90:34
                const bool land = (am && y < 0. && y > thd) || (af && y < thn && y > thsa) || (am && y < g_set.ymax && y > tha) ||
                                  (af && y < g_set.ymax && y > tha);
                if (land) for (int k = 1; k <= l; k++) LMglob(i, j, k) = LAND;
            }
        }
        break;
    }
    case 3: ocean_interior(); land_box(18, 20, 1, 16); break;          // single-hemisphere basin
    case 4: ocean_interior(); land_box(22, 24, 6, m); break;           // double-hemisphere basin
    default:
        fatal("m_global::get_landm: \"Topography\" = 0 fits bathymetry data that does not ship with the reference (its depth3land stops "
              "with 'cannot find ocean point'); use \"Read Land Mask\" or Topography 1..4");
    }
    if (g_flat) for (int k = 1; k <= l - 1; k++) for (int j = 0; j <= m + 1; j++) for (int i = 0; i <= n + 1; i++) LMglob(i, j, k) = LMglob(i, j, l);
    if (g_set.periodic)
        for (int k = 0; k <= l + 1; k++) for (int j = 0; j <= m + 1; j++)
            if (LMglob(1, j, k) == OCEAN && LMglob(n, j, k) == OCEAN) { LMglob(n + 1, j, k) = PERIO; LMglob(0, j, k) = PERIO; }
}

void __m_global_MOD_initialize(int* N, int* M, int* L, double* xmin, double* xmax, double* ymin, double* ymax, double* hdim,
                               double* qz, int* periodic, int* itopo, int* flat, int* rd_mask, int* TRES, int* SRES, int* iza,
                               int* ite, int* its, int* rd_spertm, int* coupled_T, int* coupled_S, int* forcing_type,
                               const char* maskfile, const char* spertmaskfile, const char*, const char*, const char*) {
    thcmb_default_settings(&g_set);
    g_set.N = *N; g_set.M = *M; g_set.L = *L;
    g_set.xmin = *xmin; g_set.xmax = *xmax; g_set.ymin = *ymin; g_set.ymax = *ymax; g_set.hdim = *hdim; g_set.qz = *qz;
    g_set.periodic = *periodic; g_set.TRES = *TRES; g_set.SRES = *SRES; g_set.iza = *iza; g_set.ite = *ite; g_set.its = *its;
    g_set.coupled_T = *coupled_T; g_set.coupled_S = *coupled_S; g_set.forcing_type = *forcing_type;
    g_itopo = *itopo; g_flat = *flat; g_rd_mask = *rd_mask; g_rd_spertm = *rd_spertm;
    g_maskfile = maskfile ? maskfile : ""; g_spertmaskfile = spertmaskfile ? spertmaskfile : "";
    g_landm_global.assign((size_t)(*N + 2) * (*M + 2) * (*L + 2), OCEAN);   // landm = 0 (global.F90:158)
    g_have_global = true;
    // (the mask of "Read Land Mask" is read right away as well, so that get_current_landm serves it before the first get_landm)
    if (g_rd_mask && !g_maskfile.empty()) readmask();
}
void __m_global_MOD_get_landm(int* landm) {   // global.F90:299-318: topofit, then the copy
    if (!g_have_global) fatal("m_global::get_landm before m_global::initialize");
    if (g_rd_mask) { if (!g_maskfile.empty()) readmask(); else fatal("failed to read land mask: \"Read Land Mask\" without a \"Land Mask\" file"); }
    else depth3land();
    memcpy(landm, g_landm_global.data(), sizeof(int) * g_landm_global.size());
}
void __m_global_MOD_finalize(void) { g_landm_global.clear(); g_have_global = false; }
/* m_global (global.F90:215-608): the global-domain arrays THCM.C reads on the root before it scatters them */
void __m_global_MOD_get_current_landm(int* landm) { memcpy(landm, g_landm_global.data(), sizeof(int) * g_landm_global.size()); }
void __m_global_MOD_set_landm(int* landm) {   // global.F90:349-381, incl. the land-inversion fix
    if (!g_have_global) fatal("m_global::set_landm before m_global::initialize");
    const int n = g_set.N, m = g_set.M, l = g_set.L;
    memcpy(g_landm_global.data(), landm, sizeof(int) * g_landm_global.size());
    auto LMg = [&](int i, int j, int k) -> int& { return g_landm_global[(size_t)i + (size_t)(n + 2) * (j + (size_t)(m + 2) * k)]; };
    for (int i = 1; i <= n; i++) for (int j = 1; j <= m; j++) for (int k = l; k >= 2; k--)
        if (LMg(i, j, k) == LAND && LMg(i, j, k - 1) == OCEAN) LMg(i, j, k - 1) = LAND;
}
void __m_global_MOD_set_maskfile(const char* maskfile) {   // global.F90:215-224 + the topofit of the next get_landm
    if (!g_have_global) fatal("m_global::set_maskfile before m_global::initialize");
    const int n = g_set.N, m = g_set.M, l = g_set.L;
    g_maskfile = maskfile ? maskfile : "";
    std::vector<int> lm;
    const std::string p = locate_mkmask(g_maskfile);
    if (!read_mask_file(p.c_str(), n, m, l, lm)) fatal("cannot read land mask " + g_maskfile + " (tried " + p + ")");
    __m_global_MOD_set_landm(lm.data());
}
static void need_no_datafile(bool needs_file, const char* what) {
    if (needs_file) fatal(std::string(what) + ": this option reads a data file (Levitus / Trenberth) that does not ship with the reference; "
                          "provide the field through m_inserts instead");
}
void __m_global_MOD_get_windfield(double* taux, double* tauy) {   // global.F90:425-463
    need_no_datafile(g_set.iza < 2, "m_global::get_windfield (iza < 2)");
    const size_t nm = (size_t)g_set.N * g_set.M;
    for (size_t q = 0; q < nm; q++) { taux[q] = 0.0; tauy[q] = 0.0; }
}
void __m_global_MOD_get_temforcing(double* tatm) {                // global.F90:465-506
    need_no_datafile(g_set.coupled_T == 0 && g_set.ite == 0 && g_set.TRES != 0, "m_global::get_temforcing (ite = 0, TRES = 1)");
    for (size_t q = 0; q < (size_t)g_set.N * g_set.M; q++) tatm[q] = 0.0;
}
void __m_global_MOD_get_salforcing(double* emip) {                // global.F90:534-560
    need_no_datafile(g_set.its == 0 && g_set.coupled_S == 0 && g_set.SRES != 0, "m_global::get_salforcing (its = 0, SRES = 1)");
    for (size_t q = 0; q < (size_t)g_set.N * g_set.M; q++) emip[q] = 0.0;
}
void __m_global_MOD_get_internal_temforcing(double*) { need_no_datafile(true, "m_global::get_internal_temforcing"); }
void __m_global_MOD_get_internal_salforcing(double*) { need_no_datafile(true, "m_global::get_internal_salforcing"); }
/* declared by THCM.C:170 but defined nowhere in the reference's Fortran (and never called): present so that nothing is unresolved */
void __m_global_MOD_get_land_temp(double*) { fatal("m_global::get_land_temp is declared by THCM.C but not implemented by the reference"); }
void __m_global_MOD_get_spert(double* spert) {                    // global.F90:587-608
    const int n = g_set.N, m = g_set.M, l = g_set.L;
    for (size_t q = 0; q < (size_t)n * m; q++) spert[q] = (double)g_set.SRES;
    if (!g_rd_spertm) return;
    // read_spertm (forcing.F90:372-402): rows j = m+1 .. 0 of n+2 digits; spert = (1 - digit) * (1 - landm(i,j,l))
    std::ifstream f(locate_mkmask(g_spertmaskfile));
    std::vector<std::string> rows((size_t)m + 2);
    bool ok = (bool)f;
    for (int j = m + 1; ok && j >= 0; j--) ok = (bool)std::getline(f, rows[(size_t)j]);
    if (!ok) {   // the reference prints this warning and goes on with whatever the array held; here: the value without a mask
        fprintf(stderr, "WARNING: failed to read salinity perturbation mask from file mkmask/%s\n", g_spertmaskfile.c_str());
        return;
    }
    for (int j = 1; j <= m; j++) for (int i = 1; i <= n; i++) {
        const int dum = i < (int)rows[(size_t)j].size() ? rows[(size_t)j][(size_t)i] - '0' : 0;
        spert[(size_t)(i - 1) + (size_t)n * (j - 1)] = (double)(1 - dum) * (1 - LMglob(i, j, l));
    }
}
/* global grid arrays pushed by THCM.C right after m_global::initialize (THCM.C:340-354; global.F90:241-293).  The library builds the
 * same arrays itself (grid.F90 formulas, build_grid); the caller's copies are kept and compared with them when init_ creates a model
 * that spans the whole domain (under MPI the sub-domain grids differ by construction). */
static std::vector<double> g_grid_in[6];   // x, y, z, xu, yv, zw as handed in
static void keep_grid(int which, int n, const double* a) { g_grid_in[which].assign(a, a + (n > 0 ? n : 0)); }
void set_global_x(int* n, double* a) { keep_grid(0, *n, a); }
void set_global_y(int* n, double* a) { keep_grid(1, *n, a); }
void set_global_z(int* n, double* a) { keep_grid(2, *n, a); }
void set_global_xu(int* n, double* a) { keep_grid(3, *n, a); }
void set_global_yv(int* n, double* a) { keep_grid(4, *n, a); }
void set_global_zw(int* n, double* a) { keep_grid(5, *n, a); }
static void check_global_grid(const thcmb_ctx* c) {
    // THCM.C hands x(1..N), y(1..M), z(1..L), xu(0..N), yv(0..M), zw(0..L)
    const std::vector<double>* mine[6] = {&c->x, &c->y, &c->z, &c->xu, &c->yv, &c->zw};
    const int first[6] = {1, 1, 1, 0, 0, 0};
    const int count[6] = {c->s.N, c->s.M, c->s.L, c->s.N + 1, c->s.M + 1, c->s.L + 1};
    const char* name[6] = {"x", "y", "z", "xu", "yv", "zw"};
    for (int w = 0; w < 6; w++) {
        if ((int)g_grid_in[w].size() != count[w]) continue;     // not provided (or another layout): nothing to compare
        for (int i = 0; i < count[w]; i++)
            if (std::fabs((*mine[w])[first[w] + i] - g_grid_in[w][i]) > 1e-12 * (1.0 + std::fabs(g_grid_in[w][i])))
                fatal(std::string("set_global_") + name[w] + ": the caller's grid differs from grid.F90's");
    }
}

void init_(int* n, int* m, int* l, int* nmlglob, double* xmin, double* xmax, double* ymin, double* ymax, double* alphaT, double* alphaS,
           int* ih, int* vmix, int* tap, int* rho_mixing, int* coriolis_on, int* periodic, int* landm, double* taux, double* tauy,
           double* tatm, double* emip, double* spert) {
    (void)nmlglob;
    if (!g_have_global) thcmb_default_settings(&g_set);
    thcmb_settings s = g_set;
    // the library sees the caller's (sub)domain as its whole world, like the Fortran it replaces.  NOTE: temfun/salfun
    // use the GLOBAL ymin/ymax of m_global (forcing.F90:424-449); on a single rank they coincide with the local ones.
    // The caller's sub-domain (under MPI: one block of Decomp2D incl. its ghost layers, THCM.C:566-611) with its own bounds; the
    // global latitude bounds of m_global stay available to temfun / salfun (forcing.F90:418-449)
    s.N = *n; s.M = *m; s.L = *l; s.xmin = *xmin; s.xmax = *xmax;
    if (g_have_global && (g_set.ymin != *ymin || g_set.ymax != *ymax)) { s.ymin_glob = g_set.ymin; s.ymax_glob = g_set.ymax; }
    s.ymin = *ymin; s.ymax = *ymax;
    // one process per GPU under MPI: the local rank picks the device (THCM_DEVICE overrides)
    {
        int dev = 0, ndev = 0;
        for (const char* v : {"THCM_DEVICE", "OMPI_COMM_WORLD_LOCAL_RANK", "MV2_COMM_WORLD_LOCAL_RANK", "MPI_LOCALRANKID", "SLURM_LOCALID"})
            if (const char* e = getenv(v)) { dev = atoi(e); break; }
        if (cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0 && dev >= ndev) {
            fprintf(stderr, "thcm_b200: local rank %d but only %d CUDA device(s) on this node: ranks share devices (device %d); set THCM_DEVICE "
                            "or start one rank per GPU\n", dev, ndev, dev % ndev);
            dev %= ndev;
        }
        s.device = dev;
    }
    g_dims[0] = s.N; g_dims[1] = s.M; g_dims[2] = s.L;
    s.alphaT = *alphaT; s.alphaS = *alphaS; s.ih = *ih; s.vmix = *vmix; s.tap = *tap; s.rho_mixing = *rho_mixing;
    s.coriolis_on = *coriolis_on; s.periodic = *periodic; s.rank = 0; s.nranks = 1;
    if (g_ctx) {   // THCM is a singleton that replaces the previous instance (THCM.H:76-84); the CRS staging buffers were sized for it
        thcmb_destroy(g_ctx); g_ctx = nullptr;
        unpin_crs();
        for (void* p : {(void*)g_dbeg, (void*)g_djco, (void*)g_dco}) if (p) cudaFree(p);
        g_dbeg = g_djco = nullptr; g_dco = nullptr;
    }
    g_ctx = thcmb_create(&s, landm);
    g_ctx->use_integral_callback = true;
    if (g_have_global && s.N == g_set.N && s.M == g_set.M && s.L == g_set.L && s.ymin == g_set.ymin && s.ymax == g_set.ymax && s.xmin == g_set.xmin)
        check_global_grid(g_ctx);   // the model spans the whole domain: the caller's global grid must be grid.F90's
    size_t nm = (size_t)s.N * s.M;
    memcpy(g_ctx->taux.data(), taux, sizeof(double) * nm); memcpy(g_ctx->tauy.data(), tauy, sizeof(double) * nm);
    memcpy(g_ctx->tatm.data(), tatm, sizeof(double) * nm); memcpy(g_ctx->emip.data(), emip, sizeof(double) * nm);
    memcpy(g_ctx->spert.data(), spert, sizeof(double) * nm);
    refresh_params(g_ctx);
}
thcmb_ctx* thcmb_fortran_context(void) { return g_ctx; }   // the instance behind the Fortran symbols (nullptr before init_)
void finalize_(void) {
    unpin_crs();
    if (g_ctx) { thcmb_destroy(g_ctx); g_ctx = nullptr; }
    for (void* p : {(void*)g_dbeg, (void*)g_djco, (void*)g_dco}) if (p) cudaFree(p);
    g_dbeg = g_djco = nullptr; g_dco = nullptr;
}
void __m_mat_MOD_get_array_sizes(int* nrows, int* nnz) {  // mat.F90:56-68
    *nrows = G()->blk.ndim();
    *nnz = G()->blk.ndim() * (NUN * NP + 1);
}
void __m_mat_MOD_set_pointers(int* nrows, int* nnz, int* begA, int* jcoA, double* coA, double* coB, int* begF, int* jcoF, double* coF) {
    (void)nrows;
    thcmb_ctx* c = G();
    unpin_crs();
    c->crs_cap = nnz ? (long long)*nnz : 0;
    c->begA = begA; c->jcoA = jcoA; c->coA = coA; c->coB = coB;
    c->begF = begF; c->jcoF = jcoF; c->coF = coF;
}
void rhs_(double* un, double* B) {
    thcmb_ctx* c = G();
    const int n = c->blk.ndim();
    timer_start_("nlin_rhs+boundaries+matAvec");
    THCM_CUDA(cudaMemcpyAsync(c->d_un, un, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    stage_begin(c);
    thcmb_rhs_dev(c, c->d_un, c->d_tmp);
    stage_end(c, "nlin_rhs+boundaries+matAvec");
    THCM_CUDA(cudaMemcpyAsync(B, c->d_tmp, sizeof(double) * n, cudaMemcpyDeviceToHost, c->stream));
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    timer_stop_("nlin_rhs+boundaries+matAvec");
}
void fillcolb_(void) {
    thcmb_ctx* c = G();
    if (!c->coB) fatal("fillcolb_: set_pointers was not called");
    memcpy(c->coB, c->cob_local.data(), sizeof(double) * c->cob_local.size());
}
void matrix_(double* un) {
    thcmb_ctx* c = G();
    if (!c->begA) fatal("matrix_: set_pointers was not called");
    const int n = c->blk.ndim();
    timer_start_("nlin_jac+boundaries+fillcolA");
    if (!g_dbeg) {
        THCM_CUDA(cudaMalloc(&g_dbeg, sizeof(int) * (size_t)(n + 1)));
        THCM_CUDA(cudaMalloc(&g_djco, sizeof(int) * (size_t)c->gnnz));
        THCM_CUDA(cudaMalloc(&g_dco, sizeof(double) * (size_t)c->gnnz));
    }
    fillcolb_();
    THCM_CUDA(cudaMemcpyAsync(c->d_un, un, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    stage_begin(c);
    long long nnz = thcmb_jacobian_crs_dev(c, c->d_un, g_dbeg, g_djco, g_dco);
    stage_end(c, "nlin_jac+boundaries+fillcolA");
    if (c->crs_cap > 0 && nnz > c->crs_cap) fatal("matrix_: the Jacobian holds more entries than the arrays handed to set_pointers");
    const size_t cap = c->crs_cap > 0 ? (size_t)c->crs_cap : (size_t)nnz;
    pin_prefix(0, c->begA, sizeof(int) * (size_t)(n + 1), sizeof(int) * (size_t)(n + 1));
    pin_prefix(1, c->jcoA, sizeof(int) * (size_t)nnz, sizeof(int) * cap);
    pin_prefix(2, c->coA, sizeof(double) * (size_t)nnz, sizeof(double) * cap);
    THCM_CUDA(cudaMemcpyAsync(c->begA, g_dbeg, sizeof(int) * (size_t)(n + 1), cudaMemcpyDeviceToHost, c->stream));
    THCM_CUDA(cudaMemcpyAsync(c->jcoA, g_djco, sizeof(int) * (size_t)nnz, cudaMemcpyDeviceToHost, c->stream));
    THCM_CUDA(cudaMemcpyAsync(c->coA, g_dco, sizeof(double) * (size_t)nnz, cudaMemcpyDeviceToHost, c->stream));
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    timer_stop_("nlin_jac+boundaries+fillcolA");
}
void setparcs_(int* idx, double* val) { thcmb_set_par(G(), *idx, *val); }
void getparcs_(int* idx, double* val) { if (*idx >= 1 && *idx <= NPAR) *val = G()->par[*idx]; }
void setsres_(int* sres) { thcmb_ctx* c = G(); c->s.SRES = *sres; refresh_params(c); }
void set_landmask_(int* landm, int* periodic, int* reinit) {
    thcmb_ctx* c = G();
    c->s.periodic = *periodic; c->blk.periodic = *periodic; c->blk.wrap_x = (*periodic && c->blk.npN == 1) ? 1 : 0;
    set_landmask_ctx(c, landm, *reinit == 1);
    if (g_dbeg) { cudaFree(g_dbeg); cudaFree(g_djco); cudaFree(g_dco); g_dbeg = g_djco = nullptr; g_dco = nullptr; }
}
void get_forcing_(double* frc) { thcmb_get_forcing(G(), frc); }
/* m_inserts (inserts.F90:11-281; THCM.C:85-98): n*m surface fields, i fastest; no recompute until the next setparcs */
void __m_inserts_MOD_insert_taux(double* f) { insert_surface_field(G(), SF_TAUX, f); }
void __m_inserts_MOD_insert_tauy(double* f) { insert_surface_field(G(), SF_TAUY, f); }
void __m_inserts_MOD_insert_atmosphere_t(double* f) { insert_surface_field(G(), SF_TATM, f); }
void __m_inserts_MOD_insert_atmosphere_q(double* f) { insert_surface_field(G(), SF_QATM, f); }
void __m_inserts_MOD_insert_atmosphere_a(double* f) { insert_surface_field(G(), SF_ALBE, f); }
void __m_inserts_MOD_insert_atmosphere_p(double* f) { insert_surface_field(G(), SF_PATM, f); }
void __m_inserts_MOD_insert_seaice_q(double* f) { insert_surface_field(G(), SF_QSA, f); }
void __m_inserts_MOD_insert_seaice_m(double* f) { insert_surface_field(G(), SF_MSI, f); }
void __m_inserts_MOD_insert_seaice_g(double* f) { insert_surface_field(G(), SF_GSI, f); }
void __m_inserts_MOD_insert_emip(double* f) { insert_surface_field(G(), SF_EMIP, f); }
void __m_inserts_MOD_insert_adapted_emip(double* f) { insert_surface_field(G(), SF_ADAPTED_EMIP, f); }
void __m_inserts_MOD_insert_emip_pert(double* f) { insert_surface_field(G(), SF_SPERT, f); }
/* usrc.F90:254-350 (Ocean.C:48-49, 1486, 1507): the structs are Atmosphere::CommPars (18 doubles) / SeaIce::CommPars (7) */
void set_atmos_parameters_(void* pars) { thcmb_set_atmos_parameters(G(), (const double*)pars); }
void set_seaice_parameters_(void* pars) { thcmb_set_seaice_parameters(G(), (const double*)pars); }
void __m_mix_MOD_set_vmix_fix(int* fix) { G()->vmix_fix = *fix; }
/* m_usr::set_internal_forcing (usr.F90:267-300; THCM.C:594), m_thcm_utils::get_landm / loadbal_weights (thcm_utils.F90:259, 325) */
void __m_usr_MOD_set_internal_forcing(double* temp, double* salt) { set_internal_forcing(G(), temp, salt); }
void __m_thcm_utils_MOD_get_landm(int* landm) { thcmb_ctx* c = G(); memcpy(landm, c->landm.data(), sizeof(int) * c->landm.size()); }
void __m_thcm_utils_MOD_loadbal_weights(double* weights, double*, double*, double*) { loadbal_weights(G(), weights); }
/* m_probe (probe.F90; THCM.C:1568-1763): surface diagnostics of the coupled model, n*m fields, host state vector */
void __m_probe_MOD_get_atmosphere_t(double* f) { probe_get_field(G(), SF_TATM, f); }
void __m_probe_MOD_get_atmosphere_q(double* f) { probe_get_field(G(), SF_QATM, f); }
void __m_probe_MOD_get_atmosphere_p(double* f) { probe_get_field(G(), SF_PATM, f); }
void __m_probe_MOD_get_emip(double* f) { probe_get_field(G(), SF_EMIP, f); }
void __m_probe_MOD_get_adapted_emip(double* f) { probe_get_field(G(), SF_ADAPTED_EMIP, f); }
void __m_probe_MOD_get_emip_pert(double* f) { probe_get_field(G(), SF_SPERT, f); }
void __m_probe_MOD_get_taux(double* f) { probe_get_field(G(), SF_TAUX, f); }
void __m_probe_MOD_get_tauy(double* f) { probe_get_field(G(), SF_TAUY, f); }
void __m_probe_MOD_get_suno(double* f) { probe_get_suno(G(), f); }
void __m_probe_MOD_compute_evap(double* evap, double* un) { probe_compute_evap(G(), un, evap); }
void __m_probe_MOD_get_salflux(double* un, double* salflux, double* scorr, double* qsoaflux, double* qsosflux) {
    probe_get_salflux(G(), un, salflux, scorr, qsoaflux, qsosflux);
}
void __m_probe_MOD_get_temflux(double* un, double* totflux, double* swflux, double* shflux, double* lhflux, double* siflux, double* simask) {
    probe_get_temflux(G(), un, totflux, swflux, shflux, lhflux, siflux, simask);
}
void __m_probe_MOD_get_derivatives(double* un, double* dftdm, double* dfsdq, double* dfsdm, double* dfsdg) {
    probe_get_derivatives(G(), un, dftdm, dfsdq, dfsdm, dfsdg);
}
/* m_integrals (integrals.F90:17-88; THCM.C:2133, 2155) */
void __m_integrals_MOD_salt_advection(double* un, double* check) { integrals_salt_advection(G(), un, check); }
void __m_integrals_MOD_salt_diffusion(double* un, double* check) { integrals_salt_diffusion(G(), un, check); }
/* forcing.F90:235-280 (THCM.C:848); usrc.F90:201-251, 421-431 (Ocean.C:889, 1610; THCM.C:1974); inout.F90:20 (Ocean.C:1877) */
void get_stochastic_forcing_(void) { thcmb_ctx* c = G(); stochastic_forcing(c, c->begF, c->jcoF, c->coF); refresh_params(c); }
void getdeps_(double* Ooa, double* Os, double* nus, double* eta, double* lvsc, double* qdim, double* pqsnd) {
    double o[7]; get_deps(G(), o);
    *Ooa = o[0]; *Os = o[1]; *nus = o[2]; *eta = o[3]; *lvsc = o[4]; *qdim = o[5]; *pqsnd = o[6];
}
void get_parameters_(double* r0dim, double* udim, double* hdim) { get_dim_parameters(G(), r0dim, udim, hdim); }
void get_nondimensionalization_parameters_(double* out) { for (int i = 0; i < NP; i++) out[i] = 0.0; }
void writeparams_(void) { write_params(G()); }
void write_data_(double* u, int* ofile, int* lab) { write_data(G(), u, *ofile, lab); }
void write_levitus_(const char*) { fprintf(stderr, "thcm_b200: write_levitus is a debugging dump of Levitus fields the library never reads; nothing written\n"); }
/* m_scaling (scaling.F90:29-105) on the Jacobian of the last matrix_ call, m_thcm_utils::intcond_scaling (thcm_utils.F90:285-309) */
void __m_scaling_MOD_average_block(double* db) { average_block(G(), db); }
void __m_scaling_MOD_compute(double* db, double* rowscales, double* colscales) { scaling_compute(G(), db, rowscales, colscales); }
void __m_thcm_utils_MOD_intcond_scaling(double* values, int* indices, int* len) { *len = intcond_scaling(G(), values, indices); }

}  // extern "C"
