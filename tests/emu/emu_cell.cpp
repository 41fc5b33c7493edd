// =============================================================================
// emu_cell.cpp -- CPU UNIT TEST HARNESS for the device code (test infrastructure only).
//
// Compiles the library's own per-cell device functions (i-emic_b200/csrc/thcm_cell.cuh: eval_row,
// boundaries, the usol ghost rules) and host setup (thcm_host.cpp) with g++ and runs them cell by
// cell in the order the CUDA kernels use, so that `pytest -m "not gpu"` can compare the kernel
// arithmetic, the class/slot tables, the static graph and the halo plan with the oracle on a box
// without a GPU.  It is NOT part of libthcm_b200.so and is never a fallback for it.
// =============================================================================
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../i-emic_b200/csrc/thcm_cell.cuh"

using namespace thcm;

extern "C" void thcm_throw_error_(char* msg) { fprintf(stderr, "emu: %s\n", msg); }
extern "C" void thcm_forcing_integral_(double*, double*, int*, double* out) { *out = 0.0; }   // only reached by contexts init_ creates

namespace {
struct Emu {
    thcmb_ctx c;
    std::vector<uint32_t> nbmask; std::vector<uint8_t> surf, uvlive; std::vector<int> send_idx, recv_slot;
};

AsmArgs make_args(Emu* e, const double* un, const double* halo) {
    thcmb_ctx* c = &e->c;
    const Block& b = c->blk;
    AsmArgs a;
    a.b = DevBlock{b.N, b.M, b.L, b.i0, b.j0, b.n0, b.m0, b.periodic, b.wrap_x, b.halo_w, b.halo_e, b.halo_s, b.halo_n, b.hk, b.ncell()};
    a.t = c->tab; a.t.jt = c->jt_host.data(); a.t.kt = c->kt_host.data(); a.t.msi = c->msi_local.data();
    a.un = un; a.halo = halo; a.nbmask = e->nbmask.data(); a.surf = e->surf.data(); a.uvlive = e->uvlive.data();
    a.frc = c->frc_local.data(); a.rowptr = c->rowptr_host.data();
    a.val = nullptr; a.blockcnt = nullptr; a.begA = nullptr; a.jcoA = nullptr; a.coA = nullptr; a.out = nullptr; a.sign = 1.0;
    return a;
}

// one (cell,row): the sequence of do_row() in thcm_assembly.cu
template <int R, bool JAC>
void cell_row(const AsmArgs& a, int cell, double* E, Cell& c, int& cls, uint32_t& nb) {
    const DevBlock& b = a.b;
    int li = cell % b.n0, r = cell / b.n0, lj = r % b.m0, k0 = r / b.m0;
    c.li = li; c.lj = lj; c.gi = b.i0 + li + 1; c.gj = b.j0 + lj + 1; c.k = k0 + 1;
    nb = a.nbmask[cell];
    double sm = (double)(int)(int8_t)a.surf[lj * b.n0 + li];
    cls = (c.gi == 1 ? 1 : 0) | (c.gi == b.N ? 2 : 0) | (c.gj == 1 ? 4 : 0) | (c.gj == b.M ? 8 : 0) | (c.k == 1 ? 16 : 0) | (c.k == b.L ? 32 : 0);
    if (!((nb >> 4) & 1u)) {
        eval_row<R, JAC, true>(E, a.t, a.b, c, sm, DirectTile{a, c.gi, c.gj, c.k}, DirectTabs{a.t, c.gj, c.k});
        if constexpr (JAC && (R == TT || R == SS)) vmix_jac<R>(E, a.t, c, nb, DirectTile{a, c.gi, c.gj, c.k}, DirectTabs{a.t, c.gj, c.k});
    }
    boundaries<R>(E, nb, c.gi < b.N, c.gj < b.M);
    for (int q = 0; q < RowSlots<R>::N; q++) E[q] = std::fabs(E[q]) > DROP_TOL ? E[q] : 0.0;
}

template <int R>
void jac_row(Emu* e, const AsmArgs& a, int cell, double* val, int* beg_cnt, std::vector<int>* jco, std::vector<double>* co) {
    double E[RowSlots<R>::N]; Cell c{0, 0, 0, 0, 0}; int cls; uint32_t nb;
    cell_row<R, true>(a, cell, E, c, cls, nb);
    const ClassTables& ct = class_tables(a.b.periodic);
    int base = a.rowptr[NUN * cell + R - 1];
    int cnt = 0;
    for (int q = 0; q < RowSlots<R>::N; q++) {
        int p = ct.pos[cls][ROW_OFF[R - 1] + q];
        if (val && p >= 0) val[base + p] = E[q];
        if (E[q] != 0.0) {
            cnt++;
            if (jco) {
                int loc = row_slots(R)[q].loc, col = row_slots(R)[q].col;
                int gi2 = c.gi + loc_di(loc), gj2 = c.gj + loc_dj(loc), k2 = c.k + loc_dk(loc);
                if (a.b.periodic) { if (gi2 == 0) gi2 = a.b.N; else if (gi2 == a.b.N + 1) gi2 = 1; }
                jco->push_back(NUN * ((k2 - 1) * a.b.N * a.b.M + a.b.N * (gj2 - 1) + gi2 - 1) + col);
                co->push_back(E[q]);
            }
        }
    }
    if (beg_cnt) beg_cnt[NUN * cell + R - 1] = cnt;
}

template <int R>
void rhs_row(const AsmArgs& a, int cell, double* out) {
    const DevBlock& b = a.b;
    double E[RowSlots<R>::N]; Cell c{0, 0, 0, 0, 0}; int cls; uint32_t nb;
    cell_row<R, false>(a, cell, E, c, cls, nb);
    double s = 0.0;
    for (int q = 0; q < RowSlots<R>::N; q++) {
        int loc = row_slots(R)[q].loc, col = row_slots(R)[q].col;
        int gi2 = c.gi + loc_di(loc), gj2 = c.gj + loc_dj(loc), k2 = c.k + loc_dk(loc);
        bool inside = gj2 >= 1 && gj2 <= b.M && k2 >= 1 && k2 <= b.L && (b.periodic || (gi2 >= 1 && gi2 <= b.N));
        if (E[q] != 0.0 && inside) s = E[q] * stage_value(a, SV_RAW + col - 1, gi2, gj2, k2) + s;
    }
    int row = NUN * cell + R - 1;
    double mixv = 0.0;
    if constexpr (R == TT || R == SS) mixv = vmix_rhs<R>(a.t, c, nb, DirectTile{a, c.gi, c.gj, c.k}, DirectTabs{a.t, c.gj, c.k});
    double B = -s - mixv + a.frc[row] - 0.0;
    B = B * (((nb >> 4) & 1u) ? 0.0 : 1.0);
    out[row] = a.sign * B;
}
}  // namespace

extern "C" {

double emu_tanh(double x) { return thcm::fd_tanh(x); }   // the device path's tanh (thcm_tanh.h), compiled for the host

void* emu_create(const thcmb_settings* s, const int* landm) {
    Emu* e = new Emu();
    thcmb_ctx* c = &e->c;
    c->s = *s;
    if (!setup_block(c, landm)) { delete e; return nullptr; }
    build_grid(c); stpnt(c); apply_landmask_rules(c, landm, false); vmix_init(c); init_surface_fields(c);
    build_static_host(c, e->nbmask, e->surf, e->uvlive, e->send_idx, e->recv_slot);
    compute_forcing(c); compute_tables(c); compute_cob(c);
    return e;
}
void emu_destroy(void* h) { delete (Emu*)h; }
// vmix_control (mix_imp.f:139-169) on the GLOBAL field norms the caller passes in (0/1 flags), and the fix flag
void emu_vmix_control(void* h, int temp, int salt) { thcmb_ctx* c = &((Emu*)h)->c; if (c->vmix_flag >= 2 && c->vmix_fix == 0) { vmix_set_flags(c, temp, salt); compute_tables(c); } }
void emu_set_vmix_fix(void* h, int fix) { ((Emu*)h)->c.vmix_fix = fix; }
void emu_set_par(void* h, int idx, double v) { thcmb_ctx* c = &((Emu*)h)->c; c->par[idx] = v; compute_forcing(c); compute_tables(c); compute_cob(c); }
// m_inserts / set_atmos_parameters / set_seaice_parameters (coupled mode)
void emu_set_field(void* h, int which, const double* f) { insert_surface_field(&((Emu*)h)->c, which, f); }
void emu_set_atmos(void* h, const double* p) { thcmb_ctx* c = &((Emu*)h)->c; set_atmos_parameters(c, p); compute_forcing(c); compute_tables(c); compute_cob(c); }
void emu_set_seaice(void* h, const double* p) { thcmb_ctx* c = &((Emu*)h)->c; set_seaice_parameters(c, p); compute_forcing(c); compute_tables(c); compute_cob(c); }
// host diagnostics of the B1 boundary (thcm_probe.cpp)
void emu_set_internal_forcing(void* h, const double* t, const double* s_) { thcmb_ctx* c = &((Emu*)h)->c; set_internal_forcing(c, t, s_); }
int emu_probe_field(void* h, int which, double* out) { return probe_get_field(&((Emu*)h)->c, which, out) ? 1 : 0; }
void emu_probe_suno(void* h, double* out) { probe_get_suno(&((Emu*)h)->c, out); }
void emu_compute_evap(void* h, const double* un, double* out) { probe_compute_evap(&((Emu*)h)->c, un, out); }
void emu_salflux(void* h, const double* un, double* sf, double* corr, double* qa, double* qs) { probe_get_salflux(&((Emu*)h)->c, un, sf, corr, qa, qs); }
void emu_temflux(void* h, const double* un, double* six) {
    thcmb_ctx* c = &((Emu*)h)->c; size_t nm = (size_t)c->s.N * c->s.M;
    probe_get_temflux(c, un, six, six + nm, six + 2 * nm, six + 3 * nm, six + 4 * nm, six + 5 * nm);
}
void emu_derivatives(void* h, const double* un, double* four) {
    thcmb_ctx* c = &((Emu*)h)->c; size_t nm = (size_t)c->s.N * c->s.M;
    probe_get_derivatives(c, un, four, four + nm, four + 2 * nm, four + 3 * nm);
}
void emu_salt_advection(void* h, const double* un, double* out) { integrals_salt_advection(&((Emu*)h)->c, un, out); }
void emu_salt_diffusion(void* h, const double* un, double* out) { integrals_salt_diffusion(&((Emu*)h)->c, un, out); }
void emu_stochastic_forcing(void* h, int* beg, int* jco, double* co) { thcmb_ctx* c = &((Emu*)h)->c; stochastic_forcing(c, beg, jco, co); }
int emu_ocean_block_atmosphere(void* h, double albed, const double* pdist, const int* colT, const int* colQ, const int* colA, const int* colP,
                               int* beg, int* jco, double* co) {
    return ocean_block_atmosphere(&((Emu*)h)->c, albed, pdist, colT, colQ, colA, colP, beg, jco, co);
}
int emu_ocean_block_seaice(void* h, const double* un, const int* colQ, const int* colM, const int* colG, int* beg, int* jco, double* co) {
    return ocean_block_seaice(&((Emu*)h)->c, un, colQ, colM, colG, beg, jco, co);
}
void emu_getdeps(void* h, double* out7) { get_deps(&((Emu*)h)->c, out7); }
void emu_loadbal(void* h, double* w) { loadbal_weights(&((Emu*)h)->c, w); }
// the library's grid arrays (build_grid = grid.F90): x(1..N), xu(0..N), y(1..M), yv(0..M), z(1..L), zw(0..L), dfzT(1..L), dfzW(0..L)
void emu_grid(void* h, double* x, double* xu, double* y, double* yv, double* z, double* zw, double* dfzT, double* dfzW) {
    thcmb_ctx* c = &((Emu*)h)->c; const int N = c->s.N, M = c->s.M, L = c->s.L;
    for (int i = 1; i <= N; i++) x[i - 1] = c->x[i];
    for (int i = 0; i <= N; i++) xu[i] = c->xu[i];
    for (int j = 1; j <= M; j++) y[j - 1] = c->y[j];
    for (int j = 0; j <= M; j++) yv[j] = c->yv[j];
    for (int k = 1; k <= L; k++) { z[k - 1] = c->z[k]; dfzT[k - 1] = c->dfzT[k]; }
    for (int k = 0; k <= L; k++) { zw[k] = c->zw[k]; dfzW[k] = c->dfzW[k]; }
}
// set_landmask_ (usrc.F90:353-418) exactly as thcm_api.cu runs it on the host side: mask rules with the inversion fix, static data, then
// either the re-initialisation (vmix_init + forcing + lin) or just the mass diagonal
void emu_set_landmask(void* h, const int* landm, int periodic, int reinit) {
    Emu* e = (Emu*)h; thcmb_ctx* c = &e->c;
    c->s.periodic = periodic; c->blk.periodic = periodic; c->blk.wrap_x = (periodic && c->blk.npN == 1) ? 1 : 0;
    apply_landmask_rules(c, landm, true);
    c->peers.clear(); c->send_dst_host.clear(); c->send_peer_host.clear();
    build_static_host(c, e->nbmask, e->surf, e->uvlive, e->send_idx, e->recv_slot);
    if (reinit == 1) { vmix_init(c); compute_forcing(c); compute_tables(c); compute_cob(c); }
    else compute_cob(c);
}
void emu_setsres(void* h, int sres) { thcmb_ctx* c = &((Emu*)h)->c; c->s.SRES = sres; compute_forcing(c); compute_tables(c); compute_cob(c); }   // setsres_
// cell compaction maps: number of ocean cells, then ocell[n_ocean] and ccell[ncell]
int emu_ocean_cells(void* h) { return (int)((Emu*)h)->c.ocell_host.size(); }
void emu_cell_maps(void* h, int* ocell, int* ccell) {
    thcmb_ctx* c = &((Emu*)h)->c;
    memcpy(ocell, c->ocell_host.data(), sizeof(int) * c->ocell_host.size());
    memcpy(ccell, c->ccell_host.data(), sizeof(int) * c->ccell_host.size());
}
double emu_get_par(void* h, int idx) { return ((Emu*)h)->c.par[idx]; }
int emu_ndim(void* h) { return ((Emu*)h)->c.blk.ndim(); }
long long emu_gnnz(void* h) { return ((Emu*)h)->c.gnnz; }
int emu_halo_size(void* h) { return NUN * ((Emu*)h)->c.blk.nhalo_cells(); }
// the block of settings->rank alone (decomposition incl. ocean-weighted cut lines), without building a model
int emu_block_only(const thcmb_settings* s, const int* landm, int* out) {
    thcmb_ctx c; c.s = *s;
    if (!setup_block(&c, landm)) return -1;
    const Block& b = c.blk; int v[] = {b.i0, b.j0, b.n0, b.m0, b.npN, b.npM}; memcpy(out, v, sizeof(v));
    return 0;
}
void emu_block(void* h, int* out) { const Block& b = ((Emu*)h)->c.blk; int v[] = {b.i0, b.j0, b.n0, b.m0, b.npN, b.npM, b.pidN, b.pidM, b.hk}; memcpy(out, v, sizeof(v)); }
void emu_graph(void* h, int* rowptr, int* col) {
    thcmb_ctx* c = &((Emu*)h)->c;
    memcpy(rowptr, c->rowptr_host.data(), sizeof(int) * c->rowptr_host.size());
    memcpy(col, c->col_host.data(), sizeof(int) * c->col_host.size());
}
void emu_halo_gids(void* h, int* g) { thcmb_ctx* c = &((Emu*)h)->c; memcpy(g, c->halo_gid.data(), sizeof(int) * c->halo_gid.size()); }
void emu_local_gids(void* h, int* g) { thcmb_ctx* c = &((Emu*)h)->c; memcpy(g, c->local_gid.data(), sizeof(int) * c->local_gid.size()); }
void emu_get_forcing(void* h, double* f, int masked) { thcmb_ctx* c = &((Emu*)h)->c; memcpy(f, (masked ? c->frc_local : c->frc_raw).data(), sizeof(double) * c->frc_local.size()); }
void emu_get_cob(void* h, double* f) { thcmb_ctx* c = &((Emu*)h)->c; memcpy(f, c->cob_local.data(), sizeof(double) * c->cob_local.size()); }
// halo plan: returns number of peers; arrays sized by emu_plan_sizes
void emu_plan_sizes(void* h, int* npeers, int* nsend, int* nrecv) { Emu* e = (Emu*)h; *npeers = (int)e->c.peers.size(); *nsend = e->c.nsend_cells; *nrecv = e->c.nrecv_cells; }
void emu_plan(void* h, int* peers5, int* send_idx, int* recv_slot) {
    Emu* e = (Emu*)h;
    for (size_t p = 0; p < e->c.peers.size(); p++) {
        auto& q = e->c.peers[p];
        int v[5] = {q.rank, q.send_off, q.send_cnt, q.recv_off, q.recv_cnt};
        memcpy(peers5 + 5 * p, v, sizeof(v));
    }
    memcpy(send_idx, e->send_idx.data(), sizeof(int) * e->send_idx.size());
    memcpy(recv_slot, e->recv_slot.data(), sizeof(int) * e->recv_slot.size());
}

void emu_jacobian(void* h, const double* un, const double* halo, double* val) {
    Emu* e = (Emu*)h; AsmArgs a = make_args(e, un, halo);
    for (long long q = 0; q < e->c.gnnz; q++) val[q] = 0.0;
    for (int cell = 0; cell < a.b.ncell; cell++) {
        jac_row<1>(e, a, cell, val, nullptr, nullptr, nullptr); jac_row<2>(e, a, cell, val, nullptr, nullptr, nullptr);
        jac_row<3>(e, a, cell, val, nullptr, nullptr, nullptr); jac_row<4>(e, a, cell, val, nullptr, nullptr, nullptr);
        jac_row<5>(e, a, cell, val, nullptr, nullptr, nullptr); jac_row<6>(e, a, cell, val, nullptr, nullptr, nullptr);
    }
}
// Fortran-order thresholded CRS (1-based); returns nnz
long long emu_crs(void* h, const double* un, const double* halo, int* beg, int* jco, double* co) {
    Emu* e = (Emu*)h; AsmArgs a = make_args(e, un, halo);
    std::vector<int> j; std::vector<double> v; std::vector<int> cnt(NUN * a.b.ncell);
    for (int cell = 0; cell < a.b.ncell; cell++) {
        jac_row<1>(e, a, cell, nullptr, cnt.data(), &j, &v); jac_row<2>(e, a, cell, nullptr, cnt.data(), &j, &v);
        jac_row<3>(e, a, cell, nullptr, cnt.data(), &j, &v); jac_row<4>(e, a, cell, nullptr, cnt.data(), &j, &v);
        jac_row<5>(e, a, cell, nullptr, cnt.data(), &j, &v); jac_row<6>(e, a, cell, nullptr, cnt.data(), &j, &v);
    }
    int acc = 1;
    for (int r = 0; r < NUN * a.b.ncell; r++) { beg[r] = acc; acc += cnt[r]; }
    beg[NUN * a.b.ncell] = acc;
    memcpy(jco, j.data(), sizeof(int) * j.size()); memcpy(co, v.data(), sizeof(double) * v.size());
    return (long long)j.size();
}
// stage_position (the unconditional-load form the kernels run) against stage_value (the reference semantics) for every
// owned cell, all 27 neighbour positions and all 11 staged fields; returns the number of mismatches
long long emu_check_staging(void* h, const double* un, const double* halo) {
    Emu* e = (Emu*)h; AsmArgs a = make_args(e, un, halo);
    const DevBlock& b = a.b;
    long long bad = 0;
    for (int cell = 0; cell < b.ncell; cell++) {
        int li = cell % b.n0, r = cell / b.n0, lj = r % b.m0, k0 = r / b.m0;
        int gi = b.i0 + li + 1, gj = b.j0 + lj + 1, k = k0 + 1;
        for (int dk = -1; dk <= 1; dk++) for (int dj = -1; dj <= 1; dj++) for (int di = -1; di <= 1; di++) {
            double out[SV_NRHS];
            stage_position<SV_NRHS>(a, gi + di, gj + dj, k + dk, out);
            for (int sv = 0; sv < SV_NRHS; sv++) {
                double ref = stage_value(a, sv, gi + di, gj + dj, k + dk);
                if (!(out[sv] == ref)) bad++;
            }
            // the fast path the kernels take for positions inside the owned block (thcm_assembly.cu stage_inputs)
            int ie = li + di, je = lj + dj, k2 = k + dk;
            if (ie >= 0 && ie < b.n0 && je >= 0 && je < b.m0 && k2 >= 1 && k2 <= b.L) {
                size_t cc = ((size_t)(k2 - 1) * b.m0 + je) * b.n0 + ie;
                bool live = a.uvlive[((size_t)(k2 - 1) * (b.m0 + 2) + (gj + dj - b.j0)) * (b.n0 + 2) + (gi + di - b.i0)] != 0;
                double fast[SV_NRHS];
                stage_regular<SV_NRHS>(a.un + (size_t)NUN * cc, live, k2 != b.L, fast);
                for (int sv = 0; sv < SV_NRHS; sv++) if (!(fast[sv] == out[sv])) bad++;
            }
        }
    }
    return bad;
}
// The pipelined kernel's staging, replayed on the host: the loader's bulk copies (line_plan), the in-place usol fix-up
// from the tile descriptor bits and the consumers' record accessor, against stage_value(); plus the descriptor's masks
// and graph offsets and the per-j / per-k table records.  Returns the number of mismatches.
long long emu_check_tiles(void* h, const double* un, const double* halo) {
    Emu* e = (Emu*)h; AsmArgs a = make_args(e, un, halo);
    const DevBlock& b = a.b;
    std::vector<TileDesc> td;
    build_tile_descs(&e->c, e->nbmask, e->surf, e->uvlive, td);
    const int ntile = tiles_per_row(b) * b.m0 * b.L;
    long long bad = 0;
    if ((int)td.size() != ntile) return -1;
    std::vector<double> rec(9 * TILE_W * NUN);
    for (int t = 0; t < ntile; t++) {
        const TileGeom g = tile_geom_of(b, t);
        const TileDesc& d = td[t];
        std::fill(rec.begin(), rec.end(), -7.77e77);
        long long copied = 0;
        for (int r = 0; r < 9; r++) {
            LineSeg seg[3];
            int ns = line_plan(b, g, r, seg);
            for (int q = 0; q < ns; q++) {
                const double* src = (seg[q].halo ? halo : un) + (size_t)NUN * seg[q].idx;
                if (seg[q].halo && !halo) { bad++; continue; }
                for (int i = 0; i < seg[q].n * NUN; i++) rec[((size_t)r * TILE_W + seg[q].x0) * NUN + i] = src[i];
                copied += seg[q].n;
            }
        }
        if (copied != 9ll * (g.ncell + 2)) bad++;
        for (int r = 0; r < 9; r++) for (int x = 0; x < g.ncell + 2; x++) {
            double* p = &rec[((size_t)r * TILE_W + x) * NUN];
            if (!((d.uvbits[r] >> x) & 1ull)) { p[0] = 0.0; p[1] = 0.0; }
            if (!((d.wbits[r] >> x) & 1ull)) p[2] = 0.0;
            for (int sv = 0; sv < SV_NJAC; sv++) {
                double got = p[sv <= SV_W ? sv : sv + 1];
                double ref = stage_value(a, sv, g.gi0 - 1 + x, g.gj + r % 3 - 1, g.k + r / 3 - 1);
                if (!(got == ref)) bad++;
            }
        }
        for (int x = 0; x < TILE_CELLS; x++) {
            uint32_t nb = x < g.ncell ? a.nbmask[g.cell0 + x] : 0u;
            if (d.nbmask[x] != nb) bad++;
            unsigned sf = x < g.ncell ? a.surf[(size_t)g.lj * b.n0 + (g.cell0 + x) % b.n0] : 0;
            if (((d.surfbits >> x) & 1u) != sf) bad++;
        }
        if (d.g0 != a.rowptr[NUN * g.cell0] || d.tot != a.rowptr[NUN * (g.cell0 + g.ncell)] - d.g0) bad++;
        for (int tb = 0; tb < J_COUNT; tb++) for (int dd = 0; dd < JREC; dd++)
            if (!(e->c.jrec_host[((size_t)g.gj * J_COUNT + tb) * JREC + dd] == a.t.jt[(size_t)tb * a.t.jstride + g.gj + dd - 1])) bad++;
        for (int tb = 0; tb < K_COUNT; tb++)
            if (!(e->c.krec_host[(size_t)g.k * K_COUNT + tb] == a.t.kt[(size_t)tb * a.t.kstride + g.k])) bad++;
    }
    return bad;
}
// fraction of tiles that take the TMA bulk-store path (nothing clipped, 16-byte aligned): returns count, writes total
long long emu_fast_tiles(void* h, long long* total) {
    Emu* e = (Emu*)h;
    std::vector<TileDesc> td;
    build_tile_descs(&e->c, e->nbmask, e->surf, e->uvlive, td);
    long long f = 0;
    for (auto& d : td) f += d.flags & 1u;
    *total = (long long)td.size();
    return f;
}
void emu_rhs(void* h, const double* un, const double* halo, double* B) {
    Emu* e = (Emu*)h; AsmArgs a = make_args(e, un, halo);
    for (int cell = 0; cell < a.b.ncell; cell++) {
        rhs_row<1>(a, cell, B); rhs_row<2>(a, cell, B); rhs_row<3>(a, cell, B);
        rhs_row<4>(a, cell, B); rhs_row<5>(a, cell, B); rhs_row<6>(a, cell, B);
    }
}
}  // extern "C"
