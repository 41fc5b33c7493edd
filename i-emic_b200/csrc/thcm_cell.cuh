// =============================================================================
// thcm_cell.cuh -- per-(cell,row) evaluation of the THCM dependency block: state access with the
// ghost rules of usol, row evaluation (lin + nonlinear atoms) and `boundaries`.  Pure functions of
// the kernel arguments, shared by all assembly kernels (thcm_assembly.cu).  THCM_HD lets the unit
// test tests/emu/emu_cell.cpp compile the very same functions for the host to check them against
// the oracle without a GPU (a test of the device code, not a CPU fallback: the library itself only
// ever instantiates them in __global__ kernels).
// =============================================================================
#pragma once
#include "thcm_internal.h"
#include "thcm_slots.h"
#include "thcm_tanh.h"
#ifdef __CUDACC__
#define THCM_HD __host__ __device__ __forceinline__
#else
#define THCM_HD inline
#endif
#ifdef __CUDA_ARCH__
#define THCM_LDG(p) __ldg(p)
#else
#define THCM_LDG(p) (*(p))
#endif

namespace thcm {

constexpr int CELLS_PER_BLOCK = 32;
constexpr int ASM_THREADS = 32 * NUN;
constexpr double DROP_TOL = 1.0e-10;  // assemble.F90:115

struct AsmArgs {
    DevBlock b;
    DevTables t;
    const double* un;        // owned state, 6 interleaved unknowns per cell
    const double* halo;      // halo state (nranks > 1)
    const uint32_t* nbmask;
    const uint8_t* surf;
    const uint8_t* uvlive;
    const double* frc;
    const int* rowptr;       // static graph
    double* val;
    int* blockcnt;           // JAC_COUNT out / JAC_CRS in (exclusive scan)
    int* begA; int* jcoA; double* coA;
    double* out;             // RHS
    double sign;             // RHS: +1 -> B (rhs_), -1 -> F = -B (THCM.C:1011)
    const TileDesc* tdesc;   // pipelined kernel: per-tile descriptors, per-j / per-k table records
    const double* jrec; const double* krec;
    int ntile;
    const int* tile_list;    // Jacobian kernels: tiles to work on (nullptr = all tiles in order)
};

// ---------------------------------------------------------------------------
// state access with the ghost rules of usol (usrc.F90:1014-1121), GLOBAL 1-based indices
// ---------------------------------------------------------------------------
struct Cell {
    int gi, gj, k;   // global Fortran indices of this cell (1-based)
    int li, lj;      // local 0-based
};

THCM_HD double raw(const AsmArgs& a, int gi, int gj, int k, int var /*0..5*/) {
    const DevBlock& b = a.b;
    int ie = gi - 1 - b.i0, je = gj - 1 - b.j0, kk = k - 1;
    if (b.wrap_x) { if (ie < 0) ie += b.n0; else if (ie >= b.n0) ie -= b.n0; }
    if (ie >= 0 && ie < b.n0 && je >= 0 && je < b.m0)
        return a.un[(size_t)NUN * (((size_t)kk * b.m0 + je) * b.n0 + ie) + var];
    // halo slot (mirror of thcm::halo_slot)
    int wrow = b.n0 + b.halo_w + b.halo_e, hs;
    if (je == -1) hs = kk * b.hk + (ie + b.halo_w);
    else if (je == b.m0) hs = kk * b.hk + b.halo_s * wrow + (ie + b.halo_w);
    else if (ie == -1) hs = kk * b.hk + (b.halo_s + b.halo_n) * wrow + je;
    else hs = kk * b.hk + (b.halo_s + b.halo_n) * wrow + b.halo_w * b.m0 + je;
    return a.halo[(size_t)NUN * hs + var];
}

// t, s (var 4, 5): no-flux mirror in y and (non-periodic) x, periodic copy otherwise (usrc.F90:1046-1100)
THCM_HD double ts_value(const AsmArgs& a, int gi, int gj, int k, int var) {
    const DevBlock& b = a.b;
    if (gj < 1) gj = 1; else if (gj > b.M) gj = b.M;
    if (!b.periodic) { if (gi < 1) gi = 1; else if (gi > b.N) gi = b.N; }
    if (k < 1) k = 1; else if (k > b.L) k = b.L;
    return raw(a, gi, gj, k, var);
}
// w (var 2): w(i,j,0) = w(i,j,l) = 0 for i in 1..n; the periodic ghost columns 0 and n+1 are copied BEFORE
// w(:,:,l) is zeroed, so they keep the raw top-level value (usrc.F90:1051-1052 vs :1093)
THCM_HD double w_value(const AsmArgs& a, int gi, int gj, int k) {
    const DevBlock& b = a.b;
    if (k < 1 || k > b.L || gj < 1 || gj > b.M) return 0.0;
    if (gi >= 1 && gi <= b.N) return k == b.L ? 0.0 : raw(a, gi, gj, k, 2);
    return b.periodic ? raw(a, gi, gj, k, 2) : 0.0;
}
// u, v (var 0, 1) on cell corners: usol's no-slip zeroing is precomputed as the uvlive box; corner 0 in a
// periodic domain is the raw copy of corner N (usrc.F90:1049-1050) with its OWN zeroing rule
THCM_HD double uv_value(const AsmArgs& a, int ic, int jc, int k, int var) {
    const DevBlock& b = a.b;
    if (k < 1 || k > b.L) return 0.0;   // ghost levels only feed entries that `boundaries` removes
    int bn = b.n0 + 2, bm = b.m0 + 2;
    int bi = ic - b.i0, bj = jc - b.j0;
    if (!a.uvlive[((size_t)(k - 1) * bm + bj) * bn + bi]) return 0.0;
    return raw(a, ic, jc, k, var);
}

// ---------------------------------------------------------------------------
// Staged fields.  The kernels first stage, for every cell of a tile and its 26 neighbours, the fields exactly as
// usol leaves them (u, v with the no-slip zeroing, w with its lid/bottom rule, t, s with the no-flux mirrors) --
// and, for the residual, the raw unknowns the matrix-vector product multiplies -- into shared memory with
// independent, coalesced loads; the row evaluation then reads shared memory only.
// ---------------------------------------------------------------------------
enum { SV_U = 0, SV_V, SV_W, SV_T, SV_S, SV_NJAC = 5, SV_RAW = 5, SV_NRHS = 11 };

// value of staged field sv at the global position (gi, gj, k), which may lie one cell outside the domain
THCM_HD double stage_value(const AsmArgs& a, int sv, int gi, int gj, int k) {
    const DevBlock& b = a.b;
    if (sv <= SV_V) {
        if (gi > b.N || gj > b.M) return 0.0;     // corners N+1 / M+1 do not exist (never read by a kept entry)
        return uv_value(a, gi, gj, k, sv);
    }
    if (sv == SV_W) return w_value(a, gi, gj, k);
    if (sv <= SV_S) return ts_value(a, gi, gj, k, sv - SV_T + 4);
    // raw unknown sv - SV_RAW: only positions inside the domain (with periodic wrap) are ever multiplied
    bool inside = gj >= 1 && gj <= b.M && k >= 1 && k <= b.L && (b.periodic || (gi >= 1 && gi <= b.N));
    return inside ? raw(a, gi, gj, k, sv - SV_RAW) : 0.0;
}

// Source record of ONE staged position: the clamped (no-flux mirror) / wrapped cell whose six raw unknowns feed the
// staged fields; always a valid record of the owned state (`halo` = 0, idx = cell) or of the halo buffer (`halo` = 1,
// idx = slot, mirror of thcm::halo_slot).
struct PosSrc { int halo; long long idx; };
THCM_HD PosSrc position_source(const DevBlock& b, int gi, int gj, int k) {
    int cj = gj < 1 ? 1 : (gj > b.M ? b.M : gj);
    int ck = k < 1 ? 1 : (k > b.L ? b.L : k);
    int ci = gi;
    if (!b.periodic) ci = gi < 1 ? 1 : (gi > b.N ? b.N : gi);
    int ie = ci - 1 - b.i0, je = cj - 1 - b.j0, kk = ck - 1;
    if (b.wrap_x) { if (ie < 0) ie += b.n0; else if (ie >= b.n0) ie -= b.n0; }
    if (ie >= 0 && ie < b.n0 && je >= 0 && je < b.m0) return PosSrc{0, ((long long)kk * b.m0 + je) * b.n0 + ie};
    int wrow = b.n0 + b.halo_w + b.halo_e, hs;
    if (je == -1) hs = kk * b.hk + (ie + b.halo_w);
    else if (je == b.m0) hs = kk * b.hk + b.halo_s * wrow + (ie + b.halo_w);
    else if (ie == -1) hs = kk * b.hk + (b.halo_s + b.halo_n) * wrow + je;
    else hs = kk * b.hk + (b.halo_s + b.halo_n) * wrow + b.halo_w * b.m0 + je;
    return PosSrc{1, (long long)hs};
}
// what usol keeps of the record at this position: bit 0 u,v survive (corner exists and no-slip rule, usrc.F90:1104-1119),
// bit 1 w survives (lid / bottom / ghost-column rule, usrc.F90:1051-1052, 1093-1094), bit 2 the position lies inside the
// domain (incl. the periodic wrap): only then does matAvec multiply its raw unknowns
enum { POS_UV = 1, POS_W = 2, POS_INSIDE = 4 };
THCM_HD unsigned position_flags(const DevBlock& b, const uint8_t* uvlive, int gi, int gj, int k) {
    int ck = k < 1 ? 1 : (k > b.L ? b.L : k);
    int bi = gi - b.i0, bj = gj - b.j0;
    bool corner = gi <= b.N && gj <= b.M && k >= 1 && k <= b.L;
    int bic = bi < 0 ? 0 : (bi > b.n0 + 1 ? b.n0 + 1 : bi), bjc = bj < 0 ? 0 : (bj > b.m0 + 1 ? b.m0 + 1 : bj);
    bool live = uvlive[((size_t)(ck - 1) * (b.m0 + 2) + bjc) * (b.n0 + 2) + bic] != 0;
    bool win = k >= 1 && k <= b.L && gj >= 1 && gj <= b.M;
    bool wi = (gi >= 1 && gi <= b.N) ? (k != b.L) : (b.periodic != 0);
    bool inside = gj >= 1 && gj <= b.M && k >= 1 && k <= b.L && (b.periodic || (gi >= 1 && gi <= b.N));
    return ((live && corner) ? POS_UV : 0u) | ((win && wi) ? POS_W : 0u) | (inside ? POS_INSIDE : 0u);
}

// All staged fields of ONE position with unconditional, independent loads (what the kernels execute): the six raw
// unknowns of the clamped / wrapped source cell (always a valid address) + one uvlive byte, then selects.  Must agree
// with stage_value() for every sv (checked on the host by tests/emu: emu_check_staging).
template <int NSV>
THCM_HD void stage_position(const AsmArgs& a, int gi, int gj, int k, double* out) {
    const DevBlock& b = a.b;
    const PosSrc ps = position_source(b, gi, gj, k);
    const double* src = (ps.halo ? a.halo : a.un) + (size_t)NUN * ps.idx;
    double r[NUN];
#ifdef __CUDA_ARCH__
    const double2* s2 = reinterpret_cast<const double2*>(src);   // cells are 48-byte records, 16-byte aligned
    double2 q0 = __ldg(s2), q1 = __ldg(s2 + 1), q2 = __ldg(s2 + 2);
    r[0] = q0.x; r[1] = q0.y; r[2] = q1.x; r[3] = q1.y; r[4] = q2.x; r[5] = q2.y;
#else
    for (int v = 0; v < NUN; v++) r[v] = src[v];
#endif
    const unsigned fl = position_flags(b, a.uvlive, gi, gj, k);
    out[SV_U] = (fl & POS_UV) ? r[0] : 0.0;
    out[SV_V] = (fl & POS_UV) ? r[1] : 0.0;
    out[SV_W] = (fl & POS_W) ? r[2] : 0.0;
    out[SV_T] = r[4];
    out[SV_S] = r[5];
    if constexpr (NSV > SV_NJAC) {
#pragma unroll
        for (int v = 0; v < NUN; v++) out[SV_RAW + v] = (fl & POS_INSIDE) ? r[v] : 0.0;
    }
}

// Fast path of the staging for a position that lies INSIDE the owned block (the bulk of a tile): `rec` is the cell's
// 48-byte record, `live` its uvlive byte, wlive = (k != L).  Equals stage_position()/stage_value() there (tests/emu).
template <int NSV>
THCM_HD void stage_regular(const double* rec, bool live, bool wlive, double* out) {
    double r[NUN];
#ifdef __CUDA_ARCH__
    const double2* s2 = reinterpret_cast<const double2*>(rec);
    double2 q0 = __ldg(s2), q1 = __ldg(s2 + 1), q2 = __ldg(s2 + 2);
    r[0] = q0.x; r[1] = q0.y; r[2] = q1.x; r[3] = q1.y; r[4] = q2.x; r[5] = q2.y;
#else
    for (int v = 0; v < NUN; v++) r[v] = rec[v];
#endif
    out[SV_U] = live ? r[0] : 0.0;
    out[SV_V] = live ? r[1] : 0.0;
    out[SV_W] = wlive ? r[2] : 0.0;
    out[SV_T] = r[4];
    out[SV_S] = r[5];
    if constexpr (NSV > SV_NJAC) {
#pragma unroll
        for (int v = 0; v < NUN; v++) out[SV_RAW + v] = r[v];
    }
}

// ---------------------------------------------------------------------------
// Tiles.  A tile is TI = 32 consecutive cells of one (j,k) grid line of the owned block; the kernels stage its
// 3 x 3 x (TI+2) neighbourhood.  Line r = (dk+1)*3 + (dj+1) of that neighbourhood is, in global memory, ONE contiguous
// run of 48-byte records (two or three runs at the periodic seam / block edges): what the TMA loader copies.
// ---------------------------------------------------------------------------
constexpr int TILE_CELLS = CELLS_PER_BLOCK;
constexpr int TILE_W = TILE_CELLS + 2;
struct TileGeom { int cell0, ncell, gi0, gj, k, lj; };   // first owned cell, count, global (1-based) origin
THCM_HD int tiles_per_row(const DevBlock& b) { return (b.n0 + TILE_CELLS - 1) / TILE_CELLS; }
THCM_HD TileGeom tile_geom_of(const DevBlock& b, int tile) {
    const int nbx = tiles_per_row(b);
    int ib = tile % nbx, rest = tile / nbx;
    int lj = rest % b.m0, k0 = rest / b.m0;
    TileGeom g;
    g.cell0 = (k0 * b.m0 + lj) * b.n0 + ib * TILE_CELLS;
    g.ncell = b.n0 - ib * TILE_CELLS < TILE_CELLS ? b.n0 - ib * TILE_CELLS : TILE_CELLS;
    g.gi0 = b.i0 + ib * TILE_CELLS + 1; g.gj = b.j0 + lj + 1; g.k = k0 + 1; g.lj = lj;
    return g;
}
struct LineSeg { int halo; long long idx; int x0, n; };   // n records from un / halo record idx to staged columns x0..
THCM_HD int line_plan(const DevBlock& b, const TileGeom& g, int r, LineSeg* seg) {
    const int dj = r % 3 - 1, dk = r / 3 - 1;
    const PosSrc m = position_source(b, g.gi0, g.gj + dj, g.k + dk);             // owned columns: always one run
    const PosSrc w = position_source(b, g.gi0 - 1, g.gj + dj, g.k + dk);
    const PosSrc e = position_source(b, g.gi0 + g.ncell, g.gj + dj, g.k + dk);
    int n = 0;
    LineSeg mid{m.halo, m.idx, 1, g.ncell};
    if (w.halo == m.halo && w.idx == m.idx - 1) { mid.idx = m.idx - 1; mid.x0 = 0; mid.n++; }
    else { seg[n].halo = w.halo; seg[n].idx = w.idx; seg[n].x0 = 0; seg[n].n = 1; n++; }
    if (e.halo == m.halo && e.idx == m.idx + g.ncell) mid.n++;
    else { seg[n].halo = e.halo; seg[n].idx = e.idx; seg[n].x0 = g.ncell + 1; seg[n].n = 1; n++; }
    seg[n++] = mid;
    return n;
}

// tile / table accessors used by the host emulation and as the reference semantics of the shared-memory ones:
//   tile(sv, di, dj, dk)  staged field at the neighbour offset;  tabs.jt(table, dj), tabs.kt(table)
struct DirectTile {
    const AsmArgs& a; int gi, gj, k;
    THCM_HD double operator()(int sv, int di, int dj, int dk) const { return stage_value(a, sv, gi + di, gj + dj, k + dk); }
};
struct DirectTabs {
    const DevTables& t; int gj, k;
    THCM_HD double jt(int tb, int dj) const { return THCM_LDG(t.jt + (size_t)tb * t.jstride + gj + dj); }
    THCM_HD double kt(int tb) const { return THCM_LDG(t.kt + (size_t)tb * t.kstride + k); }
};

template <int N> struct IntC { static constexpr int value = N; };
template <int I, int N, class F> THCM_HD void static_for(F&& f) {
    if constexpr (I < N) { f(IntC<I>{}); static_for<I + 1, N>(f); }
}

// E[slot_of(R,loc,col)] with a compile-time check that the entry is structural
template <int R, int LOC, int COL> THCM_HD double& entref(double* E) {
    static_assert(slot_of(R, LOC, COL) >= 0, "not a structural entry of this row");
    return E[slot_of(R, LOC, COL)];
}
#define ENT(loc, col) entref<R, loc, col>(E)

// ---------------------------------------------------------------------------
// row evaluation: An = Al (lin, usrc.F90:690-785) + nonlinear atoms (usrc.F90:842-882 | 950-1007)
// ---------------------------------------------------------------------------
// CPL: compile the coupled-mode terms in (the uncoupled kernels carry none of that code: their instruction streams are
// instruction-cache bound)
template <int R, bool JAC, bool CPL, class Tile, class Tabs>
THCM_HD void eval_row(double* E, const DevTables& t, const DevBlock& blk, const Cell& c, double sm, const Tile& tile, const Tabs& tabs) {
    const int gi = c.gi, gj = c.gj, k = c.k, N = blk.N, M = blk.M, L = blk.L;
    const int a = 0;  // (the accessors below keep the call shape of the global-memory versions)
    auto JT = [&](int tb, int j) { return tabs.jt(tb, j - gj); };
    auto KT = [&](int tb, int kk) { (void)kk; return tabs.kt(tb); };
    auto UV = [&](int, int ic, int jc, int kk, int var) { return tile(var, ic - gi, jc - gj, kk - k); };
    auto WW_ = [&](int, int ic, int jc, int kk) { return tile(SV_W, ic - gi, jc - gj, kk - k); };
    auto TS = [&](int, int ic, int jc, int kk, int var) { return tile(SV_T + var - 4, ic - gi, jc - gj, kk - k); };
    const double epsr = t.epsr;
#pragma unroll
    for (int q = 0; q < RowSlots<R>::N; q++) E[q] = 0.0;

    if constexpr (R == UU || R == VV) {
        // lin atoms exist for j = 1..m-1 only (spf.F90:34,45,62,68; 98,108,124,130), uzz/vzz for all j
        const bool jin = gj <= M - 1;
        const double tdz8 = KT(K_TDZI8, k);
        // w averages of unlin(5)/vnlin(5) (spf.F90:618-628, 741-751)
        double w23 = (WW_(a, gi, gj, k) + WW_(a, gi, gj + 1, k) + WW_(a, gi + 1, gj, k) + WW_(a, gi + 1, gj + 1, k)) * tdz8;
        double w14 = -(WW_(a, gi, gj, k - 1) + WW_(a, gi, gj + 1, k - 1) + WW_(a, gi + 1, gj, k - 1) + WW_(a, gi + 1, gj + 1, k - 1)) * tdz8;
        double w5 = w14 + w23;
        const double c2x = JT(J_C2X, gj), c2y = JT(J_C2Y, gj), tanv = JT(J_TANYV, gj);
        if constexpr (R == UU) {
            // ---- Al(UU,UU) = -EH*(uxx+uyy+ucsi) - EV*uzz ----
            double l2 = jin ? JT(J_LU2, gj) : 0.0, l4 = jin ? JT(J_LU4, gj) : 0.0, l6 = jin ? JT(J_LU6, gj) : 0.0;
            double l5 = (jin ? JT(J_LU5, gj) : 0.0) - KT(K_ZU5, k);
            double l14 = -0.0 - KT(K_ZU14, k), l23 = -0.0 - KT(K_ZU23, k);
            // ---- nonlinear: epsr*(uux|Urux + uvy1 + uwz + uvy2) ----
            double a8 = gi <= N - 1 ? (JAC ? 2 * UV(a, gi + 1, gj, k, 0) * c2x : UV(a, gi + 1, gj, k, 0) * c2x) : 0.0;
            double a2 = gi >= 2 ? (JAC ? -2 * UV(a, gi - 1, gj, k, 0) * c2x : -UV(a, gi - 1, gj, k, 0) * c2x) : 0.0;
            double a4 = gj >= 2 ? -UV(a, gi, gj - 1, k, 1) * JT(J_COSYV, gj - 1) * c2y : 0.0;
            double a6 = gj <= M - 1 ? UV(a, gi, gj + 1, k, 1) * JT(J_COSYV, gj + 1) * c2y : 0.0;
            double uvy2 = UV(a, gi, gj, k, 1) * tanv;
            ENT(2, UU) = l2 + epsr * a2;
            ENT(8, UU) = l2 + epsr * a8;
            ENT(4, UU) = l4 + epsr * a4;
            ENT(6, UU) = l6 + epsr * a6;
            ENT(5, UU) = l5 + epsr * (w5 + uvy2);
            ENT(14, UU) = l14 + epsr * w14;
            ENT(23, UU) = l23 + epsr * w23;
            // ---- Al(UU,VV) = -fv - EH*vxs ; Al(UU,PP) = px ----
            double luv2 = jin ? JT(J_LUV2, gj) : 0.0, luv8 = jin ? JT(J_LUV8, gj) : 0.0, luv5 = jin ? -JT(J_CORV, gj) : 0.0;
            double px = jin ? c2x : 0.0;
            ENT(5, PP) = -px; ENT(6, PP) = -px; ENT(8, PP) = px; ENT(9, PP) = px;
            if constexpr (JAC) {
                // An(UU,VV) += epsr*(Urvy1 + Urvy2), An(UU,WW) += epsr*Urwz (usrc.F90:951-952)
                double b4 = gj >= 2 ? -UV(a, gi, gj - 1, k, 0) * JT(J_COSYV, gj - 1) * c2y : 0.0;
                double b6 = gj <= M - 1 ? UV(a, gi, gj + 1, k, 0) * JT(J_COSYV, gj + 1) * c2y : 0.0;
                double b5 = UV(a, gi, gj, k, 0) * tanv;
                ENT(2, VV) = luv2; ENT(8, VV) = luv8;
                ENT(4, VV) = 0.0 + epsr * b4; ENT(6, VV) = 0.0 + epsr * b6;
                ENT(5, VV) = luv5 + epsr * b5;
                double u0 = UV(a, gi, gj, k, 0);
                double up = (u0 + UV(a, gi, gj, k + 1, 0)) * tdz8, dn = -(u0 + UV(a, gi, gj, k - 1, 0)) * tdz8;
                double eu = epsr * up, ed = epsr * dn;
                ENT(5, WW) = eu; ENT(6, WW) = eu; ENT(8, WW) = eu; ENT(9, WW) = eu;
                ENT(14, WW) = ed; ENT(15, WW) = ed; ENT(17, WW) = ed; ENT(18, WW) = ed;
            } else {
                ENT(2, VV) = luv2; ENT(8, VV) = luv8; ENT(5, VV) = luv5;
            }
        } else {
            // ---- Al(VV,VV) = -EH*(vxx+vyy+vcsi) - EV*vzz ----
            double l2 = jin ? JT(J_LV2, gj) : 0.0, l4 = jin ? JT(J_LV4, gj) : 0.0, l6 = jin ? JT(J_LV6, gj) : 0.0;
            double l5 = (jin ? JT(J_LV5, gj) : 0.0) - KT(K_ZU5, k);
            double l14 = -0.0 - KT(K_ZU14, k), l23 = -0.0 - KT(K_ZU23, k);
            // nonlinear: epsr*(uvx + vvy|Vrvy + vwz)
            double a8 = gi <= N - 1 ? UV(a, gi + 1, gj, k, 0) * c2x : 0.0;
            double a2 = gi >= 2 ? -UV(a, gi - 1, gj, k, 0) * c2x : 0.0;
            double a6 = gj <= M - 1 ? (JAC ? 2 * UV(a, gi, gj + 1, k, 1) * JT(J_COSYV, gj + 1) * c2y : UV(a, gi, gj + 1, k, 1) * JT(J_COSYV, gj + 1) * c2y) : 0.0;
            double a4 = gj >= 2 ? (JAC ? -2 * UV(a, gi, gj - 1, k, 1) * JT(J_COSYV, gj - 1) * c2y : -UV(a, gi, gj - 1, k, 1) * JT(J_COSYV, gj - 1) * c2y) : 0.0;
            ENT(2, VV) = l2 + epsr * a2;
            ENT(8, VV) = l2 + epsr * a8;
            ENT(4, VV) = l4 + epsr * a4;
            ENT(6, VV) = l6 + epsr * a6;
            ENT(5, VV) = l5 + epsr * w5;
            ENT(14, VV) = l14 + epsr * w14;
            ENT(23, VV) = l23 + epsr * w23;
            // ---- Al(VV,UU) = fu - EH*uxs ; Al(VV,PP) = py ----
            double lvu2 = jin ? JT(J_LVU2, gj) : 0.0, lvu8 = jin ? JT(J_LVU8, gj) : 0.0, lvu5 = jin ? JT(J_CORV, gj) : 0.0;
            double py = jin ? t.dyi : 0.0;
            ENT(5, PP) = -py; ENT(8, PP) = -py; ENT(6, PP) = py; ENT(9, PP) = py;
            double u0 = UV(a, gi, gj, k, 0);
            if constexpr (JAC) {
                // An(VV,UU) += epsr*(Urt2 + uVrx), An(VV,WW) += epsr*Vrwz (usrc.F90:965-967)
                double b8 = gi <= N - 1 ? UV(a, gi + 1, gj, k, 1) * c2x : 0.0;
                double b2 = gi >= 2 ? -UV(a, gi - 1, gj, k, 1) * c2x : 0.0;
                ENT(2, UU) = lvu2 + epsr * b2; ENT(8, UU) = lvu8 + epsr * b8;
                ENT(5, UU) = lvu5 + epsr * (2 * u0 * tanv);
                double v0 = UV(a, gi, gj, k, 1);
                double up = (v0 + UV(a, gi, gj, k + 1, 1)) * tdz8, dn = -(v0 + UV(a, gi, gj, k - 1, 1)) * tdz8;
                double eu = epsr * up, ed = epsr * dn;
                ENT(5, WW) = eu; ENT(6, WW) = eu; ENT(8, WW) = eu; ENT(9, WW) = eu;
                ENT(14, WW) = ed; ENT(15, WW) = ed; ENT(17, WW) = ed; ENT(18, WW) = ed;
            } else {
                ENT(2, UU) = lvu2; ENT(8, UU) = lvu8;
                ENT(5, UU) = lvu5 + epsr * (u0 * tanv);   // ut2 (usrc.F90:853)
            }
        }
    } else if constexpr (R == WW) {
        // Al(WW,PP) = pz ; Al(WW,TT) = -Ra(1+xes*alpt1) tbc/2 ; Al(WW,SS) = lambda Ra tbc/2 (usrc.F90:716-718)
        ENT(5, PP) = KT(K_WP5, k); ENT(23, PP) = KT(K_WP23, k);
        double lt = t.cWT * sm / 2., ls = t.cWS * sm / 2.;
        double q5 = 0.0, q23 = 0.0, r5 = 0.0, r23 = 0.0;
        if (k <= L - 1) {  // wnlin fills k = 1..l-1 only (spf.F90:503-539)
            double t0 = TS(a, gi, gj, k, 4), t1 = TS(a, gi, gj, k + 1, 4);
            if constexpr (JAC) {
                q5 = (t0 + t1) / 2.; q23 = q5;
                double s2 = t0 + t1; r5 = 0.375 * (s2 * s2); r23 = r5;
            } else {
                q23 = t1 / 4.; q5 = (t0 + 2 * t1) / 4.;
                r5 = 0.125 * (t0 * t0 + 3 * t1 * t0 + 3 * t1 * t1); r23 = 0.125 * t1 * t1;
            }
        }
        ENT(5, TT) = lt - t.c2 * q5 + t.c3 * r5;
        ENT(23, TT) = lt - t.c2 * q23 + t.c3 * r23;
        ENT(5, SS) = ls; ENT(23, SS) = ls;
    } else if constexpr (R == PP) {
        // Al(PP,UU) = uxc, Al(PP,VV) = vyc, Al(PP,WW) = wzc (usrc.F90:726-728, spf.F90:152-178)
        double cp = JT(J_CP, gj), pa = JT(J_PVA, gj), pb = JT(J_PVB, gj);
        ENT(2, UU) = -cp; ENT(4, UU) = cp; ENT(1, UU) = -cp; ENT(5, UU) = cp;
        ENT(4, VV) = -pa; ENT(2, VV) = pb; ENT(1, VV) = -pa; ENT(5, VV) = pb;
        ENT(5, WW) = KT(K_PW5, k); ENT(14, WW) = KT(K_PW14, k);
    } else {
        // ---- T / S rows: Al = -ph*(txx+tyy) - pv*tzz + RES*bi*tc (usrc.F90:758,785) ----
        constexpr int var = (R == TT) ? 4 : 5;
        const double c4x = JT(J_C4X, gj), c4y = JT(J_C4Y, gj), cv0 = JT(J_COSYV, gj - 1), cv1 = JT(J_COSYV, gj);
        const double dfz = KT(K_DFZT, k), tdzi = t.tdzi2;
        double l2 = JT(J_TT2, gj) * sm, l4 = JT(J_TT4, gj) * sm, l6 = JT(J_TT6, gj) * sm;
        double l5 = JT(J_TT5, gj) * sm - KT(K_ZT5, k) * sm;
        // restoring term TRES*bi*tc | SRES*bi*sc, or -- coupled to an external atmosphere / sea ice (usrc.F90:742-783) -- the
        // sensible + latent heat flux and sea-ice terms of the surface level (tc = sc = 1 and mc = msi at k = l, else 0, so
        // the lower levels only add exact zeros); TT,SS / SS,TT centre entries below
        bool cpl = false;
        double mc = 0.0;
        if constexpr (CPL) {
            cpl = (R == TT ? t.coupled_T : t.coupled_S) && k == L;
            if (cpl) mc = THCM_LDG(t.msi + (size_t)c.lj * blk.n0 + c.li);
        }
        if (CPL && cpl) {
            if constexpr (R == TT) l5 = l5 + t.cpl_ooa + t.cpl_dedt_t + mc * (t.cpl_qtz - t.cpl_ooa - t.cpl_dedt_t);
            else l5 = l5 - mc * t.cpl_pq * t.cpl_zeta * t.cpl_a0 / t.cpl_rl;
        } else {
            l5 = l5 + KT(R == TT ? K_RT : K_RS, k);
        }
        double l14 = -0.0 - KT(K_ZT14, k) * sm, l23 = -0.0 - KT(K_ZT23, k) * sm;
        // tnlin(3): Utrx, tnlin(5): Vtry, tnlin(7): Wtrz -- identical in rhs and jacobian (usrc.F90:869-872, 983-991)
        double x2 = -(UV(a, gi - 1, gj, k, 0) + UV(a, gi - 1, gj - 1, k, 0)) * c4x * sm;
        double x8 = (UV(a, gi, gj, k, 0) + UV(a, gi, gj - 1, k, 0)) * c4x * sm;
        double x5 = x2 + x8;
        double y4 = -(UV(a, gi, gj - 1, k, 1) + UV(a, gi - 1, gj - 1, k, 1)) * c4y * cv0 * sm;
        double y6 = (UV(a, gi, gj, k, 1) + UV(a, gi - 1, gj, k, 1)) * c4y * cv1 * sm;
        double y5 = y4 + y6;
        double z14 = -WW_(a, gi, gj, k - 1) * sm * tdzi / dfz;
        double z23 = WW_(a, gi, gj, k) * sm * tdzi / dfz;
        double z5 = z14 + z23;
        ENT(2, R) = l2 + x2; ENT(8, R) = l2 + x8;
        ENT(4, R) = l4 + y4; ENT(6, R) = l6 + y6;
        ENT(5, R) = l5 + x5 + y5 + z5;
        ENT(14, R) = l14 + z14; ENT(23, R) = l23 + z23;
        if (CPL && cpl) {
            if constexpr (R == TT) {
                ENT(5, SS) = t.cpl_ts * mc;                                  // Al(TT,SS) = -QTnd*zeta*a0*mc (usrc.F90:755)
            } else {
                const double QSoa = -t.cpl_dedt_s, QSos = t.cpl_pq * t.cpl_zeta / t.cpl_rl;
                ENT(5, TT) = QSoa + mc * (QSos - QSoa);                      // Al(SS,TT) (usrc.F90:774-783)
            }
        }
        if constexpr (JAC) {
            // tnlin(2): urTx, tnlin(4): vrTy, tnlin(6): wrTz (usrc.F90:988-990, 1004-1006)
            double tc = TS(a, gi, gj, k, var);
            double ux_w = -(tc + TS(a, gi - 1, gj, k, var)) * c4x * sm;
            double ux_e = (TS(a, gi + 1, gj, k, var) + tc) * c4x * sm;
            ENT(2, UU) = ux_w; ENT(4, UU) = ux_e; ENT(1, UU) = ux_w; ENT(5, UU) = ux_e;
            double vy_s = -c4y * (tc + TS(a, gi, gj - 1, k, var)) * cv0 * sm;
            double vy_n = c4y * (TS(a, gi, gj + 1, k, var) + tc) * cv1 * sm;
            ENT(4, VV) = vy_s; ENT(1, VV) = vy_s; ENT(5, VV) = vy_n; ENT(2, VV) = vy_n;
            ENT(14, WW) = -tdzi * sm * (tc + TS(a, gi, gj, k - 1, var)) / dfz;
            ENT(5, WW) = k <= L - 1 ? tdzi * sm * (TS(a, gi, gj, k + 1, var) + tc) / dfz : 0.0;
        }
    }
}

// ---------------------------------------------------------------------------
// Tracer mixing (mix_imp.f) without neutral physics and GM (MIXP = MKAP = 0; the reference's own THCM::evaluate cannot insert their
// entries: they fall outside the maximal graph).  What remains of vmix_fun are the two vertical schemes
//   Ftimp(k) = -tprstb(-drhodzt(k), SPL1) * P_VC * dtdzt(k)            implicit mixing / convective adjustment (mix_imp.f:489-492)
//   Ftzt(k)  =  tprstb( drhodzt(k), SPL1) * eps * dtdzt(k) / (drhodzt(k) - 1e-20), eps = (1 - ALPC) * ENER * PE_V
//                                                                      "consistent" vertical mixing, ALPC != 1 (mix_imp.f:478-487)
//   mix_T    = (Ftzt(k) - Ftzt(k-1)) / (dz * dfzT(k)) + (Ftimp(k) - Ftimp(k-1)) / (dz * dfzT(k))   (mix_imp.f:511-524)
// a function of T,S in the cell and its two vertical neighbours only.  tt / ss = T, S at k-1, k, k+1 as usol leaves them;
// oc[3] = isoc (OCEAN or PERIO) of the three cells.  Operation order follows the reference statement by statement.
// The only transcendental is tanh, taken from thcm_tanh.h: one specified algorithm (fdlibm's tanh through expm1, plain IEEE operations)
// on the device, in the host emulation (tests/emu) and in the oracle, so the term AND its forward-difference Jacobian block (one ulp of
// tanh is amplified by 1 / eps = 1e8 there) are bit-exact against the oracle on the B200.
// ---------------------------------------------------------------------------
struct MixTabs { double dfzT, dfzW, dfzWm; int k; };
THCM_HD double mix_tprstb(double grad, double fac) {   // mix_imp.f:837-857
    double a = -grad * fac;
    double th = fd_tanh(a * a * a);
    return th > 0.0 ? th : 0.0;
}
// the four vertical fluxes through ONE cell face, between the cell below (lo) and the cell above (hi) it: implicit mixing Ftimp / Fsimp
// and consistent mixing Ftzt / Fszt.  io = isoc(lo) * isoc(hi), dzw = dz * dfzW(face).  One taper evaluation per scheme and face
// (the reference evaluates tprstb twice with the same argument, for T and for S: the same bits).
struct MixFace { double Ft, Fs, Gt, Gs; };
THCM_HD double mix_rho_of(const DevTables& t, double tt, double ss) {
    constexpr double alpt1 = 2.93, alpt2 = 8.3e-02, alpt3 = 6.6e-04;   // usr.F90:151-153
    return t.mix_lambda * ss - tt - t.mix_xes * (alpt1 * tt + alpt2 * tt * tt - alpt3 * tt * tt * tt);
}
THCM_HD MixFace mix_face(const DevTables& t, double t_lo, double t_hi, double s_lo, double s_hi, double io, double dzw) {
    const double drho = io * (mix_rho_of(t, t_hi, s_hi) - mix_rho_of(t, t_lo, s_lo)) / dzw;
    const double dtz = io * (t_hi - t_lo) / dzw;
    const double dsz = io * (s_hi - s_lo) / dzw;
    MixFace F{0.0, 0.0, 0.0, 0.0};
    if (t.mix_eps != 0.0) {   // mix_imp.f:478-487: Ftzt = Ftzt + tprstb(drhodzt, SPL1) * eps * dtdzt / (drhodzt - epsln)
        const double tp = mix_tprstb(drho, t.mix_fac);
        F.Gt = F.Gt + tp * t.mix_eps * dtz / (drho - 1.0e-20);
        F.Gs = F.Gs + tp * t.mix_eps * dsz / (drho - 1.0e-20);
    }
    if (t.mix_kvc != 0.0) {   // mix_imp.f:489-492
        const double tp = mix_tprstb(-drho, t.mix_fac);
        F.Ft = -(tp * t.mix_kvc * dtz);
        F.Fs = -(tp * t.mix_kvc * dsz);
    }
    return F;
}
// divergence of the fluxes of the faces below (F0) and above (F1) a cell (mix_imp.f:495-560); var = 4: temperature row, 5: salinity row.
// The zonal and meridional differences are exact zeros without neutral physics / GM (refused, thcm_host.cpp).
THCM_HD double mix_combine(const DevTables& t, int var, const MixFace& F0, const MixFace& F1, double dfzT) {
    double mix = 0.0;
    mix = ((var == 4 ? F1.Gt : F1.Gs) - (var == 4 ? F0.Gt : F0.Gs)) / (t.mix_dz * dfzT) + mix;   // mix_imp.f:511-513, 541-543
    if (var == 4) {
        if (t.mix_rho) mix = ((F1.Ft - F0.Ft) - (F1.Fs - F0.Fs) * t.mix_lambda) / (2.0 * t.mix_dz * dfzT) + mix;
        else mix = (F1.Ft - F0.Ft) / (t.mix_dz * dfzT) + mix;
    } else {
        if (t.mix_rho) mix = ((F1.Fs - F0.Fs) - (F1.Ft - F0.Ft) / t.mix_lambda) / (2.0 * t.mix_dz * dfzT) + mix;
        else mix = (F1.Fs - F0.Fs) / (t.mix_dz * dfzT) + mix;
    }
    return mix;
}
// face below cell k: Ftimp(:,:,0) and Ftzt(:,:,0) are never set (k = 1)
THCM_HD MixFace mix_face_below(const DevTables& t, const double* tt, const double* ss, const double* oc, const MixTabs& mt) {
    if (mt.k == 1) return MixFace{0.0, 0.0, 0.0, 0.0};
    return mix_face(t, tt[0], tt[1], ss[0], ss[1], oc[1] * oc[0], t.mix_dz * mt.dfzWm);
}
THCM_HD MixFace mix_face_above(const DevTables& t, const double* tt, const double* ss, const double* oc, const MixTabs& mt) {
    return mix_face(t, tt[1], tt[2], ss[1], ss[2], oc[2] * oc[1], t.mix_dz * mt.dfzW);
}
// var = 4: temperature row, 5: salinity row
THCM_HD double vmix_value(const DevTables& t, int var, const double* tt, const double* ss, const double* oc, const MixTabs& mt) {
    if (!(var == 4 ? t.mix_temp : t.mix_salt)) return 0.0;
    return mix_combine(t, var, mix_face_below(t, tt, ss, oc, mt), mix_face_above(t, tt, ss, oc, mt), mt.dfzT);
}
// gathers the column and evaluates the mixing term of row R (TT or SS) of one cell; nb = the cell's neighbour mask
template <int R, class Tile, class Tabs>
THCM_HD double vmix_rhs(const DevTables& t, const Cell& c, uint32_t nb, const Tile& tile, const Tabs& tabs) {
    static_assert(R == TT || R == SS, "mixing acts on the tracer rows");
    if (!(t.mix_temp | t.mix_salt)) return 0.0;
    double tt[3], ss[3], oc[3];
#pragma unroll
    for (int q = 0; q < 3; q++) { tt[q] = tile(SV_T, 0, 0, q - 1); ss[q] = tile(SV_S, 0, 0, q - 1); }
    oc[0] = ((nb >> 13) & 1u) ? 0.0 : 1.0; oc[1] = ((nb >> 4) & 1u) ? 0.0 : 1.0; oc[2] = ((nb >> 22) & 1u) ? 0.0 : 1.0;
    const MixTabs mt{tabs.kt(K_DFZT), tabs.kt(K_DFZW), tabs.kt(K_DFZWM), c.k};
    return vmix_value(t, R == TT ? 4 : 5, tt, ss, oc, mt);
}
// vmix_jac (mix_imp.f:729-815): forward differences, eps = 1e-8, of vmix_fun w.r.t. the T,S unknowns of the OCEAN cells among the
// neighbours -- here k-1, k, k+1 -- added to An(loc, R, TT|SS) for loc = 14, 5, 23 BEFORE `boundaries`.  (The reference perturbs whole
// colour groups at once; no row meets two columns of a group, so the quotient is the same.)
// A perturbation of the cell below moves only the face below, one of the cell above only the face above: the other face keeps the bits
// of the unperturbed evaluation.  So the whole block of a cell -- its T row AND its S row -- is a function of TEN face evaluations:
//   0 below(base)   2 below(T[k-1]+eps)  3 below(T[k]+eps)  6 below(S[k-1]+eps)  7 below(S[k]+eps)
//   1 above(base)   4 above(T[k]+eps)    5 above(T[k+1]+eps) 8 above(S[k]+eps)   9 above(S[k+1]+eps)
// (28 taper evaluations per row when every perturbed state re-evaluates both faces and both tracers).  mix_face_needed says which of
// them exist for a cell, mix_face_eval computes one, vmix_jac_from_faces turns them into the six entries of row R: the TMA-staged
// Jacobian kernel lets the T warp and the S warp of a tile compute five faces each and exchange them through shared memory.
constexpr int MIX_NFACE = 10;
THCM_HD bool mix_face_needed(const DevTables& t, int f, const double* oc) {
    if (f < 2) return true;
    if (f < 6 ? !t.mix_temp : !t.mix_salt) return false;
    const int g = f < 6 ? f - 2 : f - 6;           // 0: cell below, 1 / 2: this cell (face below / above), 3: cell above
    return g == 0 ? oc[0] != 0.0 : (g == 3 ? oc[2] != 0.0 : oc[1] != 0.0);
}
THCM_HD MixFace mix_face_eval(const DevTables& t, int f, const double* tt, const double* ss, const double* oc, const MixTabs& mt) {
    const double eps = 1.0e-08;
    double tp[3] = {tt[0], tt[1], tt[2]}, sp[3] = {ss[0], ss[1], ss[2]};
    if (f >= 2) {
        const int g = f < 6 ? f - 2 : f - 6;
        const int q = g == 0 ? 0 : (g == 3 ? 2 : 1);
        if (f < 6) tp[q] = tt[q] + eps; else sp[q] = ss[q] + eps;
    }
    const bool below = f == 0 || f == 2 || f == 3 || f == 6 || f == 7;
    return below ? mix_face_below(t, tp, sp, oc, mt) : mix_face_above(t, tp, sp, oc, mt);
}
// F(f): the face evaluation f of this cell (only the needed ones are read)
template <int R, class Faces>
THCM_HD void vmix_jac_from_faces(double* E, const DevTables& t, const Faces& F, const double* oc, const MixTabs& mt) {
    static_assert(R == TT || R == SS, "mixing acts on the tracer rows");
    constexpr int var = R == TT ? 4 : 5;
    const double eps = 1.0e-08;
    const MixFace F0 = F(0), F1 = F(1);
    const double f0 = mix_combine(t, var, F0, F1, mt.dfzT);
    // the neighbour is a column only if it is an OCEAN cell of the domain (k-1 >= 1, k+1 <= L: the frame is LAND)
    if (t.mix_temp) {
        if (oc[0] != 0.0) entref<R, 14, TT>(E) = entref<R, 14, TT>(E) + (mix_combine(t, var, F(2), F1, mt.dfzT) - f0) / eps;
        if (oc[1] != 0.0) entref<R, 5, TT>(E) = entref<R, 5, TT>(E) + (mix_combine(t, var, F(3), F(4), mt.dfzT) - f0) / eps;
        if (oc[2] != 0.0) entref<R, 23, TT>(E) = entref<R, 23, TT>(E) + (mix_combine(t, var, F0, F(5), mt.dfzT) - f0) / eps;
    }
    if (t.mix_salt) {
        if (oc[0] != 0.0) entref<R, 14, SS>(E) = entref<R, 14, SS>(E) + (mix_combine(t, var, F(6), F1, mt.dfzT) - f0) / eps;
        if (oc[1] != 0.0) entref<R, 5, SS>(E) = entref<R, 5, SS>(E) + (mix_combine(t, var, F(7), F(8), mt.dfzT) - f0) / eps;
        if (oc[2] != 0.0) entref<R, 23, SS>(E) = entref<R, 23, SS>(E) + (mix_combine(t, var, F0, F(9), mt.dfzT) - f0) / eps;
    }
}
// the column of a cell as the mixing term sees it
template <class Tile, class Tabs>
THCM_HD MixTabs mix_column(const Cell& c, uint32_t nb, const Tile& tile, const Tabs& tabs, double* tt, double* ss, double* oc) {
#pragma unroll
    for (int q = 0; q < 3; q++) { tt[q] = tile(SV_T, 0, 0, q - 1); ss[q] = tile(SV_S, 0, 0, q - 1); }
    oc[0] = ((nb >> 13) & 1u) ? 0.0 : 1.0; oc[1] = 1.0; oc[2] = ((nb >> 22) & 1u) ? 0.0 : 1.0;
    return MixTabs{tabs.kt(K_DFZT), tabs.kt(K_DFZW), tabs.kt(K_DFZWM), c.k};
}
struct LocalFaces { const MixFace* f; THCM_HD MixFace operator()(int q) const { return f[q]; } };
// one thread does everything for its row (per-position kernels, host emulation)
template <int R, class Tile, class Tabs>
THCM_HD void vmix_jac(double* E, const DevTables& t, const Cell& c, uint32_t nb, const Tile& tile, const Tabs& tabs) {
    static_assert(R == TT || R == SS, "mixing acts on the tracer rows");
    if (!(R == TT ? t.mix_temp : t.mix_salt)) return;
    if ((nb >> 4) & 1u) return;                      // rows of OCEAN cells only (vmix_el_1/2)
    double tt[3], ss[3], oc[3];
    const MixTabs mt = mix_column(c, nb, tile, tabs, tt, ss, oc);
    MixFace F[MIX_NFACE];
#pragma unroll
    for (int f = 0; f < MIX_NFACE; f++) if (mix_face_needed(t, f, oc)) F[f] = mix_face_eval(t, f, tt, ss, oc, mt);
    vmix_jac_from_faces<R>(E, t, LocalFaces{F}, oc, mt);
}

// ---------------------------------------------------------------------------
// boundaries (boundary.F90:80-387) on the structural entries of row R.  Written as a transliteration
// of the reference's statement sequence; statements that touch structurally-zero entries vanish at
// compile time (slot_of(...) < 0).  nb: bit loc-1 = neighbour loc is LAND (bit 4: centre not OCEAN),
// bits 27..31 = southee, easteast, northee, nnorthee, landm(i,j+2,k).
// ---------------------------------------------------------------------------
template <int R, int DST, int SRC, int COL> THCM_HD void addcol(double* E) {
    if constexpr (slot_of(R, SRC, COL) >= 0) {
        static_assert(slot_of(R, DST, COL) >= 0, "fold target must be a structural entry");
        E[slot_of(R, DST, COL)] = E[slot_of(R, DST, COL)] + E[slot_of(R, SRC, COL)];
    }
}
template <int R, int LOC, int COL> THCM_HD void setent(double* E, double v) {
    if constexpr (slot_of(R, LOC, COL) >= 0) E[slot_of(R, LOC, COL)] = v;
}
template <int R, int LOC> THCM_HD void zloc(double* E) {
    setent<R, LOC, 1>(E, 0.0); setent<R, LOC, 2>(E, 0.0); setent<R, LOC, 3>(E, 0.0);
    setent<R, LOC, 4>(E, 0.0); setent<R, LOC, 5>(E, 0.0); setent<R, LOC, 6>(E, 0.0);
}
template <int R, int LOC> THCM_HD void zuv(double* E) { setent<R, LOC, UU>(E, 0.0); setent<R, LOC, VV>(E, 0.0); }
template <int R, int ROW> THCM_HD void zrow(double* E) {
    if constexpr (R == ROW) {
#pragma unroll
        for (int q = 0; q < RowSlots<R>::N; q++) E[q] = 0.0;
    }
}
// "row ROW becomes the identity row": An(:,ROW,:)=0; An(5,:,ROW)=0; An(5,ROW,ROW)=1 (boundary.F90:256-266 etc.)
template <int R, int ROW> THCM_HD void identity_row(double* E) {
    zrow<R, ROW>(E);
    setent<R, 5, ROW>(E, 0.0);
    if constexpr (R == ROW) E[slot_of(R, 5, ROW)] = 1.0;
}

template <int R>
THCM_HD void boundaries(double* E, uint32_t nb, bool i_lt_n, bool j_lt_m) {
    auto land = [&](int loc) { return (nb >> (loc - 1)) & 1u; };
    const bool southee = (nb >> 27) & 1u, easteast = (nb >> 28) & 1u, northee = (nb >> 29) & 1u, nnorthee = (nb >> 30) & 1u,
               nn2 = (nb >> 31) & 1u;
    if (!land(5)) {  // centre == OCEAN
        if (land(14)) {  // bottom
            if (land(11) && land(10) && land(13)) { addcol<R, 1, 10, UU>(E); addcol<R, 1, 10, VV>(E); }
            zuv<R, 10>(E);
            if (land(11) && land(18) && land(15)) { addcol<R, 2, 11, UU>(E); addcol<R, 2, 11, VV>(E); }  // sic: neastb, boundary.F90:91
            zuv<R, 11>(E);
            if (land(17) && land(16) && land(13)) { addcol<R, 4, 13, UU>(E); addcol<R, 4, 13, VV>(E); }
            zuv<R, 13>(E);
            if (land(17) && land(18) && land(15)) { addcol<R, 5, 14, UU>(E); addcol<R, 5, 14, VV>(E); }
            addcol<R, 5, 14, TT>(E); addcol<R, 5, 14, SS>(E);
            zloc<R, 14>(E);
        }
        if (land(10)) zloc<R, 10>(E);
        if (land(11)) zloc<R, 11>(E);
        if (land(12)) zloc<R, 12>(E);
        if (land(13)) zloc<R, 13>(E);
        if (land(15)) zloc<R, 15>(E);
        if (land(16)) zloc<R, 16>(E);
        if (land(17)) zloc<R, 17>(E);
        if (land(18)) zloc<R, 18>(E);
        if (land(23)) {  // top
            if (land(20) && land(19) && land(22)) { addcol<R, 1, 19, UU>(E); addcol<R, 1, 19, VV>(E); }
            zuv<R, 19>(E);
            if (land(20) && land(21) && land(24)) { addcol<R, 2, 20, UU>(E); addcol<R, 2, 20, VV>(E); }
            zuv<R, 20>(E);
            if (land(26) && land(25) && land(22)) { addcol<R, 4, 22, UU>(E); addcol<R, 4, 22, VV>(E); }
            zuv<R, 22>(E);
            if (land(26) && land(27) && land(24)) { addcol<R, 5, 23, UU>(E); addcol<R, 5, 23, VV>(E); }
            addcol<R, 5, 23, TT>(E); addcol<R, 5, 23, SS>(E);
            zloc<R, 23>(E);
            zrow<R, WW>(E);
            // 1e-10 placeholders that the strict threshold later drops (boundary.F90:173-176, assemble.F90:115)
            setent<R, 5, WW>(E, 1.0e-10); setent<R, 6, WW>(E, 1.0e-10); setent<R, 8, WW>(E, 1.0e-10); setent<R, 9, WW>(E, 1.0e-10);
            if constexpr (R == WW) E[slot_of(R, 5, WW)] = 1.0;
        }
        if (land(19)) zloc<R, 19>(E);
        if (land(20)) zloc<R, 20>(E);
        if (land(21)) zloc<R, 21>(E);
        if (land(22)) zloc<R, 22>(E);
        if (land(24)) zloc<R, 24>(E);
        if (land(25)) zloc<R, 25>(E);
        if (land(26)) zloc<R, 26>(E);
        if (land(27)) zloc<R, 27>(E);
        if (land(1)) zuv<R, 1>(E);
        if (land(2)) { addcol<R, 5, 2, TT>(E); addcol<R, 5, 2, SS>(E); zloc<R, 2>(E); zuv<R, 1>(E); }
        if (land(3)) { zuv<R, 2>(E); zuv<R, 3>(E); }
        else if (j_lt_m) { if (nn2) zuv<R, 3>(E); }
        if (land(4)) { addcol<R, 5, 4, SS>(E); addcol<R, 5, 4, TT>(E); zloc<R, 4>(E); zuv<R, 1>(E); }
        if (land(6)) {
            zuv<R, 2>(E);
            if constexpr (R == PP) { setent<R, 2, UU>(E, 0.0); setent<R, 2, VV>(E, 0.0); setent<R, 5, UU>(E, 0.0); setent<R, 5, VV>(E, 0.0); }
            identity_row<R, VV>(E);
            identity_row<R, UU>(E);
            addcol<R, 5, 6, SS>(E); addcol<R, 5, 6, TT>(E);
            zloc<R, 6>(E);
        } else if (j_lt_m) {
            if (nn2) { zuv<R, 3>(E); zuv<R, 6>(E); }
        }
        if (land(7)) { zuv<R, 4>(E); zuv<R, 7>(E); }
        else if (i_lt_n) { if (southee) zuv<R, 7>(E); }
        if (land(8)) {
            zuv<R, 4>(E);
            if constexpr (R == PP) { setent<R, 4, UU>(E, 0.0); setent<R, 4, VV>(E, 0.0); setent<R, 5, UU>(E, 0.0); setent<R, 5, VV>(E, 0.0); }
            identity_row<R, UU>(E);
            identity_row<R, VV>(E);
            addcol<R, 5, 8, SS>(E); addcol<R, 5, 8, TT>(E);
            zloc<R, 8>(E);
            zuv<R, 7>(E);
        } else if (i_lt_n) {
            if (easteast) { zuv<R, 7>(E); zuv<R, 8>(E); }
        }
        if (land(9)) {
            identity_row<R, UU>(E);
            identity_row<R, VV>(E);
            zuv<R, 7>(E);
        } else if (i_lt_n || j_lt_m) {
            if (i_lt_n) {
                if (northee) { zuv<R, 8>(E); zuv<R, 9>(E); }
                else if (j_lt_m) { if (nnorthee) zuv<R, 9>(E); }
            }
            if (j_lt_m) { if (nn2) { zuv<R, 6>(E); zuv<R, 9>(E); } }
        }
    } else {  // centre on land: identity rows (boundary.F90:381-386)
#pragma unroll
        for (int q = 0; q < RowSlots<R>::N; q++) E[q] = 0.0;
        E[slot_of(R, 5, R)] = 1.0;
    }
}


}  // namespace thcm
