// =============================================================================
// thcm_assembly.cu -- fused THCM residual / Jacobian kernels for sm_100a.
//
// One kernel family replaces, per Newton step, the reference's
//   An = Al; nlin_rhs|nlin_jac; boundaries; fillcolA; matAvec; B = -Au + Frc
// (usrc.F90:449-603, spf.F90, boundary.F90, assemble.F90:57-139, matetc.F90:147-166).
// The dense per-cell block An(27,6,6) is never materialised: each (cell,row) pair owns the <= 24
// structurally non-zero entries of its row (thcm_slots.h) in registers, evaluates them from the
// state and the host-built 1-D metric tables in the reference's operation order (no FMA: the
// library is compiled with --fmad=false), applies `boundaries` as compile-time-pruned folds and
// emits, depending on MODE,
//   RHS       : B(row)            (matrix-free residual, Fortran summation order, same 1e-10 drop)
//   JAC_GRAPH : values into the static maximal-graph CSR (explicit zeros kept, THCM.C:1052,1095)
//   JAC_COUNT : JAC_GRAPH + number of |a|>1e-10 entries per assembly block
//   JAC_CRS   : Fortran-order thresholded CRS begA/jcoA/coA (1-based), after the block-count scan
//
// Tile = TI = 32 consecutive cells of one (j,k) grid line.
//  Phase 1 (all threads): stage the 3x3x(TI+2) neighbourhood -- fields as usol leaves them -- and the j/k slices of
//    the metric tables in shared memory.  Positions inside the owned block take a fast path (3 x LDG.128 + one mask
//    byte), only edge columns / halo / mirror rows run the general rule; all loads of a block are issued together:
//    one latency exposure instead of ~60 dependent ones.
//  Phase 2: warp r evaluates row type(s) r for the 32 cells from shared memory only: u | v | w+p | T | S
//    (24 | 22 | 18 | 20 | 20 entries: balanced, no divergence on the row type).
//  Phase 3: the tile's values, staged in shared memory cell by cell, leave through the TMA unit: one
//    cp.async.bulk.global.shared::cta of 832 B per cell (SASS UBLKCP) -- clipped edge tiles use a plain coalesced loop.
// HBM traffic per cell: 48 B state + 5 B masks in, 832 B values out.
// =============================================================================
#include <algorithm>
#include <cstddef>
#include <cstdio>
#include <type_traits>
#include "thcm_cell.cuh"

namespace thcm {

__constant__ ClassTables c_cls;

void upload_class_tables(const ClassTables& t) { THCM_CUDA(cudaMemcpyToSymbol(c_cls, &t, sizeof(ClassTables))); }

constexpr int TI = CELLS_PER_BLOCK;   // 32 cells per tile
constexpr int TW = TI + 2;            // staged width (one neighbour column each side)
// graph-mode output staging: cell-major, stride 106 doubles = 848 B: a multiple of 16 B (TMA bulk store source alignment)
// and only a 2-way bank conflict when the 32 lanes of a warp each write entry p of their own cell (104 would be 8-way)
constexpr int VSTRIDE = NSLOT_TOTAL + 2;

constexpr __host__ __device__ int mode_warps(int mode) { return mode == MODE_JAC_CRS ? 6 : 5; }

template <int NSV> struct SmemIn {
    double st[NSV][3][3][TW];         // [field][dk+1][dj+1][x]
    double tj[J_COUNT][3];            // j-tables at gj-1, gj, gj+1
    double tk[K_COUNT];               // k-tables at k
};
template <int MODE> struct Smem;
template <> struct Smem<MODE_RHS> { SmemIn<SV_NRHS> in; };
template <> struct Smem<MODE_JAC_GRAPH> { SmemIn<SV_NJAC> in; alignas(16) double v[TI * VSTRIDE]; int cstart[TI + 1]; };
template <> struct Smem<MODE_JAC_COUNT> { SmemIn<SV_NJAC> in; alignas(16) double v[TI * VSTRIDE]; int cstart[TI + 1]; int cnt[NUN]; };
template <> struct Smem<MODE_JAC_CRS> { SmemIn<SV_NJAC> in; double v[TI * NSLOT_TOTAL]; int c[TI * NSLOT_TOTAL]; int off[TI * NUN + 1]; };

template <int NSV> struct SmemTile {
    const SmemIn<NSV>* in; int lane;
    __device__ __forceinline__ double operator()(int sv, int di, int dj, int dk) const { return in->st[sv][dk + 1][dj + 1][lane + 1 + di]; }
};
template <int NSV> struct SmemTabs {
    const SmemIn<NSV>* in;
    __device__ __forceinline__ double jt(int tb, int dj) const { return in->tj[tb][dj + 1]; }
    __device__ __forceinline__ double kt(int tb) const { return in->tk[tb]; }
};

__device__ __forceinline__ TileGeom tile_geom(const DevBlock& b) { return tile_geom_of(b, blockIdx.x); }

template <int NSV, int NT>
__device__ __forceinline__ void stage_inputs(const AsmArgs& a, const TileGeom& g, SmemIn<NSV>& in) {
    const DevBlock& b = a.b;
    // ---- positions: 9 grid lines (dj, dk) x (ncell + 2) columns.  A line inside the owned block is a run of contiguous
    //      48-byte records (fast path); the descriptor is recomputed per position (a dozen integer ops) so that every
    //      global load of the block -- records, mask bytes, table slices -- is in flight at once: ONE latency exposure ----
    const int w = g.ncell + 2;
    for (int p = threadIdx.x; p < 9 * TW; p += NT) {
        const int r = p / TW, x = p - r * TW;
        if (x >= w) continue;
        const int dj = r % 3 - 1, dk = r / 3 - 1;
        const int gj2 = g.gj + dj, k2 = g.k + dk, je = gj2 - 1 - b.j0;
        double out[NSV];
        if (je >= 0 && je < b.m0 && k2 >= 1 && k2 <= b.L && x >= 1 && x <= g.ncell) {   // inside the block => inside the domain
            const size_t cell = ((size_t)(k2 - 1) * b.m0 + je) * b.n0 + (g.gi0 - 1 - b.i0) + (x - 1);
            const uint8_t live = __ldg(a.uvlive + ((size_t)(k2 - 1) * (b.m0 + 2) + (gj2 - b.j0)) * (b.n0 + 2) + (g.gi0 - b.i0) + (x - 1));
            stage_regular<NSV>(a.un + (size_t)NUN * cell, live != 0, k2 != b.L, out);
        } else {
            stage_position<NSV>(a, g.gi0 - 1 + x, g.gj + dj, g.k + dk, out);
        }
#pragma unroll
        for (int sv = 0; sv < NSV; sv++) in.st[sv][dk + 1][dj + 1][x] = out[sv];
    }
    const DevTables& t = a.t;
    for (int q = threadIdx.x; q < J_COUNT * 3 + K_COUNT; q += NT) {
        if (q < J_COUNT * 3) { int tb = q / 3, d = q % 3; in.tj[tb][d] = __ldg(t.jt + (size_t)tb * t.jstride + g.gj + d - 1); }
        else { int tb = q - J_COUNT * 3; in.tk[tb] = __ldg(t.kt + (size_t)tb * t.kstride + g.k); }
    }
}

template <int R, int MODE, bool CPL>
__device__ __forceinline__ void do_row(const AsmArgs& a, Smem<MODE>& sh, const TileGeom& g, int lane, uint32_t nb_in, double sm) {
    constexpr int NSV = MODE == MODE_RHS ? SV_NRHS : SV_NJAC;
    const DevBlock& b = a.b;
    const int cell = g.cell0 + lane;
    const bool active = lane < g.ncell;
    double E[RowSlots<R>::N];
    Cell c{0, 0, 0, 0, 0};
    int cls = 0;
    uint32_t nb = 0;
    SmemTile<NSV> tile{&sh.in, lane};
    if (active) {
        c.li = cell % b.n0; c.lj = g.lj; c.gi = g.gi0 + lane; c.gj = g.gj; c.k = g.k;
        nb = nb_in;
        cls = (c.gi == 1 ? 1 : 0) | (c.gi == b.N ? 2 : 0) | (c.gj == 1 ? 4 : 0) | (c.gj == b.M ? 8 : 0) | (c.k == 1 ? 16 : 0) |
              (c.k == b.L ? 32 : 0);
        if (!((nb >> 4) & 1u)) {
            eval_row<R, MODE != MODE_RHS, CPL>(E, a.t, b, c, sm, tile, SmemTabs<NSV>{&sh.in});
            if constexpr (MODE != MODE_RHS && (R == TT || R == SS)) vmix_jac<R>(E, a.t, c, nb, tile, SmemTabs<NSV>{&sh.in});   // usrc.F90:489-508
        }
    }
    // open ocean away from the bottom and the lid: no LAND among the 27 (+5) neighbours of any cell of the warp, so every
    // statement of `boundaries` is a no-op -- skip its ~60 predicated blocks (warp-uniform branch)
    const bool open_ocean = __all_sync(0xffffffffu, !active || nb == 0u);
    if (active) {
        if (!open_ocean) boundaries<R>(E, nb, c.gi < b.N, c.gj < b.M);
        // strict threshold of fillcolA (assemble.F90:115); the graph keeps explicit zeros instead
#pragma unroll
        for (int q = 0; q < RowSlots<R>::N; q++) E[q] = fabs(E[q]) > DROP_TOL ? E[q] : 0.0;
    }

    if constexpr (MODE == MODE_RHS) {
        if (active) {
            // matAvec (matetc.F90:160-164): v2 = coA*v1(jcoA) + v2 in CRS order = slot order
            double s = 0.0;
            static_for<0, RowSlots<R>::N>([&](auto qc) {
                constexpr int q = decltype(qc)::value;
                constexpr int loc = row_slots(R)[q].loc, col = row_slots(R)[q].col;
                int gi2 = c.gi + loc_di(loc), gj2 = c.gj + loc_dj(loc), k2 = c.k + loc_dk(loc);
                // a kept entry always points inside the domain (the dummy LAND frame removes the others)
                bool inside = gj2 >= 1 && gj2 <= b.M && k2 >= 1 && k2 <= b.L && (b.periodic || (gi2 >= 1 && gi2 <= b.N));
                if (E[q] != 0.0 && inside) s = E[q] * tile(SV_RAW + col - 1, loc_di(loc), loc_dj(loc), loc_dk(loc)) + s;
            });
            // B = -Au - mix + Frc - p0*(1-par(RESC))*ures ; B *= (1 - landm) (usrc.F90:576-591)
            int row = NUN * cell + R - 1;
            double mixv = 0.0;
            if constexpr (R == TT || R == SS) mixv = vmix_rhs<R>(a.t, c, nb, tile, SmemTabs<NSV>{&sh.in});   // vmix_fun, usrc.F90:551-571
            double B = -s - mixv + a.frc[row] - 0.0;
            B = B * (((nb >> 4) & 1u) ? 0.0 : 1.0);
            a.out[row] = a.sign * B;
        }
        return;
    } else {
        if constexpr (MODE == MODE_JAC_GRAPH || MODE == MODE_JAC_COUNT) {
            int cnt = 0;
            // interior cells (the bulk): graph positions are compile-time constants; edge classes go through the table
            const bool interior = __all_sync(0xffffffffu, !active || cls == 0);
            if (active) {
                if (interior) {
                    double* dst = sh.v + lane * VSTRIDE + ROW_OFF[R - 1];
                    static_for<0, RowSlots<R>::N>([&](auto qc) {
                        constexpr int q = decltype(qc)::value;
                        dst[interior_pos(R, q)] = E[q];
                        if (E[q] != 0.0) cnt++;
                    });
                } else {
                    int rowoff = 0;   // offset of row R inside the cell's block of the graph = sum of the shorter rows' lengths
#pragma unroll
                    for (int r = 0; r < R - 1; r++) rowoff += c_cls.rowlen[cls][r];
                    const int base = lane * VSTRIDE + rowoff;
                    static_for<0, RowSlots<R>::N>([&](auto qc) {
                        constexpr int q = decltype(qc)::value;
                        int p = c_cls.pos[cls][ROW_OFF[R - 1] + q];
                        if (p >= 0) sh.v[base + p] = E[q];
                        if (E[q] != 0.0) cnt++;
                    });
                }
            }
            if constexpr (MODE == MODE_JAC_COUNT) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
                if (lane == 0) sh.cnt[R - 1] = cnt;
            }
        } else {  // MODE_JAC_CRS, phase 1: counts per (cell,row) into sh.off (cell-major, row-minor)
            int cnt = 0;
            if (active) {
#pragma unroll
                for (int q = 0; q < RowSlots<R>::N; q++) cnt += E[q] != 0.0 ? 1 : 0;
            }
            sh.off[lane * NUN + R - 1] = cnt;
            __syncthreads();
            // exclusive scan over the 192 counts by warp 0 (6 per lane)
            if (R == 1) {
                int loc[NUN], tot = 0;
#pragma unroll
                for (int r = 0; r < NUN; r++) { loc[r] = tot; tot += sh.off[lane * NUN + r]; }
                int incl = tot;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
                int excl = incl - tot;
#pragma unroll
                for (int r = 0; r < NUN; r++) sh.off[lane * NUN + r] = excl + loc[r];
                if (lane == 31) sh.off[TI * NUN] = incl;
            }
            __syncthreads();
            if (active) {
                int o = sh.off[lane * NUN + R - 1];
                const int bbase = a.blockcnt[blockIdx.x];
                a.begA[NUN * cell + R - 1] = bbase + o + 1;  // 1-based
                static_for<0, RowSlots<R>::N>([&](auto qc) {
                    constexpr int q = decltype(qc)::value;
                    constexpr int loc = row_slots(R)[q].loc, col = row_slots(R)[q].col;
                    if (E[q] != 0.0) {
                        int gi2 = c.gi + loc_di(loc), gj2 = c.gj + loc_dj(loc), k2 = c.k + loc_dk(loc);
                        if (b.periodic) { if (gi2 == 0) gi2 = b.N; else if (gi2 == b.N + 1) gi2 = 1; }  // shift, assemble.F90:171-177
                        sh.v[o] = E[q];
                        sh.c[o] = NUN * ((k2 - 1) * b.N * b.M + b.N * (gj2 - 1) + gi2 - 1) + col;  // find_row2
                        o++;
                    }
                });
            }
        }
    }
}

// TMA bulk store shared -> global (SASS UBLKCP); source and destination 16-byte aligned, size a multiple of 16
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, int bytes) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n" ::"l"(gdst), "r"(s), "r"(bytes) : "memory");
}

template <int MODE, bool CPL>
__global__ void __launch_bounds__(32 * mode_warps(MODE)) thcm_assemble_kernel(const AsmArgs a) {
    constexpr int NT = 32 * mode_warps(MODE);
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Smem<MODE>& sh = *reinterpret_cast<Smem<MODE>*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const TileGeom g = tile_geom(a.b);
    // per-cell mask words are fetched before the staging so that their latency overlaps it (lane <-> cell)
    uint32_t nb = 0; double sm = 0.0; int g0 = 0;
    if (lane < g.ncell) {
        nb = __ldg(a.nbmask + g.cell0 + lane);
        sm = (double)(int)(int8_t)__ldg(a.surf + g.lj * a.b.n0 + (g.cell0 + lane) % a.b.n0);
    }
    if constexpr (MODE == MODE_JAC_GRAPH || MODE == MODE_JAC_COUNT) g0 = __ldg(a.rowptr + NUN * g.cell0);
    stage_inputs<MODE == MODE_RHS ? SV_NRHS : SV_NJAC, NT>(a, g, sh.in);
    if constexpr (MODE == MODE_JAC_GRAPH || MODE == MODE_JAC_COUNT) {
        // start of every cell's entries relative to the tile's first entry (cells at domain edges hold clipped rows)
        if (threadIdx.x <= g.ncell) sh.cstart[threadIdx.x] = __ldg(a.rowptr + NUN * (g.cell0 + threadIdx.x)) - g0;
    }
    __syncthreads();
    if constexpr (MODE == MODE_JAC_CRS) {
        switch (warp) {
        case 0: do_row<1, MODE, CPL>(a, sh, g, lane, nb, sm); break;
        case 1: do_row<2, MODE, CPL>(a, sh, g, lane, nb, sm); break;
        case 2: do_row<3, MODE, CPL>(a, sh, g, lane, nb, sm); break;
        case 3: do_row<4, MODE, CPL>(a, sh, g, lane, nb, sm); break;
        case 4: do_row<5, MODE, CPL>(a, sh, g, lane, nb, sm); break;
        default: do_row<6, MODE, CPL>(a, sh, g, lane, nb, sm); break;
        }
    } else {   // 5 warps: u | v | w + p | T | S
        switch (warp) {
        case 0: do_row<1, MODE, CPL>(a, sh, g, lane, nb, sm); break;
        case 1: do_row<2, MODE, CPL>(a, sh, g, lane, nb, sm); break;
        case 2: do_row<3, MODE, CPL>(a, sh, g, lane, nb, sm); do_row<4, MODE, CPL>(a, sh, g, lane, nb, sm); break;
        case 3: do_row<5, MODE, CPL>(a, sh, g, lane, nb, sm); break;
        default: do_row<6, MODE, CPL>(a, sh, g, lane, nb, sm); break;
        }
    }
    if constexpr (MODE == MODE_JAC_GRAPH || MODE == MODE_JAC_COUNT) {
        const int tot = sh.cstart[g.ncell];
        double* gdst = a.val + g0;
        const bool bulk = tot == g.ncell * NSLOT_TOTAL && (((uintptr_t)gdst) & 15) == 0;   // nothing clipped, aligned
        if (bulk) {
            // generic-proxy writes (STS) must be visible to the async proxy before the TMA unit reads them
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            __syncthreads();
            if (threadIdx.x < g.ncell) {
                bulk_store(gdst + (size_t)threadIdx.x * NSLOT_TOTAL, sh.v + threadIdx.x * VSTRIDE, NSLOT_TOTAL * (int)sizeof(double));
                asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");   // smem must stay valid until it has been read
            }
        } else {
            __syncthreads();
            for (int q = threadIdx.x; q < tot; q += NT) {
                int cl = min(q / NSLOT_TOTAL, g.ncell - 1);      // a few steps off at clipped edges
                while (q < sh.cstart[cl]) cl--;
                while (q >= sh.cstart[cl + 1]) cl++;
                gdst[q] = sh.v[cl * VSTRIDE + (q - sh.cstart[cl])];
            }
        }
        if constexpr (MODE == MODE_JAC_COUNT) {
            if (threadIdx.x == 0) a.blockcnt[blockIdx.x] = sh.cnt[0] + sh.cnt[1] + sh.cnt[2] + sh.cnt[3] + sh.cnt[4] + sh.cnt[5];
        }
    } else if constexpr (MODE == MODE_JAC_CRS) {
        __syncthreads();
        const int bbase = a.blockcnt[blockIdx.x], tot = sh.off[TI * NUN];
        for (int q = threadIdx.x; q < tot; q += NT) { a.coA[bbase + q] = sh.v[q]; a.jcoA[bbase + q] = sh.c[q]; }
        if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) a.begA[NUN * a.b.ncell] = bbase + tot + 1;
    }
}


// =============================================================================
// TMA staging helpers of the Jacobian kernels: grid lines of raw 48-byte state records, the tile descriptor and the j / k table
// records arrive by cp.async.bulk into one Stage, completion through an mbarrier.
// (A persistent, warp-specialised variant -- loader warp, five consumer warps, storer warp, double-buffered stages -- was built and
// measured in round 1: 2x SLOWER than one block per tile, instruction-fetch bound with five row-type instruction streams per CTA;
// removed in round 2)
// =============================================================================
// LINES: bit r set = grid line r = (dk+1)*3 + (dj+1) of the 3x3 neighbourhood is staged (compacted in bit order)
constexpr __host__ __device__ int popc9(int m) { int c = 0; for (int i = 0; i < 9; i++) c += (m >> i) & 1; return c; }
template <int LINES> struct alignas(16) Stage {
    static constexpr int MASK = LINES, NL = popc9(LINES);
    static constexpr __host__ __device__ int slot(int r) { return popc9(LINES & ((1 << r) - 1)); }
    double rec[NL][TW][NUN];       // raw records; u,v,w zeroed in place where usol does
    TileDesc desc;
    double tj[J_COUNT][JREC];
    double tk[K_COUNT];
};
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(void* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(void* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(void* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(void* bar, uint32_t parity, int tag = 0) {
    uint32_t done = 0, polls = 0;
    const long long t0 = clock64();
    while (true) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
                     : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
        if (done) break;
        if ((++polls & 255u) == 0 && clock64() - t0 > (tag == 4 ? 1000000000ll : (tag == 1 || tag == 3) ? 2000000000ll : 4000000000ll)) {           // a lost hand-off must fail loudly, never hang the GPU
            if ((threadIdx.x & 31) == 0) printf("thcm_jac_tma: block %d warp %d stuck at wait %d (parity %u)\n", blockIdx.x, threadIdx.x >> 5, tag, parity);
            __trap();
        }
    }
}
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, void* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n"
                 ::"r"(smem_u32(sdst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

template <class ST> struct PipeTile {
    const ST* st; int lane;
    __device__ __forceinline__ double operator()(int sv, int di, int dj, int dk) const {
        return st->rec[ST::slot((dk + 1) * 3 + (dj + 1))][lane + 1 + di][sv <= SV_W ? sv : sv + 1];   // u v w . T S
    }
};
template <class ST> struct PipeTabs {
    const ST* st;
    __device__ __forceinline__ double jt(int tb, int dj) const { return st->tj[tb][dj + 1]; }
    __device__ __forceinline__ double kt(int tb) const { return st->tk[tb]; }
};

// the ten face evaluations of the mixing Jacobian of a tile in shared memory: face f of the cell of `lane` at mx[f * TI + lane]
struct SharedFaces {
    const MixFace* mx; int lane;
    __device__ __forceinline__ MixFace operator()(int f) const { return mx[f * TI + lane]; }
};
// T | S row group with tracer mixing: the two warps of a tile (T row, S row) need the SAME ten face evaluations per cell; warp 0
// computes the faces {0, 2, 3, 4, 5} (base below + the temperature perturbations), warp 1 {1, 6, 7, 8, 9} (base above + the salinity
// perturbations) -- five taper evaluations per thread instead of ten.  The caller separates this from the readers by a block barrier.
template <class ST>
__device__ __forceinline__ void mix_faces_to_shared(const AsmArgs& a, const ST& st, const TileGeom& g, int warp, int lane, uint32_t nb,
                                                    MixFace* mx) {
    if (lane >= g.ncell || ((nb >> 4) & 1u)) return;      // rows of OCEAN cells only (vmix_el_1/2)
    const Cell c{g.gi0 + lane, g.gj, g.k, (g.cell0 + lane) % a.b.n0, g.lj};
    double tt[3], ss[3], oc[3];
    const MixTabs mt = mix_column(c, nb, PipeTile<ST>{&st, lane}, PipeTabs<ST>{&st}, tt, ss, oc);
#pragma unroll
    for (int q = 0; q < MIX_NFACE / 2; q++) {
        const int f = warp == 0 ? (q == 0 ? 0 : q + 1) : (q == 0 ? 1 : q + 5);
        if (mix_face_needed(a.t, f, oc)) mx[f * TI + lane] = mix_face_eval(a.t, f, tt, ss, oc, mt);
    }
}
// SHARED_MIX: the mixing Jacobian takes its face evaluations from shared memory (mx; nullptr = mixing is off) instead of computing all
// ten per thread
template <int R, bool CPL, bool SHARED_MIX = false, class ST>
__device__ __forceinline__ void pipe_eval(const AsmArgs& a, const ST& st, const TileGeom& g, int lane, uint32_t nb, double sm,
                                          double* E, const MixFace* mx = nullptr) {
    if (lane < g.ncell && !((nb >> 4) & 1u)) {
        Cell c{g.gi0 + lane, g.gj, g.k, (g.cell0 + lane) % a.b.n0, g.lj};
        eval_row<R, true, CPL>(E, a.t, a.b, c, sm, PipeTile<ST>{&st, lane}, PipeTabs<ST>{&st});
        if constexpr (R == TT || R == SS) {   // usrc.F90:489-508
            if constexpr (!SHARED_MIX) vmix_jac<R>(E, a.t, c, nb, PipeTile<ST>{&st, lane}, PipeTabs<ST>{&st});
            else if (mx != nullptr && (R == TT ? a.t.mix_temp : a.t.mix_salt)) {
                double tt[3], ss[3], oc[3];
                const MixTabs mt = mix_column(c, nb, PipeTile<ST>{&st, lane}, PipeTabs<ST>{&st}, tt, ss, oc);
                vmix_jac_from_faces<R>(E, a.t, SharedFaces{mx, lane}, oc, mt);
            }
        }
    }
}
// open_ocean / interior are TILE-uniform (descriptor flag, tile geometry): no warp votes, inactive lanes of a ragged tile
// just skip the bodies
template <int R>
__device__ __forceinline__ void pipe_finish(const AsmArgs& a, const TileGeom& g, int lane, uint32_t nb, bool open_ocean, double* E) {
    if (lane < g.ncell) {
        if (!open_ocean) boundaries<R>(E, nb, g.gi0 + lane < a.b.N, g.gj < a.b.M);
#pragma unroll
        for (int q = 0; q < RowSlots<R>::N; q++) E[q] = fabs(E[q]) > DROP_TOL ? E[q] : 0.0;
    }
}
// SEG_ROW0: first row type staged in `v` (the staging holds rows SEG_ROW0..6 or a prefix of them), VS: doubles per cell
template <int R, int SEG_ROW0 = 1, int VS = VSTRIDE>
__device__ __forceinline__ void pipe_emit(const AsmArgs& a, double* v, const TileGeom& g, int lane, bool interior, const double* E) {
    if (lane < g.ncell) {
        if (interior) {
            constexpr int off = ROW_OFF[R - 1] - ROW_OFF[SEG_ROW0 - 1];
            double* dst = v + lane * VS + off;
            static_for<0, RowSlots<R>::N>([&](auto qc) {
                constexpr int q = decltype(qc)::value;
                dst[interior_pos(R, q)] = E[q];
            });
        } else {
            const int gi = g.gi0 + lane;
            const int cls = (gi == 1 ? 1 : 0) | (gi == a.b.N ? 2 : 0) | (g.gj == 1 ? 4 : 0) | (g.gj == a.b.M ? 8 : 0) | (g.k == 1 ? 16 : 0) |
                            (g.k == a.b.L ? 32 : 0);
            int rowoff = 0;   // offset of row R from the first staged row = sum of the rows' clipped lengths in between
#pragma unroll
            for (int r = SEG_ROW0 - 1; r < R - 1; r++) rowoff += c_cls.rowlen[cls][r];
            const int base = lane * VS + rowoff;
            static_for<0, RowSlots<R>::N>([&](auto qc) {
                constexpr int q = decltype(qc)::value;
                int p = c_cls.pos[cls][ROW_OFF[R - 1] + q];
                if (p >= 0) v[base + p] = E[q];
            });
        }
    }
}

// =============================================================================
// One-block-per-tile Jacobian kernel with TMA staging (MODE_JAC_GRAPH default).  Same three phases as
// thcm_assemble_kernel, but
//   * phase 1 is nine cp.async.bulk line copies (+ descriptor + table records) issued by one warp and awaited on an
//     mbarrier by everybody, followed by the in-place usol fix-up from the descriptor bits: ~30 instead of ~240
//     instructions per warp, no index arithmetic per position;
//   * tile-uniform facts (open ocean, interior, all land) come from the host-built descriptor flags instead of warp votes;
//   * tiles that are entirely LAND (identity rows: a state-independent pattern) skip staging and evaluation altogether.
// The persistent variant above measured SLOWER on the B200 (instruction-fetch bound: five row types x 2-3 CTAs do not share
// the 32 KB L1.5 instruction cache the way 4-5 co-scheduled blocks do, ncu stall_no_instruction 16 of 28 cycles per issue).
// =============================================================================
// Row groups: A = u | v | w+p (64 entries, bytes [0, 512) of an interior cell's 832-byte record), B = T | S (40 entries,
// bytes [512, 832)).  One kernel per group: every SM then runs at most three (two) distinct instruction streams, which fit
// the 32 KB L1.5 instruction cache (with all five in one kernel the top stall was no_instruction), blocks are smaller
// (more tiles in flight per SM) and the group boundary is a 32-byte sector boundary of the record.
template <int GROUP> struct RowGroup;
// LINES: the grid lines the group's rows read (eval_row): u,v rows -- u,v on level k (dj -1..1) and k+-1 (dj 0), w on levels
// k-1, k at dj 0, 1; w row -- T on k, k+1; T,S rows -- u,v on level k (dj -1, 0), w on k-1, k, T/S on the 7-point star
template <> struct RowGroup<0> { static constexpr int ROW0 = 1, ROW1 = 4, NWARP = 3, LEN = 64, VS = 66, LINES = 0xBE; };
template <> struct RowGroup<1> { static constexpr int ROW0 = 5, ROW1 = 6, NWARP = 2, LEN = 40, VS = 42, LINES = 0xBA; };

template <int GROUP> struct alignas(16) TmaSmem {
    Stage<RowGroup<GROUP>::LINES> st;
    double v[TI * RowGroup<GROUP>::VS];
    int cstart[TI + 2];
    unsigned long long bar;
};
template <int G> __device__ __forceinline__ Stage<RowGroup<G>::LINES>& smem_stage(TmaSmem<G>& s) { return s.st; }
template <int G> __device__ __forceinline__ double* smem_out(TmaSmem<G>& s) { return s.v; }

__device__ __forceinline__ constexpr int diag_pos(int R) { return ROW_OFF[R - 1] + interior_pos(R, slot_of(R, 5, R)); }

template <int GROUP, int RA, int RB, bool CPL, class SH>
__device__ __forceinline__ void tma_rows(const AsmArgs& a, SH& sh, const TileGeom& g, int lane, bool open_ocean, bool interior) {
    using G = RowGroup<GROUP>;
    constexpr int NA = RowSlots<RA>::N, NB = RowSlots<RB>::N;
    auto& st = smem_stage<GROUP>(sh);
    double* v = smem_out<GROUP>(sh);
    const uint32_t nb = st.desc.nbmask[lane];
    const double sm = (double)((st.desc.surfbits >> lane) & 1u);
    double EA[NA], EB[NB];
    pipe_eval<RA, CPL>(a, st, g, lane, nb, sm, EA);
    pipe_finish<RA>(a, g, lane, nb, open_ocean, EA);
    pipe_emit<RA, G::ROW0, G::VS>(a, v, g, lane, interior, EA);
    if constexpr (RB != RA) {
        pipe_eval<RB, CPL>(a, st, g, lane, nb, sm, EB);
        pipe_finish<RB>(a, g, lane, nb, open_ocean, EB);
        pipe_emit<RB, G::ROW0, G::VS>(a, v, g, lane, interior, EB);
    }
}
template <int GROUP, int BLOCKS_PER_SM, bool CPL>
__global__ void __launch_bounds__(32 * RowGroup<GROUP>::NWARP, BLOCKS_PER_SM) thcm_jac_tma_kernel(const AsmArgs a) {
    using G = RowGroup<GROUP>;
    constexpr int NT = 32 * G::NWARP;
    constexpr int SEG0 = ROW_OFF[G::ROW0 - 1];   // first entry of the group inside an interior cell record
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using SH = TmaSmem<GROUP>;
    SH& sh = *reinterpret_cast<SH*>(smem_raw);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = a.tile_list ? a.tile_list[blockIdx.x] : (int)blockIdx.x;
    using ST = Stage<G::LINES>;
    ST& st = smem_stage<GROUP>(sh);
    double* const vout = smem_out<GROUP>(sh);
    const TileGeom g = tile_geom_of(a.b, tile);
    const int w = g.ncell + 2;
    // the loads go out before anything is known about the tile (one latency exposure): per staged grid line one
    // cp.async.bulk run of raw records (two or three at the seam / block edges), the descriptor and the table records
    if (threadIdx.x == 0) {
        mbar_init(&sh.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0)
            mbar_expect_tx(&sh.bar, (uint32_t)(ST::NL * w * NUN * sizeof(double) + sizeof(TileDesc) + sizeof(st.tj) + sizeof(st.tk)));
        __syncwarp();
        if (lane < 9) {
            if ((G::LINES >> lane) & 1) {
                LineSeg seg[3];
                const int ns = line_plan(a.b, g, lane, seg);
#pragma unroll
                for (int q = 0; q < 3; q++)
                    if (q < ns)
                        bulk_load(&st.rec[ST::slot(lane)][seg[q].x0][0], (seg[q].halo ? a.halo : a.un) + (size_t)NUN * seg[q].idx,
                                  (uint32_t)(seg[q].n * NUN * sizeof(double)), &sh.bar);
            }
        } else if (lane == 9) bulk_load(&st.desc, a.tdesc + tile, sizeof(TileDesc), &sh.bar);
        else if (lane == 10) bulk_load(st.tj, a.jrec + (size_t)g.gj * J_COUNT * JREC, sizeof(st.tj), &sh.bar);
        else if (lane == 11) bulk_load(st.tk, a.krec + (size_t)g.k * K_COUNT, sizeof(st.tk), &sh.bar);
        mbar_wait(&sh.bar, 0, 4);   // one warp polls, the others sleep at the barrier
    }
    __syncthreads();
    const uint32_t flags = st.desc.flags;
    const int g0 = st.desc.g0;
    const bool fast = (flags & 1u) != 0, open_ocean = (flags & 2u) != 0, all_land = (flags & 4u) != 0;
    const bool interior = g.gi0 > 1 && g.gi0 + g.ncell - 1 < a.b.N && g.gj > 1 && g.gj < a.b.M && g.k > 1 && g.k < a.b.L;
    if (all_land && interior && fast) {
        // identity rows only: zeros with a one on the diagonal of every row of the group
        double2* v2 = reinterpret_cast<double2*>(vout);
        for (int i = threadIdx.x; i < g.ncell * (G::VS / 2); i += NT) v2[i] = make_double2(0.0, 0.0);
        __syncthreads();
        constexpr int NR = G::ROW1 - G::ROW0 + 1;
        for (int i = threadIdx.x; i < g.ncell * NR; i += NT) {
            const int cell = i / NR, r = G::ROW0 + (i - cell * NR);
            const int dp = r == 1 ? diag_pos(1) : r == 2 ? diag_pos(2) : r == 3 ? diag_pos(3) : r == 4 ? diag_pos(4) : r == 5 ? diag_pos(5) : diag_pos(6);
            vout[cell * G::VS + dp - SEG0] = 1.0;
        }
    } else {
        // usol in place: no-slip zeroing of u,v, lid / bottom / ghost-column rule of w (descriptor bits)
        for (int p = threadIdx.x; p < ST::NL * TW; p += NT) {
            const int sl = p / TW, x = p - sl * TW;
            int r = 0;   // grid line held by slot sl
#pragma unroll
            for (int i = 0, c = 0; i < 9; i++) if ((G::LINES >> i) & 1) { if (c == sl) r = i; c++; }
            if (x < w) {
                if (!((st.desc.uvbits[r] >> x) & 1ull)) *reinterpret_cast<double2*>(&st.rec[sl][x][0]) = make_double2(0.0, 0.0);
                if (!((st.desc.wbits[r] >> x) & 1ull)) st.rec[sl][x][2] = 0.0;
            }
        }
        __syncthreads();
        if constexpr (GROUP == 0) {
            switch (warp) {
            case 0: tma_rows<0, 1, 1, CPL>(a, sh, g, lane, open_ocean, interior); break;
            case 1: tma_rows<0, 2, 2, CPL>(a, sh, g, lane, open_ocean, interior); break;
            default: tma_rows<0, 3, 4, CPL>(a, sh, g, lane, open_ocean, interior); break;
            }
        } else {
            // T | S: the same steps as tma_rows, with the face evaluations of the mixing Jacobian shared between the two warps.  The
            // exchange area aliases the output staging (10 x 32 x 32 B <= 32 x 42 x 8 B): a barrier after the faces are written, one
            // more before the first row is staged.  mixing is block-uniform (kernel argument).
            static_assert(sizeof(MixFace) * MIX_NFACE * TI <= sizeof(double) * TI * G::VS, "the face exchange must fit the output staging");
            const bool mixing = (a.t.mix_temp | a.t.mix_salt) != 0;
            MixFace* const mx = mixing ? reinterpret_cast<MixFace*>(vout) : nullptr;
            const uint32_t nb = st.desc.nbmask[lane];
            const double sm = (double)((st.desc.surfbits >> lane) & 1u);
            double E[RowSlots<5>::N];
            static_assert(RowSlots<5>::N == RowSlots<6>::N, "T and S rows have the same length");
            if (mixing) { mix_faces_to_shared(a, st, g, warp, lane, nb, mx); __syncthreads(); }
            if (warp == 0) { pipe_eval<5, CPL, true>(a, st, g, lane, nb, sm, E, mx); pipe_finish<5>(a, g, lane, nb, open_ocean, E); }
            else { pipe_eval<6, CPL, true>(a, st, g, lane, nb, sm, E, mx); pipe_finish<6>(a, g, lane, nb, open_ocean, E); }
            if (mixing) __syncthreads();
            if (warp == 0) pipe_emit<5, G::ROW0, G::VS>(a, vout, g, lane, interior, E);
            else pipe_emit<6, G::ROW0, G::VS>(a, vout, g, lane, interior, E);
        }
    }
    if (fast) {
        // nothing clipped: the group's segment of every cell record leaves as one TMA bulk store
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
        __syncthreads();
        if (threadIdx.x < g.ncell) {
            bulk_store(a.val + g0 + (size_t)threadIdx.x * NSLOT_TOTAL + SEG0, vout + threadIdx.x * G::VS, G::LEN * (int)sizeof(double));
            asm volatile("cp.async.bulk.commit_group;\n" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;\n" ::: "memory");
        }
    } else {
        // clipped / unaligned tile: per cell, the group's rows are the entries [rowptr[6c + ROW0-1], rowptr[6c + ROW1])
        if (threadIdx.x < g.ncell) {
            sh.cstart[threadIdx.x] = __ldg(a.rowptr + NUN * (g.cell0 + threadIdx.x) + G::ROW0 - 1);
            if (threadIdx.x == 0) sh.cstart[TI + 1] = 0;
        }
        __syncthreads();
        for (int q = threadIdx.x; q < g.ncell * G::LEN; q += NT) {
            const int cl = q / G::LEN, e = q - cl * G::LEN;
            const int lo = sh.cstart[cl], hi = __ldg(a.rowptr + NUN * (g.cell0 + cl) + G::ROW1);
            if (e < hi - lo) a.val[lo + e] = vout[cl * G::VS + e];
        }
    }
}

// =============================================================================
// Matrix-free residual with TMA staging (MODE_RHS default since round 2; the per-position kernel thcm_assemble_kernel<RHS> needed
// ~240 instructions per warp for its staging and ran at 7 % of the HBM roofline).  Same prologue as the Jacobian kernels: the six
// grid lines the rows read arrive as cp.async.bulk runs of raw 48-byte records together with the tile descriptor and the table
// records.  The records STAY raw -- matAvec multiplies the raw unknowns (matetc.F90:160-164) -- and the accessor applies usol's
// no-slip / lid rules (usrc.F90:1014-1121) from the descriptor's bit masks when an atom reads u, v or w.  Five warps = five row
// types (u | v | w + p | T | S), no output staging: every lane owns one cell and writes its rows.  Tiles that are entirely LAND
// are not visited (B = 0 there, usrc.F90:580-591: the output is zeroed first and the blocks walk the list of the other tiles).
// =============================================================================
// The residual runs as the same two row-group kernels as the Jacobian (u | v | w+p and T | S): with all five instruction streams in one
// kernel the top stall was no_instruction (ncu r02m: 6.2 per issue) -- the unrolled row evaluations of several co-resident blocks do not
// fit the instruction cache -- and 70 registers x 160 threads capped the SM at 5 blocks.
template <int GROUP> struct alignas(16) RhsSmem {
    Stage<RowGroup<GROUP>::LINES> st;
    unsigned long long bar;
};
template <class ST> struct RhsTile {
    const ST* st; int lane;
    __device__ __forceinline__ double operator()(int sv, int di, int dj, int dk) const {
        const int r = (dk + 1) * 3 + (dj + 1), x = lane + 1 + di;
        if (sv >= SV_RAW) return st->rec[ST::slot(r)][x][sv - SV_RAW];
        const double v = st->rec[ST::slot(r)][x][sv <= SV_W ? sv : sv + 1];   // u v w . T S
        if (sv <= SV_V) return ((st->desc.uvbits[r] >> x) & 1ull) ? v : 0.0;
        if (sv == SV_W) return ((st->desc.wbits[r] >> x) & 1ull) ? v : 0.0;
        return v;
    }
};
template <int R, bool CPL, class ST>
__device__ __forceinline__ void rhs_row(const AsmArgs& a, const ST& st, const TileGeom& g, int lane, bool open_ocean, bool interior) {
    if (lane >= g.ncell) return;
    const DevBlock& b = a.b;
    const uint32_t nb = st.desc.nbmask[lane];
    const double sm = (double)((st.desc.surfbits >> lane) & 1u);
    const int cell = g.cell0 + lane;
    const Cell c{g.gi0 + lane, g.gj, g.k, cell % b.n0, g.lj};
    const RhsTile<ST> tile{&st, lane};
    const PipeTabs<ST> tabs{&st};
    double E[RowSlots<R>::N];
    if (!((nb >> 4) & 1u)) eval_row<R, false, CPL>(E, a.t, b, c, sm, tile, tabs);
    if (!open_ocean) boundaries<R>(E, nb, c.gi < b.N, c.gj < b.M);
#pragma unroll
    for (int q = 0; q < RowSlots<R>::N; q++) E[q] = fabs(E[q]) > DROP_TOL ? E[q] : 0.0;   // fillcolA's threshold (assemble.F90:115)
    // matAvec (matetc.F90:160-164): v2 = coA*v1(jcoA) + v2 in CRS order = slot order
    double s = 0.0;
    if (interior) {   // tile-uniform: every neighbour of every cell of the tile lies inside the domain
        static_for<0, RowSlots<R>::N>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            constexpr int loc = row_slots(R)[q].loc, col = row_slots(R)[q].col;
            if (E[q] != 0.0) s = E[q] * tile(SV_RAW + col - 1, loc_di(loc), loc_dj(loc), loc_dk(loc)) + s;
        });
    } else {
        static_for<0, RowSlots<R>::N>([&](auto qc) {
            constexpr int q = decltype(qc)::value;
            constexpr int loc = row_slots(R)[q].loc, col = row_slots(R)[q].col;
            const int gi2 = c.gi + loc_di(loc), gj2 = c.gj + loc_dj(loc), k2 = c.k + loc_dk(loc);
            const bool inside = gj2 >= 1 && gj2 <= b.M && k2 >= 1 && k2 <= b.L && (b.periodic || (gi2 >= 1 && gi2 <= b.N));
            if (E[q] != 0.0 && inside) s = E[q] * tile(SV_RAW + col - 1, loc_di(loc), loc_dj(loc), loc_dk(loc)) + s;
        });
    }
    const int row = NUN * cell + R - 1;
    double mixv = 0.0;
    if constexpr (R == TT || R == SS) mixv = vmix_rhs<R>(a.t, c, nb, tile, tabs);   // vmix_fun, usrc.F90:551-571
    double B = -s - mixv + a.frc[row] - 0.0;                                          // usrc.F90:576
    B = B * (((nb >> 4) & 1u) ? 0.0 : 1.0);                                           // usrc.F90:580-591
    a.out[row] = a.sign * B;
}
template <int GROUP, bool CPL>
__global__ void __launch_bounds__(32 * RowGroup<GROUP>::NWARP) thcm_rhs_tma_kernel(const AsmArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    using G = RowGroup<GROUP>;
    RhsSmem<GROUP>& sh = *reinterpret_cast<RhsSmem<GROUP>*>(smem_raw);
    using ST = Stage<G::LINES>;
    ST& st = sh.st;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tile = a.tile_list ? a.tile_list[blockIdx.x] : (int)blockIdx.x;
    const TileGeom g = tile_geom_of(a.b, tile);
    const int w = g.ncell + 2;
    if (threadIdx.x == 0) {
        mbar_init(&sh.bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0)
            mbar_expect_tx(&sh.bar, (uint32_t)(ST::NL * w * NUN * sizeof(double) + sizeof(TileDesc) + sizeof(st.tj) + sizeof(st.tk)));
        __syncwarp();
        if (lane < 9) {
            if ((G::LINES >> lane) & 1) {
                LineSeg seg[3];
                const int ns = line_plan(a.b, g, lane, seg);
#pragma unroll
                for (int q = 0; q < 3; q++)
                    if (q < ns)
                        bulk_load(&st.rec[ST::slot(lane)][seg[q].x0][0], (seg[q].halo ? a.halo : a.un) + (size_t)NUN * seg[q].idx,
                                  (uint32_t)(seg[q].n * NUN * sizeof(double)), &sh.bar);
            }
        } else if (lane == 9) bulk_load(&st.desc, a.tdesc + tile, sizeof(TileDesc), &sh.bar);
        else if (lane == 10) bulk_load(st.tj, a.jrec + (size_t)g.gj * J_COUNT * JREC, sizeof(st.tj), &sh.bar);
        else if (lane == 11) bulk_load(st.tk, a.krec + (size_t)g.k * K_COUNT, sizeof(st.tk), &sh.bar);
        mbar_wait(&sh.bar, 0, 4);   // one warp polls, the others sleep at the barrier
    }
    __syncthreads();
    const bool open_ocean = (st.desc.flags & 2u) != 0;
    const bool interior = g.gi0 > 1 && g.gi0 + g.ncell - 1 < a.b.N && g.gj > 1 && g.gj < a.b.M && g.k > 1 && g.k < a.b.L;
    if constexpr (GROUP == 0) {
        switch (warp) {
        case 0: rhs_row<1, CPL>(a, st, g, lane, open_ocean, interior); break;
        case 1: rhs_row<2, CPL>(a, st, g, lane, open_ocean, interior); break;
        default: rhs_row<3, CPL>(a, st, g, lane, open_ocean, interior); rhs_row<4, CPL>(a, st, g, lane, open_ocean, interior); break;
        }
    } else {
        if (warp == 0) rhs_row<5, CPL>(a, st, g, lane, open_ocean, interior);
        else rhs_row<6, CPL>(a, st, g, lane, open_ocean, interior);
    }
}
template <int GROUP, bool CPL> static void launch_rhs_tma_t(thcmb_ctx* c, const AsmArgs& a, int nblocks) {
    static bool attr_set = false;
    if (!attr_set) {
        THCM_CUDA(cudaFuncSetAttribute(thcm_rhs_tma_kernel<GROUP, CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RhsSmem<GROUP>)));
        attr_set = true;
    }
    if (nblocks > 0) thcm_rhs_tma_kernel<GROUP, CPL><<<nblocks, 32 * RowGroup<GROUP>::NWARP, sizeof(RhsSmem<GROUP>), c->stream>>>(a);
}
static void launch_rhs_tma(thcmb_ctx* c, AsmArgs a) {
    int nblocks = a.ntile;
    if (c->d_active_tiles && c->n_active_tiles < a.ntile) {   // B = 0 on the all-LAND tiles: zero everything once, visit the other tiles
        THCM_CUDA(cudaMemsetAsync(a.out, 0, sizeof(double) * (size_t)NUN * a.b.ncell, c->stream));
        a.tile_list = c->d_active_tiles; nblocks = c->n_active_tiles;
    }
    launch_rhs_tma_t<0, false>(c, a, nblocks);   // coupled mode only touches the T | S rows
    if (a.t.coupled_T || a.t.coupled_S) launch_rhs_tma_t<1, true>(c, a, nblocks);
    else launch_rhs_tma_t<1, false>(c, a, nblocks);
    c->launches++;   // two kernels per residual
}

template <int GROUP, int BLOCKS_PER_SM, bool CPL> static void launch_jac_tma_group(thcmb_ctx* c, const AsmArgs& a, int nblocks) {
    using SH = TmaSmem<GROUP>;
    static bool attr_set = false;
    if (!attr_set) {
        THCM_CUDA(cudaFuncSetAttribute(thcm_jac_tma_kernel<GROUP, BLOCKS_PER_SM, CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(SH)));
        attr_set = true;
    }
    if (nblocks > 0) thcm_jac_tma_kernel<GROUP, BLOCKS_PER_SM, CPL><<<nblocks, 32 * RowGroup<GROUP>::NWARP, sizeof(SH), c->stream>>>(a);
}
// coupled mode only touches the T | S rows (group B); group A is the same kernel either way.
// Tiles whose cells are all LAND hold identity rows whatever the state (boundary.F90:381-386): they are written by the FIRST assembly
// after the static data were built and skipped afterwards (the blocks then walk the list of the other tiles): a third of the tiles of a
// global mask, with their loads, barriers and 832-byte-per-cell stores.
template <int BA, int BB> static void launch_jac_tma(thcmb_ctx* c, AsmArgs a) {
    int nblocks = a.ntile;
    if (c->land_tiles_written && c->d_active_tiles) { a.tile_list = c->d_active_tiles; nblocks = c->n_active_tiles; }
    launch_jac_tma_group<0, BA, false>(c, a, nblocks);
    if (a.t.coupled_T || a.t.coupled_S) launch_jac_tma_group<1, BB, true>(c, a, nblocks);
    else launch_jac_tma_group<1, BB, false>(c, a, nblocks);
    c->land_tiles_written = true;
    c->launches++;   // two kernels per assembly
}

// exclusive scan of the per-block CRS counts (one block; n_blocks <= a few 1e5)
__global__ void scan_counts_kernel(int* cnt, int n) {
    __shared__ int warp_tot[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < n; base += blockDim.x) {
        int i = base + threadIdx.x;
        int v = i < n ? cnt[i] : 0, incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
        if (lane == 31) warp_tot[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int w = warp_tot[lane], wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int u = __shfl_up_sync(0xffffffffu, wi, o); if (lane >= o) wi += u; }
            warp_tot[lane] = wi - w;
        }
        __syncthreads();
        int excl = carry + warp_tot[warp] + incl - v;
        if (i < n) cnt[i] = excl;
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1) carry = excl + v;
        __syncthreads();
    }
}

int asm_block_count(const Block& b) { return ((b.n0 + TI - 1) / TI) * b.m0 * b.L; }

int scan_block_counts(thcmb_ctx* c) {
    ProfScope prof_(c, KID_SCAN);
    scan_counts_kernel<<<1, 1024, 0, c->stream>>>(c->d_blockcnt, c->n_asm_blocks);
    c->launches++;
    return 0;
}

template <int MODE, bool CPL> static void launch_mode_t(thcmb_ctx* c, const AsmArgs& a, int nblk) {
    static bool attr_set = false;
    if (!attr_set) {
        THCM_CUDA(cudaFuncSetAttribute(thcm_assemble_kernel<MODE, CPL>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem<MODE>)));
        attr_set = true;
    }
    thcm_assemble_kernel<MODE, CPL><<<nblk, 32 * mode_warps(MODE), sizeof(Smem<MODE>), c->stream>>>(a);
}
template <int MODE> static void launch_mode(thcmb_ctx* c, const AsmArgs& a, int nblk) {
    if (a.t.coupled_T || a.t.coupled_S) launch_mode_t<MODE, true>(c, a, nblk);
    else launch_mode_t<MODE, false>(c, a, nblk);
}

int launch_assembly(thcmb_ctx* c, int mode, const double* d_un, double* d_out, int* d_begA, int* d_jcoA, double* d_coA) {
    const Block& b = c->blk;
    if ((((uintptr_t)d_un) & 15) != 0) fatal("state vector must be 16-byte aligned (cells are read as 3 x 128-bit loads)");
    AsmArgs a;
    a.b = DevBlock{b.N, b.M, b.L, b.i0, b.j0, b.n0, b.m0, b.periodic, b.wrap_x, b.halo_w, b.halo_e, b.halo_s, b.halo_n, b.hk, b.ncell()};
    a.t = c->tab; a.t.jt = c->d_jt; a.t.kt = c->d_kt; a.t.msi = c->d_msi;
    a.un = d_un; a.halo = c->d_halo; a.nbmask = c->d_nbmask; a.surf = c->d_surf; a.uvlive = c->d_uvlive; a.frc = c->d_frc;
    a.rowptr = c->d_rowptr; a.val = c->d_val; a.blockcnt = c->d_blockcnt; a.begA = d_begA; a.jcoA = d_jcoA; a.coA = d_coA;
    a.out = d_out; a.sign = 1.0;
    a.tdesc = c->d_tdesc; a.jrec = c->d_jrec; a.krec = c->d_krec; a.ntile = c->n_asm_blocks; a.tile_list = nullptr;
    int nblk = c->n_asm_blocks;
    static const int kid_of_mode[4] = {KID_ASM_RHS, KID_ASM_JAC, KID_ASM_COUNT, KID_ASM_CRS};
    ProfScope prof_(c, kid_of_mode[mode & 3]);
    switch (mode & 0xff) {
    case MODE_RHS:
        a.sign = (mode & 0x100) ? -1.0 : 1.0;
        if (c->asm_pipe == 1) launch_rhs_tma(c, a);
        else launch_mode<MODE_RHS>(c, a, nblk);   // THCM_ASM_PIPE=0: per-position loads
        break;
    case MODE_JAC_GRAPH:
        if (c->asm_pipe == 1) launch_jac_tma<8, 11>(c, a);
        else { launch_mode<MODE_JAC_GRAPH>(c, a, nblk); c->land_tiles_written = true; }   // THCM_ASM_PIPE=0: per-position loads, every tile
        break;
    case MODE_JAC_COUNT: launch_mode<MODE_JAC_COUNT>(c, a, nblk); break;
    case MODE_JAC_CRS: launch_mode<MODE_JAC_CRS>(c, a, nblk); break;
    default: return -1;
    }
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) fatal(std::string("assembly kernel launch failed: ") + cudaGetErrorString(e));
    return 0;
}

}  // namespace thcm
