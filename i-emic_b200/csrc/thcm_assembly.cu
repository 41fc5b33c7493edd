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
// Mapping: block = 32 consecutive owned cells x 6 warps; warp r evaluates row type r (u,v,w,p,T,S)
// for the 32 cells (no intra-warp divergence on the row type), results are staged in shared memory
// at their final offsets and written out by the whole block as one contiguous, coalesced range.
// HBM traffic per cell: 48 B state + 5 B masks in, 832 B values out.
// =============================================================================
#include <cstdio>
#include "thcm_cell.cuh"

namespace thcm {

__constant__ ClassTables c_cls;

void upload_class_tables(const ClassTables& t) { THCM_CUDA(cudaMemcpyToSymbol(c_cls, &t, sizeof(ClassTables))); }

// ---------------------------------------------------------------------------
// the kernel
// ---------------------------------------------------------------------------
template <int MODE> struct Smem;
template <> struct Smem<MODE_RHS> { double dummy; };
template <> struct Smem<MODE_JAC_GRAPH> { double v[CELLS_PER_BLOCK * NSLOT_TOTAL]; };
template <> struct Smem<MODE_JAC_COUNT> { double v[CELLS_PER_BLOCK * NSLOT_TOTAL]; int cnt[NUN]; };
template <> struct Smem<MODE_JAC_CRS> { double v[CELLS_PER_BLOCK * NSLOT_TOTAL]; int c[CELLS_PER_BLOCK * NSLOT_TOTAL]; int off[CELLS_PER_BLOCK * NUN + 1]; };

template <int R, int MODE>
__device__ __forceinline__ void do_row(const AsmArgs& a, Smem<MODE>& sh, int cell0, int ncell_blk, int lane) {
    const DevBlock& b = a.b;
    const int cell = cell0 + lane;
    const bool active = lane < ncell_blk;
    double E[RowSlots<R>::N];
    Cell c{0, 0, 0, 0, 0};
    int cls = 0;
    uint32_t nb = 0;
    if (active) {
        int li = cell % b.n0, r = cell / b.n0, lj = r % b.m0, k0 = r / b.m0;
        c.li = li; c.lj = lj; c.gi = b.i0 + li + 1; c.gj = b.j0 + lj + 1; c.k = k0 + 1;
        nb = a.nbmask[cell];
        double sm = (double)(int)(int8_t)a.surf[lj * b.n0 + li];
        cls = (c.gi == 1 ? 1 : 0) | (c.gi == b.N ? 2 : 0) | (c.gj == 1 ? 4 : 0) | (c.gj == b.M ? 8 : 0) | (c.k == 1 ? 16 : 0) |
              (c.k == b.L ? 32 : 0);
        if (!((nb >> 4) & 1u)) eval_row<R, MODE != MODE_RHS>(E, a, c, sm);
        boundaries<R>(E, nb, c.gi < b.N, c.gj < b.M);
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
                if (E[q] != 0.0 && inside) s = E[q] * raw(a, gi2, gj2, k2, col - 1) + s;
            });
            // B = -Au - mix + Frc - p0*(1-par(RESC))*ures ; B *= (1 - landm) (usrc.F90:576-591)
            int row = NUN * cell + R - 1;
            double B = -s - 0.0 + a.frc[row] - 0.0;
            B = B * (((nb >> 4) & 1u) ? 0.0 : 1.0);
            a.out[row] = a.sign * B;
        }
        return;
    } else {
        const int g0 = a.rowptr[NUN * cell0];
        if constexpr (MODE == MODE_JAC_GRAPH || MODE == MODE_JAC_COUNT) {
            int cnt = 0;
            if (active) {
                const int base = a.rowptr[NUN * cell + R - 1] - g0;
                static_for<0, RowSlots<R>::N>([&](auto qc) {
                    constexpr int q = decltype(qc)::value;
                    int p = c_cls.pos[cls][ROW_OFF[R - 1] + q];
                    if (p >= 0) sh.v[base + p] = E[q];
                    if (E[q] != 0.0) cnt++;
                });
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
                if (lane == 31) sh.off[CELLS_PER_BLOCK * NUN] = incl;
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

template <int MODE>
__global__ void __launch_bounds__(ASM_THREADS) thcm_assemble_kernel(const AsmArgs a) {
    __shared__ Smem<MODE> sh;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cell0 = blockIdx.x * CELLS_PER_BLOCK;
    const int ncell_blk = min(CELLS_PER_BLOCK, a.b.ncell - cell0);
    switch (warp) {
    case 0: do_row<1, MODE>(a, sh, cell0, ncell_blk, lane); break;
    case 1: do_row<2, MODE>(a, sh, cell0, ncell_blk, lane); break;
    case 2: do_row<3, MODE>(a, sh, cell0, ncell_blk, lane); break;
    case 3: do_row<4, MODE>(a, sh, cell0, ncell_blk, lane); break;
    case 4: do_row<5, MODE>(a, sh, cell0, ncell_blk, lane); break;
    default: do_row<6, MODE>(a, sh, cell0, ncell_blk, lane); break;
    }
    if constexpr (MODE == MODE_JAC_GRAPH || MODE == MODE_JAC_COUNT) {
        __syncthreads();
        const int g0 = a.rowptr[NUN * cell0], g1 = a.rowptr[NUN * (cell0 + ncell_blk)];
        for (int q = threadIdx.x; q < g1 - g0; q += ASM_THREADS) a.val[g0 + q] = sh.v[q];
        if constexpr (MODE == MODE_JAC_COUNT) {
            if (threadIdx.x == 0) a.blockcnt[blockIdx.x] = sh.cnt[0] + sh.cnt[1] + sh.cnt[2] + sh.cnt[3] + sh.cnt[4] + sh.cnt[5];
        }
    } else if constexpr (MODE == MODE_JAC_CRS) {
        __syncthreads();
        const int bbase = a.blockcnt[blockIdx.x], tot = sh.off[CELLS_PER_BLOCK * NUN];
        for (int q = threadIdx.x; q < tot; q += ASM_THREADS) { a.coA[bbase + q] = sh.v[q]; a.jcoA[bbase + q] = sh.c[q]; }
        if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) a.begA[NUN * a.b.ncell] = bbase + tot + 1;
    }
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

int scan_block_counts(thcmb_ctx* c) {
    ProfScope prof_(c, KID_SCAN);
    scan_counts_kernel<<<1, 1024, 0, c->stream>>>(c->d_blockcnt, c->n_asm_blocks);
    c->launches++;
    return 0;
}

int launch_assembly(thcmb_ctx* c, int mode, const double* d_un, double* d_out, int* d_begA, int* d_jcoA, double* d_coA) {
    const Block& b = c->blk;
    AsmArgs a;
    a.b = DevBlock{b.N, b.M, b.L, b.i0, b.j0, b.n0, b.m0, b.periodic, b.wrap_x, b.halo_w, b.halo_e, b.halo_s, b.halo_n, b.hk, b.ncell()};
    a.t = c->tab; a.t.jt = c->d_jt; a.t.kt = c->d_kt;
    a.un = d_un; a.halo = c->d_halo; a.nbmask = c->d_nbmask; a.surf = c->d_surf; a.uvlive = c->d_uvlive; a.frc = c->d_frc;
    a.rowptr = c->d_rowptr; a.val = c->d_val; a.blockcnt = c->d_blockcnt; a.begA = d_begA; a.jcoA = d_jcoA; a.coA = d_coA;
    a.out = d_out; a.sign = 1.0;
    int nblk = c->n_asm_blocks;
    static const int kid_of_mode[4] = {KID_ASM_RHS, KID_ASM_JAC, KID_ASM_COUNT, KID_ASM_CRS};
    ProfScope prof_(c, kid_of_mode[mode & 3]);
    switch (mode & 0xff) {
    case MODE_RHS:
        a.sign = (mode & 0x100) ? -1.0 : 1.0;
        thcm_assemble_kernel<MODE_RHS><<<nblk, ASM_THREADS, 0, c->stream>>>(a); break;
    case MODE_JAC_GRAPH: thcm_assemble_kernel<MODE_JAC_GRAPH><<<nblk, ASM_THREADS, 0, c->stream>>>(a); break;
    case MODE_JAC_COUNT: thcm_assemble_kernel<MODE_JAC_COUNT><<<nblk, ASM_THREADS, 0, c->stream>>>(a); break;
    case MODE_JAC_CRS: thcm_assemble_kernel<MODE_JAC_CRS><<<nblk, ASM_THREADS, 0, c->stream>>>(a); break;
    default: return -1;
    }
    c->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) fatal(std::string("assembly kernel launch failed: ") + cudaGetErrorString(e));
    return 0;
}

}  // namespace thcm
