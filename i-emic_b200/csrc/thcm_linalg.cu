// =============================================================================
// thcm_linalg.cu -- FP64 CSR SpMV, Krylov vector kernels, halo pack/unpack, 6x6 block-diagonal
// preconditioner.  Replaces Epetra_CrsMatrix::Apply (Ocean.C:1369-1374), matAvec (matetc.F90:147-166)
// and Epetra_MultiVector::Dot/Norm2/Update/Scale as used by GMRESSolver.H:177-187 / IDRSolver.H.
// All kernels are HBM-bandwidth bound; none is a dense contraction (no tensor cores).
// =============================================================================
#include <cstdio>
#include <dlfcn.h>
#include "thcm_internal.h"

namespace thcm {

constexpr int NSM = 148;  // B200: 148 SMs; persistent-style grids are sized in multiples of this

// ---------------------------------------------------------------------------
// SpMV: CSR-vector with 8 lanes per row (rows hold 7..24 entries, THCM.C:2320-2325).
// Column ids >= nlocal address the halo buffer.
// ---------------------------------------------------------------------------
constexpr int SPMV_THREADS = 256;

// LANES lanes share a row; every lane first issues its first UNROLL (col, val) loads -- predicated, independent --
// then the x gathers, then the products: all loads of a row are in flight together (rows are short, so a plain
// loop exposes one memory latency per iteration and leaves the kernel latency-bound at full occupancy).
// 4 lanes x 6 entries measured best on the B200 (r01: 8x1, 8x3, 4x3, 16x2, 2x12 were the alternatives).
// LSKIP: rows of LAND cells are identity rows whatever the state (boundary.F90:381-386; explicit zeros elsewhere in the maximal graph):
// y = x for them without touching their values or column ids (landcell = one byte per owned cell).  Bit-identical to the full product
// (r02a on the 1-degree grid: 0.265 ms instead of 0.352 ms).
template <int LANES, int UNROLL, bool LSKIP>
__global__ void __launch_bounds__(SPMV_THREADS) spmv_csr_kernel(int nrow, const int* __restrict__ rp, const int* __restrict__ col,
                                                                 const double* __restrict__ val, const double* __restrict__ x,
                                                                 const double* __restrict__ halo, int nlocal, double* __restrict__ y,
                                                                 const unsigned char* __restrict__ landcell) {
    const int sub = threadIdx.x & (LANES - 1);
    constexpr int rows_per_block = SPMV_THREADS / LANES;
    // the loop bound is warp-uniform and rows that do not take part stay in the body with an empty range: the full-mask
    // shuffles below must be reached by every lane of the warp
    for (int base = blockIdx.x * rows_per_block; base < nrow; base += gridDim.x * rows_per_block) {
        const int r0 = base + (threadIdx.x / LANES);
        const bool act = r0 < nrow;
        const int row = act ? r0 : 0;
        bool land = false;
        if constexpr (LSKIP) land = act && __ldg(landcell + row / NUN) != 0;
        const bool ld = act && !land;
        const int b = ld ? __ldg(rp + row) : 0, e = ld ? __ldg(rp + row + 1) : 0;
        int cc[UNROLL]; double vv[UNROLL], xx[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int q = b + sub + u * LANES;
            const bool ok = q < e;
            cc[u] = ok ? __ldg(col + q) : -1;
            vv[u] = ok ? __ldg(val + q) : 0.0;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; u++)
            xx[u] = cc[u] < 0 ? 0.0 : (cc[u] < nlocal ? __ldg(x + cc[u]) : __ldg(halo + (cc[u] - nlocal)));
        double s = 0.0;
#pragma unroll
        for (int u = 0; u < UNROLL; u++) s += vv[u] * xx[u];
        for (int q = b + sub + UNROLL * LANES; q < e; q += LANES) {   // rows longer than UNROLL*LANES (generic CSR)
            const int cidx = __ldg(col + q);
            const double xv = cidx < nlocal ? __ldg(x + cidx) : __ldg(halo + (cidx - nlocal));
            s += __ldg(val + q) * xv;
        }
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, LANES);
        if constexpr (LSKIP) { if (sub == 0 && land) s = __ldg(x + row); }
        if (sub == 0 && act) y[row] = s;
    }
}

int spmv(thcmb_ctx* c, int nrow, const int* rp, const int* col, const double* val, const double* x, const double* halo,
         int nlocal, double* y) {
    ProfScope prof_(c, KID_SPMV);
    const int rows_per_block = SPMV_THREADS / 4;
    const int grid = (int)std::max<long long>(1, std::min<long long>(((long long)nrow + rows_per_block - 1) / rows_per_block, (long long)NSM * 64));
    if (col == c->d_col && c->d_landcell)   // the context's own graph: identity rows of LAND cells are not streamed
        spmv_csr_kernel<4, 6, true><<<grid, SPMV_THREADS, 0, c->stream>>>(nrow, rp, col, val, x, halo, nlocal, y, c->d_landcell);
    else
        spmv_csr_kernel<4, 6, false><<<grid, SPMV_THREADS, 0, c->stream>>>(nrow, rp, col, val, x, halo, nlocal, y, nullptr);
    c->launches++;
    return 0;
}

// ---------------------------------------------------------------------------
// reductions: per-block partials (warp shuffle -> shared) -> the last block to finish sums the
// partials in a fixed order (deterministic) and writes the device scalar.  No host sync.
// ---------------------------------------------------------------------------
constexpr int RED_THREADS = 256;
constexpr int RED_BLOCKS = NSM * 8;

__device__ __forceinline__ double block_sum(double v) {
    __shared__ double wsum[RED_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x < 32) {
        s = threadIdx.x < RED_THREADS / 32 ? wsum[threadIdx.x] : 0.0;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    }
    __syncthreads();
    return s;  // valid in thread 0
}

// ---------------------------------------------------------------------------
// Fused reduction + cross-GPU all-reduce over NVLink peer memory.
// Every rank owns a small mailbox (cudaMalloc + CUDA IPC, mapped by all peers).  The LAST block of a reduction kernel
// -- the one that sums the per-block partials -- stores the rank's sum straight into every peer's mailbox (P2P store over
// NVLink), then polls its own mailbox for the peers' sums and adds them in rank order (bitwise identical on all ranks).
// One kernel does the local reduction and the collective: no NCCL launch (~10 us) per Krylov dot product.
// Slots are double-buffered by the parity of a sequence number that all ranks advance in lockstep (SPMD).
// ---------------------------------------------------------------------------
// Wire format = NCCL's LL protocol: a double travels as ONE 16-byte store {lo32, flag, hi32, flag}; each 8-byte half
// carries its own copy of the flag (the low 32 bits of the sequence number), so the receiver needs no fence and no
// second round trip: it polls the 16 bytes until both flags match (only 8-byte store atomicity is assumed).
struct alignas(16) P2PSlot { unsigned int lo, f0, hi, f1; };
struct P2PArgs {
    int nranks, rank;
    P2PSlot* mine;            // scalar slots [2][MAX_RANKS], then the vector slots (multi_dot)
    P2PSlot* const* peers;    // device array [nranks] of peer mailboxes (entry `rank` unused)
    unsigned long long* seq;  // DEVICE counter of completed exchanges: advanced by the kernel that really exchanges, so a
                              // pass that every rank skips on the device (DGKS) keeps the slot parity alternating
};
constexpr int P2P_MAX_RANKS = 16;
constexpr int P2P_VEC_LEN = 64 + 1;       // MD_MAXV projections + w.w
constexpr size_t P2P_VEC_OFFSET = 1024;   // byte offset of the vector slots inside the mailbox allocation
constexpr size_t P2P_MAILBOX_BYTES = P2P_VEC_OFFSET + sizeof(P2PSlot) * 2 * P2P_MAX_RANKS * P2P_VEC_LEN;
static_assert(sizeof(P2PSlot) * 2 * P2P_MAX_RANKS <= P2P_VEC_OFFSET, "mailbox layout");
// the IPC-shared allocation continues with the halo flags [2][MAX_RANKS] (u64 sequence numbers) and the two halo buffers
constexpr size_t P2P_HALOFLAG_OFFSET = (P2P_MAILBOX_BYTES + 255) / 256 * 256;
constexpr size_t P2P_HALO_OFFSET = P2P_HALOFLAG_OFFSET + 256;
static_assert(2 * P2P_MAX_RANKS * sizeof(unsigned long long) <= 256, "halo flag area");
static inline size_t p2p_halo_bytes(const thcmb_ctx* c) { return ((size_t)NUN * std::max(c->blk.nhalo_cells(), 1) * sizeof(double) + 255) / 256 * 256; }
// ... and, behind the two plain halo buffers, two halo buffers in LL format (one 16-byte slot per double: data + flags) for the halo
// exchange that is fused into the compact SpMV -- the reader polls the slot it needs, no fence, no flag round trip, no wait kernel
static inline size_t p2p_ll_bytes(const thcmb_ctx* c) { return 2 * p2p_halo_bytes(c); }
static inline size_t p2p_ll_offset(const thcmb_ctx* c, int b) { return P2P_HALO_OFFSET + 2 * p2p_halo_bytes(c) + (size_t)b * p2p_ll_bytes(c); }
static inline size_t p2p_total_bytes(const thcmb_ctx* c) { return p2p_ll_offset(c, 2); }

__device__ __forceinline__ void ll_store(P2PSlot* dst, double v, unsigned int flag) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};\n" ::"l"(dst), "r"((unsigned int)b), "r"(flag),
                 "r"((unsigned int)(b >> 32)), "r"(flag) : "memory");
}
// A peer that never arrives (a rank that died or left the SPMD sequence) must not hang the node silently: every poll loop is bounded
// (~2^27 polls of >= 0.5 us: more than a minute, against exchanges that take microseconds), then reports and traps -- the host sees a
// CUDA error at its next call and fails through thcm_throw_error_.
constexpr unsigned int P2P_SPIN_LIMIT = 1u << 27;
__device__ __noinline__ void p2p_timeout(const void* slot, unsigned int want, unsigned int got) {
    printf("thcm_b200: timed out waiting for a peer GPU (slot %p, expected flag %u, last seen %u): a rank left the SPMD sequence\n", slot, want, got);
    __trap();
}
__device__ __forceinline__ double ll_wait(const P2PSlot* src, unsigned int flag) {
    unsigned int lo, f0, hi, f1, spins = 0;
    do {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(src) : "memory");
        if (++spins > P2P_SPIN_LIMIT) p2p_timeout(src, flag, f0);
    } while (f0 != flag || f1 != flag);
    return __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}
// the same poll without a call inside the loop (a call site keeps every live value of the caller in callee-saved registers or on the
// stack: the compact SpMV needed 58 instead of 40 registers for its six unrolled polls): returns false after the spin limit, the caller
// reports once, at a point where nothing is live
__device__ __forceinline__ bool ll_poll(const P2PSlot* src, unsigned int flag, double& out) {
    unsigned int lo, f0, hi, f1, spins = 0;
    do {
        asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];\n" : "=r"(lo), "=r"(f0), "=r"(hi), "=r"(f1) : "l"(src) : "memory");
        if (++spins > P2P_SPIN_LIMIT) return false;
    } while (f0 != flag || f1 != flag);
    out = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
    return true;
}
__device__ __forceinline__ void flag_wait(volatile unsigned long long* f, unsigned long long seq) {
    unsigned int spins = 0;
    while (*f != seq) if (++spins > P2P_SPIN_LIMIT) p2p_timeout((const void*)f, (unsigned int)seq, (unsigned int)*f);
}

// optional epilogue of a reduction (DGKS, thcmb_gmres): flag = (result < 0.5 * *ww_old), *final_out = result
// flag2_out (batched tails only, multi_tail): the norm of the vector AFTER the second update is taken from Pythagoras,
//   ||w' - V h2||^2 = ||w'||^2 - ||h2||^2   (V orthonormal, h2 = V^T w'),
// both terms being in hand (all-reduced) right here -- the second update then needs no reduction of its own and moves into the
// head kernel of the next Arnoldi step.  Guard: when the difference cancels below 1 % of ||w'||^2 the explicit update + norm
// kernel runs instead (*flag2_out = 1).
struct RedEpilogue { const double* ww_old; int* flag_out; double* final_out; int* flag2_out; };
__device__ __forceinline__ void red_epilogue(double tot, const RedEpilogue ep) {
    if (ep.flag_out) *ep.flag_out = (tot < 0.5 * (*ep.ww_old)) ? 1 : 0;
    if (ep.final_out) *ep.final_out = tot;
}
__device__ __forceinline__ void finish_reduction(double blocksum, double* partial, unsigned int* counter, double* out, const P2PArgs pa,
                                                 const RedEpilogue ep = RedEpilogue{nullptr, nullptr, nullptr, nullptr}) {
    __shared__ bool last;
    __shared__ double peer_val[P2P_MAX_RANKS];
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = blocksum;
        __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (last) {
        double v = 0.0;
        for (int i = threadIdx.x; i < (int)gridDim.x; i += RED_THREADS) v += ((volatile double*)partial)[i];
        double s = block_sum(v);
        if (pa.nranks <= 1) {
            if (threadIdx.x == 0) { *out = s; *counter = 0u; red_epilogue(s, ep); }
            return;
        }
        __shared__ double mysum;
        const unsigned long long seq = *pa.seq + 1ull;
        if (threadIdx.x == 0) { mysum = s; *counter = 0u; }
        __syncthreads();
        const int par = (int)(seq & 1ull);
        const unsigned int flag = (unsigned int)seq;
        if (threadIdx.x < pa.nranks) {
            const int r = threadIdx.x;
            if (r == pa.rank) peer_val[r] = mysum;
            else {
                ll_store(pa.peers[r] + par * P2P_MAX_RANKS + pa.rank, mysum, flag);          // push into peer r's mailbox
                peer_val[r] = ll_wait(pa.mine + par * P2P_MAX_RANKS + r, flag);              // pull peer r's sum from mine
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double tot = 0.0;
            for (int r = 0; r < pa.nranks; r++) tot += peer_val[r];   // fixed rank order: identical bits on every rank
            *out = tot;
            *pa.seq = seq;
            red_epilogue(tot, ep);
        }
    }
}

__global__ void __launch_bounds__(RED_THREADS) dot_kernel(int n, const double* __restrict__ x, const double* __restrict__ y,
                                                           double* partial, unsigned int* counter, double* out, const P2PArgs pa) {
    double v = 0.0;
    if ((((uintptr_t)x | (uintptr_t)y) & 15) == 0) {   // 16-byte aligned: 128-bit loads
        const int n2 = n >> 1;
        const double2* x2 = reinterpret_cast<const double2*>(x);
        const double2* y2 = reinterpret_cast<const double2*>(y);
        for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < n2; i += gridDim.x * RED_THREADS) {
            double2 a = x2[i], b = y2[i];
            v += a.x * b.x; v += a.y * b.y;
        }
        if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) v += x[n - 1] * y[n - 1];
    } else {
        for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += gridDim.x * RED_THREADS) v += x[i] * y[i];
    }
    double s = block_sum(v);
    finish_reduction(s, partial, counter, out, pa);
}

// fused modified Gram-Schmidt step (GMRESSolver.H:177-181): w -= h_k * v_k ; h_next = w . v_next
// (h_k read from device memory: the previous reduction result; bitwise the same operations as dot + update)
__global__ void __launch_bounds__(RED_THREADS) mgs_step_kernel(int n, const double* __restrict__ hk, const double* __restrict__ vk,
                                                                const double* __restrict__ vnext, double* __restrict__ w,
                                                                double* partial, unsigned int* counter, double* out, const P2PArgs pa) {
    const double h = *hk;
    double v = 0.0;
    if (((((uintptr_t)vk) | ((uintptr_t)vnext) | ((uintptr_t)w)) & 15) == 0) {   // 128-bit path
        const int n2 = n >> 1;
        const double2* vk2 = reinterpret_cast<const double2*>(vk);
        const double2* vn2 = reinterpret_cast<const double2*>(vnext);
        double2* w2 = reinterpret_cast<double2*>(w);
        for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < n2; i += gridDim.x * RED_THREADS) {
            double2 a = vk2[i], ww = w2[i], b = vn2[i];
            ww.x = -h * a.x + 1.0 * ww.x;   // update(-H, V[k], 1.0): this = a*A + b*this
            ww.y = -h * a.y + 1.0 * ww.y;
            w2[i] = ww;
            v += ww.x * b.x; v += ww.y * b.y;
        }
        if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
            double wi = -h * vk[n - 1] + 1.0 * w[n - 1];
            w[n - 1] = wi; v += wi * vnext[n - 1];
        }
    } else {
        for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < n; i += gridDim.x * RED_THREADS) {
            double wi = -h * vk[i] + 1.0 * w[i];
            w[i] = wi;
            v += wi * vnext[i];
        }
    }
    double s = block_sum(v);
    finish_reduction(s, partial, counter, out, pa);
}

// y = a x + b y with the BLAS / Epetra_MultiVector::Update convention: an operand whose scalar is zero is NOT read (b = 0 makes this a
// scaled copy into possibly uninitialised memory; 0 * NaN would poison it).  x and y may alias (scale in place), hence no __restrict__.
__global__ void axpby_kernel(int n, double a, const double* x, double b, double* y) {
    if (b == 0.0)      for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = a * x[i];
    else if (a == 0.0) for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = b * y[i];
    else               for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = a * x[i] + b * y[i];
}
__global__ void axpy_negdev_kernel(int n, const double* __restrict__ h, const double* __restrict__ x, double* __restrict__ y) {
    const double a = -(*h);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = a * x[i] + 1.0 * y[i];
}
// x *= 1/sqrt(nrm2) ; nrm = sqrt(nrm2)   (H[i+1][i] = w.norm(); w.scale(1.0 / H[i+1][i]), GMRESSolver.H:185-186)
__global__ void scale_invsqrt_kernel(int n, const double* __restrict__ nrm2, double* __restrict__ x, double* nrm_out) {
    const double nrm = sqrt(*nrm2);
    const double a = 1.0 / nrm;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = a * x[i];
    if (blockIdx.x == 0 && threadIdx.x == 0 && nrm_out) *nrm_out = nrm;
}
__global__ void copy_kernel(int n, const double* __restrict__ x, double* __restrict__ y) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = x[i];
}
__global__ void fill_kernel(int n, double a, double* __restrict__ x) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) x[i] = a;
}

static inline int ew_grid(int n) {
    long long want = ((long long)n + 255) / 256;
    return (int)std::max<long long>(1, std::min<long long>(want, (long long)NSM * 16));
}

static P2PArgs p2p_args(thcmb_ctx* c) {
    P2PArgs pa{1, 0, nullptr, nullptr, nullptr};
    if (c->p2p_on) {
        pa.nranks = c->blk.nranks; pa.rank = c->blk.rank;
        pa.mine = (P2PSlot*)c->d_mailbox; pa.peers = (P2PSlot* const*)c->d_peer_mailboxes;
        pa.seq = c->d_p2p_seq;
    }
    return pa;
}

// CUDA IPC plumbing of the mailboxes (one cudaMalloc per rank, mapped by every peer of the same node)
int p2p_local_handle(thcmb_ctx* c, void* handle64) {
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    if (c->blk.nranks > P2P_MAX_RANKS) return -1;
    if (!c->d_mailbox) {
        const size_t bytes = p2p_total_bytes(c);
        THCM_CUDA(cudaMalloc(&c->d_mailbox, bytes));
        THCM_CUDA(cudaMemset(c->d_mailbox, 0, bytes));
        THCM_CUDA(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, c->d_mailbox);
    if (e != cudaSuccess) { set_error(std::string("cudaIpcGetMemHandle: ") + cudaGetErrorString(e)); return -2; }
    memcpy(handle64, &h, 64);
    return 0;
}
int p2p_open(thcmb_ctx* c, const void* handles_all) {
    std::vector<void*> ptrs(c->blk.nranks, nullptr);
    for (int r = 0; r < c->blk.nranks; r++) {
        if (r == c->blk.rank) { ptrs[r] = c->d_mailbox; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, (const char*)handles_all + 64 * (size_t)r, 64);
        cudaError_t e = cudaIpcOpenMemHandle(&ptrs[r], h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) { set_error(std::string("cudaIpcOpenMemHandle: ") + cudaGetErrorString(e)); cudaGetLastError(); return -1; }
    }
    THCM_CUDA(cudaMalloc(&c->d_peer_mailboxes, sizeof(void*) * ptrs.size()));
    THCM_CUDA(cudaMemcpy(c->d_peer_mailboxes, ptrs.data(), sizeof(void*) * ptrs.size(), cudaMemcpyHostToDevice));
    c->p2p_peer_ptrs = ptrs;
    THCM_CUDA(cudaMalloc(&c->d_p2p_seq, 2 * sizeof(unsigned long long)));
    THCM_CUDA(cudaMemset(c->d_p2p_seq, 0, 2 * sizeof(unsigned long long)));
    c->p2p_on = true;
    // halo push: my two halo buffers inside the shared allocation, and for every neighbour the address of ITS buffers
    // (their size is the neighbour's own: read it from the neighbour's block geometry)
    const char* env = getenv("THCM_HALO_P2P");
    if (!env || atoi(env) != 0) {
        THCM_CUDA(cudaStreamSynchronize(c->stream));
        for (int b = 0; b < 2; b++) c->d_halo_p2p[b] = (double*)((char*)c->d_mailbox + P2P_HALO_OFFSET + b * p2p_halo_bytes(c));
        if (c->d_halo) cudaFree(c->d_halo);
        c->d_halo = c->d_halo_p2p[0];
        std::vector<double*> ph(2 * std::max<size_t>(c->peers.size(), 1), nullptr);
        std::vector<void*> pl(2 * std::max<size_t>(c->peers.size(), 1), nullptr);
        for (size_t q = 0; q < c->peers.size(); q++) {
            thcmb_ctx tmp; tmp.blk = Block();
            if (!decomp2d(c->blk.nranks, c->peers[q].rank, c->blk.N, c->blk.M, c->blk.L, c->blk.periodic, tmp.blk, c->blk.cuts)) fatal("halo push: bad peer block");
            const size_t pb = p2p_halo_bytes(&tmp);
            for (int b = 0; b < 2; b++) {
                ph[b * c->peers.size() + q] = (double*)((char*)ptrs[c->peers[q].rank] + P2P_HALO_OFFSET + b * pb);
                pl[b * c->peers.size() + q] = (void*)((char*)ptrs[c->peers[q].rank] + p2p_ll_offset(&tmp, b));
            }
        }
        THCM_CUDA(cudaMalloc(&c->d_peer_halo, sizeof(double*) * ph.size()));
        THCM_CUDA(cudaMemcpy(c->d_peer_halo, ph.data(), sizeof(double*) * ph.size(), cudaMemcpyHostToDevice));
        THCM_CUDA(cudaMalloc(&c->d_peer_ll, sizeof(void*) * pl.size()));
        THCM_CUDA(cudaMemcpy(c->d_peer_ll, pl.data(), sizeof(void*) * pl.size(), cudaMemcpyHostToDevice));
        for (int b = 0; b < 2; b++) c->d_halo_ll[b] = (char*)c->d_mailbox + p2p_ll_offset(c, b);
        c->halo_ll_seq = 0;
        THCM_CUDA(cudaMalloc(&c->d_halo_counter, sizeof(unsigned int)));
        THCM_CUDA(cudaMemset(c->d_halo_counter, 0, sizeof(unsigned int)));
        c->halo_seq = 0;
        c->halo_p2p = true;
    }
    return 0;
}
void p2p_close(thcmb_ctx* c) {
    for (int r = 0; r < (int)c->p2p_peer_ptrs.size(); r++)
        if (r != c->blk.rank && c->p2p_peer_ptrs[r]) cudaIpcCloseMemHandle(c->p2p_peer_ptrs[r]);
    c->p2p_peer_ptrs.clear();
    if (c->d_peer_mailboxes) cudaFree(c->d_peer_mailboxes);
    if (c->d_peer_halo) cudaFree(c->d_peer_halo);
    if (c->d_peer_ll) cudaFree(c->d_peer_ll);
    c->d_peer_ll = nullptr; c->d_halo_ll[0] = c->d_halo_ll[1] = nullptr;
    if (c->d_p2p_seq) cudaFree(c->d_p2p_seq);
    c->d_p2p_seq = nullptr;
    if (c->halo_p2p) c->d_halo = nullptr;   // lived inside the mailbox allocation
    if (c->d_mailbox) cudaFree(c->d_mailbox);
    c->d_peer_mailboxes = nullptr; c->d_mailbox = nullptr; c->d_peer_halo = nullptr; c->p2p_on = false; c->halo_p2p = false;
}

int dot_dev(thcmb_ctx* c, int n, const double* x, const double* y, double* d_out) {
    { ProfScope prof_(c, KID_DOT);
    dot_kernel<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(n, x, y, c->d_partial, c->d_counter, d_out, p2p_args(c)); }
    c->launches++;
    return c->p2p_on ? 0 : allreduce_dev(c, d_out, 1);
}
int mgs_step_dev(thcmb_ctx* c, int n, const double* d_hk, const double* vk, const double* vnext, double* w, double* d_out) {
    { ProfScope prof_(c, KID_MGS);
    mgs_step_kernel<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(n, d_hk, vk, vnext, w, c->d_partial, c->d_counter, d_out, p2p_args(c)); }
    c->launches++;
    return c->p2p_on ? 0 : allreduce_dev(c, d_out, 1);
}
int axpby(thcmb_ctx* c, int n, double a, const double* x, double b, double* y) {
    ProfScope prof_(c, KID_AXPBY);
    axpby_kernel<<<ew_grid(n), 256, 0, c->stream>>>(n, a, x, b, y); c->launches++; return 0;
}
int axpy_negdev(thcmb_ctx* c, int n, const double* d_h, const double* x, double* y) {
    ProfScope prof_(c, KID_AXPY_DEV);
    axpy_negdev_kernel<<<ew_grid(n), 256, 0, c->stream>>>(n, d_h, x, y); c->launches++; return 0;
}
int scale_invsqrt_dev(thcmb_ctx* c, int n, const double* d_nrm2, double* x, double* d_nrm) {
    ProfScope prof_(c, KID_SCALE);
    scale_invsqrt_kernel<<<ew_grid(n), 256, 0, c->stream>>>(n, d_nrm2, x, d_nrm); c->launches++; return 0;
}
int copy(thcmb_ctx* c, int n, const double* x, double* y) {
    ProfScope prof_(c, KID_COPY);
    copy_kernel<<<ew_grid(n), 256, 0, c->stream>>>(n, x, y); c->launches++; return 0;
}
int fill(thcmb_ctx* c, int n, double a, double* x) {
    ProfScope prof_(c, KID_FILL);
    fill_kernel<<<ew_grid(n), 256, 0, c->stream>>>(n, a, x); c->launches++; return 0;
}
// out[slot[e]] = in[e]: the stored Jacobian values (graph order) into the value layout of a caller's matrix object (Epetra bridge,
// include/thcm_epetra_bridge.hpp); slot is a permutation, so every position of `out` is written exactly once
__global__ void scatter_slots_kernel(long long n, const int* __restrict__ slot, const double* __restrict__ in, double* __restrict__ out) {
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) out[slot[e]] = in[e];
}
int scatter_slots(thcmb_ctx* c, long long n, const int* d_slot, const double* d_in, double* d_out) {
    ProfScope prof_(c, KID_COPY);
    const int grid = (int)std::min<long long>((n + 255) / 256, (long long)NSM * 32);
    if (n > 0) scatter_slots_kernel<<<grid, 256, 0, c->stream>>>(n, d_slot, d_in, d_out);
    c->launches++; return 0;
}

// ---------------------------------------------------------------------------
// Row replacements THCM::evaluate applies on top of the Fortran assembly (THCM.C:1013-1041, 1164-1172, 2180-2296): the salinity
// integral condition row (SRES = 0: F_row = sign (c.x - correction), J_row = sign c^T, a DENSE row that is not part of the
// static graph -- its product is a dot product) and the two pressure Dirichlet rows ("Fix Pressure Points": identity rows).
// rows3 = {integral row, pfix1, pfix2} as LOCAL row ids, -1 = not owned / off.
// ---------------------------------------------------------------------------
struct FixRows { int ic, p1, p2; };
__global__ void fix_residual_rows_kernel(FixRows r, double sign, const double* __restrict__ cx, double correction, double* __restrict__ F) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (r.ic >= 0) F[r.ic] = sign * (*cx - correction);
    if (r.p1 >= 0) F[r.p1] = 0.0;
    if (r.p2 >= 0) F[r.p2] = 0.0;
}
__global__ void fix_spmv_rows_kernel(FixRows r, double sign, const double* __restrict__ cx, const double* __restrict__ x, double* __restrict__ y) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    if (r.ic >= 0) y[r.ic] = sign * (*cx);
    (void)x;   // the pressure rows are identity rows of the stored matrix already
}
__global__ void fix_jacobian_rows_kernel(FixRows r, const int* __restrict__ rp, const int* __restrict__ col, double* __restrict__ val) {
    const int rows[3] = {r.ic, r.p1, r.p2};
    const int which = blockIdx.x;
    const int row = rows[which];
    if (row < 0) return;
    for (int q = rp[row] + threadIdx.x; q < rp[row + 1]; q += blockDim.x)
        val[q] = (which > 0 && col[q] == row) ? 1.0 : 0.0;   // integral row: the sparse part is empty (the dense part lives in the SpMV)
}
int fix_residual_rows(thcmb_ctx* c, const double* d_un, double* d_F) {
    if (!c->ic_on && !c->pfix_on) return 0;
    if (c->ic_on) dot_dev(c, c->blk.ndim(), c->d_iccoeff, d_un, c->d_scalars + 4010);
    fix_residual_rows_kernel<<<1, 32, 0, c->stream>>>(FixRows{c->ic_on ? c->ic_lrow : -1, c->pfix_on ? c->pfix_lrow[0] : -1, c->pfix_on ? c->pfix_lrow[1] : -1},
                                                      (double)c->ic_sign, c->d_scalars + 4010, c->ic_correction, d_F);
    c->launches++;
    return 0;
}
int fix_spmv_rows(thcmb_ctx* c, const double* d_x, double* d_y) {
    if (!c->ic_on) return 0;
    dot_dev(c, c->blk.ndim(), c->d_iccoeff, d_x, c->d_scalars + 4010);
    fix_spmv_rows_kernel<<<1, 32, 0, c->stream>>>(FixRows{c->ic_lrow, -1, -1}, (double)c->ic_sign, c->d_scalars + 4010, d_x, d_y);
    c->launches++;
    return 0;
}
int fix_jacobian_rows(thcmb_ctx* c) {
    if (!c->ic_on && !c->pfix_on) return 0;
    fix_jacobian_rows_kernel<<<3, 32, 0, c->stream>>>(FixRows{c->ic_on ? c->ic_lrow : -1, c->pfix_on ? c->pfix_lrow[0] : -1, c->pfix_on ? c->pfix_lrow[1] : -1},
                                                      c->d_rowptr, c->d_col, c->d_val);
    c->launches++;
    return 0;
}

// ---------------------------------------------------------------------------
// Theta time stepping (src/transient/ThetaModel.H:87-165): the implicit step reuses residual, Jacobian and solver and adds
//   rhs_theta = M (u_n - u_{n+1}) + dt (1 - theta) F(u_n) + dt theta F(u_{n+1})      (one fused elementwise kernel)
//   J_theta   = J - M / (theta dt)   on the diagonal of the stored graph Jacobian     (one thread per row)
// with the diagonal mass matrix M = coB (assemble.F90:18-54).  Operation order follows the Epetra Update calls.
// ---------------------------------------------------------------------------
__global__ void theta_rhs_kernel(int n, double dt, double theta, const double* __restrict__ state, const double* __restrict__ old_state,
                                 const double* __restrict__ old_rhs, const double* __restrict__ cob, double* __restrict__ F) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double xdot = -1.0 * state[i] + 1.0 * old_state[i];                    // ThetaModel.H:103-104
        const double bx = cob[i] * xdot;                                             // applyMassMat
        const double f = (dt * (1 - theta)) * old_rhs[i] + (dt * theta) * F[i];      // :108-109
        F[i] = 1.0 * bx + 1.0 * f;                                                   // :112
    }
}
__global__ void theta_jac_kernel(int nrow, double dt, double theta, const int* __restrict__ rp, const int* __restrict__ col,
                                 const double* __restrict__ cob, double* __restrict__ val, int* __restrict__ missing) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= nrow) return;
    const double value = -cob[r] / dt / theta;                                       // ThetaModel.H:139
    bool found = false;
    for (int q = rp[r]; q < rp[r + 1]; q++)
        if (col[q] == r) { val[q] = val[q] + value; found = true; break; }           // SumIntoGlobalValues on the diagonal
    if (!found) atomicAdd(missing, 1);
}
// out = M v with the diagonal mass matrix (Ocean::applyMassMat, Ocean.C:1448-1457: out.Multiply(1.0, diagB, v, 0.0); `out` is not read)
__global__ void mass_apply_kernel(int n, const double* __restrict__ cob, const double* __restrict__ v, double* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) out[i] = cob[i] * v[i];
}
int mass_apply(thcmb_ctx* c, int n, const double* d_cob, const double* v, double* out) {
    ProfScope prof_(c, KID_AXPBY);
    mass_apply_kernel<<<ew_grid(n), 256, 0, c->stream>>>(n, d_cob, v, out); c->launches++; return 0;
}
int theta_rhs(thcmb_ctx* c, int n, double theta, double dt, const double* state, const double* old_state, const double* old_rhs,
              const double* d_cob, double* F) {
    ProfScope prof_(c, KID_AXPBY);
    theta_rhs_kernel<<<ew_grid(n), 256, 0, c->stream>>>(n, dt, theta, state, old_state, old_rhs, d_cob, F); c->launches++; return 0;
}
int theta_jacobian(thcmb_ctx* c, double theta, double dt, const double* d_cob) {
    if (!c->d_flags) { THCM_CUDA(cudaMalloc(&c->d_flags, sizeof(int) * 8)); THCM_CUDA(cudaMemsetAsync(c->d_flags, 0, sizeof(int) * 8, c->stream)); }
    const int nrow = c->blk.ndim();
    THCM_CUDA(cudaMemsetAsync(c->d_flags + 4, 0, sizeof(int), c->stream));
    { ProfScope prof_(c, KID_AXPBY);
      theta_jac_kernel<<<(nrow + 255) / 256, 256, 0, c->stream>>>(nrow, dt, theta, c->d_rowptr, c->d_col, d_cob, c->d_val, c->d_flags + 4); }
    c->launches++;
    int missing = 0;
    THCM_CUDA(cudaMemcpyAsync(&missing, c->d_flags + 4, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    if (missing) fatal("theta_jacobian: a row of the maximal graph has no diagonal entry");
    return 0;
}

// ---------------------------------------------------------------------------
// Batched (classical) Gram-Schmidt pass, the orthogonalisation Belos uses on the reference's production path
// (Ocean.C:977-1024: "Orthogonalization" = "DGKS"): ALL projections h = V^T w of an iteration in ONE reduction kernel
// (+ w.w), then ONE update kernel w -= V h.  Per iteration this is 2-4 dependent global reductions instead of the i+2
// of modified Gram-Schmidt -- the latency that decides multi-GPU efficiency (hard part 5).
// The cross-GPU sum of the nv+1 partial results is fused into the reduction kernel (peer mailboxes, as above).
// ---------------------------------------------------------------------------
constexpr int MD_MAXV = 64;            // projections per kernel (GMRES restart <= 63 in this mode)
constexpr int MD_CHUNK = 8;
constexpr int MD_BLOCKS = NSM * 4;
static_assert(P2P_VEC_LEN == MD_MAXV + 1, "mailbox layout");
struct VecList { const double* v[MD_MAXV]; int nv; };

// Tail of the batched reductions, run by the LAST block of the grid: fixed-order sum of the per-slice partials
// (nslices x [MD_MAXV+1]), the cross-GPU exchange (LL stores into the peers' mailboxes), out[0..nv-1] and out[nv], epilogue.
__device__ __forceinline__ void multi_tail(int nslices, int nv, const double* partial, unsigned int* counter, double* out,
                                           const P2PArgs pa, const RedEpilogue ep) {
    constexpr int stride = MD_MAXV + 1;
    const int nthreads = blockDim.x;
    __shared__ double mine[MD_MAXV + 1];
    __shared__ double seg[3][MD_MAXV + 1];
    // three threads per column, each summing a contiguous third of the slices with eight loads in flight; the thirds are
    // then added in order: a fixed summation tree, identical on every launch
    for (int idx = threadIdx.x; idx < 3 * (MD_MAXV + 1); idx += nthreads) {
        const int t3 = idx / (MD_MAXV + 1), q = idx - t3 * (MD_MAXV + 1);
        if (q <= nv) {
            const int col = q < nv ? q : MD_MAXV;
            const int ns = nslices, b0 = (int)((long long)ns * t3 / 3), b1 = (int)((long long)ns * (t3 + 1) / 3);
            const volatile double* pc = partial + col;
            double v = 0.0;
            int b = b0;
            for (; b + 8 <= b1; b += 8) {
                double t0 = pc[(size_t)(b + 0) * stride], t1 = pc[(size_t)(b + 1) * stride], t2 = pc[(size_t)(b + 2) * stride], t3v = pc[(size_t)(b + 3) * stride];
                double t4 = pc[(size_t)(b + 4) * stride], t5 = pc[(size_t)(b + 5) * stride], t6 = pc[(size_t)(b + 6) * stride], t7 = pc[(size_t)(b + 7) * stride];
                v += t0; v += t1; v += t2; v += t3v; v += t4; v += t5; v += t6; v += t7;
            }
            for (; b < b1; b++) v += pc[(size_t)b * stride];
            seg[t3][q] = v;
        }
    }
    __syncthreads();
    for (int q = threadIdx.x; q <= nv; q += nthreads) mine[q] = (seg[0][q] + seg[1][q]) + seg[2][q];
    const unsigned long long seq = pa.nranks > 1 ? *pa.seq + 1ull : 0ull;
    if (threadIdx.x == 0) *counter = 0u;
    __syncthreads();
    if (pa.nranks > 1) {
        const int par = (int)(seq & 1ull);
        const unsigned int flag = (unsigned int)seq;
        const size_t slot0 = P2P_VEC_OFFSET / sizeof(P2PSlot) + (size_t)par * P2P_MAX_RANKS * P2P_VEC_LEN;
        // push: every (peer, value) pair is one 16-byte LL store, spread over the block
        for (int t = threadIdx.x; t < pa.nranks * (nv + 1); t += nthreads) {
            const int r = t / (nv + 1), q = t - r * (nv + 1);
            if (r != pa.rank) ll_store(pa.peers[r] + slot0 + (size_t)pa.rank * P2P_VEC_LEN + q, mine[q], flag);
        }
        // pull: value q of every peer from my own mailbox, summed in rank order (identical bits on every rank)
        for (int q = threadIdx.x; q <= nv; q += nthreads) {
            double tot = 0.0;
            for (int r = 0; r < pa.nranks; r++)
                tot += (r == pa.rank) ? mine[q] : ll_wait(pa.mine + slot0 + (size_t)r * P2P_VEC_LEN + q, flag);
            out[q] = tot;      // out[0..nv-1] = V^T w, out[nv] = w.w
            if (q == nv) red_epilogue(tot, ep);
        }
        if (threadIdx.x == 0) *pa.seq = seq;
    } else {
        for (int q = threadIdx.x; q <= nv; q += nthreads) { out[q] = mine[q]; if (q == nv) red_epilogue(mine[q], ep); }
    }
    if (ep.flag2_out) {
        __syncthreads();   // out[0..nv] and the DGKS flag were written by threads of this block
        if (threadIdx.x == 0) {
            int f2 = 0;
            if (*ep.flag_out) {
                const double wwn = out[nv];
                double s2 = 0.0;
                for (int q = 0; q < nv; q++) s2 += out[q] * out[q];
                const double r = wwn - s2;
                if (r > 0.01 * wwn) *ep.final_out = r; else f2 = 1;
            }
            *ep.flag2_out = f2;
        }
    }
}

// grid = (slices, chunks): block (x, y) accumulates the projections of chunk y (8 basis vectors, + w.w for chunk 0)
// over slice x of the vectors.  All chunks are in flight at once: one latency-bound sweep instead of nv/8 sequential ones
// (what limited this kernel on the 1/8-size vectors of an 8-GPU run).
__global__ void __launch_bounds__(RED_THREADS) multi_dot_kernel(int n, VecList vl, const double* __restrict__ w, const int* __restrict__ skip,
                                                                 double* partial, unsigned int* counter, double* out, const P2PArgs pa) {
    __shared__ double red[RED_THREADS / 32][MD_CHUNK + 1];
    __shared__ bool last;
    const int nv = vl.nv, stride = MD_MAXV + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // conditional second pass (DGKS criterion): the flag is the same on every rank (it derives from all-reduced values), so
    // a skipped pass leaves the grid -- and the cross-GPU exchange -- entirely; `out` was zeroed by the caller
    const bool skipped = skip != nullptr && *skip == 0;
    if (skipped) return;
    {
        const int c0 = blockIdx.y * MD_CHUNK;
        const int nc = min(MD_CHUNK, nv - c0);
        double acc[MD_CHUNK + 1];
#pragma unroll
        for (int q = 0; q <= MD_CHUNK; q++) acc[q] = 0.0;
        // 128-bit loads (all Krylov vectors come from cudaMalloc: 256-byte aligned; n is even: 6 unknowns per cell)
        const int n2 = n >> 1;
        const double2* w2 = reinterpret_cast<const double2*>(w);
        for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < n2; i += gridDim.x * RED_THREADS) {
            const double2 wi = w2[i];
            double2 vv[MD_CHUNK];
#pragma unroll
            for (int q = 0; q < MD_CHUNK; q++) if (q < nc) vv[q] = reinterpret_cast<const double2*>(vl.v[c0 + q])[i];
#pragma unroll
            for (int q = 0; q < MD_CHUNK; q++) if (q < nc) { acc[q] += wi.x * vv[q].x; acc[q] += wi.y * vv[q].y; }
            if (c0 == 0) { acc[MD_CHUNK] += wi.x * wi.x; acc[MD_CHUNK] += wi.y * wi.y; }
        }
        if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
            const double wi = w[n - 1];
            for (int q = 0; q < nc; q++) acc[q] += wi * vl.v[c0 + q][n - 1];
            if (c0 == 0) acc[MD_CHUNK] += wi * wi;
        }
#pragma unroll
        for (int q = 0; q <= MD_CHUNK; q++) {
            double v = acc[q];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red[warp][q] = v;
        }
        __syncthreads();
        if (threadIdx.x <= MD_CHUNK) {
            double v = 0.0;
            for (int ww = 0; ww < RED_THREADS / 32; ww++) v += red[ww][threadIdx.x];
            if (threadIdx.x < nc) partial[(size_t)blockIdx.x * stride + c0 + threadIdx.x] = v;
            if (threadIdx.x == MD_CHUNK && c0 == 0) partial[(size_t)blockIdx.x * stride + MD_MAXV] = v;
        }
    }
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        last = (t == gridDim.x * gridDim.y - 1);
    }
    __syncthreads();
    if (!last) return;
    multi_tail((int)gridDim.x, nv, partial, counter, out, pa, RedEpilogue{nullptr, nullptr, nullptr, nullptr});
}

// Fused first update + second projection of classical Gram-Schmidt with re-orthogonalisation (CGS2 / DGKS):
//   w' = w - V h1   and, in the same sweep over the basis,   h2 = V^T w',  ww' = w'.w'
// Every thread owns two elements (one double2).  The update needs all nv basis values of its elements before w' is known;
// they are parked in the thread's own shared-memory column and read back by the projection, whose nv partial sums live
// in registers (fully unrolled, NVCAP of them) across all tiles of the thread.  The basis is read from HBM ONCE for both:
// three passes over V per iteration (dot, this kernel, second update) instead of four.  No block barrier in the loop.
constexpr int FUSED_THREADS = 128;
template <int NVCAP>
__global__ void __launch_bounds__(FUSED_THREADS) fused_axpy_dot_kernel(int n, VecList vl, const double* __restrict__ h1, double* __restrict__ w,
                                                                        double* partial, unsigned int* counter, double* out,
                                                                        const P2PArgs pa, const RedEpilogue ep) {
    extern __shared__ __align__(16) unsigned char fsm_raw[];
    double2* vs = reinterpret_cast<double2*>(fsm_raw);          // [nv][FUSED_THREADS]: column t belongs to thread t
    const int nv = vl.nv, n2 = n >> 1;
    __shared__ double hs[MD_MAXV];
    __shared__ double red[FUSED_THREADS / 32][NVCAP + 1];
    __shared__ bool last;
    for (int q = threadIdx.x; q < nv; q += FUSED_THREADS) hs[q] = h1[q];
    __syncthreads();
    double acc[NVCAP];
#pragma unroll
    for (int q = 0; q < NVCAP; q++) acc[q] = 0.0;
    double wwacc = 0.0;
    double2* w2 = reinterpret_cast<double2*>(w);
    double2* mycol = vs + threadIdx.x;
    for (int i = blockIdx.x * FUSED_THREADS + threadIdx.x; i < n2; i += gridDim.x * FUSED_THREADS) {
        double2 wi = w2[i];
#pragma unroll
        for (int c0 = 0; c0 < NVCAP; c0 += MD_CHUNK) {
            if (c0 < nv) {
                double2 vv[MD_CHUNK];
#pragma unroll
                for (int q = 0; q < MD_CHUNK; q++)
                    if (c0 + q < nv) vv[q] = reinterpret_cast<const double2*>(vl.v[c0 + q])[i];
#pragma unroll
                for (int q = 0; q < MD_CHUNK; q++)
                    if (c0 + q < nv) {
                        mycol[(size_t)(c0 + q) * FUSED_THREADS] = vv[q];
                        wi.x = wi.x - hs[c0 + q] * vv[q].x;
                        wi.y = wi.y - hs[c0 + q] * vv[q].y;
                    }
            }
        }
        w2[i] = wi;
        wwacc += wi.x * wi.x; wwacc += wi.y * wi.y;
#pragma unroll
        for (int q = 0; q < NVCAP; q++)
            if (q < nv) {
                const double2 v = mycol[(size_t)q * FUSED_THREADS];
                acc[q] += wi.x * v.x; acc[q] += wi.y * v.y;
            }
    }
    // per-block results: warp shuffle tree, then the warps in order
    constexpr int stride = MD_MAXV + 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q <= NVCAP; q++) {
        double v = q < NVCAP ? acc[q < NVCAP ? q : 0] : wwacc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x <= NVCAP) {
        double v = 0.0;
        for (int ww = 0; ww < FUSED_THREADS / 32; ww++) v += red[ww][threadIdx.x];
        if (threadIdx.x < nv) partial[(size_t)blockIdx.x * stride + threadIdx.x] = v;
        if (threadIdx.x == NVCAP) partial[(size_t)blockIdx.x * stride + MD_MAXV] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    multi_tail((int)gridDim.x, nv, partial, counter, out, pa, ep);
}

// Variant 2 of the fused first update + second projection (THCM_FUSED_CGS2=2): L2-tiled instead of parked in shared memory.
// A block walks over tiles of 256 double2 elements.  Phase A: thread t owns element t of the tile, streams the nv basis values
// of that element from HBM (8 independent 16-byte loads in flight), forms w' = w - V h1, stores it to global memory and into a
// 4 KB shared tile.  Phase B: warp q' takes the basis vectors q = q', q'+8, ... and re-reads THEIR 4 KB rows of the same tile --
// now L2 hits, the rows were touched microseconds ago and the live tiles of all resident blocks (2 per SM x nv x 4 KB ~ 60 MB at
// nv = 50) fit the 126 MB L2 -- against the shared w' tile; the per-lane partial sums of a vector stay in ONE register across
// all tiles.  HBM sees the basis once; no nv-sized register arrays, no nv x 2 KB shared parking, 16 warps per SM.
constexpr int F2_THREADS = 256;
constexpr int F2_VPW = MD_MAXV / (F2_THREADS / 32);   // basis vectors per warp (8)
// BPS: resident blocks per SM the register budget is bounded for, CHUNK: independent basis loads per thread in phase A
template <int BPS, int CHUNK>
__global__ void __launch_bounds__(F2_THREADS, BPS) fused2_axpy_dot_kernel(int n, VecList vl, const double* __restrict__ h1, double* __restrict__ w,
                                                                          double* partial, unsigned int* counter, double* out,
                                                                          const P2PArgs pa, const RedEpilogue ep) {
    constexpr int NW = F2_THREADS / 32;
    __shared__ double hs[MD_MAXV];
    __shared__ double2 wt[2][F2_THREADS];
    __shared__ double wred[NW];
    __shared__ bool last;
    const int nv = vl.nv, n2 = n >> 1;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int q = threadIdx.x; q < nv; q += F2_THREADS) hs[q] = h1[q];
    __syncthreads();
    double acc[F2_VPW];
#pragma unroll
    for (int j = 0; j < F2_VPW; j++) acc[j] = 0.0;
    double wwacc = 0.0;
    double2* w2 = reinterpret_cast<double2*>(w);
    const int ntiles = (n2 + F2_THREADS - 1) / F2_THREADS;
    int buf = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, buf ^= 1) {
        const int base = tile * F2_THREADS;
        // ---- phase A: w' = w - V h1 on this thread's element (q order = the order of multi_axpy) ----
        const int i = base + threadIdx.x;
        const bool valid = i < n2;
        double2 wi = valid ? w2[i] : make_double2(0.0, 0.0);
        for (int c0 = 0; c0 < nv; c0 += CHUNK) {
            double2 vv[CHUNK];
#pragma unroll
            for (int q = 0; q < CHUNK; q++)
                vv[q] = (valid && c0 + q < nv) ? reinterpret_cast<const double2*>(vl.v[c0 + q])[i] : make_double2(0.0, 0.0);
#pragma unroll
            for (int q = 0; q < CHUNK; q++)
                if (c0 + q < nv) { wi.x = wi.x - hs[c0 + q] * vv[q].x; wi.y = wi.y - hs[c0 + q] * vv[q].y; }
        }
        if (valid) w2[i] = wi;
        wt[buf][threadIdx.x] = wi;
        wwacc += wi.x * wi.x; wwacc += wi.y * wi.y;
        __syncthreads();   // the tile is complete; also: every warp has left phase B of tile - 1, whose buffer tile + 1 will overwrite
        // ---- phase B: h2[q] += V_q[tile] . w'[tile] for this warp's vectors (L2 hits) ----
#pragma unroll
        for (int j = 0; j < F2_VPW; j++) {
            const int q = warp + NW * j;
            if (q < nv) {
                const double2* vq = reinterpret_cast<const double2*>(vl.v[q]) + base;
                double2 v[F2_THREADS / 32];
#pragma unroll
                for (int r = 0; r < F2_THREADS / 32; r++) {
                    const int e = lane + 32 * r;
                    v[r] = (base + e < n2) ? vq[e] : make_double2(0.0, 0.0);
                }
                double a = acc[j];
#pragma unroll
                for (int r = 0; r < F2_THREADS / 32; r++) {
                    const double2 wv = wt[buf][lane + 32 * r];
                    a += wv.x * v[r].x; a += wv.y * v[r].y;
                }
                acc[j] = a;
            }
        }
    }
    // per-block results: one (warp, j) pair per vector, lanes summed by a shuffle tree; w'.w' over the whole block
    constexpr int stride = MD_MAXV + 1;
#pragma unroll
    for (int j = 0; j < F2_VPW; j++) {
        double v = acc[j];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        const int q = warp + NW * j;
        if (lane == 0 && q < nv) partial[(size_t)blockIdx.x * stride + q] = v;
    }
    {
        double v = wwacc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) wred[warp] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0;
        for (int ww = 0; ww < NW; ww++) v += wred[ww];
        partial[(size_t)blockIdx.x * stride + MD_MAXV] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        unsigned int t = atomicAdd(counter, 1u);
        last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!last) return;
    multi_tail((int)gridDim.x, nv, partial, counter, out, pa, ep);
}

// (A third variant -- the basis tile staged in SHARED memory by cp.async.bulk / mbarrier, single- and double-buffered, so that phase B
// re-reads nothing through the L2 -- was built and measured in round 2 (profiles/r02/bench_r02y_cgs3.json, bench_r02z2_cgs3.json,
// ncu_full_f3_r02z_summary.json): 0.203 ms and 0.272 ms per launch against 0.205 ms of variant 2, correct but not faster; removed.
// What its profile says about this operation is in DESIGN.md section 3.2.)
// w -= sum_q h[q] v_q  (applied in q order);  skipped when *skip == 0
__global__ void __launch_bounds__(256) multi_axpy_kernel(int n, VecList vl, const double* __restrict__ h, const int* __restrict__ skip,
                                                          double* __restrict__ w) {
    __shared__ double hs[MD_MAXV];
    if (skip != nullptr && *skip == 0) return;
    for (int q = threadIdx.x; q < vl.nv; q += blockDim.x) hs[q] = h[q];
    __syncthreads();
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double wi = w[i];
        for (int c0 = 0; c0 < vl.nv; c0 += MD_CHUNK) {
            double vv[MD_CHUNK];
#pragma unroll
            for (int q = 0; q < MD_CHUNK; q++) vv[q] = (c0 + q < vl.nv) ? vl.v[c0 + q][i] : 0.0;
#pragma unroll
            for (int q = 0; q < MD_CHUNK; q++) if (c0 + q < vl.nv) wi = wi - hs[c0 + q] * vv[q];
        }
        w[i] = wi;
    }
}

// fused: w -= V h, then ww = w.w of the UPDATED w in the same sweep (+ cross-GPU sum, + the DGKS decision in the epilogue):
// one pass over w and one kernel less per orthogonalisation pass than multi_axpy followed by dot
__global__ void __launch_bounds__(RED_THREADS) multi_axpy_dot_kernel(int n, VecList vl, const double* __restrict__ h, const int* __restrict__ skip,
                                                                      double* __restrict__ w, double* partial, unsigned int* counter, double* out,
                                                                      const P2PArgs pa, const RedEpilogue ep) {
    __shared__ double hs[MD_MAXV];
    if (skip != nullptr && *skip == 0) return;
    for (int q = threadIdx.x; q < vl.nv; q += blockDim.x) hs[q] = h[q];
    __syncthreads();
    double acc = 0.0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double wi = w[i];
        for (int c0 = 0; c0 < vl.nv; c0 += MD_CHUNK) {
            double vv[MD_CHUNK];
#pragma unroll
            for (int q = 0; q < MD_CHUNK; q++) vv[q] = (c0 + q < vl.nv) ? vl.v[c0 + q][i] : 0.0;
#pragma unroll
            for (int q = 0; q < MD_CHUNK; q++) if (c0 + q < vl.nv) wi = wi - hs[c0 + q] * vv[q];
        }
        w[i] = wi;
        acc += wi * wi;
    }
    double s = block_sum(acc);
    finish_reduction(s, partial, counter, out, pa, ep);
}

// sum of squares of the T and S fields of an interleaved state vector (vmix_control, mix_imp.f:149-155); rare (once per
// continuation step with Mixing = 2), so plain atomics + a host read
__global__ void field_sumsq_kernel(int ncell, const double* __restrict__ un, double* out2) {
    double t = 0.0, s = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < ncell; c += gridDim.x * blockDim.x) {
        const double a = un[(size_t)NUN * c + 4], b = un[(size_t)NUN * c + 5];
        t += a * a; s += b * b;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { t += __shfl_xor_sync(0xffffffffu, t, o); s += __shfl_xor_sync(0xffffffffu, s, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(out2, t); atomicAdd(out2 + 1, s); }
}
int field_sumsq(thcmb_ctx* c, const double* d_un, double* h_out2) {
    double* d = c->d_scalars + 4000;
    THCM_CUDA(cudaMemsetAsync(d, 0, 2 * sizeof(double), c->stream));
    field_sumsq_kernel<<<NSM * 4, 256, 0, c->stream>>>(c->blk.ncell(), d_un, d);
    c->launches++;
    if (c->blk.nranks > 1) allreduce_dev(c, d, 2);
    THCM_CUDA(cudaMemcpyAsync(h_out2, d, 2 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    return 0;
}

// DGKS criterion on the device: need2 = (ww_new < 0.5 * ww_old)  (Belos DGKS dep_tol = 1/sqrt(2) on the norms)
__global__ void dgks_flag_kernel(const double* ww_old, const double* ww_new, int* flag) { *flag = (*ww_new < 0.5 * (*ww_old)) ? 1 : 0; }

static P2PArgs p2p_vec_args(thcmb_ctx* c) {
    P2PArgs pa{1, 0, nullptr, nullptr, nullptr};
    if (c->p2p_on) {
        pa.nranks = c->blk.nranks; pa.rank = c->blk.rank;
        pa.mine = (P2PSlot*)c->d_mailbox; pa.peers = (P2PSlot* const*)c->d_peer_mailboxes;
        pa.seq = c->d_p2p_seq + 1;
    }
    return pa;
}

int multi_dot_dev(thcmb_ctx* c, int n, int nv, double* const* vecs, const double* w, const int* d_skip, double* d_out) {
    if (nv > MD_MAXV) {
        // more basis vectors than one kernel takes: chunks of MD_MAXV in increasing order -- every chunk leaves w.w behind its last
        // projection, the next chunk overwrites that slot with its first projection, the last chunk's lands at d_out[nv]
        for (int q0 = 0; q0 < nv; q0 += MD_MAXV) multi_dot_dev(c, n, std::min(MD_MAXV, nv - q0), vecs + q0, w, d_skip, d_out + q0);
        return 0;
    }
    VecList vl; vl.nv = nv;
    for (int q = 0; q < nv; q++) vl.v[q] = vecs[q];
    if (!c->d_mdpartial) THCM_CUDA(cudaMalloc(&c->d_mdpartial, sizeof(double) * (size_t)MD_BLOCKS * (MD_MAXV + 1)));
    { ProfScope prof_(c, KID_MULTIDOT);
      const int nchunk = std::max(1, (nv + MD_CHUNK - 1) / MD_CHUNK);
      // ONE wave: slices x chunks = what the SMs hold at once (76 registers x 256 threads: 3 blocks per SM -- the 592 blocks of the
      // r02m build ran as 444 + a 148-block tail at a third of the occupancy, ~8 % of the kernel)
      static int resident = 0;
      if (!resident) {
          int occ = 0;
          THCM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, multi_dot_kernel, RED_THREADS, 0));
          resident = std::min(MD_BLOCKS, NSM * std::max(occ, 1));
      }
      const int slices = std::max(32, resident / nchunk);
      multi_dot_kernel<<<dim3(slices, nchunk), RED_THREADS, 0, c->stream>>>(n, vl, w, d_skip, c->d_mdpartial, c->d_counter, d_out, p2p_vec_args(c)); }
    c->launches++;
    return c->p2p_on ? 0 : allreduce_dev(c, d_out, nv + 1);
}
int multi_axpy_dev(thcmb_ctx* c, int n, int nv, double* const* vecs, const double* d_h, const int* d_skip, double* w) {
    if (nv > MD_MAXV) {
        for (int q0 = 0; q0 < nv; q0 += MD_MAXV) multi_axpy_dev(c, n, std::min(MD_MAXV, nv - q0), vecs + q0, d_h + q0, d_skip, w);
        return 0;
    }
    VecList vl; vl.nv = nv;
    for (int q = 0; q < nv; q++) vl.v[q] = vecs[q];
    ProfScope prof_(c, KID_MULTIAXPY);
    multi_axpy_kernel<<<ew_grid(n), 256, 0, c->stream>>>(n, vl, d_h, d_skip, w);
    c->launches++;
    return 0;
}
// rare_path: the launch almost always finds its skip flag cleared and leaves (the explicit second update behind the Pythagorean norm):
// one block per SM keeps what such a launch costs small (6.6 -> ~4 us); when it does run it is slower than the full grid
int multi_axpy_dot_dev(thcmb_ctx* c, int n, int nv, double* const* vecs, const double* d_h, const int* d_skip, double* w, double* d_ww,
                       const double* d_ww_old, int* d_flag_out, double* d_final_out, int kid, bool rare_path) {
    if (nv > MD_MAXV) {   // all but the last chunk only update; the last one also takes the norm of the fully updated vector
        const int last0 = (nv - 1) / MD_MAXV * MD_MAXV;
        multi_axpy_dev(c, n, last0, vecs, d_h, d_skip, w);
        return multi_axpy_dot_dev(c, n, nv - last0, vecs + last0, d_h + last0, d_skip, w, d_ww, d_ww_old, d_flag_out, d_final_out, kid, rare_path);
    }
    VecList vl; vl.nv = nv;
    for (int q = 0; q < nv; q++) vl.v[q] = vecs[q];
    { ProfScope prof_(c, kid >= 0 ? kid : KID_MULTIAXPY);
      const int grid = rare_path ? std::min(ew_grid(n), NSM) : std::min(ew_grid(n), RED_BLOCKS);
      multi_axpy_dot_kernel<<<grid, RED_THREADS, 0, c->stream>>>(n, vl, d_h, d_skip, w, c->d_partial, c->d_counter, d_ww, p2p_args(c),
                                                                 RedEpilogue{d_ww_old, d_flag_out, d_final_out, nullptr}); }
    c->launches++;
    if (!c->p2p_on && c->blk.nranks > 1) fatal("multi_axpy_dot needs the P2P mailboxes on multi-GPU runs (THCM_P2P=1)");
    return 0;
}
// w -= V h1 ; out[0..nv) = V^T w (updated w) ; out[nv] = w.w ; *flag = out[nv] < 0.5 * *ww_old ; *final = out[nv]
template <int NVCAP>
static void launch_fused(thcmb_ctx* c, int n, const VecList& vl, const double* d_h1, double* w, double* d_out, const RedEpilogue& ep) {
    const size_t smem = (size_t)vl.nv * FUSED_THREADS * sizeof(double2);
    static bool attr_set = false;
    if (!attr_set) {
        THCM_CUDA(cudaFuncSetAttribute(fused_axpy_dot_kernel<NVCAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((size_t)NVCAP * FUSED_THREADS * sizeof(double2))));
        attr_set = true;
    }
    int occ = 0;
    THCM_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fused_axpy_dot_kernel<NVCAP>, FUSED_THREADS, smem));
    const int n2 = n >> 1;
    const int grid = std::max(1, std::min(std::min((n2 + FUSED_THREADS - 1) / FUSED_THREADS, NSM * std::max(occ, 1)), MD_BLOCKS));
    fused_axpy_dot_kernel<NVCAP><<<grid, FUSED_THREADS, smem, c->stream>>>(n, vl, d_h1, w, c->d_mdpartial, c->d_counter, d_out, p2p_vec_args(c), ep);
}
int fused_axpy_dot_dev(thcmb_ctx* c, int n, int nv, double* const* vecs, const double* d_h1, double* w, double* d_out,
                       const double* d_ww_old, int* d_flag_out, double* d_final_out, int* d_flag2_out) {
    if (nv > MD_MAXV) fatal("fused_axpy_dot: too many vectors");
    if (n & 1) fatal("fused_axpy_dot: vector length must be even (6 unknowns per cell)");
    if (!c->p2p_on && c->blk.nranks > 1) fatal("fused_axpy_dot needs the P2P mailboxes on multi-GPU runs (THCM_P2P=1)");
    VecList vl; vl.nv = nv;
    for (int q = 0; q < nv; q++) vl.v[q] = vecs[q];
    if (!c->d_mdpartial) THCM_CUDA(cudaMalloc(&c->d_mdpartial, sizeof(double) * (size_t)MD_BLOCKS * (MD_MAXV + 1)));
    const RedEpilogue ep{d_ww_old, d_flag_out, d_final_out, d_flag2_out};
    // measured (profiles/ncu_full_r01d_fused_cgs2_summary.json): at nv = 11 the shared-memory kernel <16> needs 0.17 ms, the L2-tiled one
    // 0.21 ms; averaged over nv = 1..50 it is 0.47 vs 0.40 ms (the <32..64> instantiations park up to 131 KB per block)
    if (c->fused_cgs2 >= 2 && nv > 16) {
        ProfScope prof_(c, KID_MULTIAXPY);
        const int ntiles = ((n >> 1) + F2_THREADS - 1) / F2_THREADS;
        // resident blocks per SM x independent phase-A loads per thread, measured at 1 degree (Newton step, profiles/bench_r01g_*,
        // profiles/r02/bench_r02a_THCM_FUSED2_BPS_*): 2 x 8 (92 registers) 75.6 ms, 3 x 4 (78 registers) 74.5 ms = kept, 4 x 4 (64
        // registers) 77.5 ms, 3 x 8 74.3 ms (within noise).  The live tiles stay below the L2 size: 3 x 148 x nv x 4 KB = 89 MB at nv = 50
        // Above ~32 vectors the live tiles of 3 blocks per SM (444 x nv x 4 KB) no longer stay in the L2: ncu at nv = 49 (profiles/r02/
        // ncu_full_r02f_summary.json) shows 2.42 GB of DRAM traffic for 1.69 GB of algorithmic bytes -- phase B re-reads from HBM.  Two
        // blocks per SM (296 x nv x 4 KB = 59 MB at nv = 50) with 8 loads in flight per thread keep phase B in the L2.
        if (nv > 32) {
            const int grid = std::max(1, std::min(std::min(ntiles, NSM * 2), MD_BLOCKS));
            fused2_axpy_dot_kernel<2, 8><<<grid, F2_THREADS, 0, c->stream>>>(n, vl, d_h1, w, c->d_mdpartial, c->d_counter, d_out, p2p_vec_args(c), ep);
        } else {
            const int grid = std::max(1, std::min(std::min(ntiles, NSM * 3), MD_BLOCKS));
            fused2_axpy_dot_kernel<3, 4><<<grid, F2_THREADS, 0, c->stream>>>(n, vl, d_h1, w, c->d_mdpartial, c->d_counter, d_out, p2p_vec_args(c), ep);
        }
        c->launches++;
        return 0;
    }
    { ProfScope prof_(c, KID_MULTIAXPY);
      if (nv <= 16) launch_fused<16>(c, n, vl, d_h1, w, d_out, ep);
      else if (nv <= 32) launch_fused<32>(c, n, vl, d_h1, w, d_out, ep);
      else if (nv <= 48) launch_fused<48>(c, n, vl, d_h1, w, d_out, ep);
      else launch_fused<64>(c, n, vl, d_h1, w, d_out, ep); }
    c->launches++;
    return 0;
}
int dgks_flag_dev(thcmb_ctx* c, const double* ww_old, const double* ww_new, int* d_flag) {
    dgks_flag_kernel<<<1, 1, 0, c->stream>>>(ww_old, ww_new, d_flag);
    c->launches++;
    return 0;
}

// ---------------------------------------------------------------------------
// halo exchange (replaces Epetra_Import of TRIOS_Domain.C:599 / the column-map import of
// Epetra_CrsMatrix::Apply): pack -> grouped ncclSend/ncclRecv -> unpack, all on the context's stream.
// NCCL is bound lazily with dlopen so that single-GPU use has no NCCL dependency.
// ---------------------------------------------------------------------------
__global__ void halo_pack_kernel(int ncells, const int* __restrict__ idx, const double* __restrict__ x, double* __restrict__ buf) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ncells * NUN; t += gridDim.x * blockDim.x) {
        int cidx = t / NUN, v = t - cidx * NUN;
        buf[t] = x[(size_t)NUN * idx[cidx] + v];
    }
}
__global__ void halo_unpack_kernel(int ncells, const int* __restrict__ slot, const double* __restrict__ buf, double* __restrict__ halo) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ncells * NUN; t += gridDim.x * blockDim.x) {
        int cidx = t / NUN, v = t - cidx * NUN;
        halo[(size_t)NUN * slot[cidx] + v] = buf[t];
    }
}

// Halo exchange as ONE kernel over NVLink peer memory: every boundary cell is stored straight into the neighbour's halo
// buffer (no pack / NCCL send-recv / unpack: three launches and two staging copies less per operator application); the
// last block to finish publishes a sequence flag in every neighbour's mailbox and waits for theirs, so the halo is complete
// when the kernel ends.  Two halo buffers alternate by the parity of the sequence number: a neighbour that runs ahead
// writes exchange s+1 into the other buffer while this rank still reads exchange s.
constexpr int HALO_MAX_PEERS = 16;   // per-band longitude cuts: a north / south edge can touch several blocks
struct HaloPeers { int n; int rank[HALO_MAX_PEERS]; unsigned char send[HALO_MAX_PEERS], recv[HALO_MAX_PEERS]; };
__global__ void __launch_bounds__(256) halo_push_kernel(int ncells, const int* __restrict__ idx, const int* __restrict__ dst_slot,
                                                         const int* __restrict__ peer, const double* __restrict__ x, double* const* peer_halo,
                                                         HaloPeers hp, int myrank, char* const* peer_base, char* my_base,
                                                         unsigned long long seq, unsigned int* counter, int wait) {
    __shared__ bool last;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ncells * (NUN / 2); t += gridDim.x * blockDim.x) {
        const int cidx = t / (NUN / 2), v = t - cidx * (NUN / 2);   // 16-byte pieces: records are 48 bytes, 16-byte aligned
        const double2 val = reinterpret_cast<const double2*>(x + (size_t)NUN * idx[cidx])[v];
        reinterpret_cast<double2*>(peer_halo[peer[cidx]] + (size_t)NUN * dst_slot[cidx])[v] = val;
    }
    // one system-scope fence per block, by the thread that has observed the whole block's stores through the barrier
    // (fences are cumulative); a fence per thread made this kernel ~3x slower
    __syncthreads();
    if (threadIdx.x == 0) { __threadfence_system(); last = atomicAdd(counter, 1u) == gridDim.x - 1; }
    __syncthreads();
    if (!last) return;
    // every block's stores were performed at the neighbours (its system fence completed) before it bumped the counter this
    // block has just read: the flags can follow with a device-scope fence only -- system fences cost ~5 us each here
    // (the flag store to a PEER must be ordered after them at system scope in the PTX memory model: one fence, one thread group)
    __threadfence_system();
    const int par = (int)(seq & 1ull);
    if (threadIdx.x < hp.n) {
        const int q = threadIdx.x;
        if (hp.send[q]) {
            volatile unsigned long long* f = (volatile unsigned long long*)(peer_base[hp.rank[q]] + P2P_HALOFLAG_OFFSET) + par * P2P_MAX_RANKS + myrank;
            *f = seq;
        }
        if (wait && hp.recv[q]) {
            volatile unsigned long long* f = (volatile unsigned long long*)(my_base + P2P_HALOFLAG_OFFSET) + par * P2P_MAX_RANKS + hp.rank[q];
            flag_wait(f, seq);
        }
    }
    if (wait) __threadfence_system();
    if (threadIdx.x == 0) *counter = 0u;
}
struct Id128 { char b[128]; };  // ncclUniqueId (nccl.h: struct { char internal[128]; }), passed by value
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(void**, int, Id128, int) = nullptr;
    int (*CommDestroy)(void*) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*Send)(const void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, void*, cudaStream_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
};
static NcclApi g_nccl;
static bool nccl_load() {
    if (g_nccl.lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) { g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL); if (g_nccl.lib) break; }
    if (!g_nccl.lib) { set_error("cannot dlopen libnccl.so.2"); return false; }
    auto sym = [&](const char* s) { return dlsym(g_nccl.lib, s); };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
    g_nccl.Send = (decltype(g_nccl.Send))sym("ncclSend");
    g_nccl.Recv = (decltype(g_nccl.Recv))sym("ncclRecv");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
    return g_nccl.GetUniqueId && g_nccl.CommInitRank && g_nccl.Send && g_nccl.Recv && g_nccl.AllReduce;
}
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;  // ncclDataType_t / ncclRedOp_t values (nccl.h)

int nccl_unique_id(void* id128) { if (!nccl_load()) return -1; return g_nccl.GetUniqueId(id128); }
int nccl_init(thcmb_ctx* c, const void* id128) {
    if (!nccl_load()) return -1;
    Id128 id; memcpy(id.b, id128, 128);
    THCM_CUDA(cudaSetDevice(c->device));
    int rc = g_nccl.CommInitRank(&c->nccl_comm, c->blk.nranks, id, c->blk.rank);
    if (rc != 0) set_error("ncclCommInitRank failed rc=" + std::to_string(rc));
    return rc;
}
void nccl_destroy(thcmb_ctx* c) { if (c->nccl_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->nccl_comm); c->nccl_comm = nullptr; }

int allreduce_dev(thcmb_ctx* c, double* d_buf, int count) {
    if (c->blk.nranks == 1) return 0;
    if (!c->nccl_comm) fatal("nranks > 1 but thcmb_nccl_init was not called");
    return g_nccl.AllReduce(d_buf, d_buf, (size_t)count, NCCL_FLOAT64, NCCL_SUM, c->nccl_comm, c->stream);
}

static HaloPeers halo_peers(const thcmb_ctx* c) {
    HaloPeers hp; hp.n = (int)c->peers.size();
    for (int q = 0; q < hp.n; q++) { hp.rank[q] = c->peers[q].rank; hp.send[q] = c->peers[q].send_cnt > 0; hp.recv[q] = c->peers[q].recv_cnt > 0; }
    return hp;
}
int halo_exchange(thcmb_ctx* c, const double* d_x, bool wait) {
    if (c->blk.nranks == 1) return 0;
    if (c->halo_p2p) {
        if ((int)c->peers.size() > HALO_MAX_PEERS) fatal("halo push: more than 16 neighbours");
        if ((((uintptr_t)d_x) & 15) != 0) fatal("halo push: vector must be 16-byte aligned");
        HaloPeers hp = halo_peers(c);
        const unsigned long long seq = ++c->halo_seq;
        const int par = (int)(seq & 1ull);
        ProfScope prof_(c, KID_HALO_PACK);
        const int grid = std::max(1, std::min(ew_grid(c->nsend_cells * (NUN / 2)), NSM));
        halo_push_kernel<<<grid, 256, 0, c->stream>>>(c->nsend_cells, c->d_send_idx, c->d_send_dst, c->d_send_peer, d_x,
                                                      c->d_peer_halo + (size_t)par * c->peers.size(), hp, c->blk.rank,
                                                      (char* const*)c->d_peer_mailboxes, (char*)c->d_mailbox, seq, c->d_halo_counter, wait ? 1 : 0);
        c->launches++;
        c->d_halo = c->d_halo_p2p[par];
        return 0;
    }
    if (!c->nccl_comm) fatal("nranks > 1 but thcmb_nccl_init was not called");
    if (c->nsend_cells > 0) {
        ProfScope prof_(c, KID_HALO_PACK);
        halo_pack_kernel<<<ew_grid(c->nsend_cells * NUN), 256, 0, c->stream>>>(c->nsend_cells, c->d_send_idx, d_x, c->d_sendbuf);
        c->launches++;
    }
    g_nccl.GroupStart();
    for (auto& p : c->peers) {
        if (p.send_cnt) g_nccl.Send(c->d_sendbuf + (size_t)NUN * p.send_off, (size_t)NUN * p.send_cnt, NCCL_FLOAT64, p.rank, c->nccl_comm, c->stream);
        if (p.recv_cnt) g_nccl.Recv(c->d_recvbuf + (size_t)NUN * p.recv_off, (size_t)NUN * p.recv_cnt, NCCL_FLOAT64, p.rank, c->nccl_comm, c->stream);
    }
    g_nccl.GroupEnd();
    if (c->nrecv_cells > 0) {
        ProfScope prof_(c, KID_HALO_UNPACK);
        halo_unpack_kernel<<<ew_grid(c->nrecv_cells * NUN), 256, 0, c->stream>>>(c->nrecv_cells, c->d_recv_slot, c->d_recvbuf, c->d_halo);
        c->launches++;
    }
    return 0;
}

// ---------------------------------------------------------------------------
// 6x6 block-diagonal preconditioner (first step): extract the in-cell block
// of the stored Jacobian, invert it with partial pivoting (w and p rows have no diagonal entry on
// ocean cells, spf.F90:176,340), fall back to the identity when a block is numerically singular.
// ---------------------------------------------------------------------------
// 128 cells per block.  Gather: the 6 x 128 rows of the block are walked by all threads (thread t takes rows t, t + 128, ...: consecutive
// threads read consecutive rows); the in-cell entries of a row are its stencil-centre slots, whose positions inside the sorted graph row
// only depend on the cell's boundary class (host-built table cpos[class][row][col], -1 = no such entry): one row-pointer load and at most
// six value loads per row, no column ids, no search.  Rows of LAND cells are identity rows and are not read at all (half of a global
// grid).  Inversion: one thread per cell, every thread busy.  Store: coalesced from shared memory.
constexpr int BDB_CELLS = 128;
__global__ void __launch_bounds__(BDB_CELLS) blockdiag_build_kernel(DevBlock b, const int* __restrict__ rp, const double* __restrict__ val,
                                                                    const signed char* __restrict__ cpos,
                                                                    const unsigned char* __restrict__ landcell, double* __restrict__ minv) {
    extern __shared__ double sA[];                      // [BDB_CELLS][37] (+1: the inversion walks the rows of 128 different blocks at once)
    constexpr int LD = NUN * NUN + 1;
    const int ncell = b.ncell;
    const int cell0 = blockIdx.x * BDB_CELLS;
    for (int lr = threadIdx.x; lr < BDB_CELLS * NUN; lr += BDB_CELLS) {
        const int lc = lr / NUN, r = lr - lc * NUN, cell = cell0 + lc;
        double* dst = sA + lc * LD + r * NUN;
        if (cell >= ncell || __ldg(landcell + cell)) {
#pragma unroll
            for (int q = 0; q < NUN; q++) dst[q] = q == r ? 1.0 : 0.0;
            continue;
        }
        const int li = cell % b.n0, rest = cell / b.n0, lj = rest % b.m0, k0 = rest / b.m0;
        const int gi = b.i0 + li + 1, gj = b.j0 + lj + 1, k = k0 + 1;
        const int cls = (gi == 1 ? 1 : 0) | (gi == b.N ? 2 : 0) | (gj == 1 ? 4 : 0) | (gj == b.M ? 8 : 0) | (k == 1 ? 16 : 0) | (k == b.L ? 32 : 0);
        const signed char* cp = cpos + cls * (NUN * NUN) + r * NUN;
        const int base = __ldg(rp + NUN * cell + r);
        double v[NUN];
#pragma unroll
        for (int q = 0; q < NUN; q++) { const int pq = cp[q]; v[q] = pq >= 0 ? __ldg(val + base + pq) : 0.0; }
#pragma unroll
        for (int q = 0; q < NUN; q++) dst[q] = v[q];
    }
    __syncthreads();
    const int c = threadIdx.x;
    if (cell0 + c < ncell && !__ldg(landcell + cell0 + c)) {
        double A[NUN][NUN], B[NUN][NUN];
        for (int i = 0; i < NUN; i++) for (int q = 0; q < NUN; q++) { A[i][q] = sA[c * LD + i * NUN + q]; B[i][q] = i == q ? 1.0 : 0.0; }
        bool singular = false;
        for (int p = 0; p < NUN; p++) {
            int piv = p; double best = fabs(A[p][p]);
            for (int i = p + 1; i < NUN; i++) if (fabs(A[i][p]) > best) { best = fabs(A[i][p]); piv = i; }
            if (best < 1e-14) { singular = true; break; }
            if (piv != p) for (int q = 0; q < NUN; q++) { double t = A[p][q]; A[p][q] = A[piv][q]; A[piv][q] = t; t = B[p][q]; B[p][q] = B[piv][q]; B[piv][q] = t; }
            double d = 1.0 / A[p][p];
            for (int q = 0; q < NUN; q++) { A[p][q] *= d; B[p][q] *= d; }
            for (int i = 0; i < NUN; i++) if (i != p) {
                double f = A[i][p];
                if (f != 0.0) for (int q = 0; q < NUN; q++) { A[i][q] -= f * A[p][q]; B[i][q] -= f * B[p][q]; }
            }
        }
        for (int i = 0; i < NUN; i++) for (int q = 0; q < NUN; q++) sA[c * LD + i * NUN + q] = singular ? (i == q ? 1.0 : 0.0) : B[i][q];
    }
    __syncthreads();
    const int nloc = min(BDB_CELLS, ncell - cell0) * NUN * NUN;
    for (int i = threadIdx.x; i < nloc; i += BDB_CELLS) minv[(size_t)cell0 * NUN * NUN + i] = sA[(i / (NUN * NUN)) * LD + i % (NUN * NUN)];
}
__global__ void blockdiag_apply_kernel(int ncell, const double* __restrict__ minv, const double* __restrict__ x, double* __restrict__ y) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= ncell * NUN) return;
    int cell = t / NUN, r = t - cell * NUN;
    const double* M = minv + (size_t)cell * 36 + r * NUN;
    const double* xc = x + (size_t)cell * NUN;
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < NUN; q++) s += M[q] * xc[q];
    y[t] = s;
}
// m_scaling::average_block (scaling.F90:29-64): sum of the in-cell 6x6 blocks of the Jacobian over the OCEAN cells (+ their
// count).  Per-block partials in a fixed reduction tree; the host adds the partials in block order (setup frequency).
constexpr int AVG_THREADS = 128;
__global__ void __launch_bounds__(AVG_THREADS) average_block_kernel(int ncell, const int* __restrict__ rp, const int* __restrict__ col,
                                                                     const double* __restrict__ val, const uint32_t* __restrict__ nbmask,
                                                                     double* __restrict__ partial /* [grid][37] */) {
    __shared__ double wsum[AVG_THREADS / 32][37];
    double A[NUN * NUN + 1];
#pragma unroll
    for (int q = 0; q <= NUN * NUN; q++) A[q] = 0.0;
    for (int cell = blockIdx.x * blockDim.x + threadIdx.x; cell < ncell; cell += gridDim.x * blockDim.x) {
        if ((nbmask[cell] >> 4) & 1u) continue;            // landm(ix,iy,iz) == OCEAN only
        A[NUN * NUN] += 1.0;
        for (int r = 0; r < NUN; r++) {
            const int row = NUN * cell + r;
            for (int q = rp[row]; q < rp[row + 1]; q++) {
                const int cc = col[q] - NUN * cell;
                if (cc >= 0 && cc < NUN) {
#pragma unroll
                    for (int r2 = 0; r2 < NUN; r2++)
#pragma unroll
                        for (int c2 = 0; c2 < NUN; c2++) if (r2 == r && c2 == cc) A[r2 * NUN + c2] += val[q];
                }
            }
        }
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q <= NUN * NUN; q++) {
        double v = A[q];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) wsum[warp][q] = v;
    }
    __syncthreads();
    if (threadIdx.x <= NUN * NUN) {
        double v = 0.0;
        for (int w = 0; w < AVG_THREADS / 32; w++) v += wsum[w][threadIdx.x];
        partial[(size_t)blockIdx.x * 37 + threadIdx.x] = v;
    }
}
// db[(ii-1) + 6*(jj-1)] (Fortran layout) = local average block, as the reference's average_block leaves it
int average_block(thcmb_ctx* c, double* db36) {
    const int ncell = c->blk.ncell();
    const int grid = std::max(1, std::min((ncell + AVG_THREADS - 1) / AVG_THREADS, NSM * 4));
    double* d_part = nullptr;
    THCM_CUDA(cudaMalloc(&d_part, sizeof(double) * 37 * (size_t)grid));
    average_block_kernel<<<grid, AVG_THREADS, 0, c->stream>>>(ncell, c->d_rowptr, c->d_col, c->d_val, c->d_nbmask, d_part);
    c->launches++;
    std::vector<double> part((size_t)37 * grid);
    THCM_CUDA(cudaMemcpyAsync(part.data(), d_part, sizeof(double) * part.size(), cudaMemcpyDeviceToHost, c->stream));
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    cudaFree(d_part);
    double sum[37];
    for (int q = 0; q < 37; q++) { sum[q] = 0.0; for (int b = 0; b < grid; b++) sum[q] += part[(size_t)b * 37 + q]; }
    const double nl = sum[36];
    for (int r = 0; r < NUN; r++) for (int cc = 0; cc < NUN; cc++) db36[r + NUN * cc] = nl > 0 ? sum[r * NUN + cc] / nl : 0.0;
    return 0;
}

int build_blockdiag(thcmb_ctx* c) {
    int ncell = c->blk.ncell();
    if (!c->d_minv) THCM_CUDA(cudaMalloc(&c->d_minv, sizeof(double) * 36 * (size_t)ncell));
    ProfScope prof_(c, KID_PRECON_BUILD);
    const size_t smem = sizeof(double) * BDB_CELLS * (NUN * NUN + 1);
    blockdiag_build_kernel<<<std::max(1, (ncell + BDB_CELLS - 1) / BDB_CELLS), BDB_CELLS, smem, c->stream>>>(dev_block(c->blk), c->d_rowptr, c->d_val,
                                                                                                            c->d_cpos, c->d_landcell, c->d_minv);
    c->launches++;
    return 0;
}
// ---------------------------------------------------------------------------
// Ocean-only (cell-compacted) Krylov space, THCM_KRYLOV_COMPACT=1 (candidate, one rank): rows of LAND cells are
// identity rows and the Newton right-hand side vanishes there, so every Krylov vector is zero on LAND -- GMRES runs on vectors that
// hold the OCEAN cells only (51.5 % of the length at 1 degree).  ocell[ci] = full cell of compact cell ci, ccell[cell] = compact
// index or -1.  The compact SpMV reads the rows of the ocean cells from the SAME graph-order value / column arrays and maps a
// column to its compact position through ccell (columns on LAND carry exact zeros and are skipped).
// ---------------------------------------------------------------------------
__global__ void gather_cells_kernel(int nc, const int* __restrict__ ocell, const double* __restrict__ in, double* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nc * NUN; i += gridDim.x * blockDim.x) {
        const int ci = i / NUN, v = i - ci * NUN;
        out[i] = in[(size_t)__ldg(ocell + ci) * NUN + v];
    }
}
__global__ void scatter_cells_kernel(int ncell, const int* __restrict__ ccell, const double* __restrict__ in, double* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < ncell * NUN; i += gridDim.x * blockDim.x) {
        const int cell = i / NUN, v = i - cell * NUN;
        const int ci = __ldg(ccell + cell);
        out[i] = ci >= 0 ? in[(size_t)ci * NUN + v] : 0.0;     // identity rows: x = b = 0 on LAND
    }
}
// number of non-zero entries of x on LAND cells, summed over the ranks (the compact space is only valid when there are none, and
// every rank must take the same decision): an ordinary reduction kernel with the fused all-reduce
__global__ void __launch_bounds__(RED_THREADS) land_nonzero_kernel(int ncell, const int* __restrict__ ccell, const double* __restrict__ x,
                                                                   double* partial, unsigned int* counter, double* out, const P2PArgs pa) {
    double v = 0.0;
    for (int i = blockIdx.x * RED_THREADS + threadIdx.x; i < ncell * NUN; i += gridDim.x * RED_THREADS)
        if (__ldg(ccell + i / NUN) < 0 && x[i] != 0.0) v += 1.0;
    double s = block_sum(v);
    finish_reduction(s, partial, counter, out, pa);
}
// Halo exchange of the compact SpMV: every boundary cell's six values go straight into the neighbour's LL halo buffer, each as one
// 16-byte {lo, flag, hi, flag} store (zeros for LAND cells).  Fire and forget: no fence, no counter, no wait -- the consumer is the
// neighbour's SpMV kernel, which polls exactly the slots its rows reference.
__global__ void __launch_bounds__(256) halo_push_ll_kernel(int ncells, const int* __restrict__ cidx, const int* __restrict__ dst_slot,
                                                            const int* __restrict__ peer, const double* __restrict__ xc,
                                                            P2PSlot* const* peer_ll, unsigned int flag) {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < ncells * NUN; t += gridDim.x * blockDim.x) {
        const int q = t / NUN, v = t - q * NUN;
        const int ci = __ldg(cidx + q);
        const double val = ci >= 0 ? xc[(size_t)NUN * ci + v] : 0.0;
        ll_store(peer_ll[__ldg(peer + q)] + (size_t)NUN * __ldg(dst_slot + q) + v, val, flag);
    }
}
// y_c = J x_c on the rows of the ocean cells.  Values and row pointers are the graph's own arrays (row = 6 ocell[ci] + r), the column
// ids come from the compact column array stored at the same offsets: < 0 LAND column (skipped), < nlocal_c owned, else a halo slot.
// HALO = 0: one rank.  HALO = 1: the slot is an LL slot that is polled until the neighbour's push of THIS exchange (flag) has landed
// (interior rows never wait).  HALO = 2: the exchange was landed into a plain array by the kernel that ran before (the landing blocks
// of scale_precon_push_kernel): owned and halo columns are ONE batch of independent gathers, the address chosen by a select.  Measured
// on 2 GPUs (profiles/r02/bench_r02s_*): 0.103 ms with the halo columns ignored, 0.114 with plain loads issued after the owned gathers,
// 0.125 with the polls -- the second, dependent round of loads costs as much as the polling itself.
template <int LANES, int UNROLL, int HALO>
__global__ void __launch_bounds__(SPMV_THREADS) spmv_compact_kernel(int nrow_c, const int* __restrict__ ocell, const int* __restrict__ rp,
                                                                     const int* __restrict__ colc, const double* __restrict__ val,
                                                                     const double* __restrict__ xc, int nlocal_c, const P2PSlot* halo_ll,
                                                                     const double* __restrict__ xh, unsigned int flag, double* __restrict__ yc) {
    const int sub = threadIdx.x & (LANES - 1);
    constexpr int rows_per_block = SPMV_THREADS / LANES;
    for (int base = blockIdx.x * rows_per_block; base < nrow_c; base += gridDim.x * rows_per_block) {
        const int i = base + (threadIdx.x / LANES);
        const bool act = i < nrow_c;
        int row = 0;
        if (act) { const int ci = i / NUN; row = __ldg(ocell + ci) * NUN + (i - ci * NUN); }
        const int b = act ? __ldg(rp + row) : 0, e = act ? __ldg(rp + row + 1) : 0;
        int cc[UNROLL]; double vv[UNROLL], xx[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; u++) {
            const int q = b + sub + u * LANES;
            const bool ok = q < e;
            cc[u] = ok ? __ldg(colc + q) : -1;
            vv[u] = ok ? __ldg(val + q) : 0.0;
        }
        if constexpr (HALO == 2) {
#pragma unroll
            for (int u = 0; u < UNROLL; u++) {
                const double* p = cc[u] < nlocal_c ? xc + cc[u] : xh + (cc[u] - nlocal_c);
                xx[u] = cc[u] >= 0 ? __ldg(p) : 0.0;
            }
        } else {
            // owned columns first, as independent predicated loads (all gathers of the row in flight together); the few halo columns
            // after them -- a poll loop inside the gather would serialise it
#pragma unroll
            for (int u = 0; u < UNROLL; u++) xx[u] = (cc[u] >= 0 && (HALO == 0 || cc[u] < nlocal_c)) ? __ldg(xc + cc[u]) : 0.0;
            if constexpr (HALO == 1) {
                bool arrived = true;
#pragma unroll
                for (int u = 0; u < UNROLL; u++)
                    if (cc[u] >= nlocal_c) arrived = ll_poll(halo_ll + (cc[u] - nlocal_c), flag, xx[u]) && arrived;
                if (!arrived) p2p_timeout(halo_ll, flag, 0u);
            }
        }
        double s = 0.0;
#pragma unroll
        for (int u = 0; u < UNROLL; u++) s += vv[u] * xx[u];
        for (int q = b + sub + UNROLL * LANES; q < e; q += LANES) {   // rows longer than UNROLL*LANES (not in the THCM graph)
            const int cidx = __ldg(colc + q);
            if (cidx < 0) continue;
            double xv;
            if (HALO == 0 || cidx < nlocal_c) xv = __ldg(xc + cidx);
            else if (HALO == 2) xv = __ldg(xh + (cidx - nlocal_c));
            else xv = ll_wait(halo_ll + (cidx - nlocal_c), flag);
            s += __ldg(val + q) * xv;
        }
#pragma unroll
        for (int o = LANES / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o, LANES);
        if (sub == 0 && act) yc[i] = s;
    }
}
__global__ void blockdiag_apply_compact_kernel(int nc, const int* __restrict__ ocell, const double* __restrict__ minv,
                                               const double* __restrict__ x, double* __restrict__ y) {
    int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nc * NUN) return;
    int ci = t / NUN, r = t - ci * NUN;
    const double* M = minv + (size_t)__ldg(ocell + ci) * 36 + r * NUN;
    const double* xc = x + (size_t)ci * NUN;
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < NUN; q++) s += M[q] * xc[q];
    y[t] = s;
}
// Head of an Arnoldi step in ONE kernel (compact space, 6x6 block-diagonal preconditioner, flexible GMRES, batched orthogonalisation):
//   w'' = w' - V h2      the SECOND Gram-Schmidt update of the previous step, when its DGKS flag asked for it (its norm is known from
//                        Pythagoras, RedEpilogue::flag2_out; the rare explicit path has done the update already: *flag2 = 1),
//   v = w'' / ||w''||    (GMRESSolver.H:185-186),    z = M^-1 v    (:160-164),
//   and the halo of z pushed into the neighbours' LL buffers for the SpMV that follows.
// w is only read (the orthogonalisation works in a buffer of its own), so the blocks that push -- the FIRST blocks of the grid: their
// stores travel while the others sweep the vector; they recompute z for the cells of the send lists, ~2 % of the block -- need no
// ordering against the sweeping blocks.  Replaces multi_axpy_dot (+ its all-reduce) + scale_invsqrt + blockdiag_apply + halo_push:
// four launches, one reduction and two passes over the vector less per iteration.
constexpr int SPP_CELLS = 32, SPP_THREADS = SPP_CELLS * NUN;
__global__ void __launch_bounds__(SPP_THREADS) scale_precon_push_kernel(int nc, const int* __restrict__ ocell, const double* __restrict__ minv,
                                                                        const double* __restrict__ nrm2, const double* __restrict__ w,
                                                                        double* __restrict__ v, double* __restrict__ z, double* nrm_out,
                                                                        VecList vl, const double* __restrict__ h2, const int* __restrict__ flag,
                                                                        const int* __restrict__ flag2,
                                                                        int push_blocks, int nsend, const int* __restrict__ cidx,
                                                                        const int* __restrict__ dst_slot, const int* __restrict__ peer,
                                                                        P2PSlot* const* peer_ll, unsigned int llflag,
                                                                        int land_blocks, int nrecv, const int* __restrict__ recv_slot,
                                                                        const P2PSlot* my_ll, double* __restrict__ xh) {
    __shared__ double sv[SPP_THREADS];
    __shared__ double hs[MD_MAXV];
    if ((int)blockIdx.x >= (int)gridDim.x - land_blocks) {
        // LANDING blocks (the LAST blocks of the grid): the halo the neighbours' pushing blocks -- the FIRST blocks of their head kernel,
        // which depend on nothing of this rank -- store into my LL buffer is polled here and written to the plain array the SpMV of this
        // step gathers from.  By the time these blocks are scheduled the data have usually been there for tens of microseconds.
        const int lb = blockIdx.x - ((int)gridDim.x - land_blocks);
        for (int q = lb * SPP_CELLS + (int)threadIdx.x / NUN; q < nrecv; q += land_blocks * SPP_CELLS) {
            const size_t e = (size_t)NUN * __ldg(recv_slot + q) + threadIdx.x % NUN;
            double val;
            if (!ll_poll(my_ll + e, llflag, val)) p2p_timeout(my_ll + e, llflag, 0u);
            xh[e] = val;
        }
        return;
    }
    const double nrm = sqrt(*nrm2);
    const double a = 1.0 / nrm;
    const bool upd = flag != nullptr && *flag != 0 && *flag2 == 0;
    const int nv = upd ? vl.nv : 0;
    for (int q = threadIdx.x; q < nv; q += SPP_THREADS) hs[q] = h2[q];
    __syncthreads();
    const int r = threadIdx.x % NUN, lc = threadIdx.x / NUN;
    // w''[t] for one unknown: the update in the order of multi_axpy (q ascending), eight basis loads in flight
    auto updated = [&](size_t t) {
        double wi = w[t];
        for (int c0 = 0; c0 < nv; c0 += MD_CHUNK) {
            double vv[MD_CHUNK];
#pragma unroll
            for (int q = 0; q < MD_CHUNK; q++) vv[q] = (c0 + q < nv) ? vl.v[c0 + q][t] : 0.0;
#pragma unroll
            for (int q = 0; q < MD_CHUNK; q++) if (c0 + q < nv) wi = wi - hs[c0 + q] * vv[q];
        }
        return wi;
    };
    if ((int)blockIdx.x >= push_blocks) {
        const int mb = blockIdx.x - push_blocks, main_blocks = gridDim.x - push_blocks - land_blocks;
        if (mb == 0 && threadIdx.x == 0 && nrm_out) *nrm_out = nrm;
        for (int c0 = mb * SPP_CELLS; c0 < nc; c0 += main_blocks * SPP_CELLS) {
            const int t = c0 * NUN + threadIdx.x;
            const bool ok = t < nc * NUN;
            const double vi = ok ? a * updated((size_t)t) : 0.0;
            sv[threadIdx.x] = vi;
            __syncthreads();
            if (ok) {
                const double* M = minv + (size_t)__ldg(ocell + c0 + lc) * 36 + r * NUN;
                double s = 0.0;
#pragma unroll
                for (int q = 0; q < NUN; q++) s += M[q] * sv[lc * NUN + q];
                v[t] = vi;
                z[t] = s;
            }
            __syncthreads();
        }
    } else {
        for (int q0 = blockIdx.x * SPP_CELLS; q0 < nsend; q0 += push_blocks * SPP_CELLS) {
            const int q = q0 + lc;
            const int ci = q < nsend ? __ldg(cidx + q) : -1;
            sv[threadIdx.x] = ci >= 0 ? a * updated((size_t)NUN * ci + r) : 0.0;
            __syncthreads();
            if (q < nsend) {
                double s = 0.0;
                if (ci >= 0) {
                    const double* M = minv + (size_t)__ldg(ocell + ci) * 36 + r * NUN;
#pragma unroll
                    for (int k = 0; k < NUN; k++) s += M[k] * sv[lc * NUN + k];
                }
                ll_store(peer_ll[__ldg(peer + q)] + (size_t)NUN * __ldg(dst_slot + q) + r, s, llflag);
            }
            __syncthreads();
        }
    }
}
unsigned long long ll_exchange_begin(thcmb_ctx* c);
// returns the sequence number of the exchange it started (0 on one rank); the caller hands it to spmv_compact_rows.
// nv / vecs / d_h2 / d_flag / d_flag2: the pending second update (nullptr flag = none)
unsigned long long scale_precon_push(thcmb_ctx* c, const double* w, const double* d_nrm2, double* v, double* z, double* d_nrm_out,
                                     int nv, double* const* vecs, const double* d_h2, const int* d_flag, const int* d_flag2) {
    const unsigned long long seq = ll_exchange_begin(c);
    const int nc = c->n_ocell;
    if (d_flag && nv > MD_MAXV) fatal("scale_precon_push: the fused second update takes at most 64 basis vectors");
    VecList vl; vl.nv = d_flag ? nv : 0;
    for (int q = 0; q < vl.nv; q++) vl.v[q] = vecs[q];
    const int main_blocks = std::max(1, std::min((nc + SPP_CELLS - 1) / SPP_CELLS, NSM * 8));
    const bool push = c->blk.nranks > 1 && c->nsend_cells > 0;
    const int push_blocks = push ? std::max(1, std::min((c->nsend_cells + SPP_CELLS - 1) / SPP_CELLS, NSM)) : 0;
    const int par = (int)(seq & 1ull);
    // the halo of THIS exchange is landed by the last blocks of the same kernel (the SpMV that follows then needs no polling)
    const bool land = c->blk.nranks > 1 && c->nrecv_cells > 0 && c->d_recv_slot && !getenv("THCM_NO_HALO_LANDING");
    const int land_blocks = land ? std::max(1, std::min((c->nrecv_cells + SPP_CELLS - 1) / SPP_CELLS, NSM)) : 0;
    if (land && !c->d_halo_plain_c) THCM_CUDA(cudaMalloc(&c->d_halo_plain_c, sizeof(double) * (size_t)NUN * std::max(c->blk.nhalo_cells(), 1)));
    ProfScope prof_(c, KID_PRECON_APPLY);
    scale_precon_push_kernel<<<main_blocks + push_blocks + land_blocks, SPP_THREADS, 0, c->stream>>>(
        nc, c->d_ocell, c->d_minv, d_nrm2, w, v, z, d_nrm_out, vl, d_h2, d_flag, d_flag2, push_blocks, push ? c->nsend_cells : 0,
        c->d_send_cidx, c->d_send_dst, c->d_send_peer, push ? (P2PSlot* const*)c->d_peer_ll + (size_t)par * c->peers.size() : nullptr,
        (unsigned int)seq, land_blocks, land ? c->nrecv_cells : 0, c->d_recv_slot, (const P2PSlot*)c->d_halo_ll[par], c->d_halo_plain_c);
    c->launches++;
    c->halo_landed_seq = land ? seq : 0ull;
    return seq;
}

int gather_cells(thcmb_ctx* c, const double* in, double* out) {
    if (in == out) fatal("gather_cells: in-place gather is not supported");
    ProfScope prof_(c, KID_COPY);
    gather_cells_kernel<<<ew_grid(c->n_ocell * NUN), 256, 0, c->stream>>>(c->n_ocell, c->d_ocell, in, out); c->launches++; return 0;
}
int scatter_cells(thcmb_ctx* c, const double* in, double* out) {
    if (in == out) fatal("scatter_cells: in-place scatter is not supported");
    ProfScope prof_(c, KID_COPY);
    scatter_cells_kernel<<<ew_grid(c->blk.ndim()), 256, 0, c->stream>>>(c->blk.ncell(), c->d_ccell, in, out); c->launches++; return 0;
}
// the ocean-only Krylov space needs the LL halo exchange on more than one rank
bool compact_possible(const thcmb_ctx* c) {
    return c->krylov_compact && c->d_colc && (c->blk.nranks == 1 || (c->p2p_on && c->halo_p2p && c->d_peer_ll));
}
double land_nonzero_global(thcmb_ctx* c, const double* x) {
    double* d = c->d_scalars + 4020;
    land_nonzero_kernel<<<RED_BLOCKS, RED_THREADS, 0, c->stream>>>(c->blk.ncell(), c->d_ccell, x, c->d_partial, c->d_counter, d, p2p_args(c));
    c->launches++;
    if (!c->p2p_on) allreduce_dev(c, d, 1);
    double cnt = 0.0;
    THCM_CUDA(cudaMemcpyAsync(&cnt, d, sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    THCM_CUDA(cudaStreamSynchronize(c->stream));
    return cnt;
}
// One exchange of the LL halo = one sequence number; the push may come from halo_push_ll_kernel or from the kernel that produced
// the vector (scale_precon_push_kernel), the consumer is always the compact SpMV of the same sequence number.
unsigned long long ll_exchange_begin(thcmb_ctx* c) {
    if (c->blk.nranks <= 1) return 0ull;
    if ((int)c->peers.size() > HALO_MAX_PEERS) fatal("halo push: more than 16 neighbours");
    return ++c->halo_ll_seq;
}
int spmv_compact_rows(thcmb_ctx* c, const double* xc, double* yc, unsigned long long seq);
int spmv_compact(thcmb_ctx* c, const double* xc, double* yc) {
    const unsigned long long seq = ll_exchange_begin(c);
    if (c->blk.nranks > 1 && c->nsend_cells > 0) {
        const int par = (int)(seq & 1ull);
        ProfScope prof_(c, KID_HALO_PACK);
        const int pgrid = std::max(1, std::min(ew_grid(c->nsend_cells * NUN), NSM));
        halo_push_ll_kernel<<<pgrid, 256, 0, c->stream>>>(c->nsend_cells, c->d_send_cidx, c->d_send_dst, c->d_send_peer, xc,
                                                          (P2PSlot* const*)c->d_peer_ll + (size_t)par * c->peers.size(), (unsigned int)seq);
        c->launches++;
    }
    return spmv_compact_rows(c, xc, yc, seq);
}
int spmv_compact_rows(thcmb_ctx* c, const double* xc, double* yc, unsigned long long seq) {
    const int nrow_c = c->n_ocell * NUN, rows_per_block = SPMV_THREADS / 4;
    const int grid = (int)std::max<long long>(1, std::min<long long>(((long long)nrow_c + rows_per_block - 1) / rows_per_block, (long long)NSM * 64));
    if (c->blk.nranks > 1) {
        const int par = (int)(seq & 1ull);
        ProfScope prof_(c, KID_SPMV);
        if (seq != 0ull && c->halo_landed_seq == seq)   // the head kernel of this step landed the exchange
            spmv_compact_kernel<4, 6, 2><<<grid, SPMV_THREADS, 0, c->stream>>>(nrow_c, c->d_ocell, c->d_rowptr, c->d_colc, c->d_val, xc, nrow_c,
                                                                               nullptr, c->d_halo_plain_c, 0u, yc);
        else
            spmv_compact_kernel<4, 6, 1><<<grid, SPMV_THREADS, 0, c->stream>>>(nrow_c, c->d_ocell, c->d_rowptr, c->d_colc, c->d_val, xc, nrow_c,
                                                                               (const P2PSlot*)c->d_halo_ll[par], nullptr, (unsigned int)seq, yc);
    } else {
        ProfScope prof_(c, KID_SPMV);
        spmv_compact_kernel<4, 6, 0><<<grid, SPMV_THREADS, 0, c->stream>>>(nrow_c, c->d_ocell, c->d_rowptr, c->d_colc, c->d_val, xc, nrow_c,
                                                                           nullptr, nullptr, 0u, yc);
    }
    c->launches++;
    if (c->ic_on) {   // the dense integral-condition row (THCM.C:2180-2229) on the compact vectors
        dot_dev(c, nrow_c, c->d_iccoeff_c, xc, c->d_scalars + 4010);
        const int crow = c->ic_lrow >= 0 ? NUN * c->ccell_host[(size_t)(c->ic_lrow / NUN)] + c->ic_lrow % NUN : -1;
        fix_spmv_rows_kernel<<<1, 32, 0, c->stream>>>(FixRows{crow, -1, -1}, (double)c->ic_sign, c->d_scalars + 4010, xc, yc);
        c->launches++;
    }
    return 0;
}
int apply_blockdiag_compact(thcmb_ctx* c, const double* x, double* y) {
    ProfScope prof_(c, KID_PRECON_APPLY);
    const int n = c->n_ocell * NUN;
    blockdiag_apply_compact_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->n_ocell, c->d_ocell, c->d_minv, x, y);
    c->launches++;
    return 0;
}

int apply_blockdiag(thcmb_ctx* c, const double* x, double* y) {
    int n = c->blk.ndim();
    ProfScope prof_(c, KID_PRECON_APPLY);
    blockdiag_apply_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(c->blk.ncell(), c->d_minv, x, y);
    c->launches++;
    return 0;
}

}  // namespace thcm
