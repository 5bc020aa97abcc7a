// Dense kernels for the big fronts of the supernodal Cholesky / LDL' (sm_100a):
//
//   * an FP64 tensor-core tile engine: DMMA (mma.sync m8n8k4 f64) fed by 1-D TMA bulk copies
//     (cp.async.bulk global->shared, mbarrier completion) of panel columns through a 4-stage
//     shared-memory ring, warp-specialised (one producer warp, no block barrier in the K loop); two
//     tile shapes, 128 x 128 (one CTA per SM) and 64 x 128 (two CTAs per SM, the default); LDL' mode
//     scales the A fragment by the pivots on its way into the DMMAs;
//   * the blocked right-looking factorisation of a front's panel built on it, 128-column blocks:
//     diagonal block (Cholesky or LDL') plus its inverse in one CTA, TRSM as a GEMM with that inverse,
//     panel updates in a recursive (binary) schedule issued with deep look-ahead: every update is cut
//     into pieces by the step at which its columns are next touched, one prioritised stream per piece
//     class, the latency chain on the highest-priority stream (launch_wide_chol_level);
//   * the update block of every medium / big front, formed once from an exact tile list, children
//     merged in shared memory through the tile-cut table of the symbolic analysis (front_cb_kernel);
//   * the inverse of every big supernode's pivot block L11 in 2048-column diagonal blocks (recursive
//     block merge, two batched GEMMs per level, exact work lists), so that the triangular solves of
//     big supernodes become bandwidth-bound matrix-vector products spread over many CTAs instead of a
//     substitution chain inside one CTA;
//   * those multi-CTA forward / backward solve kernels.
//
// Replaces CHOLMOD's supernodal numeric factorisation and solve behind
// cholesky(Symmetric(Q,:L)) / ldlt(...) and F \ rhs (linear_system_solvers/julia.jl:34,52,99-113).
#include <algorithm>

#include "opb_internal.h"

namespace opb {

namespace {

// ---------------------------------------------------------------------------
// PTX helpers
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = smem_u32(bar);
    unsigned ok;
    do {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!ok);
}
// 1-D TMA bulk copy global -> shared, completion counted in bytes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Programmatic dependent launch (the solve chain of the big supernodes is ~250 short kernels per pair, each
// waiting for the one before): a kernel lets its successor start launching right away and itself waits for
// its predecessor only where it first touches data the predecessor wrote, so launch latency and the
// index preamble overlap the tail of the previous kernel.  Used with launch_pdl() below.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_release() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__device__ __forceinline__ bool stop_requested(const DeltaState* st) {
    const volatile int* d = &st->done;
    const volatile int* f = &st->fail;
    return (*d) | (*f);
}

// ---------------------------------------------------------------------------
// Tile engine:  acc(BM x 128) = sum_k A[m,k] * B[n,k],  BM = 64 * WM
//   A: m-contiguous (column-major M x K, element (m,k) at A[m + k*lda])
//   B: m-contiguous (B_KC = false, element (n,k) at B[n + k*ldb]) or
//      k-contiguous (B_KC = true,  element (n,k) at B[k + n*ldb])
// Requirements: A, B 16-byte aligned, lda/ldb even, K origin a multiple of 2.
// WM = 2: 128 x 128 tile, 8 compute warps (2 x 4), one CTA per SM -- the bulk kernels.
// WM = 1:  64 x 128 tile, 4 compute warps (1 x 4), TWO CTAs per SM (registers and shared memory
//          both allow it): the prologue / epilogue of one tile overlaps the K loop of the other and
//          the wave granularity halves -- the short-K and narrow kernels (TRSM, update of a single
//          block column, update blocks of the low levels), where a tile is mostly fixed cost.
// Warp tile 64 x 32 either way, thread accumulators 8 x 4 x 2.
// ---------------------------------------------------------------------------
constexpr int BN = 128, BK = 16, STAGES = 4;
constexpr int LDN = BN + 4;   // [BK][LDN]: fragment reads are bank-conflict free (LDN % 16 == 4)
constexpr int LDK = BK + 4;   // [BN][LDK]
template <int WM>
struct TileCfg {
    static constexpr int BM = 64 * WM;
    static constexpr int LDM = BM + 4;                    // % 16 == 4 as well
    static constexpr int A_STAGE = BK * LDM;
    // WM = 1 is used with m-contiguous B only
    static constexpr int B_STAGE = (WM == 2 && BN * LDK > BK * LDN) ? BN * LDK : BK * LDN;
    static constexpr int CWARPS = 4 * WM;                 // compute warps
    static constexpr int THREADS = (CWARPS + 1) * 32;     // + one producer warp that only issues the TMA copies
};
constexpr int BM = TileCfg<2>::BM;
static_assert(BM == 2 * CB_TILE && BN == 2 * CB_TILE, "tile-cut table of symbolic.cpp");
constexpr int GEMM_CWARPS = TileCfg<2>::CWARPS;
constexpr int GEMM_THREADS = TileCfg<2>::THREADS;

template <int WM>
struct GemmSmemT {
    double A[STAGES][TileCfg<WM>::A_STAGE];
    double B[STAGES][TileCfg<WM>::B_STAGE];
    unsigned long long full[STAGES];    // producer -> compute warps: the stage has landed (tx bytes)
    unsigned long long empty[STAGES];   // compute warps -> producer: the stage has been read
};
using GemmSmem = GemmSmemT<2>;

// true for the threads that hold accumulators after gemm_mainloop
template <int WM = 2>
__device__ __forceinline__ bool gemm_compute_warp() { return threadIdx.x < TileCfg<WM>::CWARPS * 32; }

template <int WM, bool B_KC>
__device__ __forceinline__ void gemm_issue(GemmSmemT<WM>& sm, int stage, int kb, const double* Ag, int lda, int mrows,
                                           const double* Bg, int ldb, int nrows, int K, int lane) {
    constexpr int LDM = TileCfg<WM>::LDM;
    const int k0 = kb * BK;
    const int nk = min(BK, K - k0);
    const unsigned bytesA = (unsigned)(((mrows + 1) & ~1) * 8);
    unsigned total;
    if (!B_KC) total = (unsigned)nk * (bytesA + (unsigned)(((nrows + 1) & ~1) * 8));
    else total = (unsigned)nk * bytesA + (unsigned)nrows * (unsigned)(((nk + 1) & ~1) * 8);
    if (lane == 0) mbar_expect_tx(&sm.full[stage], total);
    __syncwarp();
    for (int kk = lane; kk < nk; kk += 32)
        bulk_g2s(&sm.A[stage][kk * LDM], Ag + (size_t)(k0 + kk) * lda, bytesA, &sm.full[stage]);
    if (!B_KC) {
        const unsigned bytesB = (unsigned)(((nrows + 1) & ~1) * 8);
        for (int kk = lane; kk < nk; kk += 32)
            bulk_g2s(&sm.B[stage][kk * LDN], Bg + (size_t)(k0 + kk) * ldb, bytesB, &sm.full[stage]);
    } else {
        const unsigned bytesB = (unsigned)(((nk + 1) & ~1) * 8);
        for (int n = lane; n < nrows; n += 32)
            bulk_g2s(&sm.B[stage][n * LDK], Bg + (size_t)n * ldb + k0, bytesB, &sm.full[stage]);
    }
}

// one k-step (4 columns) of the warp tile: 12 fragment loads, 32 DMMAs
// SCALE (LDL' mode): the A fragment of column k is multiplied by d_k on its way into the DMMAs, so that
// the product is (L D) L' without a scaled copy of the panel (8 extra multiplies per 32 DMMAs)
template <int WM, bool B_KC, bool TAIL, bool SCALE = false>
__device__ __forceinline__ void gemm_kstep(const double* As, const double* Bs, int kk, bool kvalid,
                                           double (&acc)[8][4][2], double dkv = 1.0) {
    constexpr int LDM = TileCfg<WM>::LDM;
    double av[8], bv[4];
#pragma unroll
    for (int mt = 0; mt < 8; mt++) {
        const double v = SCALE ? As[kk * LDM + mt * 8] * dkv : As[kk * LDM + mt * 8];
        av[mt] = (!TAIL || kvalid) ? v : 0.0;
    }
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
        const double v = B_KC ? Bs[(nt * 8) * LDK + kk] : Bs[kk * LDN + nt * 8];
        bv[nt] = (!TAIL || kvalid) ? v : 0.0;
    }
#pragma unroll
    for (int mt = 0; mt < 8; mt++)
#pragma unroll
        for (int nt = 0; nt < 4; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], av[mt], bv[nt]);
}

// Warp-specialised main loop: the last warp streams the operands (TMA bulk copies, `full`
// barriers), the compute warps run the DMMAs and hand the stage back through the `empty` barriers; no
// block-wide barrier inside the K loop.  On return every stage has been consumed (block barrier),
// so the caller may reuse the shared memory; acc is valid on the compute warps only.
struct NoHook { __device__ __forceinline__ void operator()() const {} };
// stages before the end of the K loop at which the compute warps run `hook` (front_cb_kernel:
// L2 prefetch of the children's update-block entries the epilogue is about to merge)
constexpr int HOOK_LEAD = 8;

template <int WM, bool B_KC, class Hook = NoHook, bool SCALE = false>
__device__ __forceinline__ void gemm_mainloop(GemmSmemT<WM>& sm, const double* Ag, int lda, int mrows,
                                              const double* Bg, int ldb, int nrows, int K,
                                              double (&acc)[8][4][2], Hook hook = Hook(),
                                              const double* __restrict__ dk = nullptr) {
    constexpr int CW = TileCfg<WM>::CWARPS;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) { mbar_init(&sm.full[s], 1); mbar_init(&sm.empty[s], CW); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nkb = (K + BK - 1) / BK;
    if (warp == CW) {
        for (int kb = 0; kb < nkb; kb++) {
            const int stage = kb % STAGES;
            if (kb >= STAGES) mbar_wait(&sm.empty[stage], (unsigned)(((kb / STAGES) - 1) & 1));
            gemm_issue<WM, B_KC>(sm, stage, kb, Ag, lda, mrows, Bg, ldb, nrows, K, lane);
        }
    } else {
        const int wm = warp >> 2, wn = warp & 3;     // WM x 4 warps, warp tile 64 x 32
        const int q = lane & 3, g = lane >> 2;
        const int aoff = wm * 64 + g;
        const int boff = B_KC ? (wn * 32 + g) * LDK : wn * 32 + g;
        const int hook_at = nkb > HOOK_LEAD ? nkb - HOOK_LEAD : 0;
        for (int kb = 0; kb < nkb; kb++) {
            if (kb == hook_at) hook();
            const int stage = kb % STAGES;
            mbar_wait(&sm.full[stage], (unsigned)((kb / STAGES) & 1));
            const double* As = sm.A[stage] + aoff;
            const double* Bs = sm.B[stage] + boff;
            const int nk = K - kb * BK;
            if (nk >= BK) {
                double dv[BK / 4];
#pragma unroll
                for (int ks = 0; ks < BK / 4; ks++) dv[ks] = SCALE ? dk[kb * BK + ks * 4 + q] : 1.0;
#pragma unroll
                for (int ks = 0; ks < BK / 4; ks++) gemm_kstep<WM, B_KC, false, SCALE>(As, Bs, ks * 4 + q, true, acc, dv[ks]);
            } else {
#pragma unroll 1
                for (int ks = 0; ks * 4 < nk; ks++) {
                    const bool kv = ks * 4 + q < nk;
                    gemm_kstep<WM, B_KC, true, SCALE>(As, Bs, ks * 4 + q, kv, acc, (SCALE && kv) ? dk[kb * BK + ks * 4 + q] : 1.0);
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[stage]);
        }
    }
    __syncthreads();
}

// coordinates of accumulator element (mt, nt, e) inside the tile (warp rows 0 .. WM-1: threadIdx.x >> 7)
__device__ __forceinline__ int acc_row(int mt) { return ((threadIdx.x >> 5) >> 2) * 64 + mt * 8 + ((threadIdx.x & 31) >> 2); }
__device__ __forceinline__ int acc_col(int nt, int e) { return ((threadIdx.x >> 5) & 3) * 32 + nt * 8 + 2 * (threadIdx.x & 3) + e; }

// ---------------------------------------------------------------------------
// Front descriptors
// ---------------------------------------------------------------------------
struct Front {
    int s, first, c, r, N, ld, ldx;
    int64_t loff, cboff, xoff;
};
__device__ __forceinline__ Front get_front(const DevSym& S, int s) {
    Front d;
    d.s = s;
    d.first = S.sfirst[s];
    d.c = S.sfirst[s + 1] - d.first;
    d.r = (int)(S.rowptr[s + 1] - S.rowptr[s]);
    d.N = d.c + d.r;
    d.ld = ld_of(d.N);
    d.ldx = ld_of(d.c);
    d.loff = S.Loff[s];
    d.cboff = S.CBoff[s];
    d.xoff = S.Xoff[s];
    return d;
}
__device__ __forceinline__ double* front_elem(const Front& d, double* Lval, double* CB, int i, int j) {
    return (j < d.c) ? (Lval + d.loff + i + (size_t)j * d.ld)
                     : (CB + d.cboff + (i - d.c) + (size_t)(j - d.c) * d.r);
}

// ---------------------------------------------------------------------------
// Cholesky of a shared-memory panel (n rows, c <= 128 pivot columns) in 32-column
// sub-blocks: the diagonal 32 x 32 block is factorised and inverted by one warp
// (warp-synchronous, no block barrier), the rows below are multiplied by that
// inverse, the remaining panel columns get the rank-32 update.  When Xs != nullptr
// the inverse of the whole c x c pivot block is assembled in shared memory as
// 32 x 32 blocks (block (I,J), I >= J, at Xs + (I(I+1)/2 + J) * INVBUF, ld 33): the
// big fronts need it for the TRSM-as-GEMM and for the multi-CTA triangular solves.
// Loops are kept rolled on purpose: the code runs once per block, so its size (not
// its issue rate) is what the instruction cache sees.
// ---------------------------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;
constexpr int INVLD = 33;
constexpr int INVBUF = 32 * INVLD;           // doubles

// Cholesky of the w x w (w <= 32) block at (j0, j0) of the shared-memory panel P, in place,
// fused with the inverse of the factor (left in invbuf, ld 33).  Columns are taken eight at a
// time: warp 0 factorises the 8-column panel in registers (lane = row, shuffles, no barrier)
// and finalises the matching eight rows of X (lane = column); then the whole CTA applies the
// rank-8 update to the trailing columns of the block and to the later rows of X.  Two block
// barriers per eight columns; the register code is 8 x 8 unrolled, i.e. small.
// LDLT = true: unit lower L with D on the diagonal (julia.jl:47-90: ldlt without pivoting; a zero or NaN
// pivot fails the attempt), X = inverse of the unit lower factor.
constexpr int PW = 8;
template <bool LDLT = false>
__device__ bool cta_potrf32_inv(double* P, int ld, int j0, int w, double* invbuf, int* s_fail) {
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int i = tid & 31, q = tid >> 5, nq = nthr >> 5;
    double* B = P + j0 + (size_t)j0 * ld;
    for (int e = tid; e < INVBUF; e += nthr) invbuf[e] = ((e % INVLD) == (e / INVLD)) ? 1.0 : 0.0;
    if (tid == 0) *s_fail = 0;
    __syncthreads();
    const bool rv = i < w;
    for (int jp = 0; jp < w; jp += PW) {
        if (q == 0) {
            double p[PW], rs[PW];
#pragma unroll
            for (int t = 0; t < PW; t++) {
                const int j = jp + t;
                p[t] = (rv && j < w && i >= j) ? B[i + (size_t)j * ld] : ((i == j) ? 1.0 : 0.0);
            }
            bool ok = true;
#pragma unroll
            for (int t = 0; t < PW; t++) {
                const int j = jp + t;
                const double d = __shfl_sync(FULL, p[t], j & 31);
                if (LDLT) {
                    if (j < w && (d == 0.0 || d != d)) ok = false;     // zero or NaN pivot
                    rs[t] = 1.0;                                      // unit diagonal
                    // lane j keeps d (stored on the diagonal), the lanes below it hold L[i,j] = a / d
                    const double lt = (i > j) ? p[t] / d : ((i == j) ? d : 0.0);
                    p[t] = lt;
#pragma unroll
                    for (int u = t + 1; u < PW; u++) {
                        const double lu = __shfl_sync(FULL, lt, (jp + u) & 31);      // L[jp+u, j]
                        if (i > j) p[u] -= lt * (d * lu);
                    }
                } else {
                    if (!(d > 0.0)) ok = false;          // pivot <= 0 or NaN (julia.jl:39-41)
                    rs[t] = rsqrt(d);                    // = 1 / L[j,j]
                    const double lt = (i >= j) ? p[t] * rs[t] : 0.0;     // lane j: d * rsqrt(d) = sqrt(d)
                    p[t] = lt;
#pragma unroll
                    for (int u = t + 1; u < PW; u++) {
                        const double lu = __shfl_sync(FULL, lt, (jp + u) & 31);
                        p[u] -= lt * lu;                 // rows above jp+u hold junk that is never stored
                    }
                }
            }
            if (!ok) { if (i == 0) *s_fail = 1; }
            else {
#pragma unroll
                for (int t = 0; t < PW; t++) {
                    const int j = jp + t;
                    if (rv && j < w && i >= j) B[i + (size_t)j * ld] = p[t];
                }
                // rows jp .. jp+7 of X (lane = column): scale by 1/L[j,j], eliminate from the later panel rows
                double xr[PW];
#pragma unroll
                for (int t = 0; t < PW; t++) xr[t] = invbuf[((jp + t) & 31) + i * INVLD];
#pragma unroll
                for (int t = 0; t < PW; t++) {
                    xr[t] *= rs[t];
#pragma unroll
                    for (int u = t + 1; u < PW; u++) {
                        const double lut = __shfl_sync(FULL, p[t], (jp + u) & 31);     // L[jp+u, jp+t]
                        xr[u] -= lut * xr[t];
                    }
                }
#pragma unroll
                for (int t = 0; t < PW; t++)
                    if (jp + t < w) invbuf[(jp + t) + i * INVLD] = xr[t];
            }
        }
        __syncthreads();
        if (*s_fail) return false;
        // rank-8 update of the trailing columns of L and of the later rows of X
        const int i1 = jp + PW;
        if (rv && i >= i1) {
            double li[PW];
#pragma unroll
            for (int t = 0; t < PW; t++) li[t] = B[i + (size_t)(jp + t) * ld];
            if (LDLT) {
#pragma unroll
                for (int t = 0; t < PW; t++) li[t] *= B[(jp + t) + (size_t)(jp + t) * ld];     // (L D)[i, jp+t]
            }
            for (int k = i1 + q; k <= i; k += nq) {
                double acc = 0.0;
#pragma unroll
                for (int t = 0; t < PW; t++) acc += li[t] * B[k + (size_t)(jp + t) * ld];
                B[i + (size_t)k * ld] -= acc;
            }
            if (LDLT) {
#pragma unroll
                for (int t = 0; t < PW; t++) li[t] = B[i + (size_t)(jp + t) * ld];
            }
            for (int jc = q; jc < i1; jc += nq) {
                double acc = 0.0;
#pragma unroll
                for (int t = 0; t < PW; t++) acc += li[t] * invbuf[(jp + t) + jc * INVLD];
                invbuf[i + jc * INVLD] -= acc;
            }
        }
        __syncthreads();
    }
    return true;
}

__device__ __forceinline__ double* xs_block(double* Xs, int I, int J) { return Xs + (I * (I + 1) / 2 + J) * INVBUF; }

// Warp-level FP64 tensor-core product on shared-memory operands (DMMA m8n8k4), 16 x 16 tile:
//   acc += A[m0.., 0..K) * B^T, A element (m,k) at A[m + k*lda]; B element (n,k) at
//   B[n + k*ldb] (B_KC = false) or B[k + n*ldb] (B_KC = true).  Rows >= M / N read as zero.
// dk != nullptr (LDL'): column k of A is scaled by dk[k * dstride]
template <bool B_KC>
__device__ __forceinline__ void warp_tile16(double (&acc)[2][2][2], const double* A, int lda, int m0, int M,
                                            const double* B, int ldb, int n0, int N, int K,
                                            const double* dk = nullptr, int dstride = 0) {
    const int lane = threadIdx.x & 31, q = lane & 3, g = lane >> 2;
    const bool ma = m0 + g < M, mb = m0 + 8 + g < M, na = n0 + g < N, nb = n0 + 8 + g < N;
    for (int k0 = 0; k0 < K; k0 += 4) {
        const int kk = k0 + q;
        const bool kv = kk < K;
        const double sc = (dk && kv) ? dk[(size_t)kk * dstride] : 1.0;
        const double a0 = (kv && ma) ? A[(m0 + g) + (size_t)kk * lda] * sc : 0.0;
        const double a1 = (kv && mb) ? A[(m0 + 8 + g) + (size_t)kk * lda] * sc : 0.0;
        double b0, b1;
        if (B_KC) {
            b0 = (kv && na) ? B[kk + (size_t)(n0 + g) * ldb] : 0.0;
            b1 = (kv && nb) ? B[kk + (size_t)(n0 + 8 + g) * ldb] : 0.0;
        } else {
            b0 = (kv && na) ? B[(n0 + g) + (size_t)kk * ldb] : 0.0;
            b1 = (kv && nb) ? B[(n0 + 8 + g) + (size_t)kk * ldb] : 0.0;
        }
        dmma884(acc[0][0][0], acc[0][0][1], a0, b0);
        dmma884(acc[0][1][0], acc[0][1][1], a0, b1);
        dmma884(acc[1][0][0], acc[1][0][1], a1, b0);
        dmma884(acc[1][1][0], acc[1][1][1], a1, b1);
    }
}
// MODE 0: C = acc, 1: C += acc, 2: C -= acc, 3: C = -acc; cs != nullptr: column n divided by cs[n * cstride]
template <int MODE>
__device__ __forceinline__ void warp_tile16_store(const double (&acc)[2][2][2], double* C, int ldc, int m0, int M,
                                                  int n0, int N, const double* cs = nullptr, int cstride = 0) {
    const int lane = threadIdx.x & 31, q = lane & 3, g = lane >> 2;
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int m = m0 + 8 * a + g, n = n0 + 8 * b + 2 * q + e;
                if (m < M && n < N) {
                    double* p = C + m + (size_t)n * ldc;
                    const double v = cs ? acc[a][b][e] / cs[(size_t)n * cstride] : acc[a][b][e];
                    if (MODE == 0) *p = v; else if (MODE == 1) *p += v; else if (MODE == 2) *p -= v; else *p = -v;
                }
            }
}

// P: shared-memory panel, column-major (ld), n rows, c pivot columns (top c x c = pivot block).
// invbuf: INVBUF doubles; s_fail: one shared int.
template <bool PROF = false, bool LDLT = false>
__device__ bool panel_chol_smem(double* P, int ld, int n, int c, double* invbuf, double* Xs, int* s_fail,
                                long long* stamps = nullptr) {
    int ns = 0;
#define OPB_STAMP() do { if (PROF && threadIdx.x == 0) stamps[ns++] = clock64(); } while (0)
    const int tid = threadIdx.x, nthr = blockDim.x, warp = tid >> 5, nwarp = nthr >> 5;
    for (int j0 = 0; j0 < c; j0 += 32) {
        const int w = min(32, c - j0);
        OPB_STAMP();
        if (!cta_potrf32_inv<LDLT>(P, ld, j0, w, invbuf, s_fail)) return false;
        if (Xs) {
            double* xb = xs_block(Xs, j0 >> 5, j0 >> 5);
            for (int e = tid; e < INVBUF; e += nthr) xb[e] = invbuf[e];
        }
        OPB_STAMP();
        const int i1 = j0 + w;
        // rows below the diagonal block: L = A * inv(Ld)^T, in place; a warp owns 16 rows and all
        // w columns, so it has read its rows completely before it overwrites them
        {
            const int nrt = (n - i1 + 15) / 16;
            for (int it = warp; it < nrt; it += nwarp) {
                const int m0 = i1 + 16 * it;
                double acc0[2][2][2] = {}, acc1[2][2][2] = {};
                warp_tile16<false>(acc0, P + (size_t)j0 * ld, ld, m0, n, invbuf, INVLD, 0, w, w);
                if (w > 16) warp_tile16<false>(acc1, P + (size_t)j0 * ld, ld, m0, n, invbuf, INVLD, 16, w, w);
                __syncwarp();
                // LDL': L = A inv(Ld)' D^-1  (the pivots sit on the diagonal of the block)
                const double* dsc = LDLT ? P + j0 + (size_t)j0 * ld : nullptr;
                warp_tile16_store<0>(acc0, P + (size_t)j0 * ld, ld, m0, n, 0, w, dsc, ld + 1);
                if (w > 16) warp_tile16_store<0>(acc1, P + (size_t)j0 * ld, ld, m0, n, 16, w, dsc, ld + 1);
            }
        }
        __syncthreads();
        OPB_STAMP();
        // rank-w update of the remaining panel columns [i1, c), rows [i1, n), lower tiles only
        if (i1 < c) {
            const int ntm = (n - i1 + 15) / 16, ntn = (c - i1 + 15) / 16;
            const double* Ablk = P + (size_t)j0 * ld;
            for (int it = warp; it < ntm * ntn; it += nwarp) {
                const int mt = it % ntm, nt = it / ntm;
                if (mt < nt) continue;
                double acc[2][2][2] = {};
                warp_tile16<false>(acc, Ablk, ld, i1 + 16 * mt, n, Ablk, ld, i1 + 16 * nt, c, w,
                                   LDLT ? P + j0 + (size_t)j0 * ld : nullptr, ld + 1);
                warp_tile16_store<2>(acc, P, ld, i1 + 16 * mt, n, i1 + 16 * nt, c);
            }
        }
        __syncthreads();
    }
    OPB_STAMP();
    if (!Xs) return true;
    // off-diagonal blocks of X = inv(L11) by block forward substitution:
    //   T = sum_{K=J}^{I-1} L_IK X_KJ ;  X_IJ = - X_II T        (32 x 32 blocks, 4 warp tiles each)
    const int nb = (c + 31) >> 5;
    for (int dist = 1; dist < nb; dist++) {
        const int npair = nb - dist;
        // phase 1: T of pair p kept in the destination block X_IJ
        for (int it = warp; it < npair * 4; it += nwarp) {
            const int pr = it >> 2, mt = it & 1, nt = (it >> 1) & 1;
            const int J = pr, I = pr + dist;
            double acc[2][2][2] = {};
            for (int K = J; K < I; K++)
                warp_tile16<true>(acc, P + (size_t)(32 * K) * ld, ld, 32 * I + 16 * mt, c,
                                  xs_block(Xs, K, J), INVLD, 16 * nt, 32, 32);
            warp_tile16_store<0>(acc, xs_block(Xs, I, J), INVLD, 16 * mt, 32, 16 * nt, 32);
        }
        __syncthreads();
        // phase 2: X_IJ = - X_II * T, computed into registers by the warp that owns the whole
        // 32 x 16 half block column, then written over T
        for (int it = warp; it < npair * 2; it += nwarp) {
            const int pr = it >> 1, nt = it & 1;
            const int J = pr, I = pr + dist;
            double acc0[2][2][2] = {}, acc1[2][2][2] = {};
            warp_tile16<true>(acc0, xs_block(Xs, I, I), INVLD, 0, 32, xs_block(Xs, I, J), INVLD, 16 * nt, 32, 32);
            warp_tile16<true>(acc1, xs_block(Xs, I, I), INVLD, 16, 32, xs_block(Xs, I, J), INVLD, 16 * nt, 32, 32);
            __syncwarp();
            warp_tile16_store<3>(acc0, xs_block(Xs, I, J), INVLD, 0, 32, 16 * nt, 32);
            warp_tile16_store<3>(acc1, xs_block(Xs, I, J), INVLD, 16, 32, 16 * nt, 32);
        }
        __syncthreads();
    }
    OPB_STAMP();
#undef OPB_STAMP
    return true;
}

constexpr int PT = 512;              // threads of the panel kernels
constexpr int LDD = WB + 4;          // 132: fragment loads are bank-conflict free (LDD % 16 == 4)
constexpr int XS_BLOCKS = 10;        // lower 32 x 32 blocks of a 128 x 128 triangle

// diagonal block of outer step t of a big front: Cholesky + inverse
__global__ void __launch_bounds__(PT)
chol_diag_kernel(DevSym S, const int* __restrict__ list, double* __restrict__ Lval, double* __restrict__ Xinv,
                 int t, int ldlt, DeltaState* st) {
    extern __shared__ double D[];          // LDD * WB + INVBUF + XS_BLOCKS * INVBUF
    __shared__ int s_fail;
    if (stop_requested(st)) return;
    const Front d = get_front(S, list[blockIdx.x]);
    const int j0 = t * WB;
    if (j0 >= d.c) return;
    const int b = min(WB, d.c - j0);
    const int tid = threadIdx.x;
    double* invbuf = D + LDD * WB;
    double* Xs = invbuf + INVBUF;
    double* base = Lval + d.loff + j0 + (size_t)j0 * d.ld;
    for (int idx = tid; idx < b * b; idx += PT) {
        const int i = idx % b, j = idx / b;
        D[i + j * LDD] = (i >= j) ? base[i + (size_t)j * d.ld] : 0.0;
    }
    __syncthreads();
    const bool ok = ldlt ? panel_chol_smem<false, true>(D, LDD, b, b, invbuf, Xs, &s_fail)
                         : panel_chol_smem<false, false>(D, LDD, b, b, invbuf, Xs, &s_fail);
    if (!ok) {
        if (tid == 0) st->fail = 1;
        return;
    }
    double* X = Xinv + d.xoff + j0 + (size_t)j0 * d.ldx;
    for (int idx = tid; idx < b * b; idx += PT) {
        const int i = idx % b, j = idx / b;
        if (i >= j) {
            base[i + (size_t)j * d.ld] = D[i + j * LDD];
            X[i + (size_t)j * d.ldx] = xs_block(Xs, i >> 5, j >> 5)[(i & 31) + (j & 31) * INVLD];
        }
    }
    // LDL': the pivots as a contiguous vector (the tile engine scales its A operand with it)
    if (ldlt && tid < b) S.dvec[d.first + j0 + tid] = D[tid + tid * LDD];
}

// medium fronts: the whole N x c panel lives in shared memory.  Children's update blocks are
// added into the panel columns (ascending child order) and the panel is factorised; the update
// block of the front is produced later by front_cb_kernel.
__global__ void __launch_bounds__(PT)
mid_panel_kernel(DevSym S, const int* __restrict__ list, double* __restrict__ Lval,
                 const double* __restrict__ CB, int ldlt, DeltaState* st) {
    extern __shared__ double P[];          // N * c + INVBUF
    __shared__ int s_fail;
    if (stop_requested(st)) return;
    const Front d = get_front(S, list[blockIdx.x]);
    const int N = d.N, c = d.c, tid = threadIdx.x;
    double* invbuf = P + (size_t)N * c;
    double* panel = Lval + d.loff;
    for (int idx = tid; idx < N * c; idx += PT) {
        const int i = idx % N, j = idx / N;
        P[idx] = panel[i + (size_t)j * d.ld];
    }
    __syncthreads();
    for (int k = S.child_ptr[d.s]; k < S.child_ptr[d.s + 1]; k++) {
        const int ch = S.child_list[k];
        const int64_t rp = S.rowptr[ch];
        const int rc = (int)(S.rowptr[ch + 1] - rp);
        const int* __restrict__ relc = S.rel + rp;
        const double* __restrict__ cb = child_cb(S, CB, ch);
        int lo = 0, hi = rc;               // uc = number of child rows that land on pivot columns
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (relc[mid] < c) lo = mid + 1; else hi = mid; }
        const int uc = lo;
        for (int idx = tid; idx < uc * rc; idx += PT) {
            const int tt = idx % rc, u = idx / rc;
            if (tt >= u) P[relc[tt] + (size_t)relc[u] * N] += cb[tt + (size_t)u * rc];
        }
        __syncthreads();
    }
    const bool ok = ldlt ? panel_chol_smem<false, true>(P, N, N, c, invbuf, nullptr, &s_fail)
                         : panel_chol_smem<false, false>(P, N, N, c, invbuf, nullptr, &s_fail);
    if (!ok) {
        if (tid == 0) st->fail = 1;
        return;
    }
    for (int idx = tid; idx < N * c; idx += PT) {
        const int i = idx % N, j = idx / N;
        if (i >= j) panel[i + (size_t)j * d.ld] = P[idx];      // the strictly upper part holds scratch
    }
    if (ldlt) for (int j = tid; j < c; j += PT) S.dvec[d.first + j] = P[j + (size_t)j * N];
}

// big fronts: add the children's update blocks into the panel columns only.  A CTA owns EAP_RB
// consecutive front rows; a warp walks the child's columns for 32 of the child's rows (coalesced along
// the rows on both sides) with EAP_Q independent read-modify-writes in flight per lane -- the kernel is
// a chain of DRAM round trips, so its speed is the number of them in flight.
// Two shapes: 256 threads x 4 in flight for levels with many fronts (small CTAs start faster), 512 x 8 for the
// few huge fronts at the top of the tree (root of C5 100^3: 3.3 -> 2.0 ms).
constexpr int EAP_RB = 32;
template <int EAP_T, int EAP_Q>
__global__ void __launch_bounds__(EAP_T)
big_extend_add_panel_kernel(DevSym S, const int* __restrict__ list, double* __restrict__ Lval,
                            const double* __restrict__ CB, DeltaState* st) {
    constexpr int EAP_W = EAP_T / 32;
    if (stop_requested(st)) return;
    const Front d = get_front(S, list[blockIdx.y]);
    const int row0 = blockIdx.x * EAP_RB;
    if (row0 >= d.N) return;
    const int row1 = min(d.N, row0 + EAP_RB);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int k = S.child_ptr[d.s]; k < S.child_ptr[d.s + 1]; k++) {
        const int ch = S.child_list[k];
        const int64_t rp = S.rowptr[ch];
        const int rc = (int)(S.rowptr[ch + 1] - rp);
        const int* __restrict__ relc = S.rel + rp;
        const double* __restrict__ cb = child_cb(S, CB, ch);
        int lo = 0, hi = rc;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (relc[mid] < row0) lo = mid + 1; else hi = mid; }
        const int t0 = lo;
        hi = rc;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (relc[mid] < row1) lo = mid + 1; else hi = mid; }
        const int t1 = lo;
        if (t0 == t1) continue;
        lo = 0; hi = rc;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (relc[mid] < d.c) lo = mid + 1; else hi = mid; }
        const int uc = lo;
        for (int tt = t0 + tx; tt < t1; tt += 32) {
            const int pi = relc[tt];
            const int ue = min(uc, tt + 1);
            for (int u = ty; u < ue; u += EAP_W * EAP_Q) {
                double v[EAP_Q], l[EAP_Q];
                size_t off[EAP_Q];
#pragma unroll
                for (int q = 0; q < EAP_Q; q++) {
                    const int uu = u + EAP_W * q;
                    const bool ok = uu < ue;
                    off[q] = (size_t)d.loff + pi + (size_t)relc[ok ? uu : u] * d.ld;
                    v[q] = ok ? cb[tt + (size_t)uu * rc] : 0.0;
                    l[q] = ok ? Lval[off[q]] : 0.0;
                }
#pragma unroll
                for (int q = 0; q < EAP_Q; q++)
                    if (u + EAP_W * q < ue) Lval[off[q]] = l[q] + v[q];
            }
        }
        __syncthreads();     // the next child may hit the same panel entries from other threads
    }
}

// rows below the diagonal block:  L21 = A21 * inv(L_kk)^T  as a tensor-core GEMM (64-row tiles, two
// CTAs per SM: K is only 128, a tile is mostly prologue and epilogue)
template <int WM>
__global__ void __launch_bounds__(TileCfg<WM>::THREADS, 3 - WM)
chol_trsm_kernel(DevSym S, const int* __restrict__ list, double* __restrict__ Lval,
                 const double* __restrict__ Xinv, int t, int ldlt, DeltaState* st) {
    constexpr int BM_ = TileCfg<WM>::BM;
    extern __shared__ __align__(16) unsigned char smraw[];
    GemmSmemT<WM>& sm = *reinterpret_cast<GemmSmemT<WM>*>(smraw);
    if (stop_requested(st)) return;
    const Front d = get_front(S, list[blockIdx.y]);
    const int j0 = t * WB;
    if (j0 >= d.c) return;
    const int b = min(WB, d.c - j0);
    const int j1 = j0 + b;
    const int row0 = (j1 & ~1) + blockIdx.x * BM_;
    if (row0 >= d.N) return;
    const int mrows = min(BM_, d.N - row0);
    double* Ag = Lval + d.loff + row0 + (size_t)j0 * d.ld;
    const double* Bg = Xinv + d.xoff + j0 + (size_t)j0 * d.ldx;
    double acc[8][4][2];
    gemm_mainloop<WM, false>(sm, Ag, d.ld, mrows, Bg, d.ldx, b, b, acc);
    if (!gemm_compute_warp<WM>()) return;
#pragma unroll
    for (int mt = 0; mt < 8; mt++) {
        const int i = row0 + acc_row(mt);
        if (i < j1 || i >= d.N) continue;
#pragma unroll
        for (int nt = 0; nt < 4; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int n = acc_col(nt, e);
                // LDL': L21 = A21 inv(L_kk)' D_kk^-1
                if (n < b) Lval[d.loff + i + (size_t)(j0 + n) * d.ld] = ldlt ? acc[mt][nt][e] / S.dvec[d.first + j0 + n] : acc[mt][nt][e];
            }
    }
}

// right-looking update of PANEL columns [cbeg, min(cend, c)) with the finished columns
// [k0, k0+klen):  L[i,k] -= sum_p L[i,p] L[k,p]   (i >= k).
// Called in a recursive (binary) schedule, see launch_wide_chol_level: inside an outer block the
// last 2^j finished blocks update the next 2^j blocks (klen = 128 * 2^j), and a complete outer
// block updates all remaining columns at once (klen = outer), so the bulk of the panel flops runs
// with a long K loop and the panel is read-modify-written c/outer instead of c/WB times.
// cbeg and k0 are multiples of WB.  (The update block of the front is formed once, at the end,
// by front_cb_kernel.)
// WM = 1: 64-row tiles, two CTAs per SM.
template <int WM, bool LDLT = false>
__global__ void __launch_bounds__(TileCfg<WM>::THREADS, 3 - WM)
chol_panel_update_kernel(DevSym S, const int* __restrict__ list, double* __restrict__ Lval,
                         int k0, int klen, int cbeg, int cend, DeltaState* st) {
    constexpr int BM_ = TileCfg<WM>::BM;
    extern __shared__ __align__(16) unsigned char smraw[];
    GemmSmemT<WM>& sm = *reinterpret_cast<GemmSmemT<WM>*>(smraw);
    if (stop_requested(st)) return;
    const Front d = get_front(S, list[blockIdx.y]);
    if (cbeg >= d.c) return;               // no panel columns left
    const int ce = min(cend, d.c);
    const int kl = min(klen, d.c - k0);
    const int nrow = (d.N - cbeg + BM_ - 1) / BM_;
    int I = -1, J = 0;
    const int ncol = (ce - cbeg + BN - 1) / BN;
    if (WM == 1) {
        // 64-row tiles: column tile J (128 wide) starts at row tile 2J; J (nrow + 1) - J^2 tiles precede it
        const long long tp = blockIdx.x;
        const double b = (double)nrow + 1.0;
        const double disc = b * b - 4.0 * (double)tp;
        J = disc > 0.0 ? (int)((b - sqrt(disc)) * 0.5) : nrow / 2;
        while (J > 0 && (long long)J * (nrow + 1) - (long long)J * J > tp) J--;
        while ((long long)(J + 1) * (nrow + 1) - (long long)(J + 1) * (J + 1) <= tp && 2 * (J + 1) < nrow) J++;
        const long long before = (long long)J * (nrow + 1) - (long long)J * J;
        const long long off = tp - before;
        if (J < ncol && 2 * J + off < nrow) I = 2 * J + (int)off;
    } else {
        int tp = blockIdx.x;
        for (; J < ncol; J++) {
            if (tp < nrow - J) { I = J + tp; break; }
            tp -= nrow - J;
        }
    }
    if (I < 0) return;
    const int ri = cbeg + I * BM_, rj = cbeg + J * BN;
    const double* Ag = Lval + d.loff + ri + (size_t)k0 * d.ld;
    const double* Bg = Lval + d.loff + rj + (size_t)k0 * d.ld;
    double acc[8][4][2];
    // LDL': L[i,k] -= sum_p (L[i,p] d_p) L[k,p]
    gemm_mainloop<WM, false, NoHook, LDLT>(sm, Ag, d.ld, min(BM_, d.N - ri), Bg, d.ld, min(BN, d.N - rj), kl, acc,
                                           NoHook(), S.dvec + d.first + k0);
    if (!gemm_compute_warp<WM>()) return;
#pragma unroll
    for (int nt = 0; nt < 4; nt++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int k = rj + acc_col(nt, e);
            if (k >= ce) continue;
            double* colp = Lval + d.loff + (size_t)k * d.ld;
#pragma unroll
            for (int mt = 0; mt < 8; mt++) {
                const int i = ri + acc_row(mt);
                if (i >= d.N || i < k) continue;
                colp[i] -= acc[mt][nt][e];
            }
        }
}

// update block of a medium / big front, written once:
//   CB[I,J] = sum_children (extend-add) - L21[I,:] * L21[J,:]^T      (K = all c pivot columns)

// rows [t0, t1) and columns [u0, u1) of child `ch` (positions in its update block) that land in
// tile (I, J) of the parent's update block: read from the tile-cut table built at symbolic time
// (four independent loads; the binary searches they replace were 30-50 dependent L2 round trips per
// child and tile, the bulk of a short-K tile's epilogue)
struct ChildRange { int t0, t1, u0, u1, rc; const int* relc; };
// the tile covers the 64-row cuts [r64, r64 + nr64) and the 128-column tile J = cuts [2J, 2J + 2)
__device__ __forceinline__ ChildRange child_range(const DevSym& S, int ch, int r64, int nr64, int J) {
    ChildRange R;
    const int64_t rp = S.rowptr[ch];
    R.rc = (int)(S.rowptr[ch + 1] - rp);
    R.relc = S.rel + rp;
    const int* __restrict__ tc = S.tcut + S.tcut_ptr[ch];
    R.t0 = tc[r64]; R.t1 = tc[r64 + nr64];
    R.u0 = tc[2 * J]; R.u1 = tc[2 * J + 2];
    return R;
}

// pulls the children's entries of this tile towards L2 a few stages before the K loop ends, so
// the merge below finds them on chip instead of paying a DRAM round trip per dependent step
struct CbPrefetch {
    const DevSym& S; const double* CB; int s, r64, nr64, J;
    __device__ __forceinline__ void operator()() const {
        const int tid = threadIdx.x;        // compute warps only: 0 .. GEMM_CWARPS*32-1
        for (int k = S.child_ptr[s]; k < S.child_ptr[s + 1]; k++) {
            const int ch = S.child_list[k];
            const ChildRange R = child_range(S, ch, r64, nr64, J);
            if (R.t1 <= R.t0 || R.u1 <= R.u0) continue;
            const double* __restrict__ cb = child_cb(S, CB, ch);
            // one 128-byte line = 16 doubles; a column segment of <= 128 rows spans <= 9 lines
            const int nu = R.u1 - R.u0;
            for (int idx = tid; idx < nu * 9; idx += GEMM_CWARPS * 32) {
                const int u = R.u0 + idx / 9, ln = idx % 9;
                const int tb = max(R.t0, u);
                const int tt = tb + 16 * ln;
                if (tt < R.t1 + 15 && tt < R.rc) {
                    const double* p = cb + (size_t)u * R.rc + min(tt, R.t1 - 1);
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
                }
            }
        }
    }
};

template <int WM, bool LDLT = false>
__global__ void __launch_bounds__(TileCfg<WM>::THREADS, 3 - WM)
front_cb_kernel(DevSym S, const int* __restrict__ list, const int2* __restrict__ tiles,
                const double* __restrict__ Lval, double* __restrict__ CB, DeltaState* st) {
    constexpr int BM_ = TileCfg<WM>::BM, NTHR = TileCfg<WM>::THREADS, TLD_ = BM_ + 1;
    extern __shared__ __align__(16) unsigned char smraw[];
    GemmSmemT<WM>& sm = *reinterpret_cast<GemmSmemT<WM>*>(smraw);
    if (stop_requested(st)) return;
    // exact tile list of the level (built with the schedule): a grid sized for the largest front
    // times the number of fronts would be mostly CTAs that find nothing to do, and each of those
    // still holds a 100+ KB shared-memory slot for a few dependent loads
    const int2 wt = tiles[blockIdx.x];
    const Front d = get_front(S, list[wt.x]);
    const int ce = d.c & ~1;
    const long long tp = wt.y;
    int I, J;
    if (WM == 2) {
        I = (int)((sqrt(8.0 * (double)tp + 1.0) - 1.0) * 0.5);
        while ((long long)I * (I + 1) / 2 > tp) I--;
        while ((long long)(I + 1) * (I + 2) / 2 <= tp) I++;
        J = (int)(tp - (long long)I * (I + 1) / 2);
    } else {
        int a = (int)((sqrt(4.0 * (double)tp + 1.0) - 1.0) * 0.5);
        while ((long long)a * (a + 1) > tp) a--;
        while ((long long)(a + 1) * (a + 2) <= tp) a++;
        const int rem = (int)(tp - (long long)a * (a + 1));
        if (rem <= a) { I = 2 * a; J = rem; } else { I = 2 * a + 1; J = rem - (a + 1); }
    }
    const int ri = ce + I * BM_, rj = ce + J * BN;
    const int r64 = I * WM;                    // first 64-row cut of the tile
    const double* Ag = Lval + d.loff + ri;
    const double* Bg = Lval + d.loff + rj;
    double acc[8][4][2];
    gemm_mainloop<WM, false, CbPrefetch, LDLT>(sm, Ag, d.ld, min(BM_, d.N - ri), Bg, d.ld, min(BN, d.N - rj), d.c, acc,
                                               CbPrefetch{S, CB, d.s, r64, WM, J}, S.dvec + d.first);
    // the stage buffers are free now: reuse them as the BM_ x 128 tile (ld BM_ + 1)
    double* T = reinterpret_cast<double*>(smraw);
    static_assert((size_t)TLD_ * BN * sizeof(double) <= sizeof(GemmSmemT<WM>) - 2 * STAGES * sizeof(unsigned long long),
                  "merge tile must fit in the stage ring");
    if (gemm_compute_warp<WM>()) {
#pragma unroll
        for (int mt = 0; mt < 8; mt++)
#pragma unroll
            for (int nt = 0; nt < 4; nt++)
#pragma unroll
                for (int e = 0; e < 2; e++) T[acc_row(mt) + acc_col(nt, e) * TLD_] = -acc[mt][nt][e];
    }
    __syncthreads();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int NW = NTHR / 32;
    constexpr int NQ = 2 * WM;                 // 32-row groups per tile column
    for (int k = S.child_ptr[d.s]; k < S.child_ptr[d.s + 1]; k++) {
        const int ch = S.child_list[k];
        const ChildRange R = child_range(S, ch, r64, WM, J);
        if (R.t1 <= R.t0 || R.u1 <= R.u0) continue;          // uniform across the CTA
        const double* __restrict__ cb = child_cb(S, CB, ch);
        const int* __restrict__ relc = R.relc;
        // a warp takes two child columns per round, a lane up to NQ rows of each: 2 NQ independent
        // loads in flight per lane before the first shared-memory update.  Distinct (row, column)
        // pairs of one child land on distinct tile entries, so there are no conflicts inside a child;
        // the barrier orders the children (ascending: fixed summation order).
        for (int u = R.u0 + 2 * warp; u < R.u1; u += 2 * NW) {
            double v[2][NQ];
            int dst[2][NQ];
#pragma unroll
            for (int h = 0; h < 2; h++) {
                const int uu = u + h;
                const bool cok = uu < R.u1;
                const int tb = max(R.t0, uu);
                const int pc = cok ? (relc[uu] - rj) * TLD_ - ri : 0;
                const double* col = cb + (size_t)(cok ? uu : u) * R.rc;
#pragma unroll
                for (int q = 0; q < NQ; q++) {
                    const int tt = tb + lane + 32 * q;
                    const bool ok = cok && tt < R.t1;
                    v[h][q] = ok ? col[tt] : 0.0;
                    dst[h][q] = ok ? relc[tt] + pc : -1;
                }
            }
#pragma unroll
            for (int h = 0; h < 2; h++)
#pragma unroll
                for (int q = 0; q < NQ; q++)
                    if (dst[h][q] >= 0) T[dst[h][q]] += v[h][q];
        }
        __syncthreads();
    }
    // sharded instance: the tiles of a split front go into the OWNER's arena (peer stores over NVLink)
    double* out = (S.owner ? S.cb_peer[S.owner[d.s]] : CB) + d.cboff;
    for (int idx = tid; idx < BM_ * BN; idx += NTHR) {
        const int i = ri + idx % BM_, kk = rj + idx / BM_;
        if (i < d.N && kk >= d.c && kk < d.N && i >= kk)
            out[(i - d.c) + (size_t)(kk - d.c) * d.r] = T[(i - ri) + (kk - rj) * TLD_];
    }
}

// ---------------------------------------------------------------------------
// inverse of the pivot block L11 by recursive block merging:
//   inv [A 0; B C] = [Ai 0; -Ci B Ai, Ci]
// merge level l joins blocks of S = WB * 2^l columns.  Two batched GEMMs:
//   phase 0:  T   = B * Ai      (T kept in Twork at the coordinates of X21)
//   phase 1:  X21 = -Ci * T
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(GEMM_THREADS)
trtri_merge_kernel(DevSym S, const int* __restrict__ list, const int2* __restrict__ items,
                   const double* __restrict__ Lval, double* __restrict__ Xinv, double* __restrict__ Twork,
                   int lvl, int phase, const DeltaState* st) {
    extern __shared__ __align__(16) unsigned char smraw[];
    GemmSmem& sm = *reinterpret_cast<GemmSmem*>(smraw);
    if (stop_requested(st)) return;
    // exact work list of the merge level (built with the schedule): (supernode, pair * nsub^2 + I * nsub + J)
    const int2 it = items[blockIdx.x];
    const Front d = get_front(S, list[it.x]);
    const int nsub = 1 << lvl;              // 128-blocks per half
    const int Sz = WB * nsub;
    const int per_pair = nsub * nsub;
    const int pair = it.y / per_pair;
    const int ij = it.y % per_pair;
    const int I = ij / nsub, J = ij % nsub;
    const int a0 = pair * 2 * Sz, a1 = a0 + Sz;
    if (a1 >= d.c) return;
    const int a2 = min(a1 + Sz, d.c);
    const int row0 = a1 + I * BM, col0 = a0 + J * BN;
    if (row0 >= a2) return;
    const int mrows = min(BM, a2 - row0);
    double acc[8][4][2];
    double* out;
    if (phase == 0) {
        // T[i,j] = sum_{k in [col0, a1)} L[i,k] * X[k,j]   (X11 lower: X[k,j] = 0 for k < j)
        const double* Ag = Lval + d.loff + row0 + (size_t)col0 * d.ld;
        const double* Bg = Xinv + d.xoff + col0 + (size_t)col0 * d.ldx;
        gemm_mainloop<2, true>(sm, Ag, d.ld, mrows, Bg, d.ldx, BN, a1 - col0, acc);
        out = Twork + d.xoff;
    } else {
        // X21[i,j] = - sum_{k in [a1, row0 + mrows)} X[i,k] * T[k,j]   (X22 lower)
        const int kend = min(a2, row0 + BM);
        const double* Ag = Xinv + d.xoff + row0 + (size_t)a1 * d.ldx;
        const double* Bg = Twork + d.xoff + a1 + (size_t)col0 * d.ldx;
        gemm_mainloop<2, true>(sm, Ag, d.ldx, mrows, Bg, d.ldx, BN, kend - a1, acc);
        out = Xinv + d.xoff;
    }
    if (!gemm_compute_warp()) return;
    const double sgn = phase == 0 ? 1.0 : -1.0;
#pragma unroll
    for (int mt = 0; mt < 8; mt++) {
        const int i = row0 + acc_row(mt);
        if (i >= a2) continue;
#pragma unroll
        for (int nt = 0; nt < 4; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int j = col0 + acc_col(nt, e);
                out[i + (size_t)j * d.ldx] = sgn * acc[mt][nt][e];
            }
    }
}

// ---------------------------------------------------------------------------
// Multi-CTA triangular solves for big supernodes (Cholesky mode):
//   forward :  gather children -> x1 = X b1 -> u -= L21 x1
//   backward:  u = x[rows] -> b1' = x1 - L21' u -> x1 = X' b1'
// ---------------------------------------------------------------------------
constexpr int WT = 512;
constexpr int SLAB = 32;      // rows per CTA in the row-oriented products
#ifndef OPB_TALL_N
#define OPB_TALL_N 8192
#endif
constexpr int TALL_N = OPB_TALL_N;
// threshold of the triangular products with inv(L11) (fine-grained variants: 8-row CTAs forward, four warps per
// column backward).  Measured on C5 100^3: fine variants for EVERY big front 10.0 ms per solve pair (the low levels
// drown in tiny CTAs), 4096 rows 9.39 ms, 8192 rows 9.33 ms.
#ifndef OPB_TRI_TALL_N
#define OPB_TRI_TALL_N 8192
#endif
constexpr int TRI_TALL_N = OPB_TRI_TALL_N;  // fronts with at least this many rows take the finer-grained solve variants
constexpr int KG = WT / 32;   // k-groups (warps)
#ifndef OPB_TALL_WPC
#define OPB_TALL_WPC 4
#endif
constexpr int TALL_WPC = OPB_TALL_WPC;   // warps per column of the backward products on tall fronts

// children's update vectors into this supernode's right-hand side; a CTA owns a range of
// destination rows, a thread one destination: it sums that destination's sources in ascending
// child order (gather lists built at symbolic time: deterministic, no atomics, no searches)
constexpr int GR = 512;       // destination rows per CTA (one per thread)
__global__ void __launch_bounds__(WT)
wide_fwd_gather_kernel(DevSym S, const int* __restrict__ list, double* __restrict__ x, double* __restrict__ u) {
    const int s = list[blockIdx.y];
    pdl_release();
    pdl_wait();
    const int first = S.sfirst[s];
    const int c = S.sfirst[s + 1] - first;
    const int64_t rp = S.rowptr[s];
    const int r = (int)(S.rowptr[s + 1] - rp);
    const int row0 = blockIdx.x * GR;
    if (row0 >= c + r) return;
    const int row1 = min(c + r, row0 + GR);
    double* xs = x + first;
    double* us = u + rp;
    const int64_t gb = rp + first;
    for (int i = row0 + threadIdx.x; i < row1; i += WT) {
        const double acc = gather_dest(S, u, gb + i);
        if (i < c) xs[i] += acc; else us[i - c] = acc;
    }
}

// The pivot block of a big supernode is inverted in diagonal blocks of XB columns (the recursive
// merge stops there: inverting a 20000-column separator in one piece would cost 2/3 c^3 flops).
// The solves walk those blocks: per block one triangular product with its inverse and one
// rectangular product with the columns of L below / right of it.

// xnew[i] = sum_{k in [b0, i]} X[i,k] * xs[k]   for the pivot rows i of block blk.
// A CTA owns RS consecutive rows; its 512 threads split the k range 512/RS ways (a lane group of
// RS lanes reads RS consecutive rows of one column: full 32-byte sectors for RS >= 4).  Fronts
// with N >= TALL_N take RS = 8 (four times the CTAs, a quarter of the serial work per CTA: a level
// with one tall front would otherwise run on 64 CTAs), the others RS = 32.
template <int RS>
__global__ void __launch_bounds__(WT)
wide_fwd_tri_kernel(DevSym S, const int* __restrict__ list, const double* __restrict__ Xinv,
                    const double* __restrict__ x, double* __restrict__ xnew, int blk) {
    constexpr int KS = WT / RS;              // k-groups per CTA
    __shared__ double red[KS][RS];
    const Front d = get_front(S, list[blockIdx.y]);
    pdl_release();
    pdl_wait();
    if ((d.N >= TRI_TALL_N) != (RS < 32)) return;     // the other variant's front
    const int b0 = blk * XB;
    const int b1 = min(d.c, b0 + XB);
    const int i0 = b0 + blockIdx.x * RS;
    if (i0 >= b1) return;
    const int r = threadIdx.x % RS, w = threadIdx.x / RS;
    const int i = i0 + r;
    const double* X = Xinv + d.xoff;
    const double* xs = x + d.first;
    const int kend = min(b1, i0 + RS);
    double acc = 0.0;
    if (i < b1) {
        int k = b0 + w;
        for (; k + 7 * KS < kend; k += 8 * KS) {
            double v[8];
#pragma unroll
            for (int t = 0; t < 8; t++) v[t] = X[i + (size_t)(k + t * KS) * d.ldx];
#pragma unroll
            for (int t = 0; t < 8; t++) acc += v[t] * xs[k + t * KS];
        }
        for (; k < kend; k += KS) acc += X[i + (size_t)k * d.ldx] * xs[k];   // upper part of X is zero
    }
    red[w][r] = acc;
    __syncthreads();
    if (w == 0 && i < b1) {
        double v = 0.0;
#pragma unroll 8
        for (int g = 0; g < KS; g++) v += red[g][r];
        xnew[d.first + i] = v;
    }
}

// rows of block blk: xs[i] = xnew[i];  rows below it (later pivot rows and the rows >= c):
// rhs[i] -= sum_{k in block} L[i,k] * xnew[k]
// NT threads per CTA = NT/32 k-groups (NT = 256 for tall fronts was measured 3 % slower than 512).
template <int NT>
__global__ void __launch_bounds__(NT)
wide_fwd_upd_kernel(DevSym S, const int* __restrict__ list, const double* __restrict__ Lval,
                    double* __restrict__ x, const double* __restrict__ xnew, double* __restrict__ u, int blk) {
    constexpr int KGU = NT / 32;
    __shared__ double red[KGU][SLAB];
    const Front d = get_front(S, list[blockIdx.y]);
    pdl_release();
    pdl_wait();
    const int b0 = blk * XB;
    if (b0 >= d.c) return;
    const int b1 = min(d.c, b0 + XB);
    const int i0 = b0 + blockIdx.x * SLAB;
    if (i0 >= d.N) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = i0 + lane;
    const double* xn = xnew + d.first;
    if (i0 + SLAB <= b1) {           // rows inside the block: publish the forward solution
        if (w == 0) x[d.first + i] = xn[i];
        return;
    }
    const double* L = Lval + d.loff;
    double acc = 0.0;
    if (i >= b1 && i < d.N) {
        int k = b0 + w;
        for (; k + 7 * KGU < b1; k += 8 * KGU) {
            double v[8];
#pragma unroll
            for (int t = 0; t < 8; t++) v[t] = L[i + (size_t)(k + t * KGU) * d.ld];
#pragma unroll
            for (int t = 0; t < 8; t++) acc += v[t] * xn[k + t * KGU];
        }
        for (; k < b1; k += KGU) acc += L[i + (size_t)k * d.ld] * xn[k];
    }
    red[w][lane] = acc;
    __syncthreads();
    if (w == 0 && i < d.N) {
        if (i < b1) x[d.first + i] = xn[i];
        else {
            double v = 0.0;
#pragma unroll
            for (int g = 0; g < KGU; g++) v += red[g][lane];
            if (i < d.c) x[d.first + i] -= v;
            else u[S.rowptr[d.s] + (i - d.c)] -= v;
        }
    }
}

__global__ void __launch_bounds__(WT)
wide_bwd_gather_kernel(DevSym S, const int* __restrict__ list, double* __restrict__ x,
                       double* __restrict__ u, int ldlt) {
    const int s = list[blockIdx.y];
    pdl_release();
    pdl_wait();
    const int64_t rp = S.rowptr[s];
    const int r = (int)(S.rowptr[s + 1] - rp);
    const int t = blockIdx.x * WT + threadIdx.x;
    if (t < r) u[rp + t] = x[S.rowidx[rp + t]];
    // LDL': x = L^-T D^-1 y -- the pivots are applied to the whole pivot block before its backward sweep
    if (ldlt) {
        const int first = S.sfirst[s];
        if (t < S.sfirst[s + 1] - first) x[first + t] /= S.dvec[first + t];
    }
}

// xnew[k] = xs[k] - sum_{i >= b1} L[i,k] * f[i]  for the columns k of block blk, where f is the
// final solution on the later pivot rows and the ancestors' values on the rows >= c.
// WPC warps share one column (interleaved 128-row chunks, partial sums combined in shared memory
// in a fixed order): a warp per column of a few tall fronts leaves most SMs idle.  Fronts with
// N >= TALL_N take the 4-warp variant, the others the 1-warp variant.
template <int WPC>
__global__ void __launch_bounds__(WT)
wide_bwd_upd_kernel(DevSym S, const int* __restrict__ list, const double* __restrict__ Lval,
                    const double* __restrict__ x, double* __restrict__ xnew, const double* __restrict__ u, int blk) {
    __shared__ double red[KG];
    const Front d = get_front(S, list[blockIdx.y]);
    pdl_release();
    pdl_wait();
    if ((d.N >= TALL_N) != (WPC > 1)) return;     // the other variant's front (fixed per front: reproducible)
    const int b0 = blk * XB;
    const int b1 = min(d.c, b0 + XB);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int part = w % WPC;
    const int k = b0 + blockIdx.x * (KG / WPC) + w / WPC;
    const bool valid = k < b1;
    double acc = 0.0;
    if (valid) {
        const double* col = Lval + d.loff + (size_t)k * d.ld;
        const double* xs = x + d.first;
        const double* us = u + S.rowptr[d.s];
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        // later pivot rows [b1, c): groups of 128 rows dealt round-robin to the WPC warps; two groups
        // (eight independent loads per lane) are in flight where the column is long enough
        int base = b1 + 128 * part;
        for (; base + 128 * (WPC + 1) <= d.c; base += 256 * WPC) {
            const int i = base + lane, j = i + 128 * WPC;
            const double l0 = col[i], l1 = col[i + 32], l2 = col[i + 64], l3 = col[i + 96];
            const double m0 = col[j], m1 = col[j + 32], m2 = col[j + 64], m3 = col[j + 96];
            a0 += l0 * xs[i]; a1 += l1 * xs[i + 32]; a2 += l2 * xs[i + 64]; a3 += l3 * xs[i + 96];
            a0 += m0 * xs[j]; a1 += m1 * xs[j + 32]; a2 += m2 * xs[j + 64]; a3 += m3 * xs[j + 96];
        }
        for (; base + 128 <= d.c; base += 128 * WPC) {
            const int i = base + lane;
            a0 += col[i] * xs[i]; a1 += col[i + 32] * xs[i + 32];
            a2 += col[i + 64] * xs[i + 64]; a3 += col[i + 96] * xs[i + 96];
        }
        for (int i = base + lane; i < d.c && i < base + 128; i += 32) a0 += col[i] * xs[i];
        // rows below the pivot block
        const double* colr = col + d.c;
        base = 128 * part;
        for (; base + 128 * (WPC + 1) <= d.r; base += 256 * WPC) {
            const int t = base + lane, v = t + 128 * WPC;
            const double l0 = colr[t], l1 = colr[t + 32], l2 = colr[t + 64], l3 = colr[t + 96];
            const double m0 = colr[v], m1 = colr[v + 32], m2 = colr[v + 64], m3 = colr[v + 96];
            a0 += l0 * us[t]; a1 += l1 * us[t + 32]; a2 += l2 * us[t + 64]; a3 += l3 * us[t + 96];
            a0 += m0 * us[v]; a1 += m1 * us[v + 32]; a2 += m2 * us[v + 64]; a3 += m3 * us[v + 96];
        }
        for (; base + 128 <= d.r; base += 128 * WPC) {
            const int t = base + lane;
            a0 += colr[t] * us[t]; a1 += colr[t + 32] * us[t + 32];
            a2 += colr[t + 64] * us[t + 64]; a3 += colr[t + 96] * us[t + 96];
        }
        for (int t = base + lane; t < d.r && t < base + 128; t += 32) a0 += colr[t] * us[t];
        acc = (a0 + a1) + (a2 + a3);
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if (WPC == 1) {
        if (valid && lane == 0) xnew[d.first + k] = x[d.first + k] - acc;
        return;
    }
    if (lane == 0) red[w] = acc;
    __syncthreads();
    if (valid && part == 0 && lane == 0) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < WPC; q++) v += red[w + q];
        xnew[d.first + k] = x[d.first + k] - v;
    }
}

// xs[k] = sum_{i in [k, b1)} X[i,k] * xnew[i]  for the columns k of block blk
template <int WPC>
__global__ void __launch_bounds__(WT)
wide_bwd_tri_kernel(DevSym S, const int* __restrict__ list, const double* __restrict__ Xinv,
                    double* __restrict__ x, const double* __restrict__ xnew, int blk) {
    __shared__ double red[KG];
    const Front d = get_front(S, list[blockIdx.y]);
    pdl_release();
    pdl_wait();
    if ((d.N >= TRI_TALL_N) != (WPC > 1)) return;     // the other variant's front (fixed per front: reproducible)
    const int b0 = blk * XB;
    const int b1 = min(d.c, b0 + XB);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int part = w % WPC;
    const int k = b0 + blockIdx.x * (KG / WPC) + w / WPC;
    const bool valid = k < b1;
    double acc = 0.0;
    if (valid) {
        const double* col = Xinv + d.xoff + (size_t)k * d.ldx;
        const double* xn = xnew + d.first;
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        const int i0 = (k & ~31);            // aligned start; entries above the diagonal are zero
        int base = i0 + 128 * part;
        for (; base + 128 <= b1; base += 128 * WPC) {
            const int i = base + lane;
            a0 += col[i] * xn[i]; a1 += col[i + 32] * xn[i + 32];
            a2 += col[i + 64] * xn[i + 64]; a3 += col[i + 96] * xn[i + 96];
        }
        for (int i = base + lane; i < b1 && i < base + 128; i += 32) a0 += col[i] * xn[i];
        acc = (a0 + a1) + (a2 + a3);
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    }
    if (WPC == 1) {
        if (valid && lane == 0) x[d.first + k] = acc;
        return;
    }
    if (lane == 0) red[w] = acc;
    __syncthreads();
    if (valid && part == 0 && lane == 0) {
        double v = 0.0;
#pragma unroll
        for (int q = 0; q < WPC; q++) v += red[w + q];
        x[d.first + k] = v;
    }
}

inline size_t diag_smem() { return (size_t)(LDD * WB + INVBUF + XS_BLOCKS * INVBUF) * sizeof(double); }
inline size_t mid_smem(int panel) { return (size_t)(panel + INVBUF) * sizeof(double); }

}  // namespace

int g_occ_small_tiles = -1;   // resident CTAs per SM of the 64-row-tile kernels (opb_get_info "occ_small_tiles")

cudaError_t dense_configure() {
    cudaError_t e;
    e = cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)diag_smem());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(mid_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mid_smem(MIDL_PANEL));
    if (e != cudaSuccess) return e;
    auto gemm_attr = [](const void* f, size_t smem) -> cudaError_t {
        cudaError_t e2 = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e2 != cudaSuccess) return e2;
        // the whole L1/shared array as shared memory: two 64-row-tile CTAs (2 x 100 KB) must fit on an SM
        return cudaFuncSetAttribute(f, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    };
    e = gemm_attr((const void*)chol_trsm_kernel<1>, sizeof(GemmSmemT<1>)); if (e != cudaSuccess) return e;
    e = gemm_attr((const void*)chol_panel_update_kernel<1, false>, sizeof(GemmSmemT<1>)); if (e != cudaSuccess) return e;
    e = gemm_attr((const void*)chol_panel_update_kernel<2, false>, sizeof(GemmSmem)); if (e != cudaSuccess) return e;
    e = gemm_attr((const void*)front_cb_kernel<1, false>, sizeof(GemmSmemT<1>)); if (e != cudaSuccess) return e;
    e = gemm_attr((const void*)front_cb_kernel<2, false>, sizeof(GemmSmem)); if (e != cudaSuccess) return e;
    e = gemm_attr((const void*)chol_panel_update_kernel<1, true>, sizeof(GemmSmemT<1>)); if (e != cudaSuccess) return e;
    e = gemm_attr((const void*)chol_panel_update_kernel<2, true>, sizeof(GemmSmem)); if (e != cudaSuccess) return e;
    e = gemm_attr((const void*)front_cb_kernel<1, true>, sizeof(GemmSmemT<1>)); if (e != cudaSuccess) return e;
    e = gemm_attr((const void*)front_cb_kernel<2, true>, sizeof(GemmSmem)); if (e != cudaSuccess) return e;
    e = gemm_attr((const void*)trtri_merge_kernel, sizeof(GemmSmem)); if (e != cudaSuccess) return e;
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, front_cb_kernel<1, false>, TileCfg<1>::THREADS, sizeof(GemmSmemT<1>)) == cudaSuccess)
        g_occ_small_tiles = occ;
    else cudaGetLastError();
    return cudaSuccess;
}

void launch_wide_chol_level(const DevSym& S, const LevelPlan& L, const int* d_sched, double* Lval,
                            double* CB, double* Xinv, DeltaState* st_d, int ldlt, int outer_block, int cb_small_k,
                            const SideStream* side, KernelTimer* timer, int phase, cudaStream_t st) {
    // phase 0: the whole level; 1: panels only; 2: update blocks only (sharded instance, levels with
    // split fronts: the helpers' tiles come in between, see launch_factor_levels)
    const bool do_panels = phase != 2 && L.wide_count > 0, do_cb = phase != 1;
    if (!do_panels && !(do_cb && (L.cbt_count[0] > 0 || L.cbt_count[1] > 0))) return;
    const bool upd_small_tiles = (cb_small_k & 1) == 0;     // odd cb_small_k (tuning): bulk panel updates in 128-row tiles
    KernelTimer* phase_timer = (timer && timer->phases) ? timer : nullptr;
    if (phase_timer) timer = nullptr;
    // medium fronts: panel in shared memory
    for (int fc = FC_MID; do_panels && fc <= FC_MIDL; fc++) {
        if (!L.count[fc]) continue;
        mid_panel_kernel<<<L.count[fc], PT, mid_smem(L.maxPanel[fc]), st>>>(S, d_sched + L.begin[fc], Lval, CB, ldlt, st_d);
        count_launch();
    }
    // big fronts: blocked right-looking factorisation of the panel columns
    if (do_panels && L.count[FC_BIG]) {
        const int* list = d_sched + L.begin[FC_BIG];
        dim3 gea((L.maxN[FC_BIG] + EAP_RB - 1) / EAP_RB, L.count[FC_BIG]);
        if (L.count[FC_BIG] <= 4) big_extend_add_panel_kernel<512, 8><<<gea, 512, 0, st>>>(S, list, Lval, CB, st_d);
        else big_extend_add_panel_kernel<256, 4><<<gea, 256, 0, st>>>(S, list, Lval, CB, st_d);
        count_launch();
        if (phase_timer) phase_timer->put_mark(1, st);
        int sub = 1;                                            // WB blocks per outer block (power of two)
        while (sub * 2 * WB <= outer_block) sub *= 2;
        const int nsteps = (int)L.step_count.size();
        auto update = [&](cudaStream_t sx, int k0, int klen, int cbeg, int cend) -> bool {
            // fronts with panel columns beyond cbeg: a prefix of the list (sorted by c descending)
            const int tb = cbeg / WB;
            if (cend <= cbeg || tb >= nsteps || L.step_count[tb] <= 0) return false;
            const int cnt2 = L.step_count[tb];
            if (cend - cbeg <= WB || upd_small_tiles) {
                // 64-row tiles, two CTAs per SM
                const int nrow64 = (L.step_maxN[tb] - cbeg + 63) / 64;
                const int ncol = (std::min(cend, L.maxC[FC_BIG]) - cbeg + BN - 1) / BN;
                long long tiles = 0;
                for (int J = 0; J < ncol && 2 * J < nrow64; J++) tiles += nrow64 - 2 * J;
                if (tiles <= 0) return false;
                dim3 gu((unsigned)tiles, cnt2);
                if (timer) cudaEventRecord(timer->next(1), sx);
                if (ldlt) chol_panel_update_kernel<1, true><<<gu, TileCfg<1>::THREADS, sizeof(GemmSmemT<1>), sx>>>(S, list, Lval, k0, klen, cbeg, cend, st_d);
                else chol_panel_update_kernel<1, false><<<gu, TileCfg<1>::THREADS, sizeof(GemmSmemT<1>), sx>>>(S, list, Lval, k0, klen, cbeg, cend, st_d);
                if (timer) cudaEventRecord(timer->next(1), sx);
                count_launch();
                return true;
            }
            const int nrow = (L.step_maxN[tb] - cbeg + BM - 1) / BM;
            const int ncol = (std::min(cend, L.maxC[FC_BIG]) - cbeg + BN - 1) / BN;
            long long tiles = 0;
            for (int J = 0; J < ncol && J < nrow; J++) tiles += nrow - J;
            if (tiles <= 0) return false;
            dim3 gu((unsigned)tiles, cnt2);
            if (timer) cudaEventRecord(timer->next(1), sx);
            if (ldlt) chol_panel_update_kernel<2, true><<<gu, GEMM_THREADS, sizeof(GemmSmem), sx>>>(S, list, Lval, k0, klen, cbeg, cend, st_d);
            else chol_panel_update_kernel<2, false><<<gu, GEMM_THREADS, sizeof(GemmSmem), sx>>>(S, list, Lval, k0, klen, cbeg, cend, st_d);
            if (timer) cudaEventRecord(timer->next(1), sx);
            count_launch();
            return true;
        };
        if (side && side->deep && sub <= (1 << LA_CLASSES)) {
            // Deep look-ahead.  With tb blocks finished and w the largest power of two dividing tb, the
            // update U(tb) targets the block columns [tb, tb + w) (w < sub) or everything from tb on
            // (w >= sub).  Its columns are next touched as follows: block tb right away (diagonal block
            // tb), the blocks [tb + 2^i, tb + 2^(i+1)) by U(tb + 2^i), the blocks from tb + sub on by
            // U(tb + sub).  So U(tb) is issued as: block tb on the chain stream, piece i on cls[i],
            // the rest on the lowest-priority stream; and U(tb) itself waits for exactly one earlier
            // piece, the last writer of its own target range: piece log2(w) of U(tb - w), or the rest
            // of U(tb - sub).  At most one piece per class is outstanding (the next producer of class
            // i comes at tb + 2^(i+1), its consumer at tb + 2^i), so one event per class is enough.
            cudaStream_t C = side->stream;
            cudaEventRecord(side->start, st); cudaStreamWaitEvent(C, side->start, 0);
            bool out_cls[LA_CLASSES] = {false}, out_rest = false;
            for (int t = 0; t < nsteps; t++) {
                const int cnt = L.step_count[t];
                if (cnt <= 0) break;
                const int maxN = L.step_maxN[t];
                chol_diag_kernel<<<cnt, PT, diag_smem(), C>>>(S, list, Lval, Xinv, t, ldlt, st_d);
                count_launch();
                const int rem = maxN - t * WB;
                if (rem <= 0) continue;
                const int nrow = (rem + 63) / 64 + 1;
                dim3 gt(nrow, cnt);
                chol_trsm_kernel<1><<<gt, TileCfg<1>::THREADS, sizeof(GemmSmemT<1>), C>>>(S, list, Lval, Xinv, t, ldlt, st_d);
                count_launch();
                const int tb = t + 1;
                if (tb >= nsteps || L.step_count[tb] <= 0) continue;       // no panel columns left
                const int w = tb & -tb;
                const bool outer = w >= sub;
                const int k0 = (tb - (outer ? sub : w)) * WB, klen = (outer ? sub : w) * WB;
                const int near_end = tb + (outer ? sub : w);               // blocks [tb, near_end) in pieces
                // the last writer of the target range
                if (outer) { if (out_rest) { cudaStreamWaitEvent(C, side->rest_done, 0); out_rest = false; } }
                else {
                    const int i = __builtin_ctz((unsigned)w);
                    if (out_cls[i]) { cudaStreamWaitEvent(C, side->cls_done[i], 0); out_cls[i] = false; }
                }
                cudaEventRecord(side->fork, C);
                update(C, k0, klen, tb * WB, (tb + 1) * WB);
                for (int i = 0; tb + (1 << i) < near_end; i++) {
                    const int lo = tb + (1 << i), hi = std::min(tb + (2 << i), near_end);
                    if (lo >= nsteps || L.step_count[lo] <= 0) break;
                    cudaStreamWaitEvent(side->cls[i], side->fork, 0);
                    update(side->cls[i], k0, klen, lo * WB, hi * WB);
                    cudaEventRecord(side->cls_done[i], side->cls[i]);
                    out_cls[i] = true;
                }
                if (outer && near_end < nsteps && L.step_count[near_end] > 0) {
                    cudaStreamWaitEvent(side->rest, side->fork, 0);
                    update(side->rest, k0, klen, near_end * WB, 1 << 30);
                    cudaEventRecord(side->rest_done, side->rest);
                    out_rest = true;
                }
            }
            for (int i = 0; i < LA_CLASSES; i++) if (out_cls[i]) cudaStreamWaitEvent(C, side->cls_done[i], 0);
            if (out_rest) cudaStreamWaitEvent(C, side->rest_done, 0);
            cudaEventRecord(side->join, C); cudaStreamWaitEvent(st, side->join, 0);
        } else {
        // Two streams: C carries the latency chain (diagonal block, TRSM, update of the next block
        // column), B the bulk of the right-looking updates.  Classic mode: C = the caller's stream,
        // B = the side stream.  With a high-priority side stream (side->chain_on_side) the roles
        // are swapped: the chain's short kernels then win every SM a bulk tile gives back instead
        // of queueing behind the whole bulk grid.
        cudaStream_t C = st, B = st;
        if (side) { if (side->chain_on_side) C = side->stream; else B = side->stream; }
        if (C != st) { cudaEventRecord(side->start, st); cudaStreamWaitEvent(C, side->start, 0); }
        bool pending_join = false;       // a bulk update is still running on B
        for (int t = 0; t < nsteps; t++) {
            const int cnt = L.step_count[t];
            if (cnt <= 0) break;
            const int maxN = L.step_maxN[t];
            chol_diag_kernel<<<cnt, PT, diag_smem(), C>>>(S, list, Lval, Xinv, t, ldlt, st_d);
            count_launch();
            const int rem = maxN - t * WB;     // rows from the start of the block (upper bound)
            if (rem <= 0) continue;
            const int nrow = (rem + 63) / 64 + 1;
            dim3 gt(nrow, cnt);
            chol_trsm_kernel<1><<<gt, TileCfg<1>::THREADS, sizeof(GemmSmemT<1>), C>>>(S, list, Lval, Xinv, t, ldlt, st_d);
            count_launch();
            // Recursive (binary) schedule of the right-looking updates: with tb blocks finished and
            // 2^j the largest power of two dividing tb, the last 2^j blocks update the next 2^j
            // blocks (K = 128 * 2^j); when 2^j reaches the outer block they update ALL remaining
            // columns instead.  Every block column has received exactly the blocks before it when
            // its turn comes, half of the flops inside an outer block run at K = outer/2, a
            // quarter at outer/4, ..., and the panel is read-modify-written c/outer times.
            const int tb = t + 1;
            const int wblk = tb & -tb;
            int k0, klen, cend;
            const int cbeg = tb * WB;
            if (wblk >= sub) { k0 = (tb - sub) * WB; klen = sub * WB; cend = 1 << 30; }
            else { k0 = (tb - wblk) * WB; klen = wblk * WB; cend = (tb + wblk) * WB; }
            // the previous step's bulk update wrote the columns this step updates
            if (pending_join) { cudaStreamWaitEvent(C, side->join, 0); pending_join = false; }
            const bool has_rest = cend > cbeg + WB && tb + 1 < nsteps && L.step_count[tb + 1] > 0;
            if (side && has_rest) {
                // look-ahead: block column tb (all the next diagonal block and TRSM need) on the chain
                // stream, the columns beyond it on the bulk stream, concurrently with those kernels
                cudaEventRecord(side->fork, C);
                cudaStreamWaitEvent(B, side->fork, 0);
                update(B, k0, klen, cbeg + WB, cend);
                cudaEventRecord(side->join, B);      // always rejoin (stream capture demands it)
                pending_join = true;
                update(C, k0, klen, cbeg, cbeg + WB);
            } else {
                update(C, k0, klen, cbeg, cend);
            }
        }
        if (pending_join) cudaStreamWaitEvent(C, side->join, 0);
        if (C != st) { cudaEventRecord(side->fork, C); cudaStreamWaitEvent(st, side->fork, 0); }
        }
    }
    // update blocks of all medium and big fronts, written once
    if (do_cb) {
        // short K (low levels): 64-row tiles, two CTAs per SM; otherwise 128 x 128 tiles
        const int v = L.wide_maxC <= cb_small_k ? 0 : 1;
        if (phase_timer) phase_timer->put_mark(2, st);
        if (L.cbt_count[v] > 0) {
            const int2* tiles = reinterpret_cast<const int2*>(d_sched + L.cbt_begin[v]);
            if (timer) cudaEventRecord(timer->next(0), st);
            const int* wl = d_sched + L.wide_begin;
            if (v == 0 && !ldlt) front_cb_kernel<1, false><<<L.cbt_count[v], TileCfg<1>::THREADS, sizeof(GemmSmemT<1>), st>>>(S, wl, tiles, Lval, CB, st_d);
            else if (v == 0) front_cb_kernel<1, true><<<L.cbt_count[v], TileCfg<1>::THREADS, sizeof(GemmSmemT<1>), st>>>(S, wl, tiles, Lval, CB, st_d);
            else if (!ldlt) front_cb_kernel<2, false><<<L.cbt_count[v], GEMM_THREADS, sizeof(GemmSmem), st>>>(S, wl, tiles, Lval, CB, st_d);
            else front_cb_kernel<2, true><<<L.cbt_count[v], GEMM_THREADS, sizeof(GemmSmem), st>>>(S, wl, tiles, Lval, CB, st_d);
            if (timer) cudaEventRecord(timer->next(0), st);
            count_launch();
        }
    }
}

void launch_trtri(const DevSym& S, const TrtriPlan& T, const int* d_sched, const double* Lval,
                  double* Xinv, double* Twork, const DeltaState* st_d, cudaStream_t st) {
    const int* list = d_sched + T.list_begin;
    for (size_t l = 0; l < T.items_count.size(); l++) {
        if (T.items_count[l] <= 0) continue;
        const int2* items = reinterpret_cast<const int2*>(d_sched + T.items_begin[l]);
        for (int phase = 0; phase < 2; phase++) {
            trtri_merge_kernel<<<T.items_count[l], GEMM_THREADS, sizeof(GemmSmem), st>>>(S, list, items, Lval, Xinv, Twork, (int)l, phase, st_d);
            count_launch();
        }
    }
}

// launch with the programmatic-serialisation attribute: the kernel may start while its predecessor in the
// stream is still draining; it calls pdl_wait() before it reads anything the predecessor wrote
template <class... KArgs, class... Args>
static void launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

void launch_solve_wide_fwd(const DevSym& S, const LevelPlan& L, const int* d_sched, const double* Lval,
                           const double* Xinv, double* x, double* xnew, double* u, cudaStream_t st) {
    const int cnt = L.count[FC_BIG];
    if (!cnt) return;
    const int* list = d_sched + L.begin[FC_BIG];
    dim3 gg((L.maxN[FC_BIG] + GR - 1) / GR, cnt);
    launch_pdl(wide_fwd_gather_kernel, gg, dim3(WT), st, S, list, x, u);
    count_launch();
    const int nblk = (L.maxC[FC_BIG] + XB - 1) / XB;
    for (int blk = 0; blk < nblk; blk++) {
        const int cb = std::min(XB, L.maxC[FC_BIG] - blk * XB);
        if (L.maxN[FC_BIG] >= TRI_TALL_N) {
            dim3 g1((cb + 7) / 8, cnt);
            launch_pdl(wide_fwd_tri_kernel<8>, g1, dim3(WT), st, S, list, Xinv, x, xnew, blk);
            count_launch();
        }
        if (L.minN[FC_BIG] < TRI_TALL_N) {
            dim3 g1((cb + SLAB - 1) / SLAB, cnt);
            launch_pdl(wide_fwd_tri_kernel<SLAB>, g1, dim3(WT), st, S, list, Xinv, x, xnew, blk);
            count_launch();
        }
        dim3 g2((L.maxN[FC_BIG] - blk * XB + SLAB - 1) / SLAB, cnt);
        launch_pdl(wide_fwd_upd_kernel<WT>, g2, dim3(WT), st, S, list, Lval, x, xnew, u, blk);
        count_launch();
    }
}

void launch_solve_wide_bwd(const DevSym& S, const LevelPlan& L, const int* d_sched, const double* Lval,
                           const double* Xinv, double* x, double* xnew, double* u, int ldlt, cudaStream_t st) {
    const int cnt = L.count[FC_BIG];
    if (!cnt) return;
    const int* list = d_sched + L.begin[FC_BIG];
    dim3 g0((L.maxN[FC_BIG] + WT) / WT, cnt);
    launch_pdl(wide_bwd_gather_kernel, g0, dim3(WT), st, S, list, x, u, ldlt);
    count_launch();
    const int nblk = (L.maxC[FC_BIG] + XB - 1) / XB;
    for (int blk = nblk - 1; blk >= 0; blk--) {
        const int cb = std::min(XB, L.maxC[FC_BIG] - blk * XB);
        constexpr int WPC = TALL_WPC;
        const dim3 gt((cb + KG / WPC - 1) / (KG / WPC), cnt), g1((cb + KG - 1) / KG, cnt);
        if (L.maxN[FC_BIG] >= TALL_N) { launch_pdl(wide_bwd_upd_kernel<WPC>, gt, dim3(WT), st, S, list, Lval, x, xnew, u, blk); count_launch(); }
        if (L.minN[FC_BIG] < TALL_N) { launch_pdl(wide_bwd_upd_kernel<1>, g1, dim3(WT), st, S, list, Lval, x, xnew, u, blk); count_launch(); }
        if (L.maxN[FC_BIG] >= TRI_TALL_N) { launch_pdl(wide_bwd_tri_kernel<WPC>, gt, dim3(WT), st, S, list, Xinv, x, xnew, blk); count_launch(); }
        if (L.minN[FC_BIG] < TRI_TALL_N) { launch_pdl(wide_bwd_tri_kernel<1>, g1, dim3(WT), st, S, list, Xinv, x, xnew, blk); count_launch(); }
    }
}

// Force-load every kernel of this translation unit (CUDA loads kernels lazily, and a load may
// synchronise the context: that must not happen while another stream waits in a cross-rank barrier).
cudaError_t preload_dense() {
    cudaFuncAttributes a;
    cudaError_t e;
    e = cudaFuncGetAttributes(&a, big_extend_add_panel_kernel<256, 4>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, big_extend_add_panel_kernel<512, 8>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, chol_diag_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, chol_panel_update_kernel<1, false>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, chol_panel_update_kernel<2, false>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, chol_panel_update_kernel<1, true>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, chol_panel_update_kernel<2, true>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, chol_trsm_kernel<1>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, front_cb_kernel<1, false>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, front_cb_kernel<2, false>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, front_cb_kernel<1, true>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, front_cb_kernel<2, true>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, mid_panel_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, trtri_merge_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, wide_bwd_gather_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, wide_bwd_tri_kernel<1>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, wide_bwd_tri_kernel<TALL_WPC>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, wide_bwd_upd_kernel<1>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, wide_bwd_upd_kernel<TALL_WPC>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, wide_fwd_gather_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, wide_fwd_tri_kernel<8>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, wide_fwd_tri_kernel<SLAB>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, wide_fwd_upd_kernel<WT>); if (e != cudaSuccess) return e;
    return cudaSuccess;
}

}  // namespace opb
