// Dense kernels for the big fronts of the supernodal Cholesky (sm_100a):
//
//   * a 128x128 FP64 tensor-core tile engine: DMMA (mma.sync m8n8k4 f64) fed by
//     1-D TMA bulk copies (cp.async.bulk global->shared, mbarrier completion) of
//     panel columns through a 4-stage shared-memory ring;
//   * the blocked right-looking Cholesky of a front built on it, outer block
//     WB = 128:  potrf of the diagonal block in one CTA (plus its inverse),
//     TRSM as a GEMM with that inverse, rank-128 SYRK/GEMM trailing update;
//   * the inverse of every big supernode's pivot block L11 (recursive block
//     merge, two batched GEMMs per level) so that the triangular solves of big
//     supernodes become two bandwidth-bound matrix-vector products spread over
//     many CTAs instead of a substitution chain inside one CTA;
//   * those multi-CTA forward / backward solve kernels.
//
// Replaces CHOLMOD's supernodal numeric factorisation and solve behind
// cholesky(Symmetric(Q,:L)) and F \ rhs (linear_system_solvers/julia.jl:34,99-113).
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

__device__ __forceinline__ bool stop_requested(const DeltaState* st) {
    const volatile int* d = &st->done;
    const volatile int* f = &st->fail;
    return (*d) | (*f);
}

// ---------------------------------------------------------------------------
// Tile engine:  acc(128 x 128) = sum_k A[m,k] * B[n,k]
//   A: m-contiguous (column-major M x K, element (m,k) at A[m + k*lda])
//   B: m-contiguous (B_KC = false, element (n,k) at B[n + k*ldb]) or
//      k-contiguous (B_KC = true,  element (n,k) at B[k + n*ldb])
// Requirements: A, B 16-byte aligned, lda/ldb even, K origin a multiple of 2.
// 256 threads = 8 warps (2 x 4), warp tile 64 x 32, thread accumulators 8 x 4 x 2.
// ---------------------------------------------------------------------------
constexpr int BM = 128, BN = 128, BK = 16, STAGES = 4;
constexpr int LDM = BM + 4;   // [BK][LDM]: fragment reads are bank-conflict free (LDM % 16 == 4)
constexpr int LDK = BK + 4;   // [BN][LDK]
constexpr int A_STAGE = BK * LDM;                                  // 2112 doubles
constexpr int B_STAGE = (BN * LDK > BK * LDM) ? BN * LDK : BK * LDM;  // 2560 doubles
constexpr int GEMM_THREADS = 256;

struct GemmSmem {
    double A[STAGES][A_STAGE];
    double B[STAGES][B_STAGE];
    unsigned long long full[STAGES];
};

template <bool B_KC>
__device__ __forceinline__ void gemm_issue(GemmSmem& sm, int stage, int kb, const double* Ag, int lda, int mrows,
                                           const double* Bg, int ldb, int nrows, int K, int lane) {
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
            bulk_g2s(&sm.B[stage][kk * LDM], Bg + (size_t)(k0 + kk) * ldb, bytesB, &sm.full[stage]);
    } else {
        const unsigned bytesB = (unsigned)(((nk + 1) & ~1) * 8);
        for (int n = lane; n < nrows; n += 32)
            bulk_g2s(&sm.B[stage][n * LDK], Bg + (size_t)n * ldb + k0, bytesB, &sm.full[stage]);
    }
}

template <bool B_KC>
__device__ __forceinline__ void gemm_mainloop(GemmSmem& sm, const double* Ag, int lda, int mrows,
                                              const double* Bg, int ldb, int nrows, int K,
                                              double (&acc)[8][4][2]) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp >> 2, wn = warp & 3;     // 2 x 4 warps
    const int q = lane & 3, g = lane >> 2;
#pragma unroll
    for (int a = 0; a < 8; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; s++) mbar_init(&sm.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int nkb = (K + BK - 1) / BK;
    if (warp == 0)
        for (int s = 0; s < STAGES - 1 && s < nkb; s++)
            gemm_issue<B_KC>(sm, s, s, Ag, lda, mrows, Bg, ldb, nrows, K, lane);
    for (int kb = 0; kb < nkb; kb++) {
        const int stage = kb % STAGES;
        if (warp == 0 && kb + STAGES - 1 < nkb)
            gemm_issue<B_KC>(sm, (kb + STAGES - 1) % STAGES, kb + STAGES - 1, Ag, lda, mrows, Bg, ldb, nrows, K, lane);
        mbar_wait(&sm.full[stage], (unsigned)((kb / STAGES) & 1));
        const double* As = sm.A[stage];
        const double* Bs = sm.B[stage];
        const int nk = min(BK, K - kb * BK);
        const int nsteps = (nk + 3) >> 2;
        for (int ks = 0; ks < nsteps; ks++) {
            const int kk = ks * 4 + q;
            const bool kvalid = kk < nk;
            double av[8], bv[4];
#pragma unroll
            for (int mt = 0; mt < 8; mt++) {
                double v = As[kk * LDM + wm * 64 + mt * 8 + g];
                av[mt] = kvalid ? v : 0.0;
            }
#pragma unroll
            for (int nt = 0; nt < 4; nt++) {
                double v = B_KC ? Bs[(wn * 32 + nt * 8 + g) * LDK + kk] : Bs[kk * LDM + wn * 32 + nt * 8 + g];
                bv[nt] = kvalid ? v : 0.0;
            }
#pragma unroll
            for (int mt = 0; mt < 8; mt++)
#pragma unroll
                for (int nt = 0; nt < 4; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], av[mt], bv[nt]);
        }
        __syncthreads();
    }
}

// coordinates of accumulator element (mt, nt, e) inside the 128 x 128 tile
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
// Diagonal block: Cholesky of the b x b block of outer step t in shared memory,
// then its inverse (one warp per column, forward substitution in registers).
// ---------------------------------------------------------------------------
constexpr int PT = 1024;             // threads of the diagonal-block kernel
constexpr int LDD = WB + 1;          // odd leading dimension: conflict-free row and column sweeps

__global__ void __launch_bounds__(PT)
chol_diag_kernel(DevSym S, const int* __restrict__ list, double* __restrict__ Lval, double* __restrict__ Xinv,
                 int t, DeltaState* st) {
    extern __shared__ double D[];          // LDD * WB + 2 * WB
    __shared__ int s_fail;
    if (stop_requested(st)) return;
    const Front d = get_front(S, list[blockIdx.x]);
    const int j0 = t * WB;
    if (j0 >= d.c) return;
    const int b = min(WB, d.c - j0);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double* colv = D + LDD * WB;
    double* rinv = colv + WB;
    double* base = Lval + d.loff + j0 + (size_t)j0 * d.ld;
    for (int idx = tid; idx < b * b; idx += PT) {
        const int i = idx % b, j = idx / b;
        D[i + j * LDD] = (i >= j) ? base[i + (size_t)j * d.ld] : 0.0;
    }
    if (tid == 0) s_fail = 0;
    __syncthreads();
    // right-looking Cholesky, one column per step
    for (int j = 0; j < b; j++) {
        const double dj = D[j + j * LDD];
        if (!(dj > 0.0)) {       // pivot <= 0 or NaN: not positive definite (julia.jl:39-41)
            if (tid == 0) st->fail = 1;
            return;
        }
        const double ljj = sqrt(dj);
        __syncthreads();
        double* cj = D + j * LDD;
        for (int i = j + 1 + tid; i < b; i += PT) cj[i] = cj[i] / ljj;
        if (tid == 0) cj[j] = ljj;
        __syncthreads();
        for (int k = j + 1 + warp; k < b; k += PT / 32) {
            const double lkj = cj[k];
            double* ck = D + k * LDD;
            for (int i = k + lane; i < b; i += 32) ck[i] -= cj[i] * lkj;
        }
        __syncthreads();
    }
    for (int idx = tid; idx < b * b; idx += PT) {
        const int i = idx % b, j = idx / b;
        if (i >= j) base[i + (size_t)j * d.ld] = D[i + j * LDD];
    }
    for (int k = tid; k < b; k += PT) rinv[k] = 1.0 / D[k + k * LDD];
    __syncthreads();
    // inverse: column j of X = L^-1 solves L x = e_j; lane owns rows lane + 32*qq
    double* X = Xinv + d.xoff + j0 + (size_t)j0 * d.ldx;
    for (int j = warp; j < b; j += PT / 32) {
        double x[WB / 32];
#pragma unroll
        for (int qq = 0; qq < WB / 32; qq++) x[qq] = (lane + 32 * qq == j) ? 1.0 : 0.0;
#pragma unroll
        for (int qk = 0; qk < WB / 32; qk++) {
            if (qk * 32 + 31 < j || qk * 32 >= b) continue;
            for (int kk = 0; kk < 32; kk++) {
                const int k = qk * 32 + kk;
                if (k < j || k >= b) continue;
                const double xk = __shfl_sync(0xffffffffu, x[qk], kk) * rinv[k];
                if (lane == kk) x[qk] = xk;
                const double* Lk = D + k * LDD;
#pragma unroll
                for (int qq = 0; qq < WB / 32; qq++) {
                    const int i = lane + 32 * qq;
                    if (i > k && i < b) x[qq] -= Lk[i] * xk;
                }
            }
        }
#pragma unroll
        for (int qq = 0; qq < WB / 32; qq++) {
            const int i = lane + 32 * qq;
            if (i >= j && i < b) X[i + (size_t)j * d.ldx] = x[qq];
        }
    }
    (void)s_fail;
}

// rows below the diagonal block:  L21 = A21 * inv(L_kk)^T  as a tensor-core GEMM
__global__ void __launch_bounds__(GEMM_THREADS)
chol_trsm_kernel(DevSym S, const int* __restrict__ list, double* __restrict__ Lval,
                 const double* __restrict__ Xinv, int t, DeltaState* st) {
    extern __shared__ __align__(16) unsigned char smraw[];
    GemmSmem& sm = *reinterpret_cast<GemmSmem*>(smraw);
    if (stop_requested(st)) return;
    const Front d = get_front(S, list[blockIdx.y]);
    const int j0 = t * WB;
    if (j0 >= d.c) return;
    const int b = min(WB, d.c - j0);
    const int j1 = j0 + b;
    const int row0 = (j1 & ~1) + blockIdx.x * BM;
    if (row0 >= d.N) return;
    const int mrows = min(BM, d.N - row0);
    double* Ag = Lval + d.loff + row0 + (size_t)j0 * d.ld;
    const double* Bg = Xinv + d.xoff + j0 + (size_t)j0 * d.ldx;
    double acc[8][4][2];
    gemm_mainloop<false>(sm, Ag, d.ld, mrows, Bg, d.ldx, b, b, acc);
#pragma unroll
    for (int mt = 0; mt < 8; mt++) {
        const int i = row0 + acc_row(mt);
        if (i < j1 || i >= d.N) continue;
#pragma unroll
        for (int nt = 0; nt < 4; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int n = acc_col(nt, e);
                if (n < b) Lval[d.loff + i + (size_t)(j0 + n) * d.ld] = acc[mt][nt][e];
            }
    }
}

// trailing update  C -= L[:,blk] * L[:,blk]^T  over the lower 128 x 128 tiles of [j1, N)^2
__global__ void __launch_bounds__(GEMM_THREADS)
chol_syrk_kernel(DevSym S, const int* __restrict__ list, double* __restrict__ Lval, double* __restrict__ CB,
                 int t, DeltaState* st) {
    extern __shared__ __align__(16) unsigned char smraw[];
    GemmSmem& sm = *reinterpret_cast<GemmSmem*>(smraw);
    if (stop_requested(st)) return;
    const Front d = get_front(S, list[blockIdx.y]);
    const int j0 = t * WB;
    if (j0 >= d.c) return;
    const int b = min(WB, d.c - j0);
    const int j1 = j0 + b;
    const int j1e = j1 & ~1;
    const int rem = d.N - j1e;
    if (d.N - j1 <= 0) return;
    const int nt_ = (rem + BM - 1) / BM;
    const long long tp = blockIdx.x;
    if (tp >= (long long)nt_ * (nt_ + 1) / 2) return;
    int I = (int)((sqrt(8.0 * (double)tp + 1.0) - 1.0) * 0.5);
    while ((long long)I * (I + 1) / 2 > tp) I--;
    while ((long long)(I + 1) * (I + 2) / 2 <= tp) I++;
    const int J = (int)(tp - (long long)I * (I + 1) / 2);
    const int ri = j1e + I * BM, rj = j1e + J * BM;
    const double* Ag = Lval + d.loff + ri + (size_t)j0 * d.ld;
    const double* Bg = Lval + d.loff + rj + (size_t)j0 * d.ld;
    double acc[8][4][2];
    gemm_mainloop<false>(sm, Ag, d.ld, min(BM, d.N - ri), Bg, d.ld, min(BN, d.N - rj), b, acc);
#pragma unroll
    for (int nt = 0; nt < 4; nt++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int k = rj + acc_col(nt, e);
            if (k < j1 || k >= d.N) continue;
#pragma unroll
            for (int mt = 0; mt < 8; mt++) {
                const int i = ri + acc_row(mt);
                if (i >= d.N || i < k) continue;
                *front_elem(d, Lval, CB, i, k) -= acc[mt][nt][e];
            }
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
trtri_merge_kernel(DevSym S, const int* __restrict__ list, const double* __restrict__ Lval,
                   double* __restrict__ Xinv, double* __restrict__ Twork, int lvl, int phase,
                   const DeltaState* st) {
    extern __shared__ __align__(16) unsigned char smraw[];
    GemmSmem& sm = *reinterpret_cast<GemmSmem*>(smraw);
    if (stop_requested(st)) return;
    const Front d = get_front(S, list[blockIdx.y]);
    const int nsub = 1 << lvl;              // 128-blocks per half
    const int Sz = WB * nsub;
    const int per_pair = nsub * nsub;
    const int pair = blockIdx.x / per_pair;
    const int ij = blockIdx.x % per_pair;
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
        gemm_mainloop<true>(sm, Ag, d.ld, mrows, Bg, d.ldx, BN, a1 - col0, acc);
        out = Twork + d.xoff;
    } else {
        // X21[i,j] = - sum_{k in [a1, row0 + mrows)} X[i,k] * T[k,j]   (X22 lower)
        const int kend = min(a2, row0 + BM);
        const double* Ag = Xinv + d.xoff + row0 + (size_t)a1 * d.ldx;
        const double* Bg = Twork + d.xoff + a1 + (size_t)col0 * d.ldx;
        gemm_mainloop<true>(sm, Ag, d.ldx, mrows, Bg, d.ldx, BN, kend - a1, acc);
        out = Xinv + d.xoff;
    }
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
constexpr int WT = 256;
constexpr int SLAB = 32;      // rows per CTA in the row-oriented products
constexpr int KG = WT / 32;   // k-groups (warps)

__global__ void __launch_bounds__(WT)
wide_fwd_gather_kernel(DevSym S, const int* __restrict__ list, double* __restrict__ x, double* __restrict__ u) {
    const int s = list[blockIdx.x];
    const int first = S.sfirst[s];
    const int c = S.sfirst[s + 1] - first;
    const int64_t rp = S.rowptr[s];
    const int r = (int)(S.rowptr[s + 1] - rp);
    double* xs = x + first;
    double* us = u + rp;
    const int tid = threadIdx.x;
    for (int t = tid; t < r; t += WT) us[t] = 0.0;
    __syncthreads();
    for (int k = S.child_ptr[s]; k < S.child_ptr[s + 1]; k++) {
        const int ch = S.child_list[k];
        const int64_t rpc = S.rowptr[ch];
        const int rc = (int)(S.rowptr[ch + 1] - rpc);
        const int* __restrict__ relc = S.rel + rpc;
        const double* uc = u + rpc;
        for (int t = tid; t < rc; t += WT) {
            const int dst = relc[t];
            const double v = uc[t];
            if (dst < c) xs[dst] += v; else us[dst - c] += v;
        }
        __syncthreads();
    }
}

// xnew[i] = sum_{k <= i} X[i,k] * xs[k]   for the pivot rows of the slab
__global__ void __launch_bounds__(WT)
wide_fwd_tri_kernel(DevSym S, const int* __restrict__ list, const double* __restrict__ Xinv,
                    const double* __restrict__ x, double* __restrict__ xnew) {
    __shared__ double red[KG][SLAB];
    const Front d = get_front(S, list[blockIdx.y]);
    const int i0 = blockIdx.x * SLAB;
    if (i0 >= d.c) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = i0 + lane;
    const double* X = Xinv + d.xoff;
    const double* xs = x + d.first;
    const int kend = min(d.c, i0 + SLAB);
    double acc = 0.0;
    if (i < d.c) {
        int k = w;
        for (; k + 3 * KG < kend; k += 4 * KG) {
            const double v0 = X[i + (size_t)k * d.ldx], v1 = X[i + (size_t)(k + KG) * d.ldx];
            const double v2 = X[i + (size_t)(k + 2 * KG) * d.ldx], v3 = X[i + (size_t)(k + 3 * KG) * d.ldx];
            acc += v0 * xs[k] + v1 * xs[k + KG] + v2 * xs[k + 2 * KG] + v3 * xs[k + 3 * KG];
        }
        for (; k < kend; k += KG) acc += X[i + (size_t)k * d.ldx] * xs[k];   // upper part of X is zero
    }
    red[w][lane] = acc;
    __syncthreads();
    if (w == 0 && i < d.c) {
        double v = 0.0;
#pragma unroll
        for (int g = 0; g < KG; g++) v += red[g][lane];
        xnew[d.first + i] = v;
    }
}

// rows i < c: xs[i] = xnew[i];  rows i >= c: us[i-c] -= sum_k L[i,k] * xnew[k]
__global__ void __launch_bounds__(WT)
wide_fwd_upd_kernel(DevSym S, const int* __restrict__ list, const double* __restrict__ Lval,
                    double* __restrict__ x, const double* __restrict__ xnew, double* __restrict__ u) {
    __shared__ double red[KG][SLAB];
    const Front d = get_front(S, list[blockIdx.y]);
    const int i0 = blockIdx.x * SLAB;
    if (i0 >= d.N) return;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int i = i0 + lane;
    const double* xn = xnew + d.first;
    if (i0 + SLAB <= d.c) {          // pure pivot rows: publish the forward solution
        if (w == 0) x[d.first + i] = xn[i];
        return;
    }
    const double* L = Lval + d.loff;
    double acc = 0.0;
    if (i >= d.c && i < d.N) {
        int k = w;
        for (; k + 3 * KG < d.c; k += 4 * KG) {
            const double v0 = L[i + (size_t)k * d.ld], v1 = L[i + (size_t)(k + KG) * d.ld];
            const double v2 = L[i + (size_t)(k + 2 * KG) * d.ld], v3 = L[i + (size_t)(k + 3 * KG) * d.ld];
            acc += v0 * xn[k] + v1 * xn[k + KG] + v2 * xn[k + 2 * KG] + v3 * xn[k + 3 * KG];
        }
        for (; k < d.c; k += KG) acc += L[i + (size_t)k * d.ld] * xn[k];
    }
    red[w][lane] = acc;
    __syncthreads();
    if (w == 0 && i < d.N) {
        if (i < d.c) x[d.first + i] = xn[i];
        else {
            double v = 0.0;
#pragma unroll
            for (int g = 0; g < KG; g++) v += red[g][lane];
            u[S.rowptr[d.s] + (i - d.c)] -= v;
        }
    }
}

__global__ void __launch_bounds__(WT)
wide_bwd_gather_kernel(DevSym S, const int* __restrict__ list, const double* __restrict__ x,
                       double* __restrict__ u) {
    const int s = list[blockIdx.y];
    const int64_t rp = S.rowptr[s];
    const int r = (int)(S.rowptr[s + 1] - rp);
    const int t = blockIdx.x * WT + threadIdx.x;
    if (t < r) u[rp + t] = x[S.rowidx[rp + t]];
}

// xnew[k] = xs[k] - sum_t L[c+t, k] * us[t]      (one warp per pivot column)
__global__ void __launch_bounds__(WT)
wide_bwd_upd_kernel(DevSym S, const int* __restrict__ list, const double* __restrict__ Lval,
                    const double* __restrict__ x, double* __restrict__ xnew, const double* __restrict__ u) {
    const Front d = get_front(S, list[blockIdx.y]);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int k = blockIdx.x * KG + w;
    if (k >= d.c) return;
    const double* col = Lval + d.loff + d.c + (size_t)k * d.ld;
    const double* us = u + S.rowptr[d.s];
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int t = lane;
    for (; t + 96 < d.r; t += 128) {
        a0 += col[t] * us[t]; a1 += col[t + 32] * us[t + 32];
        a2 += col[t + 64] * us[t + 64]; a3 += col[t + 96] * us[t + 96];
    }
    for (; t < d.r; t += 32) a0 += col[t] * us[t];
    double acc = (a0 + a1) + (a2 + a3);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) xnew[d.first + k] = x[d.first + k] - acc;
}

// xs[k] = sum_{i >= k} X[i,k] * xnew[i]
__global__ void __launch_bounds__(WT)
wide_bwd_tri_kernel(DevSym S, const int* __restrict__ list, const double* __restrict__ Xinv,
                    double* __restrict__ x, const double* __restrict__ xnew) {
    const Front d = get_front(S, list[blockIdx.y]);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int k = blockIdx.x * KG + w;
    if (k >= d.c) return;
    const double* col = Xinv + d.xoff + (size_t)k * d.ldx;
    const double* xn = xnew + d.first;
    double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
    int i = (k & ~31) + lane;      // aligned start; entries above the diagonal are zero
    for (; i + 96 < d.c; i += 128) {
        a0 += col[i] * xn[i]; a1 += col[i + 32] * xn[i + 32];
        a2 += col[i + 64] * xn[i + 64]; a3 += col[i + 96] * xn[i + 96];
    }
    for (; i < d.c; i += 32) a0 += col[i] * xn[i];
    double acc = (a0 + a1) + (a2 + a3);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) x[d.first + k] = acc;
}

inline size_t diag_smem() { return (size_t)(LDD * WB + 2 * WB) * sizeof(double); }

}  // namespace

cudaError_t dense_configure() {
    cudaError_t e;
    e = cudaFuncSetAttribute(chol_diag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)diag_smem());
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(chol_trsm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GemmSmem));
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(chol_syrk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GemmSmem));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(trtri_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GemmSmem));
}

void launch_big_chol_level(const DevSym& S, const LevelPlan& L, const int* d_sched, double* Lval,
                           double* CB, double* Xinv, DeltaState* st_d, cudaStream_t st) {
    if (!L.big_count) return;
    const int* list = d_sched + L.big_begin;
    launch_big_extend_add(S, L, d_sched, Lval, CB, st_d, st);
    for (size_t t = 0; t < L.step_count.size(); t++) {
        const int cnt = L.step_count[t];
        if (cnt <= 0) break;
        const int maxN = L.step_maxN[t];
        chol_diag_kernel<<<cnt, PT, diag_smem(), st>>>(S, list, Lval, Xinv, (int)t, st_d);
        count_launch();
        const int rem = maxN - (int)t * WB;     // rows from the start of the block (upper bound)
        if (rem <= 0) continue;
        dim3 gt((rem + BM - 1) / BM + 1, cnt);
        chol_trsm_kernel<<<gt, GEMM_THREADS, sizeof(GemmSmem), st>>>(S, list, Lval, Xinv, (int)t, st_d);
        count_launch();
        const long long nt = (rem + BM - 1) / BM + 1;
        dim3 gu((unsigned)(nt * (nt + 1) / 2), cnt);
        chol_syrk_kernel<<<gu, GEMM_THREADS, sizeof(GemmSmem), st>>>(S, list, Lval, CB, (int)t, st_d);
        count_launch();
    }
}

void launch_trtri(const DevSym& S, const TrtriPlan& T, const int* d_sched, const double* Lval,
                  double* Xinv, double* Twork, const DeltaState* st_d, cudaStream_t st) {
    const int* list = d_sched + T.list_begin;
    for (size_t l = 0; l < T.level_count.size(); l++) {
        const int cnt = T.level_count[l];
        if (cnt <= 0) break;
        const int nsub = 1 << l;
        dim3 g((unsigned)(T.level_pairs[l] * nsub * nsub), cnt);
        for (int phase = 0; phase < 2; phase++) {
            trtri_merge_kernel<<<g, GEMM_THREADS, sizeof(GemmSmem), st>>>(S, list, Lval, Xinv, Twork, (int)l, phase, st_d);
            count_launch();
        }
    }
}

void launch_solve_wide_fwd(const DevSym& S, const LevelPlan& L, const int* d_sched, const double* Lval,
                           const double* Xinv, double* x, double* xnew, double* u, cudaStream_t st) {
    if (!L.big_count) return;
    const int* list = d_sched + L.big_begin;
    wide_fwd_gather_kernel<<<L.big_count, WT, 0, st>>>(S, list, x, u);
    dim3 g1((L.big_maxC + SLAB - 1) / SLAB, L.big_count);
    wide_fwd_tri_kernel<<<g1, WT, 0, st>>>(S, list, Xinv, x, xnew);
    dim3 g2((L.big_maxN + SLAB - 1) / SLAB, L.big_count);
    wide_fwd_upd_kernel<<<g2, WT, 0, st>>>(S, list, Lval, x, xnew, u);
    count_launch(3);
}

void launch_solve_wide_bwd(const DevSym& S, const LevelPlan& L, const int* d_sched, const double* Lval,
                           const double* Xinv, double* x, double* xnew, double* u, cudaStream_t st) {
    if (!L.big_count) return;
    const int* list = d_sched + L.big_begin;
    const int maxR = L.big_maxN;   // upper bound on r
    dim3 g0((maxR + WT - 1) / WT, L.big_count);
    wide_bwd_gather_kernel<<<g0, WT, 0, st>>>(S, list, x, u);
    dim3 g1((L.big_maxC + KG - 1) / KG, L.big_count);
    wide_bwd_upd_kernel<<<g1, WT, 0, st>>>(S, list, Lval, x, xnew, u);
    wide_bwd_tri_kernel<<<g1, WT, 0, st>>>(S, list, Xinv, x, xnew);
    count_launch(3);
}

}  // namespace opb
