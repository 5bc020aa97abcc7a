// Supernodal triangular solves  x = (L L')^-1 b  /  (L D L')^-1 b, level-scheduled
// over the supernodal elimination tree.  Replaces `F \ rhs` of
// linear_system_solvers/julia.jl:99-113.
//
// Forward: every supernode gathers its children's update vectors in ascending
// child order (no atomics), solves with its diagonal block and emits its own
// update vector u_s = -L21 * y_s (+ inherited part).  Backward: reads the
// already-final entries of its ancestors.  Each entry of L is read once per sweep.
#include "opb_internal.h"

namespace opb {

namespace {

constexpr int ST = 256;        // threads per CTA
constexpr int SB = 32;         // column block

constexpr int DLD = SB + 1;

// stage the b x b diagonal block at (j0, j0) of a panel in shared memory, plus 1/diag
template <int NT>
__device__ __forceinline__ void stage_diag(const double* __restrict__ panel, int ld, int j0, int b,
                                           double* Dd, double* rdiag, int mode) {
    for (int e = threadIdx.x; e < SB * SB; e += NT) {
        const int i = e & (SB - 1), j = e >> 5;
        if (i < b && j < b && i >= j) Dd[i + j * DLD] = panel[(j0 + i) + (size_t)(j0 + j) * ld];
    }
    __syncthreads();
    if (threadIdx.x < b) rdiag[threadIdx.x] = (mode == 0) ? 1.0 / Dd[threadIdx.x * (DLD + 1)] : 1.0;
    __syncthreads();
}

// NT threads per supernode (256 is what runs, see launch_cta_classes)
template <int NT>
__global__ void __launch_bounds__(NT)
fwd_kernel(DevSym S, const int* __restrict__ list, const double* __restrict__ Lval,
           double* __restrict__ x, double* __restrict__ u, int mode) {
    __shared__ double yb[SB];
    __shared__ double Dd[SB * DLD];
    __shared__ double rdiag[SB];
    const int s = list[blockIdx.x];
    const int first = S.sfirst[s];
    const int c = S.sfirst[s + 1] - first;
    const int64_t rp = S.rowptr[s];
    const int r = (int)(S.rowptr[s + 1] - rp);
    const int N = c + r, ld = ld_of(N);
    const double* __restrict__ panel = Lval + S.Loff[s];
    double* xs = x + first;
    double* us = u + rp;
    const int tid = threadIdx.x;
    // children's update vectors: every destination sums its sources (ascending child order)
    {
        const int64_t gb = rp + first;
        for (int dd = tid; dd < N; dd += NT) {
            const double acc = gather_dest(S, u, gb + dd);
            if (dd < c) xs[dd] += acc; else us[dd - c] = acc;
        }
    }
    __syncthreads();
    for (int j0 = 0; j0 < c; j0 += SB) {
        const int b = min(SB, c - j0);
        stage_diag<NT>(panel, ld, j0, b, Dd, rdiag, mode);
        if (tid < 32) {
            const int lane = tid;
            double xv = (lane < b) ? xs[j0 + lane] : 0.0;
            for (int q = 0; q < b; q++) {
                const double val = __shfl_sync(0xffffffffu, xv, q) * rdiag[q];
                if (lane == q) xv = val;
                else if (lane > q && lane < b) xv -= Dd[lane + q * DLD] * val;
            }
            if (lane < b) { xs[j0 + lane] = xv; yb[lane] = xv; }
        }
        __syncthreads();
        for (int i = j0 + b + tid; i < N; i += NT) {
            const double* pr = panel + i + (size_t)j0 * ld;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
            int q = 0;
            for (; q + 3 < b; q += 4) {
                a0 += pr[(size_t)q * ld] * yb[q];
                a1 += pr[(size_t)(q + 1) * ld] * yb[q + 1];
                a2 += pr[(size_t)(q + 2) * ld] * yb[q + 2];
                a3 += pr[(size_t)(q + 3) * ld] * yb[q + 3];
            }
            for (; q < b; q++) a0 += pr[(size_t)q * ld] * yb[q];
            const double acc = (a0 + a1) + (a2 + a3);
            if (i < c) xs[i] -= acc; else us[i - c] -= acc;
        }
        __syncthreads();
    }
}

template <int NT>
__global__ void __launch_bounds__(NT)
bwd_kernel(DevSym S, const int* __restrict__ list, const double* __restrict__ Lval,
           double* __restrict__ x, double* __restrict__ u, int mode) {
    __shared__ double yb[SB];
    __shared__ double Dd[SB * DLD];
    __shared__ double rdiag[SB];
    const int s = list[blockIdx.x];
    const int first = S.sfirst[s];
    const int c = S.sfirst[s + 1] - first;
    const int64_t rp = S.rowptr[s];
    const int r = (int)(S.rowptr[s + 1] - rp);
    const int N = c + r, ld = ld_of(N);
    const double* __restrict__ panel = Lval + S.Loff[s];
    double* xs = x + first;
    double* us = u + rp;
    const int* __restrict__ rows = S.rowidx + rp;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int t = tid; t < r; t += NT) us[t] = x[rows[t]];
    if (mode == 1)
        for (int j = tid; j < c; j += NT) xs[j] = xs[j] / panel[j + (size_t)j * ld];
    __syncthreads();
    const int nblk = (c + SB - 1) / SB;
    for (int blk = nblk - 1; blk >= 0; blk--) {
        const int j0 = blk * SB;
        const int b = min(SB, c - j0);
        for (int q = warp; q < b; q += NT / 32) {
            const double* col = panel + (size_t)(j0 + q) * ld;
            double a0 = 0.0, a1 = 0.0;
            int i = j0 + b + lane;
            for (; i + 32 < N; i += 64) {
                const double f0 = (i < c) ? xs[i] : us[i - c];
                const double f1 = (i + 32 < c) ? xs[i + 32] : us[i + 32 - c];
                a0 += col[i] * f0;
                a1 += col[i + 32] * f1;
            }
            for (; i < N; i += 32) a0 += col[i] * ((i < c) ? xs[i] : us[i - c]);
            double acc = a0 + a1;
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) yb[q] = xs[j0 + q] - acc;
        }
        stage_diag<NT>(panel, ld, j0, b, Dd, rdiag, mode);     // ends with a barrier: yb is complete too
        if (tid < 32) {
            double xv = (lane < b) ? yb[lane] : 0.0;
            for (int q = b - 1; q >= 0; q--) {
                const double val = __shfl_sync(0xffffffffu, xv, q) * rdiag[q];
                if (lane == q) xv = val;
                else if (lane < q) xv -= Dd[q + lane * DLD] * val;      // L[j0+q, j0+lane]
            }
            if (lane < b) xs[j0 + lane] = xv;
        }
        __syncthreads();
    }
}

// Tiny supernodes (class T32: c + r <= 32, the bulk of the lowest level): one WARP per supernode,
// lane l owns the rows l, l + 32, ... (R per lane; R = 1 is what runs) of the front.  The solution /
// update entries live in registers, the substitution runs on shuffles, every panel entry is read
// once; eight supernodes per CTA.
constexpr int TW = 8;      // warps (supernodes) per CTA

template <int R>
__device__ __forceinline__ double pick(const double (&v)[R], int j) {
    double r = v[0];
#pragma unroll
    for (int t = 1; t < R; t++) if (j == t) r = v[t];
    return r;
}

template <int R>
__global__ void __launch_bounds__(TW * 32)
fwd_small_kernel(DevSym S, const int* __restrict__ list, int count, const double* __restrict__ Lval,
                 double* __restrict__ x, double* __restrict__ u, int mode) {
    const int lane = threadIdx.x & 31;
    const int idx = blockIdx.x * TW + (threadIdx.x >> 5);
    if (idx >= count) return;
    const int s = list[idx];
    const int first = S.sfirst[s];
    const int c = S.sfirst[s + 1] - first;
    const int64_t rp = S.rowptr[s];
    const int r = (int)(S.rowptr[s + 1] - rp);
    const int N = c + r, ld = ld_of(N);
    const double* __restrict__ panel = Lval + S.Loff[s];
    double v[R];
#pragma unroll
    for (int j = 0; j < R; j++) {
        const int row = lane + 32 * j;
        v[j] = 0.0;
        if (row < N) {
            v[j] = gather_dest(S, u, rp + first + row);
            if (row < c) v[j] += x[first + row];
        }
    }
    for (int q = 0; q < c; q++) {
        const double* __restrict__ col = panel + (size_t)q * ld;
        double lq[R];
#pragma unroll
        for (int j = 0; j < R; j++) { const int row = lane + 32 * j; lq[j] = (row > q && row < N) ? col[row] : 0.0; }
        double val = __shfl_sync(0xffffffffu, pick<R>(v, q >> 5), q & 31);
        if (mode == 0) val = val / col[q];
#pragma unroll
        for (int j = 0; j < R; j++) {
            const int row = lane + 32 * j;
            if (row == q) v[j] = val; else v[j] -= lq[j] * val;     // lq = 0 above the diagonal
        }
    }
#pragma unroll
    for (int j = 0; j < R; j++) {
        const int row = lane + 32 * j;
        if (row < c) x[first + row] = v[j];
        else if (row < N) u[rp + row - c] = v[j];
    }
}

template <int R>
__global__ void __launch_bounds__(TW * 32)
bwd_small_kernel(DevSym S, const int* __restrict__ list, int count, const double* __restrict__ Lval,
                 double* __restrict__ x, int mode) {
    const int lane = threadIdx.x & 31;
    const int idx = blockIdx.x * TW + (threadIdx.x >> 5);
    if (idx >= count) return;
    const int s = list[idx];
    const int first = S.sfirst[s];
    const int c = S.sfirst[s + 1] - first;
    const int64_t rp = S.rowptr[s];
    const int r = (int)(S.rowptr[s + 1] - rp);
    const int N = c + r, ld = ld_of(N);
    const double* __restrict__ panel = Lval + S.Loff[s];
    double v[R];
#pragma unroll
    for (int j = 0; j < R; j++) {
        const int row = lane + 32 * j;
        v[j] = 0.0;
        if (row < c) {
            v[j] = x[first + row];
            if (mode == 1) v[j] = v[j] / panel[row + (size_t)row * ld];
        } else if (row < N) {
            v[j] = x[S.rowidx[rp + row - c]];          // the ancestors' entries are final
        }
    }
    for (int q = c - 1; q >= 0; q--) {
        const double* __restrict__ col = panel + (size_t)q * ld;
        double dot = 0.0;
#pragma unroll
        for (int j = 0; j < R; j++) { const int row = lane + 32 * j; if (row > q && row < N) dot += col[row] * v[j]; }
        for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        const double dq = (mode == 0) ? col[q] : 1.0;
#pragma unroll
        for (int j = 0; j < R; j++) if (lane + 32 * j == q) v[j] = (v[j] - dot) / dq;
    }
#pragma unroll
    for (int j = 0; j < R; j++) { const int row = lane + 32 * j; if (row < c) x[first + row] = v[j]; }
}

template <int R>
void launch_small(bool forward, const DevSym& S, const int* list, int count, const double* Lval, double* x,
                  double* u, int mode, cudaStream_t st) {
    if (!count) return;
    const unsigned g = (unsigned)((count + TW - 1) / TW);
    if (forward) fwd_small_kernel<R><<<g, TW * 32, 0, st>>>(S, list, count, Lval, x, u, mode);
    else bwd_small_kernel<R><<<g, TW * 32, 0, st>>>(S, list, count, Lval, x, mode);
    count_launch();
}
// Only the tiny class runs a warp per supernode: for the wider classes (R = 2..5 rows per lane) the
// serial substitution of a single warp over up to ~100 columns was measured slower than the
// CTA-per-supernode kernels with their 32-column blocks (C3: 2.28 vs 1.69 ms per solve pair).
void launch_small_classes(bool forward, const DevSym& S, const LevelPlan& L, const int* d_sched,
                          const double* Lval, double* x, double* u, int mode, cudaStream_t st) {
    launch_small<1>(forward, S, d_sched + L.begin[FC_T32], L.count[FC_T32], Lval, x, u, mode, st);
}

template <int NT>
void launch_cta(bool forward, const DevSym& S, const int* list, int count, const double* Lval, double* x,
                double* u, int mode, cudaStream_t st) {
    if (count <= 0) return;
    if (forward) fwd_kernel<NT><<<count, NT, 0, st>>>(S, list, Lval, x, u, mode);
    else bwd_kernel<NT><<<count, NT, 0, st>>>(S, list, Lval, x, u, mode);
    count_launch();
}
// CTA-per-supernode classes of a level (`solo` supernodes starting at the S64 class): ONE launch
// with 256 threads.  Per-class thread counts (64 / 128 / 256 in three launches) were measured
// slower: the extra launches per level cost more than the higher residency gains (C3: 2.02 vs
// 1.67 ms per solve pair).
void launch_cta_classes(bool forward, const DevSym& S, const LevelPlan& L, const int* d_sched, int solo,
                        const double* Lval, double* x, double* u, int mode, cudaStream_t st) {
    launch_cta<ST>(forward, S, d_sched + L.begin[FC_S64], solo, Lval, x, u, mode, st);
}

__global__ void permute_in_kernel(const double* __restrict__ b, const int* __restrict__ perm,
                                  double* __restrict__ x, int n) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) x[k] = b[perm[k]];
}

__global__ void permute_out_kernel(const double* __restrict__ x, const int* __restrict__ perm,
                                   double* __restrict__ dst, int n, int accumulate) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < n) {
        int o = perm[k];
        dst[o] = accumulate ? dst[o] + x[k] : x[k];
    }
}

}  // namespace

cudaError_t solve_configure() { return cudaSuccess; }

void launch_solve(const DevSym& S, const std::vector<LevelPlan>& plan, const int* d_sched,
                  const double* Lval, const double* Xinv, double* x, double* xnew, double* u, int fmode,
                  const ShardCtx* shard, const int* colowner, const SideStream* side, cudaStream_t st) {
    // BIG supernodes take the multi-CTA path through inv(L11) (kernels_dense.cu), the other classes
    // (front or panel fits in shared memory) are solved by one CTA each; LDL': L11 is unit lower and
    // the pivots are applied once per supernode at the start of its backward step.  (With the
    // scalar LDL' path of option "ldlt_scalar" every supernode is solved by one CTA.)
    // Sharded instance: a rank runs the supernodes it owns.  Forward, a level whose supernodes
    // have children on other ranks waits at a barrier and reads those children's update vectors
    // from the owners' HBM; backward, the owner of a top supernode pushes its part of the solution
    // to every peer before the ranks below continue, and at the end every rank publishes the
    // columns it owns.
    const bool wide = fmode != FMODE_LDLT_SCALAR;      // LDL' on the tensor path has inv(L11) too
    const int mode = fmode == FMODE_CHOLESKY ? 0 : 1;   // what the kernels distinguish: Cholesky / LDL'
    // The supernodes of one level are independent of each other: where a level has both
    // CTA-per-supernode classes (latency-bound) and BIG supernodes (bandwidth-bound), the two
    // groups run concurrently on two streams and meet again before the next level.
    if (shard) side = nullptr;
    for (size_t l = 0; l < plan.size(); l++) {
        const LevelPlan& L = plan[l];
        if (shard && L.barrier_before) launch_shard_barrier(*shard, L.barrier_mask, st);
        // warp per supernode for the tiny class, CTA per supernode above
        const int solo = (wide ? L.all_count - L.count[FC_BIG] : L.all_count) - L.count[FC_T32];
        const bool par = side && wide && L.count[FC_BIG] > 0 && (solo + L.count[FC_T32]) > 0;
        cudaStream_t s2 = par ? side->stream : st;
        if (par) { cudaEventRecord(side->fork, st); cudaStreamWaitEvent(s2, side->fork, 0); }
        launch_small_classes(true, S, L, d_sched, Lval, x, u, mode, s2);
        launch_cta_classes(true, S, L, d_sched, solo, Lval, x, u, mode, s2);
        if (wide) launch_solve_wide_fwd(S, L, d_sched, Lval, Xinv, x, xnew, u, st);
        if (par) { cudaEventRecord(side->join, s2); cudaStreamWaitEvent(st, side->join, 0); }
    }
    for (size_t l = plan.size(); l-- > 0;) {
        const LevelPlan& L = plan[l];
        const int solo = (wide ? L.all_count - L.count[FC_BIG] : L.all_count) - L.count[FC_T32];
        const bool par = side && wide && L.count[FC_BIG] > 0 && (solo + L.count[FC_T32]) > 0;
        cudaStream_t s2 = par ? side->stream : st;
        if (par) { cudaEventRecord(side->fork, st); cudaStreamWaitEvent(s2, side->fork, 0); }
        launch_small_classes(false, S, L, d_sched, Lval, x, u, mode, s2);
        launch_cta_classes(false, S, L, d_sched, solo, Lval, x, u, mode, s2);
        if (wide) launch_solve_wide_bwd(S, L, d_sched, Lval, Xinv, x, xnew, u, mode, st);
        if (par) { cudaEventRecord(side->join, s2); cudaStreamWaitEvent(st, side->join, 0); }
        if (shard) {
            launch_push_supernodes(S, d_sched + L.push_begin, L.push_count, L.push_maxc, x, st);
            if (L.barrier_before) launch_shard_barrier(*shard, L.barrier_mask, st);
        }
    }
    if (shard) {
        launch_push_owned(S, colowner, x, st);
        launch_shard_barrier(*shard, shard_all(*shard), st);
    }
}

void launch_permute_in(const double* b, const int* perm, double* x, int n, cudaStream_t st) {
    permute_in_kernel<<<(n + 255) / 256, 256, 0, st>>>(b, perm, x, n);
    count_launch();
}
void launch_permute_out_add(const double* x, const int* perm, double* dst, int n, int accumulate, cudaStream_t st) {
    permute_out_kernel<<<(n + 255) / 256, 256, 0, st>>>(x, perm, dst, n, accumulate);
    count_launch();
}

// Force-load every kernel of this translation unit (CUDA loads kernels lazily, and a load may
// synchronise the context: that must not happen while another stream waits in a cross-rank barrier).
cudaError_t preload_solve() {
    cudaFuncAttributes a;
    cudaError_t e;
    e = cudaFuncGetAttributes(&a, fwd_kernel<ST>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, fwd_small_kernel<1>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, bwd_small_kernel<1>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, bwd_kernel<ST>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, permute_in_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, permute_out_kernel); if (e != cudaSuccess) return e;
    return cudaSuccess;
}

}  // namespace opb
