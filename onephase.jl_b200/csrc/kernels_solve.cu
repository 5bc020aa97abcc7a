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

__global__ void __launch_bounds__(ST)
fwd_kernel(DevSym S, const int* __restrict__ list, const double* __restrict__ Lval,
           double* __restrict__ x, double* __restrict__ u, int mode) {
    __shared__ double yb[SB];
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
    for (int t = tid; t < r; t += ST) us[t] = 0.0;
    __syncthreads();
    for (int k = S.child_ptr[s]; k < S.child_ptr[s + 1]; k++) {
        const int ch = S.child_list[k];
        const int64_t rpc = S.rowptr[ch];
        const int rc = (int)(S.rowptr[ch + 1] - rpc);
        const int* __restrict__ relc = S.rel + rpc;
        const double* uc = u + rpc;
        for (int t = tid; t < rc; t += ST) {
            const int dst = relc[t];
            const double v = uc[t];
            if (dst < c) xs[dst] += v; else us[dst - c] += v;
        }
        __syncthreads();
    }
    for (int j0 = 0; j0 < c; j0 += SB) {
        const int b = min(SB, c - j0);
        if (tid < 32) {
            const int lane = tid;
            double xv = (lane < b) ? xs[j0 + lane] : 0.0;
            for (int q = 0; q < b; q++) {
                double lpq = (lane >= q && lane < b) ? panel[(j0 + lane) + (size_t)(j0 + q) * ld] : 0.0;
                double dq = __shfl_sync(0xffffffffu, lpq, q);
                double xq = __shfl_sync(0xffffffffu, xv, q);
                double val = (mode == 0) ? xq / dq : xq;
                if (lane == q) xv = val;
                else if (lane > q) xv -= lpq * val;
            }
            if (lane < b) { xs[j0 + lane] = xv; yb[lane] = xv; }
        }
        __syncthreads();
        for (int i = j0 + b + tid; i < N; i += ST) {
            double acc = 0.0;
            const double* pr = panel + i + (size_t)j0 * ld;
            for (int q = 0; q < b; q++) acc += pr[(size_t)q * ld] * yb[q];
            if (i < c) xs[i] -= acc; else us[i - c] -= acc;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(ST)
bwd_kernel(DevSym S, const int* __restrict__ list, const double* __restrict__ Lval,
           double* __restrict__ x, double* __restrict__ u, int mode) {
    __shared__ double yb[SB];
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
    for (int t = tid; t < r; t += ST) us[t] = x[rows[t]];
    if (mode == 1)
        for (int j = tid; j < c; j += ST) xs[j] = xs[j] / panel[j + (size_t)j * ld];
    __syncthreads();
    const int nblk = (c + SB - 1) / SB;
    for (int blk = nblk - 1; blk >= 0; blk--) {
        const int j0 = blk * SB;
        const int b = min(SB, c - j0);
        for (int q = warp; q < b; q += ST / 32) {
            const double* col = panel + (size_t)(j0 + q) * ld;
            double acc = 0.0;
            for (int i = j0 + b + lane; i < N; i += 32) {
                const double f = (i < c) ? xs[i] : us[i - c];
                acc += col[i] * f;
            }
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) yb[q] = xs[j0 + q] - acc;
        }
        __syncthreads();
        if (tid < 32) {
            double xv = (lane < b) ? yb[lane] : 0.0;
            for (int q = b - 1; q >= 0; q--) {
                // L[j0+q, j0+lane], lane <= q
                double lql = (lane <= q) ? panel[(j0 + q) + (size_t)(j0 + lane) * ld] : 0.0;
                double dq = __shfl_sync(0xffffffffu, lql, q);
                double xq = __shfl_sync(0xffffffffu, xv, q);
                double val = (mode == 0) ? xq / dq : xq;
                if (lane == q) xv = val;
                else if (lane < q) xv -= lql * val;
            }
            if (lane < b) xs[j0 + lane] = xv;
        }
        __syncthreads();
    }
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
                  const double* Lval, const double* Xinv, double* x, double* xnew, double* u, int mode,
                  cudaStream_t st) {
    // Cholesky: BIG supernodes take the multi-CTA path through inv(L11) (kernels_dense.cu), the
    // other classes (front or panel fits in shared memory) are solved by one CTA each;
    // LDL': every supernode is solved by one CTA.
    const bool wide = (mode == 0);
    for (size_t l = 0; l < plan.size(); l++) {
        const LevelPlan& L = plan[l];
        const int solo = wide ? L.all_count - L.count[FC_BIG] : L.all_count;
        if (solo) { fwd_kernel<<<solo, ST, 0, st>>>(S, d_sched + L.all_begin, Lval, x, u, mode); count_launch(); }
        if (wide) launch_solve_wide_fwd(S, L, d_sched, Lval, Xinv, x, xnew, u, st);
    }
    for (size_t l = plan.size(); l-- > 0;) {
        const LevelPlan& L = plan[l];
        const int solo = wide ? L.all_count - L.count[FC_BIG] : L.all_count;
        if (solo) { bwd_kernel<<<solo, ST, 0, st>>>(S, d_sched + L.all_begin, Lval, x, u, mode); count_launch(); }
        if (wide) launch_solve_wide_bwd(S, L, d_sched, Lval, Xinv, x, xnew, u, st);
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

}  // namespace opb
