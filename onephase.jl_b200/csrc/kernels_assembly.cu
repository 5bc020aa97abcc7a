// Assembly of the primal Schur complement M_L = tril(J' diag(y/s) J + H) with a
// structure-fixed gather map (replaces schur.jl:55 / eval.jl:85-87), scatter of
// M_L + delta*I into the supernodal panels (replaces the setindex! loop of
// schur.jl:75-77) and the device-side controller of the delta loop
// (delta_strategy.jl:37-114).
//
// Arithmetic follows SURVEY.md 9.2 bit for bit: sigma = y/s, t = fl(J[k,i]*sigma_k),
// p = fl(t*J[k,j]), accumulated over k ascending without FMA, then + H[i,j].
#include "opb_internal.h"

namespace opb {

namespace {

// One pass over J's entries: sigma = y ./ s, T = J .* sigma[row] (rounded separately: eval.jl:85-87), the
// CSR copy of J's values for the J*x products, and the reset of the diagonal-minimum scratch
__global__ void prep_kernel(const double* __restrict__ Jv, const int* __restrict__ Jrow, const int* __restrict__ Rpos,
                            const double* __restrict__ y, const double* __restrict__ s, double* __restrict__ sigma,
                            double* __restrict__ T, double* __restrict__ Rval, int64_t nnzJ, int m,
                            unsigned long long* scratch) {
    const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (p == 0) { scratch[0] = 0xffffffffffffffffull; scratch[1] = 0ull; scratch[2] = 0ull; }
    if (p < m) sigma[p] = y[p] / s[p];
    if (p < nnzJ) {
        const int r = Jrow[p];
        T[p] = __dmul_rn(Jv[p], y[r] / s[r]);
        Rval[p] = Jv[Rpos[p]];
    }
}

__global__ void assemble_M_kernel(const int64_t* __restrict__ pair_ptr, const int* __restrict__ pairA,
                                  const int* __restrict__ pairB, const int* __restrict__ hmap,
                                  const double* __restrict__ T, const double* __restrict__ Jv,
                                  const double* __restrict__ Hv, double* __restrict__ Mval, int64_t nnzM) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnzM) return;
    int64_t t0 = pair_ptr[e], t1 = pair_ptr[e + 1];
    double acc = 0.0;
    bool have = false;
    if (t0 < t1) {
        acc = __dmul_rn(T[pairA[t0]], Jv[pairB[t0]]);
        have = true;
        for (int64_t t = t0 + 1; t < t1; t++)
            acc = __dadd_rn(acc, __dmul_rn(T[pairA[t]], Jv[pairB[t]]));
    }
    int h = hmap[e];
    if (h >= 0) acc = have ? __dadd_rn(acc, Hv[h]) : Hv[h];
    Mval[e] = acc;
}

__device__ __forceinline__ unsigned long long dbl_sortable(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double sortable_dbl(unsigned long long b) {
    b = (b & 0x8000000000000000ull) ? (b & 0x7fffffffffffffffull) : ~b;
    return __longlong_as_double((long long)b);
}

// schur_diag = diag(Q) and its minimum (diag_min, kkt_system_solver.jl:291-294; NaN wins); the last CTA
// to finish publishes the result (scratch[2] = ticket counter, reset by prep_kernel)
__global__ void diag_extract_kernel(const int64_t* __restrict__ Mp, const double* __restrict__ Mval,
                                    double* __restrict__ sdiag, unsigned long long* scratch, DeltaState* st, int n) {
    int j = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long key = 0xffffffffffffffffull;
    unsigned nanf = 0;
    if (j < n) {
        double v = Mval[Mp[j]];
        sdiag[j] = v;
        if (v != v) nanf = 1; else key = dbl_sortable(v);
    }
    // warp reduce
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other < key ? other : key;
        nanf |= __shfl_xor_sync(0xffffffffu, nanf, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (key != 0xffffffffffffffffull) atomicMin(&scratch[0], key);
        if (nanf) atomicOr(&scratch[1], 1ull);
    }
    __shared__ int s_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_last = atomicAdd(&scratch[2], 1ull) == (unsigned long long)(gridDim.x - 1);
    __syncthreads();
    if (s_last && threadIdx.x == 0) {
        __threadfence();
        double v = sortable_dbl(atomicMin(&scratch[0], 0xffffffffffffffffull));
        if (atomicOr(&scratch[1], 0ull)) v = __longlong_as_double(0x7ff8000000000000ll);
        st->diag_min = v;
    }
}

__global__ void zero_kernel(double* __restrict__ p, int64_t n, const DeltaState* st) {
    if (st->done) return;
    int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t n2 = n >> 1;
    double2* p2 = reinterpret_cast<double2*>(p);
    for (int64_t k = i; k < n2; k += stride) p2[k] = make_double2(0.0, 0.0);
    if (i == 0 && (n & 1)) p[n - 1] = 0.0;
}

constexpr int64_t DIAG_FLAG = (int64_t)1 << 62;

__global__ void scatter_kernel(const double* __restrict__ Mval, const int64_t* __restrict__ amap,
                               double* __restrict__ Lval, int64_t nnzM, const DeltaState* st) {
    if (st->done) return;
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nnzM) return;
    int64_t a = amap[e];
    double v = Mval[e];
    if (a & DIAG_FLAG) { a &= ~DIAG_FLAG; v = v + st->delta; }  // Q[i,i] = schur_diag[i] + delta
    Lval[a] = v;
}

__device__ double next_delta(DeltaState* s) {
    if (s->it == 1) {
        if (s->delta_prev != 0.0) {
            double a = s->delta_min - s->tau, b = s->delta_prev * s->dec;
            if (a != a || b != b) return a + b;   // Julia's max propagates NaN
            return a > b ? a : b;   // max(DELTA_MIN - tau, get_delta(iter) * dec)
        }
        return s->delta_start - s->tau;
    }
    return s->delta * s->inc;
}

__global__ void ctl_init_kernel(DeltaState* s, double delta_prev, double delta_zero, double delta_min,
                                double delta_max, double delta_start, double inc, double dec,
                                int max_it, int mode) {
    s->delta_prev = delta_prev; s->delta_zero = delta_zero; s->delta_min = delta_min;
    s->delta_max = delta_max; s->delta_start = delta_start; s->inc = inc; s->dec = dec;
    s->max_it = max_it; s->mode = mode;
    s->done = 0; s->fail = 0; s->num_fac = 0; s->status = 2;
    s->n_pos = s->n_neg = s->n_zero = s->n_bad = 0;
    double tau = 1.5 * s->diag_min;   // tau = 1.5 * diag_min(kkt_solver)
    s->delta = delta_zero;
    if (tau > 0.0) { s->tau = 0.0; s->it = 0; }      // probe with delta = DELTA_ZERO first
    else { s->tau = tau; s->it = 1; s->delta = next_delta(s); }
}

__global__ void ctl_single_kernel(DeltaState* s, double delta, int mode) {
    s->delta = delta; s->mode = mode; s->done = 0; s->fail = 0; s->num_fac = 0; s->status = 2;
    s->it = -1; s->max_it = 0; s->tau = 0.0;
    s->n_pos = s->n_neg = s->n_zero = s->n_bad = 0;
}

__global__ void ctl_begin_kernel(DeltaState* s) {
    if (s->done) return;
    s->fail = 0;
}

__device__ void ctl_end_rule(DeltaState* s) {
    if (s->done) return;
    s->num_fac += 1;
    if (!s->fail) { s->done = 1; s->status = 1; return; }
    if (s->it < 0) { s->done = 1; s->status = 0; return; }   // single-shot attempt
    if (s->it >= 1 && s->delta > s->delta_max) { s->done = 1; s->status = 0; return; }
    s->it += 1;
    if (s->it > s->max_it) { s->done = 1; s->status = -1; return; }
    s->delta = next_delta(s);
}
__global__ void ctl_end_kernel(DeltaState* s) { ctl_end_rule(s); }
// the same rule as the last node of the body of a CUDA-graph WHILE node: the device decides
// whether the body (one factorisation attempt) runs again -- the host is not involved
__global__ void ctl_end_loop_kernel(DeltaState* s, cudaGraphConditionalHandle loop) {
    ctl_end_rule(s);
    cudaGraphSetConditional(loop, s->done ? 0u : 1u);
}

// eval_diag_J_T_J (utils/eval.jl:89-100): one thread per column, the reference's operation
// order (a^2 first, times diag_vals[row], summed over the rows ascending)
__global__ void diag_JtDJ_kernel(const int64_t* __restrict__ Jp, const int64_t* __restrict__ Ji,
                                 const double* __restrict__ Jx, const double* __restrict__ dv,
                                 double* __restrict__ out, int n, int base) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double acc = 0.0;
    for (int64_t p = Jp[i] - base; p < Jp[i + 1] - base; p++) {
        const double a = Jx[p];
        acc = __dadd_rn(acc, __dmul_rn(__dmul_rn(a, a), dv[Ji[p] - base]));
    }
    out[i] = acc;
}

inline unsigned nblk(int64_t n, int b) { return (unsigned)((n + b - 1) / b); }

}  // namespace

void launch_diag_JtDJ(const int64_t* Jp, const int64_t* Ji, const double* Jx, const double* dv, double* out,
                      int n, int base, cudaStream_t st) {
    diag_JtDJ_kernel<<<nblk(n, 128), 128, 0, st>>>(Jp, Ji, Jx, dv, out, n, base);
    count_launch();
}

static unsigned long long* diag_scratch(DeltaState* st_d) {
    // scratch lives right behind the DeltaState (allocated with 128 spare bytes): [0] sortable minimum
    // bits, [1] NaN flag, [2] ticket counter
    return reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(st_d) + ((sizeof(DeltaState) + 15) / 16) * 16);
}

void launch_prep(const double* Jv, const int* Jrow, const int* Rpos, const double* y, const double* s,
                 double* sigma, double* T, double* Rval, int64_t nnzJ, int m, DeltaState* st_d, cudaStream_t st) {
    const int64_t cnt = nnzJ > m ? nnzJ : m;
    prep_kernel<<<nblk(cnt > 0 ? cnt : 1, 256), 256, 0, st>>>(Jv, Jrow, Rpos, y, s, sigma, T, Rval, nnzJ, m, diag_scratch(st_d));
    count_launch();
}

void launch_assemble_M(const int64_t* pair_ptr, const int* pairA, const int* pairB, const int* hmap,
                       const double* T, const double* Jv, const double* Hv, double* Mval,
                       int64_t nnzM, cudaStream_t st) {
    assemble_M_kernel<<<nblk(nnzM, 256), 256, 0, st>>>(pair_ptr, pairA, pairB, hmap, T, Jv, Hv, Mval, nnzM);
    count_launch();
}

void launch_diag_extract(const int64_t* Mp, const double* Mval, double* sdiag, DeltaState* st_d,
                         int n, cudaStream_t st) {
    unsigned long long* scratch = diag_scratch(st_d);
    diag_extract_kernel<<<nblk(n > 0 ? n : 1, 256), 256, 0, st>>>(Mp, Mval, sdiag, scratch, st_d, n);
    count_launch();
}

void launch_scatter_fronts(const double* Mval, const int64_t* amap, const int64_t* dpos,
                           const double* sdiag, double* Lval, int64_t nnzL, int64_t nnzM, int n,
                           const DeltaState* st_d, int use_sdiag, cudaStream_t st) {
    (void)dpos; (void)sdiag; (void)n; (void)use_sdiag;
    int64_t want = (nnzL / 2 + 255) / 256;
    unsigned g = (unsigned)(want < 1 ? 1 : (want > 148 * 16 ? 148 * 16 : want));
    zero_kernel<<<g, 256, 0, st>>>(Lval, nnzL, st_d);
    count_launch();
    scatter_kernel<<<nblk(nnzM, 256), 256, 0, st>>>(Mval, amap, Lval, nnzM, st_d);
    count_launch();
}

void launch_ctl_begin(DeltaState* st_d, cudaStream_t st) { ctl_begin_kernel<<<1, 1, 0, st>>>(st_d); }
void launch_ctl_end(DeltaState* st_d, cudaStream_t st) { ctl_end_kernel<<<1, 1, 0, st>>>(st_d); }
void launch_ctl_end_loop(DeltaState* st_d, unsigned long long loop_handle, cudaStream_t st) {
    ctl_end_loop_kernel<<<1, 1, 0, st>>>(st_d, (cudaGraphConditionalHandle)loop_handle);
}
void launch_ctl_init(DeltaState* st_d, double delta_prev, double delta_zero, double delta_min,
                     double delta_max, double delta_start, double inc, double dec, int max_it,
                     int mode, cudaStream_t st) {
    ctl_init_kernel<<<1, 1, 0, st>>>(st_d, delta_prev, delta_zero, delta_min, delta_max, delta_start,
                                     inc, dec, max_it, mode);
                                     count_launch();
}
void launch_ctl_single(DeltaState* st_d, double delta, int mode, cudaStream_t st) {
    ctl_single_kernel<<<1, 1, 0, st>>>(st_d, delta, mode);
    count_launch();
}

// Force-load every kernel of this translation unit (CUDA loads kernels lazily, and a load may
// synchronise the context: that must not happen while another stream waits in a cross-rank barrier).
cudaError_t preload_assembly() {
    cudaFuncAttributes a;
    cudaError_t e;
    e = cudaFuncGetAttributes(&a, prep_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, assemble_M_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, diag_extract_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, zero_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, scatter_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, ctl_begin_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, ctl_end_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, ctl_end_loop_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, ctl_init_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, ctl_single_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, diag_JtDJ_kernel); if (e != cudaSuccess) return e;
    return cudaSuccess;
}

}  // namespace opb
