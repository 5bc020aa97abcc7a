// Vector-level pieces of compute_direction_implementation! (schur.jl:89-128),
// the refinement residual of solver_schur_rhs (schur.jl:165-169) and
// update_kkt_error! (kkt_system_solver.jl:67-96).  All products use the cached
// matrices of the factorisation iterate: J by rows (CSR copy of the values), J by
// columns (the caller's CSC values) and the symmetric view of the lower-triangular
// H (eval.jl:221-230: L v + L' v - diag(L) v).
#include "opb_internal.h"

namespace opb {

namespace {

__device__ __forceinline__ void max_abs_to(unsigned long long* slot, double v) {
    // |v| as an ordered unsigned key; NaN (0x7ff8...) sorts above +inf so it propagates
    unsigned long long key = (unsigned long long)__double_as_longlong(fabs(v));
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    if ((threadIdx.x & 31) == 0 && key) atomicMax(slot, key);
}

__global__ void rhs_m_kernel(DirBuffers B) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k < B.m) B.tm[k] = B.primal_r[k] * B.sigma[k] + B.comp_r[k] / B.s[k];
}

// Sparse dot products of one row with a dense vector.  W = 1: one thread per row;
// W = 32: one warp per row (long rows, e.g. the dense Hessian of COPS elec), lanes stride
// the row and the partial sums are combined with shuffles (every lane gets the result).
template <int W>
__device__ __forceinline__ double wsum(double acc) {
    if (W == 32)
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    return acc;
}
template <int W>
__device__ __forceinline__ double jt_dot(const DirBuffers& B, int j, const double* v) {
    double acc = 0.0;
    const int l = (W == 32) ? (threadIdx.x & 31) : 0;
    for (int64_t p = B.Jp[j] + l; p < B.Jp[j + 1]; p += W) acc += B.Jv[p] * v[B.Jrow[p]];
    return wsum<W>(acc);
}
template <int W>
__device__ __forceinline__ double j_dot(const DirBuffers& B, int k, const double* v) {
    double acc = 0.0;
    const int l = (W == 32) ? (threadIdx.x & 31) : 0;
    for (int64_t q = B.Rp[k] + l; q < B.Rp[k + 1]; q += W) acc += B.Rval[q] * v[B.Rcol[q]];
    return wsum<W>(acc);
}
template <int W>
__device__ __forceinline__ double h_dot(const DirBuffers& B, int i, const double* v) {
    double acc = 0.0;
    const int l = (W == 32) ? (threadIdx.x & 31) : 0;
    for (int64_t q = B.Sp[i] + l; q < B.Sp[i + 1]; q += W) acc += B.Hv[B.Spos[q]] * v[B.Scol[q]];
    return wsum<W>(acc);
}
template <int W>
__device__ __forceinline__ int row_of() { return (int)((blockIdx.x * (unsigned)blockDim.x + threadIdx.x) / W); }
template <int W>
__device__ __forceinline__ bool writer() { return W == 1 || (threadIdx.x & 31) == 0; }

template <int W>
__global__ void rhs_n_kernel(DirBuffers B) {
    const int j = row_of<W>();
    if (j >= B.n) return;
    double v = B.dual_r[j] + jt_dot<W>(B, j, B.tm);
    if (writer<W>()) { B.b[j] = v; B.res[j] = v; B.dx[j] = 0.0; }
}

template <int W>
__global__ void res_m_kernel(DirBuffers B) {
    const int k = row_of<W>();
    if (k >= B.m) return;
    const double v = B.sigma[k] * j_dot<W>(B, k, B.dx);
    if (writer<W>()) B.tm[k] = v;
}

template <int W>
__global__ void res_n_kernel(DirBuffers B) {
    const int j = row_of<W>();
    if (j >= B.n) return;
    const double delta = B.st_d->delta;
    double jac_res = jt_dot<W>(B, j, B.tm);
    double hess_res = h_dot<W>(B, j, B.dx) + delta * B.dx[j];
    if (writer<W>()) B.res[j] = B.b[j] - (jac_res + hess_res);
}

__global__ void red_reset_kernel(unsigned long long* red) { if (threadIdx.x < 8) red[threadIdx.x] = 0ull; }

template <int W>
__global__ void recover_m_kernel(DirBuffers B) {
    const int k = row_of<W>();
    double eP = 0.0, eM = 0.0, rP = 0.0, rC = 0.0;
    if (k < B.m) {
        const double jd = j_dot<W>(B, k, B.dx);
        const double pr = B.primal_r[k], cr = B.comp_r[k], yk = B.y[k], sk = B.s[k];
        const double sym_p = pr + cr / yk;
        const double dy = -(jd - sym_p) * B.sigma[k];
        const double ds = jd - pr;
        if (writer<W>()) { B.dy[k] = dy; B.ds[k] = ds; }
        eP = jd - ds - pr;
        eM = sk * dy + yk * ds - cr;
        rP = pr; rC = cr;
    }
    max_abs_to(B.red + 1, eP);
    max_abs_to(B.red + 2, eM);
    max_abs_to(B.red + 4, rP);
    max_abs_to(B.red + 5, rC);
}

template <int W>
__global__ void recover_n_kernel(DirBuffers B) {
    const int j = row_of<W>();
    double eD = 0.0, rD = 0.0;
    if (j < B.n) {
        const double delta = B.st_d->delta;
        const double dxj = B.dx[j];
        const double delta_err = delta * dxj;
        const double H_err = h_dot<W>(B, j, B.dx);
        const double J_err = jt_dot<W>(B, j, B.dy);
        rD = B.dual_r[j];
        eD = (delta_err + H_err - J_err) - rD;
    }
    max_abs_to(B.red + 0, eD);
    max_abs_to(B.red + 3, rD);
}

__global__ void kkt_err_finish_kernel(DirBuffers B) {
    auto val = [&](int i) { return __longlong_as_double((long long)B.red[i]); };
    auto mx = [](double a, double b) { return (a != a || b != b) ? __longlong_as_double(0x7ff8000000000000ll) : (a > b ? a : b); };
    double eD = val(0), eP = val(1), eM = val(2);
    double overall = mx(mx(eD, eP), eM);
    double rhs_norm = mx(mx(val(3), val(4)), val(5));
    double* o = B.st_d->kkt_err;
    o[0] = eD; o[1] = eP; o[2] = eM; o[3] = overall; o[4] = rhs_norm; o[5] = overall / rhs_norm;
}

// ---- System_rhs on the device (kkt_system_solver/system_rhs.jl:57-73) -------------------------
// dual_r = -(grad - J'y + mu_t * (a_pen * J'1)) * (1 - eta_D), mu_t = mu * eta_mu, in the reference's
// operation order (eval.jl:59-63,136-142; the products J'y and J'1 accumulate over the rows ascending)
__global__ void sysrhs_n_kernel(DirBuffers B, const double* __restrict__ grad, double mu_t, double a_pen,
                                double one_minus_eta_D) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= B.n) return;
    double jty = 0.0, jt1 = 0.0;
    for (int64_t p = B.Jp[j]; p < B.Jp[j + 1]; p++) {
        const double a = B.Jv[p];
        jty = __dadd_rn(jty, __dmul_rn(a, B.y[B.Jrow[p]]));
        jt1 = __dadd_rn(jt1, a);                       // a * 1.0
    }
    const double gl = __dadd_rn(__dsub_rn(grad[j], jty), __dmul_rn(mu_t, __dmul_rn(a_pen, jt1)));
    B.tm2[j] = __dmul_rn(-gl, one_minus_eta_D);        // tm2 aliases the resident dual_r buffer
}
// primal_r = -(cons - s) * (1 - eta_P);  comp_r = mu_t - s .* y
__global__ void sysrhs_m_kernel(DirBuffers B, const double* __restrict__ cons, double* __restrict__ primal_r,
                                double* __restrict__ comp_r, double mu_t, double one_minus_eta_P) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= B.m) return;
    primal_r[k] = __dmul_rn(-__dsub_rn(cons[k], B.s[k]), one_minus_eta_P);
    comp_r[k] = __dsub_rn(mu_t, __dmul_rn(B.s[k], B.y[k]));
}

// ---- fraction-to-the-boundary quantities (line_search/frac_boundary.jl:3-35) -------------------
// red[0] = |dx|_inf, red[1] = |dy|_inf, red[2] = |ds|_inf
__global__ void step_norms_kernel(DirBuffers B) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    max_abs_to(B.red + 0, i < B.n ? B.dx[i] : 0.0);
    max_abs_to(B.red + 1, i < B.m ? B.dy[i] : 0.0);
    max_abs_to(B.red + 2, i < B.m ? B.ds[i] : 0.0);
}
// red[3] = max(1, max_i -ds_i / (s_i - lb_i)) as an ordered key, lb = frac_bd * min(s, |dx| * |dx|^ex)
// (lb_s, frac_boundary.jl:28-33; simple_max_step :35-40 returns 1 / that maximum)
__global__ void step_ratio_kernel(DirBuffers B, double frac_bd, double ex) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    const double ndx = __longlong_as_double((long long)B.red[0]);
    const double thres = ndx * pow(ndx, ex);
    double ratio = 1.0;
    if (k < B.m) {
        const double sk = B.s[k];
        const double lb = frac_bd * fmin(sk, thres);
        ratio = -B.ds[k] / (sk - lb);
        if (!(ratio > 1.0) && ratio == ratio) ratio = 1.0;       // keep NaN visible, clamp the rest at 1
    }
    // max over positive doubles (and NaN, which sorts above +inf as a key)
    unsigned long long key = (unsigned long long)__double_as_longlong(ratio);
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
        key = other > key ? other : key;
    }
    if ((threadIdx.x & 31) == 0) atomicMax(B.red + 3, key);
}

inline unsigned nblk(int64_t n) { return (unsigned)((n + 255) / 256); }

}  // namespace

void launch_system_rhs(const DirBuffers& B, const double* grad, const double* cons, double mu_t, double a_pen,
                       double eta_P, double eta_D, cudaStream_t st) {
    DirBuffers D = B;
    D.tm2 = const_cast<double*>(B.dual_r);
    sysrhs_n_kernel<<<nblk(B.n), 256, 0, st>>>(D, grad, mu_t, a_pen, 1.0 - eta_D);
    count_launch();
    if (B.m > 0) {
        sysrhs_m_kernel<<<nblk(B.m), 256, 0, st>>>(D, cons, const_cast<double*>(B.primal_r), const_cast<double*>(B.comp_r),
                                                 mu_t, 1.0 - eta_P);
        count_launch();
    }
}

void launch_step_bounds(const DirBuffers& B, double frac_bd, double ex, cudaStream_t st) {
    red_reset_kernel<<<1, 32, 0, st>>>(B.red);
    const int nm = B.n > B.m ? B.n : B.m;
    step_norms_kernel<<<nblk(nm), 256, 0, st>>>(B);
    step_ratio_kernel<<<nblk(B.m > 0 ? B.m : 1), 256, 0, st>>>(B, frac_bd, ex);
    count_launch(3);
}

// rows * W threads
inline unsigned nblkw(int64_t rows, int W) { return (unsigned)((rows * W + 255) / 256); }

void launch_schur_rhs(const DirBuffers& B, cudaStream_t st) {
    if (B.m > 0) { rhs_m_kernel<<<nblk(B.m), 256, 0, st>>>(B); count_launch(); }
    if (B.wide_n) rhs_n_kernel<32><<<nblkw(B.n, 32), 256, 0, st>>>(B);
    else rhs_n_kernel<1><<<nblkw(B.n, 1), 256, 0, st>>>(B);
    count_launch();
}

void launch_residual(const DirBuffers& B, cudaStream_t st) {
    if (B.m > 0) {
        if (B.wide_m) res_m_kernel<32><<<nblkw(B.m, 32), 256, 0, st>>>(B);
        else res_m_kernel<1><<<nblkw(B.m, 1), 256, 0, st>>>(B);
        count_launch();
    }
    if (B.wide_n) res_n_kernel<32><<<nblkw(B.n, 32), 256, 0, st>>>(B);
    else res_n_kernel<1><<<nblkw(B.n, 1), 256, 0, st>>>(B);
    count_launch();
}

void launch_recover_and_error(const DirBuffers& B, cudaStream_t st) {
    red_reset_kernel<<<1, 32, 0, st>>>(B.red);
    if (B.m > 0) {
        if (B.wide_m) recover_m_kernel<32><<<nblkw(B.m, 32), 256, 0, st>>>(B);
        else recover_m_kernel<1><<<nblkw(B.m, 1), 256, 0, st>>>(B);
    }
    if (B.wide_n) recover_n_kernel<32><<<nblkw(B.n, 32), 256, 0, st>>>(B);
    else recover_n_kernel<1><<<nblkw(B.n, 1), 256, 0, st>>>(B);
    kkt_err_finish_kernel<<<1, 1, 0, st>>>(B);
    count_launch(4);
}

// Force-load every kernel of this translation unit (CUDA loads kernels lazily, and a load may
// synchronise the context: that must not happen while another stream waits in a cross-rank barrier).
cudaError_t preload_vec() {
    cudaFuncAttributes a;
    cudaError_t e;
    e = cudaFuncGetAttributes(&a, rhs_m_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, rhs_n_kernel<1>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, rhs_n_kernel<32>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, res_m_kernel<1>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, res_m_kernel<32>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, res_n_kernel<1>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, res_n_kernel<32>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, red_reset_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, recover_m_kernel<1>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, recover_m_kernel<32>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, recover_n_kernel<1>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, recover_n_kernel<32>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, kkt_err_finish_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, sysrhs_n_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, sysrhs_m_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, step_norms_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, step_ratio_kernel); if (e != cudaSuccess) return e;
    return cudaSuccess;
}

}  // namespace opb
