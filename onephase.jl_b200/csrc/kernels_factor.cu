// Numeric supernodal (multifrontal) factorisation of M + delta*I on the device.
// Replaces cholesky(Symmetric(Q,:L)) / ldlt(...) of linear_system_solvers/julia.jl:34,52.
//
// Layout: supernode s owns a dense panel (c+r) x c (column-major, ld = c+r) in
// Lval at Loff[s] and an update block r x r (lower part, ld = r) in CB at
// CBoff[s].  A level of the supernodal elimination tree is processed by
//   * front_small_kernel : one CTA per front, whole front in shared memory
//   * kernels_dense.cu   : fronts that do not fit -- panels in shared memory or blocked in HBM on
//                          the FP64 tensor pipe, Cholesky and LDL' alike (launch_wide_chol_level)
//   * big_* kernels      : the first, scalar LDL' path of those fronts (32-column blocks, batched
//                          over the fronts of the level); kept behind option "ldlt_scalar" for A/B runs
// Children's update blocks are added into the parent front by the CTA that owns
// the destination rows, children in ascending order: no atomics, the summation
// order is fixed, so the PD decision is reproducible run to run.
//
// A pivot <= 0 or NaN (CHOLMOD's "not positive definite", julia.jl:39-41) sets
// DeltaState::fail on the device; all later kernels of the attempt exit at once.
#include "opb_internal.h"

namespace opb {

namespace {

__device__ __forceinline__ bool stop_requested(const DeltaState* st) {
    const volatile int* d = &st->done;
    const volatile int* f = &st->fail;
    return (*d) | (*f);
}

struct FrontDims {
    int s, first, c, r, N, ld;
    int64_t loff, cboff;
};

__device__ __forceinline__ FrontDims front_dims(const DevSym& S, int s) {
    FrontDims d;
    d.s = s;
    d.first = S.sfirst[s];
    d.c = S.sfirst[s + 1] - d.first;
    d.r = (int)(S.rowptr[s + 1] - S.rowptr[s]);
    d.N = d.c + d.r;
    d.ld = ld_of(d.N);
    d.loff = S.Loff[s];
    d.cboff = S.CBoff[s];
    return d;
}

// Dense factorisation of the leading c columns of the N x N shared-memory front F
// (column-major, ld = N), right-looking.  mode 0: Cholesky, mode 1: LDL' (unit L,
// D on the diagonal).  colv: N doubles of scratch.  Returns false on a bad pivot.
template <int THREADS>
__device__ bool factor_in_smem(double* F, double* colv, int N, int c, int mode) {
    const int tid = threadIdx.x;
    const int tx = tid & 31, ty = tid >> 5;
    constexpr int NW = THREADS / 32;
    for (int j = 0; j < c; j++) {
        const double d = F[j + (size_t)j * N];
        double* cj = F + (size_t)j * N;
        if (mode == 0) {
            if (!(d > 0.0)) return false;
            const double ljj = sqrt(d);
            __syncthreads();   // everyone has read d
            for (int i = j + 1 + tid; i < N; i += THREADS) cj[i] = cj[i] / ljj;
            if (tid == 0) cj[j] = ljj;
            __syncthreads();
            for (int k = j + 1 + ty; k < N; k += NW) {
                const double lkj = cj[k];
                double* ck = F + (size_t)k * N;
                for (int i = k + tx; i < N; i += 32) ck[i] -= cj[i] * lkj;
            }
        } else {
            if (d == 0.0 || d != d) return false;
            __syncthreads();
            for (int i = j + 1 + tid; i < N; i += THREADS) { double v = cj[i]; colv[i] = v; cj[i] = v / d; }
            __syncthreads();
            for (int k = j + 1 + ty; k < N; k += NW) {
                const double wkj = colv[k];
                double* ck = F + (size_t)k * N;
                for (int i = k + tx; i < N; i += 32) ck[i] -= cj[i] * wkj;
            }
        }
        __syncthreads();
    }
    return true;
}

template <int THREADS>
__global__ void __launch_bounds__(THREADS)
front_small_kernel(DevSym S, const int* __restrict__ list, double* __restrict__ Lval,
                   double* __restrict__ CB, DeltaState* st) {
    extern __shared__ double F[];
    if (stop_requested(st)) return;
    const int tid = threadIdx.x;
    const FrontDims d = front_dims(S, list[blockIdx.x]);
    const int N = d.N, c = d.c, r = d.r;
    double* panel = Lval + d.loff;
    double* colv = F + (size_t)N * N;
    const int ld = d.ld;
    for (int idx = tid; idx < N * c; idx += THREADS) { const int i = idx % N, j = idx / N; F[idx] = panel[i + (size_t)j * ld]; }
    for (int idx = tid; idx < N * r; idx += THREADS) F[N * c + idx] = 0.0;
    __syncthreads();
    // extend-add the children's update blocks (ascending child order)
    for (int k = S.child_ptr[d.s]; k < S.child_ptr[d.s + 1]; k++) {
        const int ch = S.child_list[k];
        const int64_t rp = S.rowptr[ch];
        const int rc = (int)(S.rowptr[ch + 1] - rp);
        const int* __restrict__ relc = S.rel + rp;
        const double* __restrict__ cb = child_cb(S, CB, ch);
        // one flat loop over the child's block (not column by column: a column is a few dozen entries, and
        // a dependent round of global loads per column made this the longest phase of a small front);
        // distinct (i, j) of one child land on distinct entries of F
#pragma unroll 4
        for (int idx = tid; idx < rc * rc; idx += THREADS) {
            const int i = idx % rc, j = idx / rc;
            if (i >= j) F[relc[i] + (size_t)relc[j] * N] += cb[idx];
        }
        __syncthreads();
    }
    const int mode = st->mode;
    if (!factor_in_smem<THREADS>(F, colv, N, c, mode)) {
        if (tid == 0) st->fail = 1;
        return;
    }
    for (int idx = tid; idx < N * c; idx += THREADS) { const int i = idx % N, j = idx / N; panel[i + (size_t)j * ld] = F[idx]; }
    double* cbo = CB + d.cboff;
    for (int j = 0; j < r; j++)
        for (int i = j + tid; i < r; i += THREADS)
            cbo[i + (size_t)j * r] = F[(c + i) + (size_t)(c + j) * N];
}

// ---------------------------------------------------------------------------
// Big fronts: blocked right-looking factorisation in global memory.
// ---------------------------------------------------------------------------
__device__ __forceinline__ double* front_elem(const FrontDims& d, double* Lval, double* CB, int i, int j) {
    return (j < d.c) ? (Lval + d.loff + i + (size_t)j * d.ld)
                     : (CB + d.cboff + (i - d.c) + (size_t)(j - d.c) * d.r);
}

constexpr int EA_RB = 32;   // rows of the parent front owned by one CTA

__global__ void __launch_bounds__(256)
big_extend_add_kernel(DevSym S, const int* __restrict__ list, double* __restrict__ Lval,
                      double* __restrict__ CB, DeltaState* st) {
    if (stop_requested(st)) return;
    const FrontDims d = front_dims(S, list[blockIdx.y]);
    const int row0 = blockIdx.x * EA_RB;
    if (row0 >= d.N) return;
    const int row1 = min(d.N, row0 + EA_RB);
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    // zero the rows of the update block owned by this CTA
    {
        const int i = row0 + tx;
        if (i >= d.c && i < row1)
            for (int j = d.c + ty; j <= i; j += 8) CB[d.cboff + (i - d.c) + (size_t)(j - d.c) * d.r] = 0.0;
    }
    __syncthreads();
    for (int k = S.child_ptr[d.s]; k < S.child_ptr[d.s + 1]; k++) {
        const int ch = S.child_list[k];
        const int64_t rp = S.rowptr[ch];
        const int rc = (int)(S.rowptr[ch + 1] - rp);
        const int* __restrict__ relc = S.rel + rp;
        const double* __restrict__ cb = child_cb(S, CB, ch);
        // rows t of the child with row0 <= rel[t] < row1 (rel is ascending)
        int lo = 0, hi = rc;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (relc[mid] < row0) lo = mid + 1; else hi = mid; }
        const int t0 = lo;
        hi = rc;
        while (lo < hi) { int mid = (lo + hi) >> 1; if (relc[mid] < row1) lo = mid + 1; else hi = mid; }
        const int t1 = lo;
        for (int t = t0 + tx; t < t1; t += 32) {
            const int pi = relc[t];
            for (int j = ty; j <= t; j += 8) {
                const int pj = relc[j];
                *front_elem(d, Lval, CB, pi, pj) += cb[t + (size_t)j * rc];
            }
        }
        __syncthreads();
    }
}

// diagonal block of block column t: one CTA per front
__global__ void __launch_bounds__(256)
big_potrf_kernel(DevSym S, const int* __restrict__ list, double* __restrict__ Lval, int t, DeltaState* st) {
    __shared__ double D[NB * NB + NB];
    if (stop_requested(st)) return;
    const FrontDims d = front_dims(S, list[blockIdx.x]);
    const int j0 = t * NB;
    if (j0 >= d.c) return;
    const int b = min(NB, d.c - j0);
    double* base = Lval + d.loff + j0 + (size_t)j0 * d.ld;
    for (int idx = threadIdx.x; idx < b * b; idx += 256) {
        int i = idx % b, j = idx / b;
        D[idx] = base[i + (size_t)j * d.ld];
    }
    __syncthreads();
    if (!factor_in_smem<256>(D, D + NB * NB, b, b, st->mode)) {
        if (threadIdx.x == 0) st->fail = 1;
        return;
    }
    for (int idx = threadIdx.x; idx < b * b; idx += 256) {
        int i = idx % b, j = idx / b;
        if (i >= j) base[i + (size_t)j * d.ld] = D[idx];
    }
}

// rows below the diagonal block: X = A21 * L11^-T  (LDL': then scaled by D^-1)
__global__ void __launch_bounds__(128)
big_trsm_kernel(DevSym S, const int* __restrict__ list, double* __restrict__ Lval, int t, DeltaState* st) {
    __shared__ double D[NB * NB];
    if (stop_requested(st)) return;
    const FrontDims d = front_dims(S, list[blockIdx.y]);
    const int j0 = t * NB;
    if (j0 >= d.c) return;
    const int b = min(NB, d.c - j0);
    const int i0 = j0 + b + blockIdx.x * 128;
    if (i0 >= d.N) return;
    const double* base = Lval + d.loff + j0 + (size_t)j0 * d.ld;
    for (int idx = threadIdx.x; idx < b * b; idx += 128) {
        int i = idx % b, j = idx / b;
        D[i + j * NB] = base[i + (size_t)j * d.ld];
    }
    __syncthreads();
    const int i = i0 + threadIdx.x;
    if (i >= d.N) return;
    const int mode = st->mode;
    double* row = Lval + d.loff + i + (size_t)j0 * d.ld;
    double x[NB];
#pragma unroll
    for (int q = 0; q < NB; q++) x[q] = (q < b) ? row[(size_t)q * d.ld] : 0.0;
    if (mode == 0) {
#pragma unroll
        for (int q = 0; q < NB; q++) {
            if (q < b) {
                double acc = x[q];
#pragma unroll
                for (int p = 0; p < q; p++) acc -= x[p] * D[q + p * NB];
                x[q] = acc / D[q + q * NB];
            }
        }
    } else {
        // w_q = a_q - sum_{p<q} w_p l_qp ; l_q = w_q / d_q
#pragma unroll
        for (int q = 0; q < NB; q++) {
            if (q < b) {
                double acc = x[q];
#pragma unroll
                for (int p = 0; p < q; p++) acc -= x[p] * D[q + p * NB];
                x[q] = acc;
            }
        }
#pragma unroll
        for (int q = 0; q < NB; q++) if (q < b) x[q] = x[q] / D[q + q * NB];
    }
#pragma unroll
    for (int q = 0; q < NB; q++) if (q < b) row[(size_t)q * d.ld] = x[q];
}

constexpr int UT = 64;   // tile edge of the trailing update

// trailing update C -= P * P'  (LDL': P * D * P') over the lower tiles of the
// region [j1, N) x [j1, N), j1 = end of block column t.
__global__ void __launch_bounds__(256)
big_update_kernel(DevSym S, const int* __restrict__ list, double* __restrict__ Lval,
                  double* __restrict__ CB, int t, DeltaState* st) {
    __shared__ double As[NB * UT];
    __shared__ double Bs[NB * UT];
    if (stop_requested(st)) return;
    const FrontDims d = front_dims(S, list[blockIdx.y]);
    const int j0 = t * NB;
    if (j0 >= d.c) return;
    const int b = min(NB, d.c - j0);
    const int j1 = j0 + b;
    const int rem = d.N - j1;
    if (rem <= 0) return;
    const int nt = (rem + UT - 1) / UT;
    const long long npairs = (long long)nt * (nt + 1) / 2;
    long long tp = blockIdx.x;
    if (tp >= npairs) return;
    // tp -> (I, J), I >= J, row-major over the lower triangle
    int I = (int)((sqrt(8.0 * (double)tp + 1.0) - 1.0) * 0.5);
    while ((long long)I * (I + 1) / 2 > tp) I--;
    while ((long long)(I + 1) * (I + 2) / 2 <= tp) I++;
    const int J = (int)(tp - (long long)I * (I + 1) / 2);
    const int ri = j1 + I * UT, rj = j1 + J * UT;
    const int mode = st->mode;
    const double* pan = Lval + d.loff + (size_t)j0 * d.ld;
    for (int idx = threadIdx.x; idx < NB * UT; idx += 256) {
        const int i = idx % UT, p = idx / UT;
        double a = 0.0, bb = 0.0;
        if (p < b) {
            if (ri + i < d.N) a = pan[(ri + i) + (size_t)p * d.ld];
            if (rj + i < d.N) {
                bb = pan[(rj + i) + (size_t)p * d.ld];
                if (mode == 1) bb *= pan[(j0 + p) + (size_t)p * d.ld];
            }
        }
        As[idx] = a; Bs[idx] = bb;
    }
    __syncthreads();
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    double acc[4][4];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int q = 0; q < 4; q++) acc[a][q] = 0.0;
#pragma unroll 8
    for (int p = 0; p < NB; p++) {
        double av[4], bv[4];
#pragma unroll
        for (int a = 0; a < 4; a++) av[a] = As[p * UT + tx + 16 * a];
#pragma unroll
        for (int q = 0; q < 4; q++) bv[q] = Bs[p * UT + ty + 16 * q];
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int q = 0; q < 4; q++) acc[a][q] += av[a] * bv[q];
    }
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int k = rj + ty + 16 * q;
        if (k >= d.N) continue;
#pragma unroll
        for (int a = 0; a < 4; a++) {
            const int i = ri + tx + 16 * a;
            if (i >= d.N || i < k) continue;
            *front_elem(d, Lval, CB, i, k) -= acc[a][q];
        }
    }
}

__global__ void ldlt_inertia_kernel(const double* __restrict__ Lval, const int64_t* __restrict__ dpos,
                                    int n, DeltaState* st) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    int pos = 0, neg = 0, zer = 0, bad = 0;
    if (i < n) {
        const double dv = Lval[dpos[i]];
        const double tol = 1e-20;   // julia.jl:73
        if (dv != dv || isinf(dv)) bad = 1;
        else if (dv > tol) pos = 1;
        else if (dv < -tol) neg = 1;
        else zer = 1;
    }
    pos = __reduce_add_sync(0xffffffffu, pos);
    neg = __reduce_add_sync(0xffffffffu, neg);
    zer = __reduce_add_sync(0xffffffffu, zer);
    bad = __reduce_add_sync(0xffffffffu, bad);
    if ((threadIdx.x & 31) == 0) {
        if (pos) atomicAdd(&st->n_pos, pos);
        if (neg) atomicAdd(&st->n_neg, neg);
        if (zer) atomicAdd(&st->n_zero, zer);
        if (bad) atomicAdd(&st->n_bad, bad);
    }
}

inline size_t small_smem(int N) { return ((size_t)N * N + N) * sizeof(double); }

}  // namespace

cudaError_t factor_configure() {
    cudaError_t e = cudaFuncSetAttribute(front_small_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)small_smem(SMALL_N));
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(front_small_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)small_smem(FC_MAXN[FC_T32]));
}

void launch_factor_levels(const DevSym& S, const std::vector<LevelPlan>& plan, const int* d_sched,
                          double* Lval, double* CB, double* Xinv, DeltaState* st_d, int mode,
                          int outer_block, int cb_small_k, const ShardCtx* shard, const SideStream* side,
                          KernelTimer* timer, const std::vector<TrtriPlan>* trtri, double* Twork, cudaStream_t st) {
    unsigned prev_mask = 0;       // group of the previous level when it read peers' update blocks (see below)
    size_t next_trtri = 0;
    bool aux_busy = false;
    int lvl = -1;
    // pivot-block inverses of the levels factorised so far: on the auxiliary stream when there is one
    // and more levels follow (they hide behind those), otherwise in line
    auto flush_trtri = [&](bool last) {
        while (trtri && next_trtri < trtri->size() && ((*trtri)[next_trtri].after_level <= lvl || last)) {
            const TrtriPlan& T = (*trtri)[next_trtri++];
            if (side && side->aux && !last) {
                cudaEventRecord(side->aux_fork, st);
                cudaStreamWaitEvent(side->aux, side->aux_fork, 0);
                launch_trtri(S, T, d_sched, Lval, Xinv, Twork, st_d, side->aux);
                cudaEventRecord(side->aux_done, side->aux);
                aux_busy = true;
            } else {
                launch_trtri(S, T, d_sched, Lval, Xinv, Twork, st_d, st);
            }
        }
    };
    for (const LevelPlan& L : plan) {
        flush_trtri(false);
        lvl++;
        // Sharded instance: before a level with children on other ranks those children's update
        // blocks must be complete; AFTER such a level nobody may go on before every rank has
        // finished reading them, because the update-block arena reuses their memory from the
        // next level on (symbolic.cpp, level-lifetime allocator).  The barrier that ends the
        // attempt covers the last level.
        if (shard && prev_mask) launch_shard_barrier(*shard, prev_mask, st);
        if (shard && L.barrier_before && L.barrier_mask) launch_shard_barrier(*shard, L.barrier_mask, st);
        prev_mask = (L.barrier_before || L.split) ? L.barrier_mask : 0u;
        // children on other ranks (complete now): bulk copies of their update blocks into this rank's arena
        if (shard && L.pullcb_count) launch_pull_cb(S, d_sched + L.pullcb_begin, L.pullcb_count, L.pullcb_maxR, CB, st_d, st);
        if (timer && timer->phases) timer->put_mark(0, st);
        if (L.count[FC_T32]) {
            front_small_kernel<64><<<L.count[FC_T32], 64, small_smem(L.maxN[FC_T32]), st>>>(
                S, d_sched + L.begin[FC_T32], Lval, CB, st_d);
            count_launch();
        }
        for (int fc = FC_S64; fc <= FC_S152; fc++) {
            if (!L.count[fc]) continue;
            front_small_kernel<256><<<L.count[fc], 256, small_smem(L.maxN[fc]), st>>>(
                S, d_sched + L.begin[fc], Lval, CB, st_d);
            count_launch();
        }
        if (mode != FMODE_LDLT_SCALAR) {
            // panels in shared memory / blocked on the FP64 tensor pipe (kernels_dense.cu); LDL': the same
            // schedule with the pivots D applied inside the tile engine
            const int ldlt = mode == FMODE_LDLT ? 1 : 0;
            if (shard && L.split) {
                // Sharded instance, level with split fronts: the owners factorise the panels; then every rank
                // of a split front's range pulls the panel over NVLink and forms its share of the update
                // block's tiles, storing them into the owner's arena; nobody goes on before all tiles are in.
                launch_wide_chol_level(S, L, d_sched, Lval, CB, Xinv, st_d, ldlt, outer_block, cb_small_k, side, timer, 1, st);
                launch_shard_barrier(*shard, L.barrier_mask, st);
                launch_pull_panels(S, d_sched + L.help_begin, L.help_count, L.help_maxN, Lval, st_d, st);
                launch_wide_chol_level(S, L, d_sched, Lval, CB, Xinv, st_d, ldlt, outer_block, cb_small_k, side, timer, 2, st);
                launch_shard_barrier(*shard, L.barrier_mask, st);
                prev_mask = 0;      // that barrier already covers "nobody goes on while a peer still reads" 
            } else {
                launch_wide_chol_level(S, L, d_sched, Lval, CB, Xinv, st_d, ldlt, outer_block, cb_small_k, side, timer, 0, st);
            }
            continue;
        }
        if (!L.wide_count) continue;
        // LDL' fallback: scalar blocked path over every front that does not fit in shared memory
        const int* list = d_sched + L.wide_begin;
        dim3 gea((L.wide_maxN + EA_RB - 1) / EA_RB, L.wide_count);
        big_extend_add_kernel<<<gea, 256, 0, st>>>(S, list, Lval, CB, st_d);
        count_launch();
        const int nsteps = (L.wide_maxC + NB - 1) / NB;
        for (int t = 0; t < nsteps; t++) {
            big_potrf_kernel<<<L.wide_count, 256, 0, st>>>(S, list, Lval, t, st_d);
            count_launch();
            const int rem = L.wide_maxN - t * NB;   // upper bound on rows below
            if (rem <= 0) continue;
            dim3 gt((rem + 127) / 128, L.wide_count);
            big_trsm_kernel<<<gt, 128, 0, st>>>(S, list, Lval, t, st_d);
            count_launch();
            const long long nt = (rem + UT - 1) / UT;
            dim3 gu((unsigned)(nt * (nt + 1) / 2), L.wide_count);
            big_update_kernel<<<gu, 256, 0, st>>>(S, list, Lval, CB, t, st_d);
            count_launch();
        }
    }
    if (timer && timer->phases) timer->put_mark(3, st);
    flush_trtri(true);
    if (aux_busy) cudaStreamWaitEvent(st, side->aux_done, 0);
}

void launch_ldlt_inertia(const DevSym& S, const double* Lval, const int64_t* dpos, int n,
                         DeltaState* st_d, cudaStream_t st) {
    (void)S;
    ldlt_inertia_kernel<<<(n + 255) / 256, 256, 0, st>>>(Lval, dpos, n, st_d);
    count_launch();
}

// Force-load every kernel of this translation unit (CUDA loads kernels lazily, and a load may
// synchronise the context: that must not happen while another stream waits in a cross-rank barrier).
cudaError_t preload_factor() {
    cudaFuncAttributes a;
    cudaError_t e;
    e = cudaFuncGetAttributes(&a, front_small_kernel<64>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, front_small_kernel<256>); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, big_extend_add_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, big_potrf_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, big_trsm_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, big_update_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, ldlt_inertia_kernel); if (e != cudaSuccess) return e;
    return cudaSuccess;
}

}  // namespace opb
