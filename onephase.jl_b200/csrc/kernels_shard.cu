// One KKT instance sharded over several B200s: the pieces that cross GPUs.
//
// Independent subtrees of the supernodal elimination tree are factorised and solved by
// different ranks (one process per GPU); only the top separators couple them.  Nothing here
// calls a collective library: the buffers that cross ranks (update blocks, forward update
// vectors, the solution vector, a few flag words) are peer-mapped (CUDA IPC over NVLink), the
// consuming kernels read their children's data straight from the owner's HBM (child_cb /
// child_u in opb_internal.h), and the ordering between ranks is a flag barrier on the stream:
//
//   shard_barrier_kernel   every rank publishes (epoch, fail) into each peer's flag array with
//                          system-scope release stores, then spins on its own array until every
//                          peer has published the same epoch.  The fail bits are OR-ed into the
//                          local DeltaState, so a non-positive pivot on any rank stops the
//                          attempt everywhere and the delta rule takes the same decision on
//                          every rank without a host round trip.
//   push_* kernels         backward sweep: the owner of a top supernode writes its part of the
//                          solution into every peer's vector; at the end every rank publishes
//                          the columns it owns, so the full solution is resident everywhere.
//
// The kernels are stream-ordered and capturable in CUDA graphs (the epoch lives in device memory).
#include "opb_internal.h"

namespace opb {

namespace {

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}


__global__ void __launch_bounds__(32)
shard_barrier_kernel(ShardCtx C, unsigned mask) {
    DeltaState* st = C.state;
    const int lane = threadIdx.x;
    const int myfail = *(volatile int*)&st->fail ? 1 : 0;
    __threadfence_system();                      // everything this rank wrote before the barrier
    int peerfail = 0;
    if (lane < C.world && lane != C.rank && ((mask >> lane) & 1u)) {
        // pairwise epoch: this rank and peer `lane` have met e - 1 times before (any subset of the ranks
        // may take part in a barrier, so there is no common count)
        const unsigned long long e = C.epoch[lane] + 1;
        C.epoch[lane] = e;
        const int slot = (int)(e & 1) * MAX_SHARD;
        st_release_sys(C.flags_peer[lane] + slot + C.rank, (e << 1) | (unsigned long long)myfail);
        const long long t0 = clock64();
        unsigned long long v;
        for (;;) {
            v = ld_acquire_sys(C.flags_local + slot + lane);
            if ((v >> 1) >= e) break;
            if (clock64() - t0 > C.timeout_clocks) { *C.error = 1; peerfail = 1; break; }
            __nanosleep(200);
        }
        if ((v >> 1) == e && (v & 1)) peerfail = 1;
    }
    peerfail = __any_sync(0xffffffffu, peerfail);
    if (lane == 0 && peerfail) st->fail = 1;
    __threadfence_system();
}

// x of the listed supernodes that are flagged `top` -> every peer's vector
__global__ void __launch_bounds__(256)
push_supernodes_kernel(DevSym S, const int* __restrict__ list, const double* __restrict__ x) {
    const int s = list[blockIdx.y];
    const int first = S.sfirst[s];
    const int c = S.sfirst[s + 1] - first;
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= c) return;
    const double v = x[first + k];
    for (int p = 0; p < S.world; p++)
        if (p != S.rank) S.x_peer[p][first + k] = v;
}

// every column this rank owns -> every peer's vector (end of the backward sweep)
__global__ void __launch_bounds__(256)
push_owned_kernel(DevSym S, const int* __restrict__ colowner, const double* __restrict__ x) {
    const int k = blockIdx.x * 256 + threadIdx.x;
    if (k >= S.n || colowner[k] != S.rank) return;
    const double v = x[k];
    for (int p = 0; p < S.world; p++)
        if (p != S.rank) S.x_peer[p][k] = v;
}

// split fronts of other ranks: the rows below the pivot block of the finished panel, from the owner's HBM
// into this rank's copy of the factor storage (same offsets: every rank keeps the global layout)
constexpr int PULL_ROWS = 256, PULL_COLS = 64;
__global__ void __launch_bounds__(PULL_ROWS)
pull_panels_kernel(DevSym S, const int* __restrict__ list, double* __restrict__ Lval, const DeltaState* st) {
    if (*(volatile const int*)&st->done | *(volatile const int*)&st->fail) return;
    const int s = list[blockIdx.z];
    const int c = S.sfirst[s + 1] - S.sfirst[s];
    const int N = c + (int)(S.rowptr[s + 1] - S.rowptr[s]);
    const int ld = ld_of(N);
    const int i = (c & ~1) + blockIdx.x * PULL_ROWS + threadIdx.x;
    const int j0 = blockIdx.y * PULL_COLS;
    if (i >= N || j0 >= c) return;
    const double* __restrict__ src = S.l_peer[S.owner[s]] + S.Loff[s] + i;
    double* __restrict__ dst = Lval + S.Loff[s] + i;
    const int j1 = min(c, j0 + PULL_COLS);
    int j = j0;
    for (; j + 4 <= j1; j += 4) {
        const double v0 = src[(size_t)j * ld], v1 = src[(size_t)(j + 1) * ld];
        const double v2 = src[(size_t)(j + 2) * ld], v3 = src[(size_t)(j + 3) * ld];
        dst[(size_t)j * ld] = v0; dst[(size_t)(j + 1) * ld] = v1; dst[(size_t)(j + 2) * ld] = v2; dst[(size_t)(j + 3) * ld] = v3;
    }
    for (; j < j1; j++) dst[(size_t)j * ld] = src[(size_t)j * ld];
}

// update blocks of children on other ranks: the lower triangle of the r x r block from the owner's arena into
// this rank's arena, same offset
__global__ void __launch_bounds__(PULL_ROWS)
pull_cb_kernel(DevSym S, const int* __restrict__ list, double* __restrict__ CB, const DeltaState* st) {
    if (*(volatile const int*)&st->done | *(volatile const int*)&st->fail) return;
    const int s = list[blockIdx.z];
    const int r = (int)(S.rowptr[s + 1] - S.rowptr[s]);
    const int i = blockIdx.x * PULL_ROWS + threadIdx.x;
    const int j0 = blockIdx.y * PULL_COLS;
    if (i >= r || j0 >= r || j0 > i) return;
    const double* __restrict__ src = S.cb_peer[S.owner[s]] + S.CBoff[s] + i;
    double* __restrict__ dst = CB + S.CBoff[s] + i;
    const int j1 = min(min(r, j0 + PULL_COLS), i + 1);       // lower triangle: columns <= row
    int j = j0;
    for (; j + 4 <= j1; j += 4) {
        const double v0 = src[(size_t)j * r], v1 = src[(size_t)(j + 1) * r];
        const double v2 = src[(size_t)(j + 2) * r], v3 = src[(size_t)(j + 3) * r];
        dst[(size_t)j * r] = v0; dst[(size_t)(j + 1) * r] = v1; dst[(size_t)(j + 2) * r] = v2; dst[(size_t)(j + 3) * r] = v3;
    }
    for (; j < j1; j++) dst[(size_t)j * r] = src[(size_t)j * r];
}

}  // namespace

void launch_pull_cb(const DevSym& S, const int* list, int count, int maxR, double* CB, const DeltaState* st_d,
                    cudaStream_t st) {
    if (count <= 0 || maxR <= 0) return;
    dim3 g((maxR + PULL_ROWS - 1) / PULL_ROWS, (maxR + PULL_COLS - 1) / PULL_COLS, count);
    pull_cb_kernel<<<g, PULL_ROWS, 0, st>>>(S, list, CB, st_d);
    count_launch();
}

void launch_pull_panels(const DevSym& S, const int* list, int count, int maxN, double* Lval, const DeltaState* st_d,
                        cudaStream_t st) {
    if (count <= 0 || maxN <= 0) return;
    // columns: every front has at most maxN of them
    dim3 g((maxN + PULL_ROWS - 1) / PULL_ROWS, (maxN + PULL_COLS - 1) / PULL_COLS, count);
    pull_panels_kernel<<<g, PULL_ROWS, 0, st>>>(S, list, Lval, st_d);
    count_launch();
}

void launch_shard_barrier(const ShardCtx& C, unsigned mask, cudaStream_t st) {
    mask &= (1u << C.world) - 1u;
    if (!((mask >> C.rank) & 1u) || !(mask & ~(1u << C.rank))) return;      // not taking part / nobody else
    shard_barrier_kernel<<<1, 32, 0, st>>>(C, mask);
    count_launch();
}

void launch_push_supernodes(const DevSym& S, const int* list, int count, int maxc, const double* x, cudaStream_t st) {
    if (count <= 0 || maxc <= 0) return;
    dim3 g((maxc + 255) / 256, count);
    push_supernodes_kernel<<<g, 256, 0, st>>>(S, list, x);
    count_launch();
}

void launch_push_owned(const DevSym& S, const int* colowner, const double* x, cudaStream_t st) {
    push_owned_kernel<<<(S.n + 255) / 256, 256, 0, st>>>(S, colowner, x);
    count_launch();
}

// Force-load every kernel of this translation unit (CUDA loads kernels lazily, and a load may
// synchronise the context: that must not happen while another stream waits in a cross-rank barrier).
cudaError_t preload_shard() {
    cudaFuncAttributes a;
    cudaError_t e;
    e = cudaFuncGetAttributes(&a, shard_barrier_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, push_supernodes_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, push_owned_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, pull_panels_kernel); if (e != cudaSuccess) return e;
    e = cudaFuncGetAttributes(&a, pull_cb_kernel); if (e != cudaSuccess) return e;
    return cudaSuccess;
}

}  // namespace opb
