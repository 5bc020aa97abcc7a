// Internal declarations shared by the CUDA translation units of libonephase_b200.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <string>
#include <vector>

#include "symbolic.h"

namespace opb {

// number of kernels launched by this library (bench.py reports it as gpu_launches)
extern std::atomic<long long> g_launches;
inline void count_launch(int k = 1) { g_launches.fetch_add(k, std::memory_order_relaxed); }

// Device-resident controller state of the delta loop (delta_strategy.jl:37-114).
// Every factorisation kernel reads `done`/`fail` first and exits when set, so a
// whole attempt that is no longer needed costs only empty launches.
struct DeltaState {
    double delta;        // shift of the attempt in flight / accepted shift
    double tau;
    double delta_prev, delta_zero, delta_min, delta_max, delta_start, inc, dec;
    double diag_min;
    int done;            // loop finished (success, failure or max_it)
    int fail;            // current attempt hit a non-positive pivot
    int num_fac;
    int status;          // 1 success, 0 failure (delta > delta_max), -1 max it, 2 running
    int it;              // index i of the reference's for-loop (0 = the delta_zero probe)
    int max_it;
    int mode;            // 0 Cholesky, 1 LDL'
    int pad;
    // LDL' inertia (julia.jl:72-80)
    int n_pos, n_neg, n_zero, n_bad;
    double kkt_err[6];
};

constexpr int MAX_SHARD = 8;          // GPUs one instance can be sharded over

// leading dimension of a panel / inverse block with N rows (even: 16-byte aligned columns)
__host__ __device__ inline int ld_of(int N) { return (N + 1) & ~1; }

// Flat view of the symbolic structures in device memory.
struct DevSym {
    int n, nsuper;
    const int* sfirst;
    const int64_t* rowptr;
    const int* rowidx;
    const int* rel;
    const int64_t* Loff;
    const int64_t* CBoff;
    const int* sparent;
    const int* child_ptr;
    const int* child_list;
    const int* perm;   // perm[new] = old
    const int64_t* Xoff;  // per supernode: offset of inv(L11) (c x c, ld = ld_of(c)) in Xinv, or -1
    double* dvec;         // LDL' mode: the pivots D as a vector over the permuted columns (per handle)
    const int64_t* gptr;  // forward-solve gather lists (symbolic.h)
    const int64_t* gsrc;
    const int* gch;
    const int* tcut_ptr;  // tile cuts of every update block inside its parent's (symbolic.h)
    const int* tcut;
    // ---- one instance sharded over several GPUs; owner == nullptr on a single GPU
    const int* owner;     // per supernode: rank that owns it
    const unsigned char* cb_mirror;   // per supernode: its update block is mirrored into this rank's arena before its parent's level
    int rank, world;
    double* cb_peer[MAX_SHARD];   // update-block buffer of every rank (peer-mapped; [rank] = own)
    double* l_peer[MAX_SHARD];    // factor storage of every rank (helpers pull the panels of split fronts)
    double* u_peer[MAX_SHARD];    // forward-solve update vectors of every rank
    double* x_peer[MAX_SHARD];    // solution vector of every rank (top supernodes are pushed to all)
};

// update block / forward update vector of child `ch`: in the owner's HBM, read over NVLink when
// the child belongs to another rank (the reduction of the subtree roots' update blocks onto the
// separator front happens inside the consuming extend-add, not as a separate collective)
// (cb_mirror[ch] != 0: this rank has pulled a copy of the block into its own arena -- same offset, the slot
// is free on every rank that does not own the child -- before the level started: pull_cb_kernel)
__device__ __forceinline__ const double* child_cb(const DevSym& S, const double* CB, int ch) {
    return ((S.owner && !S.cb_mirror[ch]) ? S.cb_peer[S.owner[ch]] : CB) + S.CBoff[ch];
}
__device__ __forceinline__ const double* child_u(const DevSym& S, const double* u, int ch) {
    return (S.owner ? S.u_peer[S.owner[ch]] : u) + S.rowptr[ch];
}
// sum of the children's update-vector entries that land on destination g (ascending child order)
__device__ __forceinline__ double gather_dest(const DevSym& S, const double* u, int64_t g) {
    double acc = 0.0;
    const int64_t e1 = S.gptr[g + 1];
    for (int64_t e = S.gptr[g]; e < e1; e++)
        acc += (S.owner ? S.u_peer[S.owner[S.gch[e]]] : u)[S.gsrc[e]];
    return acc;
}

// Cross-GPU barrier state of a sharded handle: every rank owns 2 x MAX_SHARD flag words that
// its peers write over NVLink ((epoch << 1) | fail, slot = epoch & 1) and MAX_SHARD local epoch counters
// (one per peer: a barrier may involve a subset of the ranks, so the counts are kept pair by pair).
struct ShardCtx {
    int rank, world;
    unsigned long long* flags_local;            // [2][MAX_SHARD]
    unsigned long long* flags_peer[MAX_SHARD];  // the same array on every rank
    unsigned long long* epoch;                  // local counters, one per peer: barriers done together with that peer
    int* error;                                 // local: set to 1 when a wait timed out
    long long timeout_clocks;                   // a wait gives up after this many SM clocks
    DeltaState* state;                          // local controller state (fail bit exchanged at every barrier)
};

// bytes of device memory held by the library in this process (info key device_bytes)
inline std::atomic<long long> g_device_bytes{0};

template <class T>
struct DBuf {
    T* p = nullptr;
    size_t n = 0;
    cudaError_t alloc(size_t count) {
        if (count <= n && p) return cudaSuccess;
        release();
        cudaError_t e = cudaMalloc((void**)&p, (count ? count : 1) * sizeof(T));
        if (e == cudaSuccess) { n = count; g_device_bytes += (long long)(n * sizeof(T)); } else p = nullptr;
        return e;
    }
    cudaError_t upload(const std::vector<T>& h, cudaStream_t st) {
        cudaError_t e = alloc(h.size());
        if (e != cudaSuccess) return e;
        if (h.empty()) return cudaSuccess;
        return cudaMemcpyAsync(p, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, st);
    }
    void release() { if (p) { cudaFree(p); g_device_bytes -= (long long)(n * sizeof(T)); } p = nullptr; n = 0; }
};

// Front classes of a level; the supernodes of a level are stored in d_sched in this order.
//   T32 .. S152 : whole front (N x N) in shared memory, one CTA per front
//   MID, MIDL   : N x c panel in shared memory (<= 96 KB / <= 192 KB), update block by tiles
//   BIG         : blocked right-looking in HBM, outer block WB, FP64 tensor-core tiles
enum FrontClass { FC_T32 = 0, FC_S64, FC_S104, FC_S152, FC_MID, FC_MIDL, FC_BIG, NFC };
constexpr int FC_MAXN[4] = {32, 64, 104, 152};
constexpr int SMALL_N = 152;          // 152*152*8 = 184,832 B of shared memory
constexpr int MID_PANEL = 12000;      // doubles
constexpr int MIDL_PANEL = 26000;     // doubles
constexpr int NB = 32;                // block-column width of the LDL' big-front path
constexpr int WB = 128;               // block width of the Cholesky big-front path (DMMA)
constexpr int OUTER_BLOCK = 4096;     // default outer block of the panel update (WB times a power of two)
constexpr int CB_SMALL_K = 1 << 30;      // levels whose fronts have at most this many pivot columns form their update blocks in 64-row tiles
constexpr int XB = 2048;              // pivot blocks are inverted in diagonal blocks of this many columns

// Tiles of the update-block kernel (BM rows x 128 columns over the lower triangle of the r x r block,
// origin at front position ce = c & ~1).  WM = 2 (BM = 128): tile (I, J), I >= J, linear index
// I(I+1)/2 + J.  WM = 1 (BM = 64): row tile I meets column tile J when I >= 2J; the row tiles 2a and
// 2a+1 hold a+1 tiles each, a(a+1) tiles precede them.
__host__ __device__ inline long long cb_tiles(int WM, int N, int ce) {
    if (WM == 2) { const long long nt = (N - ce + 127) / 128; return nt * (nt + 1) / 2; }
    const long long n64 = (N - ce + 63) / 64, a = n64 >> 1;
    return a * (a + 1) + ((n64 & 1) ? a + 1 : 0);
}

// Per-level schedule built on the host from Symbolic.
struct LevelPlan {
    int begin[NFC] = {0}, count[NFC] = {0}, maxN[NFC] = {0}, maxC[NFC] = {0}, maxPanel[NFC] = {0};
    int minN[NFC] = {1 << 30, 1 << 30, 1 << 30, 1 << 30, 1 << 30, 1 << 30, 1 << 30};
    int all_begin = 0, all_count = 0;          // every supernode of the level
    int solo_count = 0;                        // classes T32 .. S152 (a prefix of the level)
    int wide_begin = 0, wide_count = 0;        // classes MID .. BIG (the rest)
    int wide_maxN = 0, wide_maxC = 0, wide_maxR = 0;
    // exact tile lists of the update-block kernel, as (position in the wide list, tile) pairs inside
    // d_sched: [0] 64-row tiles, [1] 128-row tiles; biggest pivot blocks first
    int cbt_begin[2] = {0, 0}, cbt_count[2] = {0, 0};
    // big fronts are sorted by pivot-column count (descending); outer step t of the blocked
    // factorisation touches the first step_count[t] of them
    std::vector<int> step_count, step_maxN;
    // sharded instance: a cross-GPU barrier precedes this level (a supernode of the level has a
    // child on another rank); push_count = top supernodes of this rank on the level, whose solution
    // is pushed to every peer in the backward sweep
    int barrier_before = 0, push_count = 0, push_begin = 0, push_maxc = 0;
    unsigned barrier_mask = 0;                 // ... with these ranks (bit mask incl. this rank; 0: nobody to wait for)
    // sharded instance: the level has split fronts (ShardMap::split): two more barriers, and the split
    // fronts of other ranks whose update-block tiles this rank forms (positions right behind the wide list)
    int split = 0, help_begin = 0, help_count = 0, help_maxN = 0;
    // children on other ranks whose update blocks this rank copies into its own arena before the level
    // (fine-grained peer loads inside the consuming kernels run at a few percent of the link rate)
    int pullcb_begin = 0, pullcb_count = 0, pullcb_maxR = 0;
};

// ---- kernels_assembly.cu
void launch_prep(const double* Jv, const int* Jrow, const int* Rpos, const double* y, const double* s,
                 double* sigma, double* T, double* Rval, int64_t nnzJ, int m, DeltaState* st_d, cudaStream_t st);
void launch_assemble_M(const int64_t* pair_ptr, const int* pairA, const int* pairB, const int* hmap,
                       const double* T, const double* Jv, const double* Hv, double* Mval,
                       int64_t nnzM, cudaStream_t st);
void launch_diag_extract(const int64_t* Mp, const double* Mval, double* sdiag, DeltaState* st_d,
                         int n, cudaStream_t st);
void launch_scatter_fronts(const double* Mval, const int64_t* amap, const int64_t* dpos,
                           const double* sdiag, double* Lval, int64_t nnzL, int64_t nnzM, int n,
                           const DeltaState* st_d, int use_sdiag, cudaStream_t st);
void launch_ctl_begin(DeltaState* st_d, cudaStream_t st);
void launch_ctl_end(DeltaState* st_d, cudaStream_t st);
void launch_ctl_end_loop(DeltaState* st_d, unsigned long long loop_handle, cudaStream_t st);
void launch_ctl_init(DeltaState* st_d, double delta_prev, double delta_zero, double delta_min,
                     double delta_max, double delta_start, double inc, double dec, int max_it,
                     int mode, cudaStream_t st);
void launch_ctl_single(DeltaState* st_d, double delta, int mode, cudaStream_t st);
void launch_diag_JtDJ(const int64_t* Jp, const int64_t* Ji, const double* Jx, const double* dv, double* out,
                      int n, int base, cudaStream_t st);

// Extra streams of a handle for the blocked panel factorisation of the big fronts.
// Legacy look-ahead (deep == false): the part of a panel update that does not touch the next block
// column runs on `stream`, concurrently with the (latency-bound) diagonal-block and TRSM kernels of
// the next step.  Deep look-ahead (deep == true): `stream` (highest priority) carries the latency
// chain; every panel update is cut into pieces by the step at which its columns are next touched --
// piece i = block columns [tb + 2^i, tb + 2^(i+1)) is needed 2^i steps later and runs on cls[i]; the
// part of an outer update beyond the next outer block runs on `rest` (lowest priority) -- and the
// chain waits only for the one piece that last wrote the columns it is about to update.  Fork / join
// through events (also while the main stream is captured into a graph).
constexpr int LA_CLASSES = 6;      // outer blocks of up to 2^6 block columns
struct SideStream {
    cudaStream_t stream = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr, start = nullptr;
    bool chain_on_side = false;   // the side stream has the highest priority and runs the latency chain
    bool deep = false;
    cudaStream_t cls[LA_CLASSES] = {nullptr}, rest = nullptr;
    cudaEvent_t cls_done[LA_CLASSES] = {nullptr}, rest_done = nullptr;
    // pivot-block inverses of finished levels run here (lowest priority) while the levels above are factorised
    cudaStream_t aux = nullptr;
    cudaEvent_t aux_fork = nullptr, aux_done = nullptr;
};

// Optional per-kernel timing of one factorisation attempt (opb_profile_factor): CUDA events
// around every launch of the two tensor-pipe kernels, on the launching stream.
struct KernelTimer {
    std::vector<cudaEvent_t> ev;       // pairs (begin, end)
    std::vector<int> kind;             // per pair: 0 = front_cb_kernel, 1 = chol_panel_update_kernel
    size_t used = 0;
    // phases == true (opb_profile_levels): no per-kernel events (the attempt runs with the look-ahead
    // streams as configured); instead single marks on the main stream at the phase boundaries of
    // every level: level start, panels of the big fronts start, update blocks start
    bool phases = false;
    std::vector<cudaEvent_t> mark;
    std::vector<int> mark_kind;        // 0 level start, 1 big panels start, 2 update blocks start, 3 end of the levels
    size_t marks_used = 0;
    void put_mark(int k, cudaStream_t st) {
        if (marks_used == mark.size()) { cudaEvent_t e; cudaEventCreate(&e); mark.push_back(e); mark_kind.push_back(k); }
        mark_kind[marks_used] = k;
        cudaEventRecord(mark[marks_used++], st);
    }
    cudaEvent_t next(int k) {
        if (used == ev.size()) { cudaEvent_t e; cudaEventCreate(&e); ev.push_back(e); }
        if ((used & 1) == 0) { if (kind.size() <= used / 2) kind.push_back(k); else kind[used / 2] = k; }
        return ev[used++];
    }
    void release() {
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
        for (cudaEvent_t e : mark) cudaEventDestroy(e);
        ev.clear(); kind.clear(); used = 0; mark.clear(); mark_kind.clear(); marks_used = 0;
    }
};

// host-side factorisation mode of the launchers: the kernels (DeltaState::mode) only distinguish 0 / 1
enum { FMODE_CHOLESKY = 0, FMODE_LDLT = 1, FMODE_LDLT_SCALAR = 2 };   // 2: option "ldlt_scalar", the first (scalar, K = 32) LDL' path

// ---- kernels_factor.cu
struct TrtriPlan;
cudaError_t factor_configure();
void launch_factor_levels(const DevSym& S, const std::vector<LevelPlan>& plan, const int* d_sched,
                          double* Lval, double* CB, double* Xinv, DeltaState* st_d, int mode,
                          int outer_block, int cb_small_k, const ShardCtx* shard, const SideStream* side,
                          KernelTimer* timer, const std::vector<TrtriPlan>* trtri, double* Twork, cudaStream_t st);

// ---- kernels_dense.cu  (Cholesky of big fronts on the FP64 tensor pipe, pivot-block inverses,
//                         multi-CTA triangular solves for big supernodes)
struct TrtriPlan {
    int after_level = 0;              // the batch may start once this elimination-tree level is factorised
    int count = 0;                    // big supernodes with more than one WB block (sorted by c descending)
    int list_begin = 0;               // position in d_sched
    // merge level l: exact work list inside d_sched, (position in the list, pair * nsub^2 + I * nsub + J) pairs
    std::vector<int> items_begin, items_count;
};
cudaError_t dense_configure();
extern int g_occ_small_tiles;
// medium + big fronts of one level (Cholesky)
void launch_wide_chol_level(const DevSym& S, const LevelPlan& L, const int* d_sched, double* Lval,
                            double* CB, double* Xinv, DeltaState* st_d, int ldlt, int outer_block, int cb_small_k,
                            const SideStream* side, KernelTimer* timer, int phase, cudaStream_t st);
void launch_trtri(const DevSym& S, const TrtriPlan& T, const int* d_sched, const double* Lval,
                  double* Xinv, double* Twork, const DeltaState* st_d, cudaStream_t st);
void launch_solve_wide_fwd(const DevSym& S, const LevelPlan& L, const int* d_sched, const double* Lval,
                           const double* Xinv, double* x, double* xnew, double* u, cudaStream_t st);
void launch_solve_wide_bwd(const DevSym& S, const LevelPlan& L, const int* d_sched, const double* Lval,
                           const double* Xinv, double* x, double* xnew, double* u, int ldlt, cudaStream_t st);
void launch_ldlt_inertia(const DevSym& S, const double* Lval, const int64_t* dpos, int n,
                         DeltaState* st_d, cudaStream_t st);

// every translation unit: load its kernels now instead of at first launch
cudaError_t preload_assembly();
cudaError_t preload_vec();
cudaError_t preload_solve();
cudaError_t preload_factor();
cudaError_t preload_dense();
cudaError_t preload_shard();

// ---- kernels_shard.cu  (cross-GPU barrier over peer-mapped flags, solution pushes)
void launch_shard_barrier(const ShardCtx& C, unsigned mask, cudaStream_t st);      // mask: participating ranks
inline unsigned shard_all(const ShardCtx& C) { return (1u << C.world) - 1u; }
void launch_push_supernodes(const DevSym& S, const int* list, int count, int maxc, const double* x, cudaStream_t st);
void launch_push_owned(const DevSym& S, const int* colowner, const double* x, cudaStream_t st);
void launch_pull_cb(const DevSym& S, const int* list, int count, int maxR, double* CB, const DeltaState* st_d, cudaStream_t st);
void launch_pull_panels(const DevSym& S, const int* list, int count, int maxN, double* Lval, const DeltaState* st_d, cudaStream_t st);

// ---- kernels_solve.cu
cudaError_t solve_configure();
// x (permuted order, length n) is overwritten with the solution; u = workspace (len rowidx)
void launch_solve(const DevSym& S, const std::vector<LevelPlan>& plan, const int* d_sched,
                  const double* Lval, const double* Xinv, double* x, double* xnew, double* u, int mode,
                  const ShardCtx* shard, const int* colowner, const SideStream* side, cudaStream_t st);
void launch_permute_in(const double* b, const int* perm, double* x, int n, cudaStream_t st);
void launch_permute_out_add(const double* x, const int* perm, double* dst, int n, int accumulate, cudaStream_t st);

// ---- kernels_vec.cu  (direction, refinement residual, KKT error)
struct DirBuffers {
    int n, m;
    // matrices
    const int64_t* Rp; const int* Rcol; const double* Rval;      // J by rows
    const int64_t* Jp; const int* Jrow; const double* Jv;         // J by columns
    const int64_t* Sp; const int* Scol; const int* Spos; const double* Hv;  // H symmetric view
    const double *y, *s, *sigma;
    const double *dual_r, *primal_r, *comp_r;
    double *b, *res, *dx, *dy, *ds, *tm, *tm2;
    unsigned long long* red;   // 8 slots of max-reduction scratch
    DeltaState* st_d;
    int wide_n, wide_m;        // long rows: one warp per row in the n- / m-sized product kernels
};
void launch_schur_rhs(const DirBuffers& B, cudaStream_t st);
void launch_residual(const DirBuffers& B, cudaStream_t st);          // res = b - (J'(S.(J dx)) + Hsym dx + delta dx)
void launch_recover_and_error(const DirBuffers& B, cudaStream_t st);  // dy, ds, kkt_err[6]
// System_rhs(iter, reduct_factors) into the resident rhs buffers (system_rhs.jl:57-73)
void launch_system_rhs(const DirBuffers& B, const double* grad, const double* cons, double mu_t, double a_pen,
                       double eta_P, double eta_D, cudaStream_t st);
// |dx|, |dy|, |ds| (inf norms) and the fraction-to-the-boundary ratio into B.red[0..3] (frac_boundary.jl:3-35)
void launch_step_bounds(const DirBuffers& B, double frac_bd, double ex, cudaStream_t st);

}  // namespace opb
