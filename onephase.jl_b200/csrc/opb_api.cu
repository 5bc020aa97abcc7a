// C ABI of libonephase_b200.so (include/onephase_b200.h).  Host orchestration
// only: symbolic cache, device buffers, kernel sequencing.  No numeric work is
// done on the host; without a CUDA device the numeric entry points fail.
#include <cuda_runtime.h>

#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <chrono>

#include "../../include/onephase_b200.h"
#include "opb_internal.h"

namespace opb {
std::atomic<long long> g_launches{0};

// Immutable per-pattern data (host + device), shared between handles.
struct Bundle {
    bool schur = false;          // built from (J,H) patterns; else from a user CSC matrix
    SchurPattern P;              // schur only
    std::vector<int64_t> Mp;     // pattern of the lower triangle that is factorised
    std::vector<int> Mi;
    std::vector<int64_t> src;    // csc path: position in the caller's nzval (or -1)
    std::vector<std::pair<std::string, double>> timing;   // host seconds: pattern, analyze, shard_map, plan, upload
    // the caller's index arrays (0-based copies): a cache hit is accepted only when they are equal,
    // so a 64-bit hash collision cannot hand out a wrong analysis
    std::vector<int64_t> in_p[2];
    std::vector<int> in_i[2];
    void keep_pattern(int k, int64_t n, const int64_t* p, const int64_t* i, int base) {
        in_p[k].resize(n + 1);
        for (int64_t c = 0; c <= n; c++) in_p[k][c] = p[c] - base;
        const int64_t nnz = in_p[k][n];
        in_i[k].resize(nnz);
        for (int64_t e = 0; e < nnz; e++) in_i[k][e] = (int)(i[e] - base);
    }
    bool same_pattern(int k, int64_t n, const int64_t* p, const int64_t* i, int base) const {
        if ((int64_t)in_p[k].size() != n + 1) return false;
        for (int64_t c = 0; c <= n; c++) if (in_p[k][c] != p[c] - base) return false;
        const int64_t nnz = in_p[k][n];
        for (int64_t e = 0; e < nnz; e++) if (in_i[k][e] != (int)(i[e] - base)) return false;
        return true;
    }
    Symbolic S;
    std::vector<LevelPlan> plan;
    std::vector<int> sched;
    std::vector<int64_t> Xoff;   // per supernode: offset of inv(L11) in Xinv (big supernodes) or -1
    int64_t x_total = 0;
    std::vector<TrtriPlan> trtri;   // batches of pivot-block inverses, ascending after_level
    int n_tiny = 0, n_small = 0, n_big = 0;
    // one instance sharded over `world` GPUs: this bundle holds the schedule of `rank`
    int rank = 0, world = 1;
    ShardMap shard;
    std::vector<int> colowner;   // per permuted column: owning rank
    std::vector<unsigned char> cb_mirror;   // per supernode: this rank mirrors its update block before the parent's level
    DBuf<unsigned char> d_cb_mirror;
    DBuf<int> d_owner, d_colowner;
    int device = -1;
    size_t device_bytes = 0;
    // device copies
    DBuf<int> d_sfirst, d_rowidx, d_rel, d_sparent, d_child_ptr, d_child_list, d_perm, d_sched;
    DBuf<int64_t> d_rowptr, d_Loff, d_CBoff, d_amap, d_dpos, d_Mp, d_src, d_Xoff, d_gptr, d_gsrc;
    DBuf<int> d_gch, d_tcut_ptr, d_tcut;
    DBuf<int64_t> d_pair_ptr, d_Jp, d_Rp, d_Sp;
    DBuf<int> d_pairA, d_pairB, d_hmap, d_Jrow, d_Rcol, d_Rpos, d_Scol, d_Spos;
    DevSym dev{};
    ~Bundle() {
        if (device >= 0) cudaSetDevice(device);
        d_sfirst.release(); d_rowidx.release(); d_rel.release(); d_sparent.release();
        d_child_ptr.release(); d_child_list.release(); d_perm.release(); d_sched.release();
        d_rowptr.release(); d_Loff.release(); d_CBoff.release(); d_amap.release(); d_dpos.release();
        d_Mp.release(); d_src.release(); d_Xoff.release(); d_pair_ptr.release(); d_Jp.release(); d_Rp.release(); d_Sp.release();
        d_pairA.release(); d_pairB.release(); d_hmap.release(); d_Jrow.release(); d_Rcol.release();
        d_Rpos.release(); d_Scol.release(); d_Spos.release();
        d_owner.release(); d_colowner.release(); d_cb_mirror.release(); d_gptr.release(); d_gsrc.release(); d_gch.release();
        d_tcut_ptr.release(); d_tcut.release();
    }
};

static std::mutex g_cache_mu;
static std::map<std::string, std::shared_ptr<Bundle>> g_cache;
static std::vector<std::string> g_cache_order;
constexpr size_t CACHE_MAX = 8;

static void build_plan(Bundle& B) {
    const Symbolic& S = B.S;
    B.plan.assign(S.nlevels, LevelPlan());
    B.sched.clear();
    B.n_tiny = B.n_small = B.n_big = 0;
    B.Xoff.assign(S.nsuper, -1);
    B.x_total = 0;
    B.cb_mirror.assign(S.nsuper, 0);
    auto cols = [&](int s) { return S.sfirst[s + 1] - S.sfirst[s]; };
    auto rows = [&](int s) { return cols(s) + (int)(S.rowptr[s + 1] - S.rowptr[s]); };
    std::vector<int> all_big;
    const bool sharded = B.world > 1;
    for (int l = 0; l < S.nlevels; l++) {
        LevelPlan& L = B.plan[l];
        std::vector<int> cls[NFC];
        std::vector<int> push;
        if (sharded) {
            L.barrier_before = B.shard.level_barrier[l]; L.split = B.shard.level_split[l];
            L.barrier_mask = B.shard.level_mask[(size_t)l * B.world + B.rank];
        }
        std::vector<int> helped;          // split fronts of other ranks this rank forms update-block tiles of
        for (int t = S.level_ptr[l]; t < S.level_ptr[l + 1]; t++) {
            const int s = S.level_list[t];
            if (sharded && B.shard.owner[s] != B.rank) {              // another rank's supernode
                if (B.shard.split[s] && B.shard.ra[s] <= B.rank && B.rank < B.shard.rb[s]) helped.push_back(s);
                continue;
            }
            const int c = cols(s), N = rows(s);
            if (sharded && B.shard.top[s]) { push.push_back(s); L.push_maxc = std::max(L.push_maxc, c); }
            int fc;
            if (N <= FC_MAXN[FC_T32]) fc = FC_T32;
            else if (N <= FC_MAXN[FC_S64]) fc = FC_S64;
            else if (N <= FC_MAXN[FC_S104]) fc = FC_S104;
            else if (N <= FC_MAXN[FC_S152]) fc = FC_S152;
            else if (c <= WB && (int64_t)N * c <= MID_PANEL) fc = FC_MID;
            else if (c <= WB && (int64_t)N * c <= MIDL_PANEL) fc = FC_MIDL;
            else fc = FC_BIG;
            cls[fc].push_back(s);
            L.maxN[fc] = std::max(L.maxN[fc], N);
            L.minN[fc] = std::min(L.minN[fc], N);
            L.maxC[fc] = std::max(L.maxC[fc], c);
            L.maxPanel[fc] = std::max(L.maxPanel[fc], N * std::min(c, WB));
            if (fc >= FC_MID) {
                L.wide_maxN = std::max(L.wide_maxN, N);
                L.wide_maxC = std::max(L.wide_maxC, c);
                L.wide_maxR = std::max(L.wide_maxR, N - c);
            }
        }
        // big fronts by pivot-column count, descending: outer step t touches a prefix
        std::vector<int>& big = cls[FC_BIG];
        std::stable_sort(big.begin(), big.end(), [&](int a, int b) { return cols(a) > cols(b); });
        const int nsteps = (L.maxC[FC_BIG] + WB - 1) / WB;
        L.step_count.assign(nsteps, 0); L.step_maxN.assign(nsteps, 0);
        for (int s : big) {
            const int c = cols(s), N = rows(s);
            for (int t = 0; t * WB < c; t++) { L.step_count[t]++; L.step_maxN[t] = std::max(L.step_maxN[t], N); }
            all_big.push_back(s);
        }
        L.all_begin = (int)B.sched.size();
        for (int fc = 0; fc < NFC; fc++) {
            L.begin[fc] = (int)B.sched.size();
            L.count[fc] = (int)cls[fc].size();
            B.sched.insert(B.sched.end(), cls[fc].begin(), cls[fc].end());
            if (fc == FC_BIG)
                for (int s : cls[fc]) {           // pivot-block inverse for the multi-CTA solves
                    B.Xoff[s] = B.x_total;
                    B.x_total += (int64_t)ld_of(cols(s)) * cols(s);
                }
        }
        L.all_count = (int)B.sched.size() - L.all_begin;
        // helped fronts sit right behind the level's wide list, so the update-block kernel addresses
        // them through the same list pointer (positions wide_count .. wide_count + help_count - 1)
        L.help_begin = (int)B.sched.size();
        L.help_count = (int)helped.size();
        B.sched.insert(B.sched.end(), helped.begin(), helped.end());
        for (int s : helped) L.help_maxN = std::max(L.help_maxN, rows(s));
        if (sharded) {
            // remote children of the fronts this rank factorises or helps with at this level: mirrored into the
            // local arena, except under long-K split fronts (c > 6000 with an update block), where the fine-grained
            // peer loads hide behind the K loop and the copies would only add link traffic
            L.pullcb_begin = (int)B.sched.size();
            auto consider = [&](int p) {
                if (cols(p) > 6000 && rows(p) > cols(p)) return;
                for (int k = S.child_ptr[p]; k < S.child_ptr[p + 1]; k++) {
                    const int c = S.child_list[k];
                    const int rc = rows(c) - cols(c);
                    if (B.shard.owner[c] == B.rank || rc <= 0) continue;
                    B.sched.push_back(c);
                    B.cb_mirror[c] = 1;
                    L.pullcb_maxR = std::max(L.pullcb_maxR, rc);
                }
            };
            for (int q = 0; q < L.all_count; q++) consider(B.sched[L.all_begin + q]);
            for (int s : helped) consider(s);
            L.pullcb_count = (int)B.sched.size() - L.pullcb_begin;
        }
        L.push_begin = (int)B.sched.size();
        L.push_count = (int)push.size();
        B.sched.insert(B.sched.end(), push.begin(), push.end());
        L.solo_count = L.count[FC_T32] + L.count[FC_S64] + L.count[FC_S104] + L.count[FC_S152];
        L.wide_begin = L.begin[FC_MID];
        L.wide_count = L.all_count - L.solo_count;
        // tile lists of the update-block kernel: big fronts first (descending pivot-column count:
        // the long tiles start first), then the medium ones
        for (int v = 0; v < 2; v++) {
            if (B.sched.size() & 1) B.sched.push_back(0);          // int2 alignment
            L.cbt_begin[v] = (int)B.sched.size();
            auto add_tiles = [&](int pos, int s) {
                const int c = cols(s), N = rows(s);
                if (N - c <= 0) return;
                const long long nt = cb_tiles(v + 1, N, c & ~1);
                // split front (sharded instance): tile t belongs to rank ra + t mod (rb - ra)
                const bool sp = sharded && B.shard.split[s];
                const int G = sp ? B.shard.rb[s] - B.shard.ra[s] : 1, me = sp ? B.rank - B.shard.ra[s] : 0;
                for (long long t = me; t < nt; t += G) { B.sched.push_back(pos); B.sched.push_back((int)t); }
            };
            for (int fc = FC_BIG; fc >= FC_MID; fc--)
                for (int q = 0; q < L.count[fc]; q++)
                    add_tiles(L.begin[fc] + q - L.wide_begin, B.sched[L.begin[fc] + q]);
            for (int q = 0; q < L.help_count; q++) add_tiles(L.help_begin + q - L.wide_begin, B.sched[L.help_begin + q]);
            L.cbt_count[v] = ((int)B.sched.size() - L.cbt_begin[v]) / 2;
        }
        B.n_tiny += L.count[FC_T32];
        B.n_small += L.solo_count - L.count[FC_T32];
        B.n_big += L.wide_count;
    }
    // pivot-block inverses: every big supernode with more than one WB block takes part in the
    // recursive merge.  Four batches -- everything below the top three levels, then each of the top
    // three levels -- so that a batch can run on a low-priority stream while the levels above it are
    // being factorised (only the root's batch has nothing left to hide behind).
    B.trtri.clear();
    const int nl = S.nlevels;
    for (int g = 0; g < 4; g++) {
        const int lo = g == 0 ? 0 : nl - 4 + g, hi = nl - 3 + g;      // levels [lo, hi)
        if (hi <= 0 || lo < 0) continue;
        TrtriPlan T;
        T.after_level = hi - 1;
        std::vector<int> multi;
        for (int s : all_big) if (cols(s) > WB && S.level[s] >= lo && S.level[s] < hi) multi.push_back(s);
        if (multi.empty()) continue;
        std::stable_sort(multi.begin(), multi.end(), [&](int a, int b) { return cols(a) > cols(b); });
        T.count = (int)multi.size();
        T.list_begin = (int)B.sched.size();
        B.sched.insert(B.sched.end(), multi.begin(), multi.end());
        const int maxc = cols(multi[0]);
        for (int l = 0; (WB << l) < std::min(maxc, XB); l++) {
            const int Sz = WB << l, nsub = 1 << l;
            if (B.sched.size() & 1) B.sched.push_back(0);          // int2 alignment
            T.items_begin.push_back((int)B.sched.size());
            for (int q = 0; q < (int)multi.size(); q++) {
                const int c = cols(multi[q]);
                for (int pair = 0; (2 * pair + 1) * Sz < c; pair++) {
                    const int a1 = (2 * pair + 1) * Sz, a2 = std::min(a1 + Sz, c);
                    const int ni = (a2 - a1 + WB - 1) / WB;
                    for (int I = 0; I < ni; I++)
                        for (int J = 0; J < nsub; J++) {
                            B.sched.push_back(q);
                            B.sched.push_back(pair * nsub * nsub + I * nsub + J);
                        }
                }
            }
            T.items_count.push_back(((int)B.sched.size() - T.items_begin.back()) / 2);
        }
        B.trtri.push_back(T);
    }
}

}  // namespace opb

using namespace opb;

struct opb_handle {
    int device = -1;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    SymOptions opt;
    std::vector<int64_t> user_perm;
    int attempts_per_sync = 2;
    int outer_block = OUTER_BLOCK;
    int cb_small_k = CB_SMALL_K;
    bool ldlt_scalar = false;            // option "ldlt_scalar"
    int fmode() const { return mode == OPB_MODE_CHOLESKY ? FMODE_CHOLESKY : (ldlt_scalar ? FMODE_LDLT_SCALAR : FMODE_LDLT); }
    double barrier_timeout_s = 20.0;     // sharded instance: a rank that waits longer reports an error
    std::shared_ptr<Bundle> B;
    bool cached_hit = false;
    // numeric state
    enum Ready { NOT_READY, SYSTEM_FORMED, FACTORED } ready = NOT_READY;
    int mode = OPB_MODE_CHOLESKY;
    DBuf<double> Jv, Hv, y, s, sigma, T, Rval, Mval, sdiag, Lval, CB;
    DBuf<double> dual_r, primal_r, comp_r, b, res, dx, dy, ds, tm, xw, xw2, uw, userval, Xinv, Twork, Dvec;
    DBuf<unsigned long long> red;
    DBuf<int64_t> scr_i0, scr_i1;        // scratch of opb_eval_diag_JtDJ
    DBuf<double> scr_d0, scr_d1, scr_d2;
    char* d_state_raw = nullptr;
    DeltaState* d_state = nullptr;
    DeltaState h_state{};
    size_t numeric_bytes = 0;
    uint64_t csc_hash = 0;
    // CUDA graphs of the launch-bound sequences (one factorisation attempt, one direction, one
    // solve): captured once per structure, replayed afterwards
    struct GraphSlot { cudaGraphExec_t exec = nullptr; long long launches = 0; int key = -1; };
    GraphSlot g_attempt, g_direction, g_solve;
    // the whole delta loop as ONE graph: a WHILE conditional node whose body is an attempt and
    // whose condition is set on the device by the last kernel of the body (ctl_end_loop_kernel)
    GraphSlot g_loop;
    bool loop_graph = true;            // option "loop_graph"
    std::string loop_diag = "not built";
    long long loop_pending = 0;        // launches per attempt of a loop whose attempt count is not known yet
    bool use_graphs = true;
    void drop_graphs() {
        for (GraphSlot* g : {&g_attempt, &g_direction, &g_solve, &g_loop}) {
            if (g->exec) cudaGraphExecDestroy(g->exec);
            *g = GraphSlot();
        }
    }

    // ---- one instance sharded over several GPUs (opb_shard_*)
    int shard_rank = 0, shard_world = 1;
    DevSym dev{};                        // B->dev plus this handle's peer pointers
    ShardCtx sctx{};
    unsigned long long* d_flags = nullptr;   // [2][MAX_SHARD] flags, [16 .. 24) pairwise epochs, [24] error
    bool peer_ok[MAX_SHARD] = {false};
    void* peer_ipc[MAX_SHARD][5] = {{nullptr}};
    std::string peer_blob[MAX_SHARD];
    bool sharded() const { return shard_world > 1; }
    const ShardCtx* shard_ctx() const { return shard_world > 1 ? &sctx : nullptr; }
    void close_peer(int p) {
        for (int k = 0; k < 5; k++) if (peer_ipc[p][k]) { cudaIpcCloseMemHandle(peer_ipc[p][k]); peer_ipc[p][k] = nullptr; }
        peer_ok[p] = false; peer_blob[p].clear();
    }

    SideStream side;                     // look-ahead stream of the blocked panel factorisation
    bool lookahead = true;
    bool solve_overlap = true;           // option "solve_overlap": solo and BIG supernodes of a level on two streams
    KernelTimer ktimer;                  // opb_profile_factor

    int fail(int code, const std::string& msg) { err = msg; return code; }
    int cuda_fail(cudaError_t e, const char* where) {
        err = std::string(where) + ": " + cudaGetErrorString(e);
        cudaGetLastError();
        return OPB_ERR_CUDA;
    }
};

#define CK(call)                                              \
    do {                                                      \
        cudaError_t e__ = (call);                             \
        if (e__ != cudaSuccess) return h->cuda_fail(e__, #call); \
    } while (0)

// Replay `enqueue` through a CUDA graph captured on first use (key distinguishes variants).
template <class F>
static void run_captured(opb_handle* h, opb_handle::GraphSlot& slot, int key, F&& enqueue) {
    if (!h->use_graphs) { enqueue(); return; }
    if (slot.exec && slot.key != key) { cudaGraphExecDestroy(slot.exec); slot = opb_handle::GraphSlot(); }
    if (!slot.exec) {
        const long long l0 = g_launches.load();
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
            cudaGetLastError(); h->use_graphs = false; enqueue(); return;
        }
        enqueue();
        cudaError_t e = cudaStreamEndCapture(h->stream, &graph);
        if (e == cudaSuccess) e = cudaGraphInstantiate(&slot.exec, graph, 0);
        if (graph) cudaGraphDestroy(graph);
        slot.launches = g_launches.load() - l0;
        g_launches.store(l0);
        slot.key = key;
        if (e != cudaSuccess) {      // capture not possible here: fall back to plain launches
            cudaGetLastError(); slot = opb_handle::GraphSlot(); h->use_graphs = false; enqueue(); return;
        }
    }
    cudaGraphLaunch(slot.exec, h->stream);
    count_launch((int)slot.launches);
}

static const char* kVersion = "onephase_b200 0.1 (sm_100a)";

extern "C" {

const char* opb_version(void) { return kVersion; }

int opb_cache_clear(void) {
    std::vector<std::shared_ptr<Bundle>> dropped;      // released outside the lock (cudaFree)
    {
        std::lock_guard<std::mutex> g(g_cache_mu);
        for (auto& kv : g_cache) dropped.push_back(std::move(kv.second));
        g_cache.clear();
        g_cache_order.clear();
    }
    return (int)dropped.size();
}
long long opb_launch_count(void) { return g_launches.load(); }

int opb_create(opb_handle** out, int device_id, unsigned flags) {
    (void)flags;
    if (!out) return OPB_ERR_INVALID;
    opb_handle* h = new opb_handle();
    *out = h;
    h->device = device_id;
    if (device_id >= 0) {
        cudaError_t e = cudaSetDevice(device_id);
        if (e != cudaSuccess) return h->cuda_fail(e, "cudaSetDevice");
        e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
        if (e != cudaSuccess) return h->cuda_fail(e, "cudaStreamCreate");
        h->own_stream = true;
        e = cudaMalloc((void**)&h->d_state_raw, sizeof(DeltaState) + 128);
        if (e != cudaSuccess) return h->cuda_fail(e, "cudaMalloc(state)");
        h->d_state = reinterpret_cast<DeltaState*>(h->d_state_raw);
        e = cudaMemsetAsync(h->d_state_raw, 0, sizeof(DeltaState) + 128, h->stream);
        if (e != cudaSuccess) return h->cuda_fail(e, "cudaMemset(state)");
        e = factor_configure();
        if (e != cudaSuccess) return h->cuda_fail(e, "factor_configure (is this an sm_100a device?)");
        e = dense_configure();
        if (e != cudaSuccess) return h->cuda_fail(e, "dense_configure (is this an sm_100a device?)");
        e = h->red.alloc(8);
        if (e != cudaSuccess) return h->cuda_fail(e, "cudaMalloc(red)");
        {
            // the side stream gets the highest priority the device offers: it carries the latency
            // chain of the blocked panel factorisation (option "chain_priority", default on)
            int least = 0, greatest = 0;
            cudaDeviceGetStreamPriorityRange(&least, &greatest);
            e = cudaStreamCreateWithPriority(&h->side.stream, cudaStreamNonBlocking, greatest);
            h->side.chain_on_side = greatest < least;
            // deep look-ahead: one stream per piece class, the sooner a piece is needed the higher its
            // priority; the rest of an outer update runs at the lowest
            for (int i = 0; i < LA_CLASSES && e == cudaSuccess; i++) {
                e = cudaStreamCreateWithPriority(&h->side.cls[i], cudaStreamNonBlocking, std::min(least, greatest + 1 + i));
                if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->side.cls_done[i], cudaEventDisableTiming);
            }
            if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->side.rest, cudaStreamNonBlocking, least);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->side.rest_done, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaStreamCreateWithPriority(&h->side.aux, cudaStreamNonBlocking, least);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->side.aux_fork, cudaEventDisableTiming);
            if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->side.aux_done, cudaEventDisableTiming);
            h->side.deep = true;
        }
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->side.fork, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->side.join, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&h->side.start, cudaEventDisableTiming);
        if (e != cudaSuccess) return h->cuda_fail(e, "side stream");
        for (auto pre : {preload_assembly, preload_vec, preload_solve, preload_factor, preload_dense, preload_shard}) {
            e = pre();
            if (e != cudaSuccess) return h->cuda_fail(e, "kernel preload (is this an sm_100a device?)");
        }

    }
    return OPB_OK;
}

int opb_destroy(opb_handle* h) {
    if (!h) return OPB_OK;
    if (h->device >= 0) {
        cudaSetDevice(h->device);
        if (h->stream) cudaStreamSynchronize(h->stream);
        DBuf<double>* bufs[] = {&h->Jv, &h->Hv, &h->y, &h->s, &h->sigma, &h->T, &h->Rval, &h->Mval, &h->sdiag,
                                &h->Lval, &h->CB, &h->dual_r, &h->primal_r, &h->comp_r, &h->b, &h->res,
                                &h->dx, &h->dy, &h->ds, &h->tm, &h->xw, &h->xw2, &h->uw, &h->userval, &h->Xinv, &h->Twork, &h->Dvec};
        for (auto* b : bufs) b->release();
        h->red.release();
        h->scr_i0.release(); h->scr_i1.release(); h->scr_d0.release(); h->scr_d1.release(); h->scr_d2.release();
        h->drop_graphs();
        for (int p = 0; p < MAX_SHARD; p++) h->close_peer(p);
        h->ktimer.release();
        if (h->side.stream) { cudaStreamSynchronize(h->side.stream); cudaStreamDestroy(h->side.stream); }
        for (int i = 0; i < LA_CLASSES; i++) {
            if (h->side.cls[i]) { cudaStreamSynchronize(h->side.cls[i]); cudaStreamDestroy(h->side.cls[i]); }
            if (h->side.cls_done[i]) cudaEventDestroy(h->side.cls_done[i]);
        }
        if (h->side.rest) { cudaStreamSynchronize(h->side.rest); cudaStreamDestroy(h->side.rest); }
        if (h->side.aux) { cudaStreamSynchronize(h->side.aux); cudaStreamDestroy(h->side.aux); }
        if (h->side.aux_fork) cudaEventDestroy(h->side.aux_fork);
        if (h->side.aux_done) cudaEventDestroy(h->side.aux_done);
        if (h->side.rest_done) cudaEventDestroy(h->side.rest_done);
        if (h->side.fork) cudaEventDestroy(h->side.fork);
        if (h->side.join) cudaEventDestroy(h->side.join);
        if (h->side.start) cudaEventDestroy(h->side.start);
        if (h->d_flags) cudaFree(h->d_flags);
        if (h->d_state_raw) cudaFree(h->d_state_raw);
        if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
    }
    h->B.reset();
    delete h;
    return OPB_OK;
}

const char* opb_last_error(const opb_handle* h) { return h ? h->err.c_str() : "null handle"; }

int opb_set_stream(opb_handle* h, void* cuda_stream) {
    if (!h) return OPB_ERR_INVALID;
    if (h->device < 0) return h->fail(OPB_ERR_NO_DEVICE, "host-only handle");
    if (h->own_stream && h->stream) { cudaStreamSynchronize(h->stream); cudaStreamDestroy(h->stream); }
    h->own_stream = false;
    h->stream = (cudaStream_t)cuda_stream;
    return OPB_OK;
}

int opb_set_option(opb_handle* h, const char* key, double v) {
    if (!h || !key) return OPB_ERR_INVALID;
    std::string k(key);
    if (k == "nd_leaf") h->opt.nd_leaf = (int)v;
    else if (k == "nd_balance") h->opt.nd_balance = v;
    else if (k == "ordering") h->opt.ordering = (int)v;
    else if (k == "relax") h->opt.relax_enable = (int)v;
    else if (k == "metis_max_n") h->opt.metis_max_n = (int)v;
    else if (k == "relax_small") h->opt.relax_small = v;
    else if (k == "shard_split_flops") h->opt.shard_split_flops = v;
    else if (k == "attempts_per_sync") h->attempts_per_sync = std::max(1, (int)v);
    else if (k == "outer_block") { h->outer_block = std::max(WB, ((int)v / WB) * WB); h->drop_graphs(); }
    else if (k == "ldlt_scalar") { h->ldlt_scalar = v != 0; h->drop_graphs(); }
    else if (k == "cb_small_k") { h->cb_small_k = (int)v; h->drop_graphs(); }
    else if (k == "barrier_timeout_s") { h->barrier_timeout_s = v; h->sctx.timeout_clocks = (long long)(v * 2.0e9); h->drop_graphs(); }
    else if (k == "lookahead") { h->lookahead = v != 0; h->side.deep = v >= 2; h->drop_graphs(); }
    else if (k == "chain_priority") { h->side.chain_on_side = v != 0; h->drop_graphs(); }
    else if (k == "graphs") { h->use_graphs = v != 0; h->drop_graphs(); }
    else if (k == "loop_graph") { h->loop_graph = v != 0; h->drop_graphs(); }
    else if (k == "solve_overlap") { h->solve_overlap = v != 0; h->drop_graphs(); }
    else return h->fail(OPB_ERR_INVALID, "unknown option " + k);
    return OPB_OK;
}

int opb_set_permutation(opb_handle* h, int64_t n, const int64_t* perm) {
    if (!h) return OPB_ERR_INVALID;
    if (!perm || n <= 0) { h->user_perm.clear(); return OPB_OK; }
    h->user_perm.assign(perm, perm + n);
    return OPB_OK;
}

// ---- one instance over several GPUs ---------------------------------------
// blob layout: [0] pid, [8] CB, [16] u, [24] x, [32] flags, [40] L (raw device pointers, valid inside the
// exporting process), [64 + 64 k] cudaIpcMemHandle_t of the same five buffers
constexpr int NPEERBUF = 5;
static_assert(sizeof(cudaIpcMemHandle_t) == 64, "blob layout");
static_assert(OPB_SHARD_BLOB_BYTES >= 64 + NPEERBUF * 64, "blob layout");

int opb_shard_init(opb_handle* h, int rank, int world) {
    if (!h) return OPB_ERR_INVALID;
    if (world < 1 || world > MAX_SHARD || rank < 0 || rank >= world) return h->fail(OPB_ERR_INVALID, "bad rank / world");
    if (h->B) return h->fail(OPB_ERR_STATE, "opb_shard_init must precede opb_set_structure");
    h->shard_rank = rank; h->shard_world = world;
    if (h->device >= 0 && world > 1) {
        cudaSetDevice(h->device);
        if (!h->d_flags) CK(cudaMalloc((void**)&h->d_flags, 32 * sizeof(unsigned long long)));
        CK(cudaMemsetAsync(h->d_flags, 0, 32 * sizeof(unsigned long long), h->stream));
        CK(cudaStreamSynchronize(h->stream));
        h->sctx = ShardCtx();
        h->sctx.rank = rank; h->sctx.world = world;
        h->sctx.flags_local = h->d_flags;
        h->sctx.flags_peer[rank] = h->d_flags;
        h->sctx.epoch = h->d_flags + 2 * MAX_SHARD;
        h->sctx.error = reinterpret_cast<int*>(h->d_flags + 3 * MAX_SHARD);
        h->sctx.state = h->d_state;
        h->sctx.timeout_clocks = (long long)(h->barrier_timeout_s * 2.0e9);
    }
    return OPB_OK;
}

int opb_shard_export(opb_handle* h, unsigned char* blob) {
    if (!h || !blob) return OPB_ERR_INVALID;
    if (h->device < 0) return h->fail(OPB_ERR_NO_DEVICE, "host-only handle");
    if (!h->sharded() || !h->B) return h->fail(OPB_ERR_STATE, "opb_shard_init + opb_set_structure first");
    cudaSetDevice(h->device);
    memset(blob, 0, OPB_SHARD_BLOB_BYTES);
    const uint64_t pid = (uint64_t)getpid();
    void* ptrs[NPEERBUF] = {h->CB.p, h->uw.p, h->xw.p, h->d_flags, h->Lval.p};
    memcpy(blob, &pid, 8);
    for (int k = 0; k < NPEERBUF; k++) {
        memcpy(blob + 8 + 8 * k, &ptrs[k], 8);
        cudaIpcMemHandle_t ih;
        CK(cudaIpcGetMemHandle(&ih, ptrs[k]));
        memcpy(blob + 64 + 64 * k, &ih, 64);
    }
    return OPB_OK;
}

int opb_shard_attach(opb_handle* h, int peer, const unsigned char* blob) {
    if (!h || !blob) return OPB_ERR_INVALID;
    if (h->device < 0) return h->fail(OPB_ERR_NO_DEVICE, "host-only handle");
    if (!h->sharded() || !h->B) return h->fail(OPB_ERR_STATE, "opb_shard_init + opb_set_structure first");
    if (peer < 0 || peer >= h->shard_world || peer == h->shard_rank) return h->fail(OPB_ERR_INVALID, "bad peer");
    cudaSetDevice(h->device);
    h->drop_graphs();
    const std::string nb(reinterpret_cast<const char*>(blob), OPB_SHARD_BLOB_BYTES);
    void* ptrs[NPEERBUF];
    uint64_t pid;
    memcpy(&pid, blob, 8);
    if (pid == (uint64_t)getpid()) {
        // same process (several handles on one or more devices): the raw pointers are usable.
        // Graph instantiation / first launch may synchronise the whole context; with a peer of
        // the same context waiting in a barrier for THIS thread's next launch that would stall
        // until the barrier times out, so same-process peers use plain launches.  (Between
        // processes the ranks' contexts are independent and the graphs stay on.)
        h->use_graphs = false;
        // ... and the two-stream look-ahead instead of the deep one: every stream of every virtual rank
        // needs its own hardware queue (a barrier kernel spinning in a shared queue blocks the peer it
        // waits for), and ten streams per handle exhaust the connections of one context
        h->side.deep = false;
        h->close_peer(peer);
        for (int k = 0; k < NPEERBUF; k++) memcpy(&ptrs[k], blob + 8 + 8 * k, 8);
    } else if (h->peer_blob[peer] == nb && h->peer_ipc[peer][0]) {
        for (int k = 0; k < NPEERBUF; k++) ptrs[k] = h->peer_ipc[peer][k];     // same allocations as before
    } else {
        h->close_peer(peer);
        for (int k = 0; k < NPEERBUF; k++) {
            cudaIpcMemHandle_t ih;
            memcpy(&ih, blob + 64 + 64 * k, 64);
            CK(cudaIpcOpenMemHandle(&ptrs[k], ih, cudaIpcMemLazyEnablePeerAccess));
            h->peer_ipc[peer][k] = ptrs[k];
        }
    }
    h->peer_blob[peer] = nb;
    h->dev.cb_peer[peer] = static_cast<double*>(ptrs[0]);
    h->dev.u_peer[peer] = static_cast<double*>(ptrs[1]);
    h->dev.x_peer[peer] = static_cast<double*>(ptrs[2]);
    h->sctx.flags_peer[peer] = static_cast<unsigned long long*>(ptrs[3]);
    h->dev.l_peer[peer] = static_cast<double*>(ptrs[4]);
    h->peer_ok[peer] = true;
    return OPB_OK;
}

static int upload_bundle(opb_handle* h, Bundle& B) {
    cudaStream_t st = h->stream;
    Symbolic& S = B.S;
    B.device = h->device;
    CK(B.d_sfirst.upload(S.sfirst, st)); CK(B.d_rowptr.upload(S.rowptr, st));
    CK(B.d_rowidx.upload(S.rowidx, st)); CK(B.d_rel.upload(S.rel, st));
    CK(B.d_Loff.upload(S.Loff, st)); CK(B.d_CBoff.upload(S.CBoff, st));
    CK(B.d_sparent.upload(S.sparent, st)); CK(B.d_child_ptr.upload(S.child_ptr, st));
    CK(B.d_child_list.upload(S.child_list, st)); CK(B.d_perm.upload(S.perm, st));
    CK(B.d_sched.upload(B.sched, st));
    CK(B.d_Xoff.upload(B.Xoff, st));
    CK(B.d_gptr.upload(S.gptr, st)); CK(B.d_gsrc.upload(S.gsrc, st)); CK(B.d_gch.upload(S.gch, st));
    CK(B.d_tcut_ptr.upload(S.tcut_ptr, st)); CK(B.d_tcut.upload(S.tcut, st));
    if (B.world > 1) { CK(B.d_owner.upload(B.shard.owner, st)); CK(B.d_colowner.upload(B.colowner, st)); CK(B.d_cb_mirror.upload(B.cb_mirror, st)); }
    {
        // diagonal entries carry a flag so the scatter kernel adds delta to them
        std::vector<int64_t> amap = S.amap;
        const int64_t flag = (int64_t)1 << 62;
        for (int j = 0; j < S.n; j++) amap[B.Mp[j]] |= flag;   // first entry of each column = diagonal
        CK(B.d_amap.upload(amap, st));
        CK(cudaStreamSynchronize(st));
    }
    CK(B.d_dpos.upload(S.dpos, st)); CK(B.d_Mp.upload(B.Mp, st));
    if (B.schur) {
        SchurPattern& P = B.P;
        CK(B.d_pair_ptr.upload(P.pair_ptr, st)); CK(B.d_pairA.upload(P.pairA, st));
        CK(B.d_pairB.upload(P.pairB, st)); CK(B.d_hmap.upload(P.hmap, st));
        CK(B.d_Jrow.upload(P.Jrow, st)); CK(B.d_Jp.upload(P.Jp, st));
        CK(B.d_Rp.upload(P.Rp, st)); CK(B.d_Rcol.upload(P.Rcol, st)); CK(B.d_Rpos.upload(P.Rpos, st));
        CK(B.d_Sp.upload(P.Sp, st)); CK(B.d_Scol.upload(P.Scol, st)); CK(B.d_Spos.upload(P.Spos, st));
    } else {
        CK(B.d_src.upload(B.src, st));
    }
    CK(cudaStreamSynchronize(st));
    B.dev.n = S.n; B.dev.nsuper = S.nsuper;
    B.dev.sfirst = B.d_sfirst.p; B.dev.rowptr = B.d_rowptr.p; B.dev.rowidx = B.d_rowidx.p;
    B.dev.rel = B.d_rel.p; B.dev.Loff = B.d_Loff.p; B.dev.CBoff = B.d_CBoff.p;
    B.dev.sparent = B.d_sparent.p; B.dev.child_ptr = B.d_child_ptr.p; B.dev.child_list = B.d_child_list.p;
    B.dev.perm = B.d_perm.p; B.dev.Xoff = B.d_Xoff.p;
    B.dev.gptr = B.d_gptr.p; B.dev.gsrc = B.d_gsrc.p; B.dev.gch = B.d_gch.p;
    B.dev.tcut_ptr = B.d_tcut_ptr.p; B.dev.tcut = B.d_tcut.p;
    B.dev.owner = B.world > 1 ? B.d_owner.p : nullptr;
    B.dev.cb_mirror = B.world > 1 ? B.d_cb_mirror.p : nullptr;
    B.dev.rank = B.rank; B.dev.world = B.world;
    return OPB_OK;
}

static int alloc_numeric(opb_handle* h) {
    h->drop_graphs();       // buffers and schedules may change
    Bundle& B = *h->B;
    const Symbolic& S = B.S;
    const int n = S.n;
    CK(h->Mval.alloc(B.Mp[n])); CK(h->sdiag.alloc(n));
    CK(h->Lval.alloc(S.nnzL)); CK(h->CB.alloc(S.cb_total));
    CK(h->xw.alloc(n)); CK(h->xw2.alloc(n)); CK(h->uw.alloc(S.rowidx.size()));
    {
        // inverses of the big supernodes' pivot blocks; the strictly upper parts must stay zero
        const size_t had = h->Xinv.n;
        CK(h->Xinv.alloc((size_t)B.x_total)); CK(h->Twork.alloc((size_t)B.x_total));
        if (B.x_total && (h->Xinv.n != had || true))
            CK(cudaMemsetAsync(h->Xinv.p, 0, (size_t)B.x_total * sizeof(double), h->stream));
    }
    CK(h->res.alloc(n)); CK(h->dx.alloc(n)); CK(h->b.alloc(n));
    if (B.schur) {
        const int m = B.P.m;
        CK(h->Jv.alloc(B.P.nnzJ)); CK(h->Hv.alloc(B.P.nnzH)); CK(h->y.alloc(m)); CK(h->s.alloc(m));
        CK(h->sigma.alloc(m)); CK(h->T.alloc(B.P.nnzJ)); CK(h->Rval.alloc(B.P.nnzJ));
        CK(h->dual_r.alloc(n)); CK(h->primal_r.alloc(m)); CK(h->comp_r.alloc(m));
        CK(h->dy.alloc(m)); CK(h->ds.alloc(m)); CK(h->tm.alloc(m));
    }
    CK(h->Dvec.alloc(n));
    h->dev = B.dev;
    h->dev.dvec = h->Dvec.p;
    if (h->sharded()) {
        // own buffers; the peers' are attached by opb_shard_attach (again after every new structure)
        h->dev.cb_peer[h->shard_rank] = h->CB.p; h->dev.u_peer[h->shard_rank] = h->uw.p;
        h->dev.x_peer[h->shard_rank] = h->xw.p;
        h->dev.l_peer[h->shard_rank] = h->Lval.p;
        for (int p = 0; p < h->shard_world; p++) if (p != h->shard_rank) h->peer_ok[p] = false;
    }
    return OPB_OK;
}

static std::string cache_key(const opb_handle* h, uint64_t h1, uint64_t h2) {
    const SymOptions& o = h->opt;
    char buf[240];
    snprintf(buf, sizeof buf, "%d|%d|%.4f|%d|%d|%d|%.3f|%d|%d/%d|%.3g|%016llx|%016llx", h->device, o.nd_leaf, o.nd_balance, o.ordering,
             o.metis_max_n, o.relax_enable, o.relax_small, h->user_perm.empty() ? 0 : 1, h->shard_rank, h->shard_world,
             o.shard_split_flops, (unsigned long long)h1, (unsigned long long)h2);
    return buf;
}

static std::shared_ptr<Bundle> cache_get(const std::string& key) {
    std::lock_guard<std::mutex> g(g_cache_mu);
    auto it = g_cache.find(key);
    return it == g_cache.end() ? nullptr : it->second;
}
static void cache_put(const std::string& key, std::shared_ptr<Bundle> b) {
    std::lock_guard<std::mutex> g(g_cache_mu);
    if (g_cache.count(key)) return;
    if (g_cache_order.size() >= CACHE_MAX) {
        g_cache.erase(g_cache_order.front());
        g_cache_order.erase(g_cache_order.begin());
    }
    g_cache[key] = b;
    g_cache_order.push_back(key);
}

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int finish_structure(opb_handle* h, std::shared_ptr<Bundle> B, const std::string& key) {
    double t0 = now_s();
    auto mark = [&](const char* name) { const double t = now_s(); B->timing.emplace_back(name, t - t0); t0 = t; };
    const int64_t* up = h->user_perm.empty() ? nullptr : h->user_perm.data();
    if (up && (int64_t)h->user_perm.size() != (int64_t)(B->Mp.size() - 1))
        return h->fail(OPB_ERR_INVALID, "permutation length does not match n");
    {
        // METIS keeps process-wide random-number state: concurrent analyses (several handles driven
        // by host threads) would not be reproducible, and the ranks of a sharded instance must
        // derive the SAME ordering.  One analysis at a time.
        static std::mutex analyze_mu;
        std::lock_guard<std::mutex> g(analyze_mu);
        if (!analyze((int)(B->Mp.size() - 1), B->Mp, B->Mi, h->opt, up, B->S))
            return h->fail(OPB_ERR_INTERNAL, "symbolic analysis failed: " + B->S.error);
    }
    mark("analyze");
    B->rank = h->shard_rank; B->world = h->shard_world;
    if (B->world > 1) {
        shard_map(B->S, B->world, h->opt.shard_split_flops, B->shard);
        B->colowner.resize(B->S.n);
        for (int j = 0; j < B->S.n; j++) B->colowner[j] = B->shard.owner[B->S.col2super[j]];
    }
    mark("shard_map");
    build_plan(*B);
    mark("plan");
    if (h->device >= 0) {
        cudaSetDevice(h->device);
        int rc = upload_bundle(h, *B);
        if (rc) return rc;
        mark("upload");
    }
    cache_put(key, B);
    h->B = B;
    h->cached_hit = false;
    h->ready = opb_handle::NOT_READY;
    if (h->device >= 0) return alloc_numeric(h);
    return OPB_OK;
}

int opb_set_structure(opb_handle* h, int64_t n, int64_t m, const int64_t* Jp, const int64_t* Ji,
                      const int64_t* Hp, const int64_t* Hi, int base) {
    if (!h) return OPB_ERR_INVALID;
    if (!Jp || !Hp || (base != 0 && base != 1) || n <= 0 || m < 0) return h->fail(OPB_ERR_INVALID, "bad arguments");
    if (h->device >= 0) cudaSetDevice(h->device);
    const int64_t nnzJ = Jp[n] - base, nnzH = Hp[n] - base;
    if (nnzJ < 0 || nnzH < 0) return h->fail(OPB_ERR_INVALID, "bad colptr");
    uint64_t h1 = pattern_hash(n, Jp, Ji, nnzJ) ^ (uint64_t)m * 0x9e3779b97f4a7c15ull;
    uint64_t h2 = pattern_hash(n, Hp, Hi, nnzH) + (uint64_t)base;
    if (!h->user_perm.empty()) h2 ^= pattern_hash(0, h->user_perm.data(), h->user_perm.data(), (int64_t)h->user_perm.size());
    std::string key = "S|" + cache_key(h, h1, h2);
    if (auto B = cache_get(key)) {
        if (B->S.n == n && B->P.m == m && B->P.nnzJ == nnzJ && B->P.nnzH == nnzH &&
            B->same_pattern(0, n, Jp, Ji, base) && B->same_pattern(1, n, Hp, Hi, base)) {
            const bool same = h->B == B;
            h->B = B; h->cached_hit = true; h->ready = opb_handle::NOT_READY;
            if (h->device >= 0 && !same) return alloc_numeric(h);
            return OPB_OK;
        }
    }
    auto B = std::make_shared<Bundle>();
    B->schur = true;
    std::string err;
    const double tp = now_s();
    if (!build_schur_pattern(n, m, Jp, Ji, Hp, Hi, base, B->P, err)) return h->fail(OPB_ERR_INVALID, err);
    B->timing.emplace_back("pattern", now_s() - tp);
    B->keep_pattern(0, n, Jp, Ji, base); B->keep_pattern(1, n, Hp, Hi, base);
    B->Mp = B->P.Mp; B->Mi = B->P.Mi;
    return finish_structure(h, B, key);
}

// ---- staged (asynchronous) building blocks ---------------------------------
static int need_device(opb_handle* h) {
    if (!h) return OPB_ERR_INVALID;
    if (h->device < 0) return h->fail(OPB_ERR_NO_DEVICE, "numeric call on a host-only handle (no CPU fallback)");
    cudaSetDevice(h->device);
    if (h->sharded() && h->B)
        for (int p = 0; p < h->shard_world; p++)
            if (p != h->shard_rank && !h->peer_ok[p])
                return h->fail(OPB_ERR_STATE, "sharded handle: peer buffers not attached (opb_shard_export / opb_shard_attach after opb_set_structure)");
    return OPB_OK;
}

static int stage_form(opb_handle* h) {
    Bundle& B = *h->B;
    cudaStream_t st = h->stream;
    launch_prep(h->Jv.p, B.d_Jrow.p, B.d_Rpos.p, h->y.p, h->s.p, h->sigma.p, h->T.p, h->Rval.p, B.P.nnzJ, B.P.m, h->d_state, st);
    launch_assemble_M(B.d_pair_ptr.p, B.d_pairA.p, B.d_pairB.p, B.d_hmap.p, h->T.p, h->Jv.p, h->Hv.p,
                      h->Mval.p, B.Mp[B.S.n], st);
    launch_diag_extract(B.d_Mp.p, h->Mval.p, h->sdiag.p, h->d_state, B.S.n, st);
    CK(cudaGetLastError());
    h->ready = opb_handle::SYSTEM_FORMED;
    return OPB_OK;
}

static void enqueue_attempt_raw(opb_handle* h, KernelTimer* timer = nullptr, unsigned long long loop_handle = 0) {
    Bundle& B = *h->B;
    cudaStream_t st = h->stream;
    launch_ctl_begin(h->d_state, st);
    launch_scatter_fronts(h->Mval.p, B.d_amap.p, B.d_dpos.p, h->sdiag.p, h->Lval.p, B.S.nnzL,
                          B.Mp[B.S.n], B.S.n, h->d_state, 1, st);
    launch_factor_levels(h->dev, B.plan, B.d_sched.p, h->Lval.p, h->CB.p, h->Xinv.p, h->d_state, h->fmode(),
                         h->outer_block, h->cb_small_k, h->shard_ctx(), (h->lookahead && (!timer || timer->phases)) ? &h->side : nullptr, timer,
                         h->fmode() != FMODE_LDLT_SCALAR ? &B.trtri : nullptr, h->Twork.p, st);
    // sharded: every rank learns about a failed pivot anywhere before the delta rule is applied
    if (h->sharded()) launch_shard_barrier(h->sctx, shard_all(h->sctx), st);
    if (loop_handle) launch_ctl_end_loop(h->d_state, loop_handle, st);
    else launch_ctl_end(h->d_state, st);
}

// Build (once per structure) the graph  WHILE(!done) { attempt }  -- delta_strategy.jl:37-114 with
// the loop itself on the device.  Returns false when this runtime cannot do it (the caller falls
// back to host-driven chunks of attempts).
static bool build_loop_graph(opb_handle* h) {
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaGraphCreate(&graph, 0);
    const char* where = "cudaGraphCreate";
    cudaGraphConditionalHandle ch = 0;
    cudaGraphNodeParams np = {};
    cudaGraphNode_t node;
    const long long l0 = g_launches.load();
    if (e == cudaSuccess) { where = "cudaGraphConditionalHandleCreate"; e = cudaGraphConditionalHandleCreate(&ch, graph, 1, cudaGraphCondAssignDefault); }
    if (e == cudaSuccess) {
        np.type = cudaGraphNodeTypeConditional;
        np.conditional.handle = ch;
        np.conditional.type = cudaGraphCondTypeWhile;
        np.conditional.size = 1;
        where = "cudaGraphAddNode(conditional)";
        e = cudaGraphAddNode(&node, graph, nullptr, 0, &np);
    }
    if (e == cudaSuccess) {
        cudaGraph_t body = np.conditional.phGraph_out[0];
        where = "cudaStreamBeginCaptureToGraph";
        e = cudaStreamBeginCaptureToGraph(h->stream, body, nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed);
        if (e == cudaSuccess) {
            enqueue_attempt_raw(h, nullptr, (unsigned long long)ch);
            cudaGraph_t out = nullptr;
            where = "cudaStreamEndCapture";
            e = cudaStreamEndCapture(h->stream, &out);
        }
    }
    h->g_loop.launches = g_launches.load() - l0;
    g_launches.store(l0);
    if (e == cudaSuccess) { where = "cudaGraphInstantiate"; e = cudaGraphInstantiate(&h->g_loop.exec, graph, 0); }
    if (graph) cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        h->loop_diag = std::string(where) + ": " + cudaGetErrorString(e);
        cudaGetLastError();
        h->g_loop = opb_handle::GraphSlot();
        return false;
    }
    h->loop_diag = "active";
    h->g_loop.key = h->mode;
    return true;
}

static void enqueue_attempt(opb_handle* h) {
    run_captured(h, h->g_attempt, h->mode, [&] { enqueue_attempt_raw(h); });
}

static int read_state(opb_handle* h) {
    CK(cudaMemcpyAsync(&h->h_state, h->d_state, sizeof(DeltaState), cudaMemcpyDeviceToHost, h->stream));
    int shard_err = 0;
    if (h->sharded()) CK(cudaMemcpyAsync(&shard_err, h->sctx.error, sizeof(int), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (h->loop_pending) {       // the device ran num_fac attempts of the loop graph; one was counted at launch
        if (h->h_state.num_fac > 1) count_launch((int)((h->h_state.num_fac - 1) * h->loop_pending));
        h->loop_pending = 0;
    }
    if (shard_err) {
        unsigned long long f[32] = {0};
        cudaMemcpy(f, h->d_flags, sizeof f, cudaMemcpyDeviceToHost);
        char buf[640];
        int o = snprintf(buf, sizeof buf, "sharded instance: a peer GPU did not reach the barrier (timeout); rank %d/%d, per peer (own epoch | peer's flags (epoch,fail) in slots 0/1):",
                         h->shard_rank, h->shard_world);
        for (int p = 0; p < h->shard_world && o < 600; p++)
            o += snprintf(buf + o, sizeof buf - o, " [%d] %llu | (%llu,%llu) (%llu,%llu)", p, f[2 * MAX_SHARD + p],
                          f[p] >> 1, f[p] & 1, f[MAX_SHARD + p] >> 1, f[MAX_SHARD + p] & 1);
        return h->fail(OPB_ERR_INTERNAL, buf);
    }
    return OPB_OK;
}

// enqueue attempts until the device controller reports done
static int run_delta_loop(opb_handle* h, bool first_chunk_only) {
    // one graph launch runs the whole loop on the device; nothing to wait for here
    if (h->use_graphs && h->loop_graph && !h->sharded()) {
        if (h->g_loop.exec && h->g_loop.key != h->mode) { cudaGraphExecDestroy(h->g_loop.exec); h->g_loop = opb_handle::GraphSlot(); }
        if (h->g_loop.exec || build_loop_graph(h)) {
            CK(cudaGraphLaunch(h->g_loop.exec, h->stream));
            count_launch((int)h->g_loop.launches);
            h->loop_pending = h->g_loop.launches;
            return OPB_OK;
        }
        h->loop_graph = false;
    }
    for (;;) {
        for (int a = 0; a < h->attempts_per_sync; a++) enqueue_attempt(h);
        CK(cudaGetLastError());
        if (first_chunk_only) return OPB_OK;
        int rc = read_state(h);
        if (rc) return rc;
        if (h->h_state.done) return OPB_OK;
    }
}

int opb_upload_values(opb_handle* h, const double* Jx, const double* Hx, const double* y, const double* s) {
    int rc = need_device(h); if (rc) return rc;
    if (!h->B || !h->B->schur) return h->fail(OPB_ERR_STATE, "opb_set_structure has not been called");
    Bundle& B = *h->B;
    cudaStream_t st = h->stream;
    if (B.P.nnzJ) CK(cudaMemcpyAsync(h->Jv.p, Jx, B.P.nnzJ * sizeof(double), cudaMemcpyHostToDevice, st));
    if (B.P.nnzH) CK(cudaMemcpyAsync(h->Hv.p, Hx, B.P.nnzH * sizeof(double), cudaMemcpyHostToDevice, st));
    if (B.P.m) {
        CK(cudaMemcpyAsync(h->y.p, y, B.P.m * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->s.p, s, B.P.m * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    return OPB_OK;
}

int opb_form_resident(opb_handle* h) {
    int rc = need_device(h); if (rc) return rc;
    if (!h->B || !h->B->schur) return h->fail(OPB_ERR_STATE, "opb_set_structure has not been called");
    return stage_form(h);
}

int opb_form(opb_handle* h, const double* Jx, const double* Hx, const double* y, const double* s,
             double* schur_diag_out, double* diag_min_out) {
    int rc = opb_upload_values(h, Jx, Hx, y, s); if (rc) return rc;
    rc = stage_form(h); if (rc) return rc;
    if (schur_diag_out)
        CK(cudaMemcpyAsync(schur_diag_out, h->sdiag.p, h->B->S.n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    rc = read_state(h); if (rc) return rc;
    if (diag_min_out) *diag_min_out = h->h_state.diag_min;
    return OPB_OK;
}

int opb_get_M_pattern(opb_handle* h, int64_t* cp, int64_t* ri) {
    if (!h || !h->B) return OPB_ERR_STATE;
    const Bundle& B = *h->B;
    if (cp) memcpy(cp, B.Mp.data(), B.Mp.size() * sizeof(int64_t));
    if (ri) for (size_t k = 0; k < B.Mi.size(); k++) ri[k] = B.Mi[k];
    return OPB_OK;
}

int opb_get_M_values(opb_handle* h, double* out) {
    int rc = need_device(h); if (rc) return rc;
    if (!h->B || h->ready == opb_handle::NOT_READY) return h->fail(OPB_ERR_STATE, "system not formed");
    CK(cudaMemcpyAsync(out, h->Mval.p, h->B->Mp[h->B->S.n] * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return OPB_OK;
}

int opb_delta_loop_resident(opb_handle* h, double delta_prev, double delta_zero, double delta_min,
                            double delta_max, double delta_start, double inc, double dec, int max_it) {
    int rc = need_device(h); if (rc) return rc;
    if (!h->B || h->ready == opb_handle::NOT_READY) return h->fail(OPB_ERR_STATE, "kkt solver not ready to factor (form_system first)");
    h->mode = OPB_MODE_CHOLESKY;
    launch_ctl_init(h->d_state, delta_prev, delta_zero, delta_min, delta_max, delta_start, inc, dec,
                    max_it, OPB_MODE_CHOLESKY, h->stream);
    rc = run_delta_loop(h, false);
    if (rc) return rc;
    h->ready = opb_handle::FACTORED;
    return OPB_OK;       // with the loop graph the call is asynchronous: opb_sync_state / the next read waits
}

int opb_factor_delta_loop(opb_handle* h, double delta_prev, double delta_zero, double delta_min,
                          double delta_max, double delta_start, double inc, double dec, int max_it,
                          double* delta_out, int* num_fac_out, int* status_out) {
    int rc = opb_delta_loop_resident(h, delta_prev, delta_zero, delta_min, delta_max, delta_start, inc, dec, max_it);
    if (rc) return rc;
    rc = read_state(h); if (rc) return rc;
    if (delta_out) *delta_out = h->h_state.delta;
    if (num_fac_out) *num_fac_out = h->h_state.num_fac;
    if (status_out) *status_out = h->h_state.status;
    return OPB_OK;
}

static int single_factor(opb_handle* h, double delta, int mode, int* ok) {
    h->mode = mode;
    launch_ctl_single(h->d_state, delta, mode, h->stream);
    enqueue_attempt(h);
    if (mode == OPB_MODE_LDLT)
        launch_ldlt_inertia(h->dev, h->Lval.p, h->B->d_dpos.p, h->B->S.n, h->d_state, h->stream);
    CK(cudaGetLastError());
    int rc = read_state(h); if (rc) return rc;
    h->ready = opb_handle::FACTORED;
    if (ok) *ok = (h->h_state.status == 1) ? 1 : 0;
    return OPB_OK;
}

int opb_factor(opb_handle* h, double delta, int* inertia_ok) {
    int rc = need_device(h); if (rc) return rc;
    if (!h->B || h->ready == opb_handle::NOT_READY) return h->fail(OPB_ERR_STATE, "kkt solver not ready to factor (form_system first)");
    return single_factor(h, delta, OPB_MODE_CHOLESKY, inertia_ok);
}

int opb_profile_factor(opb_handle* h, double delta, double* total_ms, double* cb_ms, double* update_ms,
                       double* cb_flops, double* update_flops, int* inertia_ok) {
    int rc = need_device(h); if (rc) return rc;
    if (!h->B || h->ready == opb_handle::NOT_READY) return h->fail(OPB_ERR_STATE, "kkt solver not ready to factor (form_system first)");
    if (h->sharded()) return h->fail(OPB_ERR_INVALID, "opb_profile_factor is not available on a sharded handle");
    h->mode = OPB_MODE_CHOLESKY;
    cudaStream_t st = h->stream;
    KernelTimer& T = h->ktimer;
    T.used = 0;
    cudaEvent_t e0 = T.next(2), e1 = T.next(2);            // the whole attempt
    launch_ctl_single(h->d_state, delta, OPB_MODE_CHOLESKY, st);
    CK(cudaEventRecord(e0, st));
    enqueue_attempt_raw(h, &T);                              // plain launches, no look-ahead: the timed kernels do not overlap
    CK(cudaEventRecord(e1, st));
    CK(cudaGetLastError());
    rc = read_state(h); if (rc) return rc;
    h->ready = opb_handle::FACTORED;
    if (inertia_ok) *inertia_ok = (h->h_state.status == 1) ? 1 : 0;
    float ms = 0.f;
    double sum[2] = {0.0, 0.0};
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (total_ms) *total_ms = ms;
    for (size_t p = 1; 2 * p + 1 < T.used; p++) {
        CK(cudaEventElapsedTime(&ms, T.ev[2 * p], T.ev[2 * p + 1]));
        if (T.kind[p] == 0 || T.kind[p] == 1) sum[T.kind[p]] += ms;
    }
    if (cb_ms) *cb_ms = sum[0];
    if (update_ms) *update_ms = sum[1];
    // algorithmic flops of the two kernels over the fronts they serve (MID, MIDL and BIG classes):
    // update block r^2 c (lower half, multiply-add), panel updates of the BIG fronts
    double fcb = 0.0, fup = 0.0;
    const Symbolic& S = h->B->S;
    for (int s = 0; s < S.nsuper; s++) {
        const double c = S.sfirst[s + 1] - S.sfirst[s], r = (double)(S.rowptr[s + 1] - S.rowptr[s]), N = c + r;
        if (N <= SMALL_N) continue;
        fcb += r * r * c;
        // BIG fronts: panel flops N c^2 - 2 c^3 / 3, minus what other kernels execute -- the TRSMs
        // (rows below each 128-column block times 128^2) and the diagonal blocks (128^3 / 3 each)
        if (!(c <= WB && N * c <= MIDL_PANEL))
            fup += N * c * c - 2.0 * c * c * c / 3.0 - WB * (N * c - 0.5 * c * c) - c * (double)WB * WB / 3.0;
    }
    if (cb_flops) *cb_flops = fcb;
    if (update_flops) *update_flops = fup;
    return OPB_OK;
}

int opb_profile_levels(opb_handle* h, double delta, double* out, int cap, int* nlevels_out, double* total_ms) {
    int rc = need_device(h); if (rc) return rc;
    if (!h->B || h->ready == opb_handle::NOT_READY) return h->fail(OPB_ERR_STATE, "kkt solver not ready to factor (form_system first)");
    h->mode = OPB_MODE_CHOLESKY;
    cudaStream_t st = h->stream;
    KernelTimer& T = h->ktimer;
    T.used = 0; T.marks_used = 0; T.phases = true;
    launch_ctl_single(h->d_state, delta, OPB_MODE_CHOLESKY, st);
    T.put_mark(4, st);
    enqueue_attempt_raw(h, &T);          // plain launches (no graph) on the look-ahead streams as configured
    T.put_mark(5, st);
    T.phases = false;
    CK(cudaGetLastError());
    rc = read_state(h); if (rc) return rc;
    h->ready = opb_handle::FACTORED;
    // per level: [before the big panels (small fronts, medium panels, extend-add), big panels, update blocks]
    const int nl = h->B->S.nlevels;
    if (nlevels_out) *nlevels_out = nl;
    std::vector<double> acc((size_t)nl * 3 + 2, 0.0);     // + [scatter before the levels, pivot-block inverses after them]
    int lvl = -1, phase = 0;
    float ms = 0.f;
    for (size_t k = 0; k + 1 < T.marks_used; k++) {
        const int kd = T.mark_kind[k];
        CK(cudaEventElapsedTime(&ms, T.mark[k], T.mark[k + 1]));
        if (kd == 4) { acc[(size_t)nl * 3] += ms; continue; }
        if (kd == 3) { acc[(size_t)nl * 3 + 1] += ms; continue; }
        if (kd == 0) { lvl++; phase = 0; } else if (kd == 1) phase = 1; else if (kd == 2) phase = 2; else continue;
        if (lvl < 0 || lvl >= nl) continue;
        acc[(size_t)lvl * 3 + phase] += ms;
    }
    for (int k = 0; k < nl * 3 + 2 && k < cap; k++) out[k] = acc[k];
    if (total_ms) { CK(cudaEventElapsedTime(&ms, T.mark[0], T.mark[T.marks_used - 1])); *total_ms = ms; }
    return OPB_OK;
}

int opb_upload_rhs(opb_handle* h, const double* dual_r, const double* primal_r, const double* comp_r) {
    int rc = need_device(h); if (rc) return rc;
    if (!h->B || !h->B->schur) return h->fail(OPB_ERR_STATE, "opb_set_structure has not been called");
    const int n = h->B->S.n, m = h->B->P.m;
    cudaStream_t st = h->stream;
    CK(cudaMemcpyAsync(h->dual_r.p, dual_r, n * sizeof(double), cudaMemcpyHostToDevice, st));
    if (m) {
        CK(cudaMemcpyAsync(h->primal_r.p, primal_r, m * sizeof(double), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->comp_r.p, comp_r, m * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    return OPB_OK;
}

static DirBuffers dir_buffers(opb_handle* h) {
    Bundle& B = *h->B;
    DirBuffers D{};
    D.n = B.S.n; D.m = B.P.m;
    D.Rp = B.d_Rp.p; D.Rcol = B.d_Rcol.p; D.Rval = h->Rval.p;
    D.Jp = B.d_Jp.p; D.Jrow = B.d_Jrow.p; D.Jv = h->Jv.p;
    D.Sp = B.d_Sp.p; D.Scol = B.d_Scol.p; D.Spos = B.d_Spos.p; D.Hv = h->Hv.p;
    D.y = h->y.p; D.s = h->s.p; D.sigma = h->sigma.p;
    D.dual_r = h->dual_r.p; D.primal_r = h->primal_r.p; D.comp_r = h->comp_r.p;
    D.b = h->b.p; D.res = h->res.p; D.dx = h->dx.p; D.dy = h->dy.p; D.ds = h->ds.p;
    D.tm = h->tm.p; D.tm2 = nullptr; D.red = h->red.p; D.st_d = h->d_state;
    D.wide_n = ((double)(B.P.nnzJ + (int64_t)B.P.Scol.size()) >= 24.0 * D.n) ? 1 : 0;
    D.wide_m = (D.m > 0 && (double)B.P.nnzJ >= 24.0 * D.m) ? 1 : 0;
    return D;
}

int opb_direction_resident(opb_handle* h, int n_refine) {
    int rc = need_device(h); if (rc) return rc;
    if (!h->B || !h->B->schur) return h->fail(OPB_ERR_STATE, "opb_set_structure has not been called");
    if (h->ready != opb_handle::FACTORED) return h->fail(OPB_ERR_STATE, "kkt solver not ready to compute direction!");
    Bundle& B = *h->B;
    cudaStream_t st = h->stream;
    DirBuffers D = dir_buffers(h);
    run_captured(h, h->g_direction, n_refine * 2 + h->mode, [&] {
        launch_schur_rhs(D, st);
        for (int it = 0; it < n_refine; it++) {
            launch_permute_in(h->res.p, B.d_perm.p, h->xw.p, D.n, st);
            launch_solve(h->dev, B.plan, B.d_sched.p, h->Lval.p, h->Xinv.p, h->xw.p, h->xw2.p, h->uw.p, h->fmode(),
                         h->shard_ctx(), B.d_colowner.p, h->solve_overlap ? &h->side : nullptr, st);
            launch_permute_out_add(h->xw.p, B.d_perm.p, h->dx.p, D.n, 1, st);
            // the reference also evaluates the residual after the last correction but only
            // prints it (schur.jl:177-179); it does not influence the direction
            if (it + 1 < n_refine) launch_residual(D, st);
        }
        launch_recover_and_error(D, st);
    });
    CK(cudaGetLastError());
    return OPB_OK;
}

int opb_direction(opb_handle* h, const double* dual_r, const double* primal_r, const double* comp_r,
                  int n_refine, double* dx, double* dy, double* ds, double* kkt_err) {
    int rc = opb_upload_rhs(h, dual_r, primal_r, comp_r); if (rc) return rc;
    rc = opb_direction_resident(h, n_refine); if (rc) return rc;
    const int n = h->B->S.n, m = h->B->P.m;
    cudaStream_t st = h->stream;
    if (dx) CK(cudaMemcpyAsync(dx, h->dx.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (dy && m) CK(cudaMemcpyAsync(dy, h->dy.p, m * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (ds && m) CK(cudaMemcpyAsync(ds, h->ds.p, m * sizeof(double), cudaMemcpyDeviceToHost, st));
    rc = read_state(h); if (rc) return rc;
    if (kkt_err) memcpy(kkt_err, h->h_state.kkt_err, 6 * sizeof(double));
    return OPB_OK;
}

// ---- the iterate stays resident, only (grad, cons) go up and scalars come back ----
int opb_system_rhs(opb_handle* h, const double* grad, const double* cons, double mu, double a_norm_penalty,
                   double eta_P, double eta_D, double eta_mu, double* dual_r_out, double* primal_r_out,
                   double* comp_r_out) {
    int rc = need_device(h); if (rc) return rc;
    if (!h->B || !h->B->schur) return h->fail(OPB_ERR_STATE, "opb_set_structure has not been called");
    if (h->ready == opb_handle::NOT_READY) return h->fail(OPB_ERR_STATE, "no resident iterate (opb_form / opb_upload_values + opb_form_resident first)");
    if (!grad || (h->B->P.m && !cons)) return h->fail(OPB_ERR_INVALID, "bad arguments");
    const int n = h->B->S.n, m = h->B->P.m;
    cudaStream_t st = h->stream;
    // grad -> b (n), cons -> tm (m): both are scratch of the direction, rewritten before they are read there
    CK(cudaMemcpyAsync(h->b.p, grad, n * sizeof(double), cudaMemcpyHostToDevice, st));
    if (m) CK(cudaMemcpyAsync(h->tm.p, cons, m * sizeof(double), cudaMemcpyHostToDevice, st));
    DirBuffers D = dir_buffers(h);
    launch_system_rhs(D, h->b.p, h->tm.p, mu * eta_mu, a_norm_penalty, eta_P, eta_D, st);
    CK(cudaGetLastError());
    if (dual_r_out) CK(cudaMemcpyAsync(dual_r_out, h->dual_r.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (primal_r_out && m) CK(cudaMemcpyAsync(primal_r_out, h->primal_r.p, m * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (comp_r_out && m) CK(cudaMemcpyAsync(comp_r_out, h->comp_r.p, m * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (dual_r_out || primal_r_out || comp_r_out) CK(cudaStreamSynchronize(st));
    return OPB_OK;
}

int opb_step_bounds(opb_handle* h, double frac_bd, double predict_exp, double* out4) {
    int rc = need_device(h); if (rc) return rc;
    if (!h->B || !h->B->schur) return h->fail(OPB_ERR_STATE, "opb_set_structure has not been called");
    if (h->ready != opb_handle::FACTORED || !out4) return h->fail(OPB_ERR_STATE, "no direction");
    DirBuffers D = dir_buffers(h);
    launch_step_bounds(D, frac_bd, predict_exp, h->stream);
    CK(cudaGetLastError());
    unsigned long long red[4];
    CK(cudaMemcpyAsync(red, h->red.p, sizeof red, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < 4; k++) memcpy(&out4[k], &red[k], 8);
    out4[3] = 1.0 / (out4[3] > 1.0 || out4[3] != out4[3] ? out4[3] : 1.0);       // simple_max_step = 1 / max(1, ratios)
    return OPB_OK;
}

int opb_get_direction(opb_handle* h, double* dx, double* dy, double* ds, double* kkt_err) {
    int rc = need_device(h); if (rc) return rc;
    if (!h->B || !h->B->schur) return h->fail(OPB_ERR_STATE, "opb_set_structure has not been called");
    if (h->ready != opb_handle::FACTORED) return h->fail(OPB_ERR_STATE, "no direction");
    const int n = h->B->S.n, m = h->B->P.m;
    cudaStream_t st = h->stream;
    if (dx) CK(cudaMemcpyAsync(dx, h->dx.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (dy && m) CK(cudaMemcpyAsync(dy, h->dy.p, m * sizeof(double), cudaMemcpyDeviceToHost, st));
    if (ds && m) CK(cudaMemcpyAsync(ds, h->ds.p, m * sizeof(double), cudaMemcpyDeviceToHost, st));
    rc = read_state(h); if (rc) return rc;
    if (kkt_err) memcpy(kkt_err, h->h_state.kkt_err, 6 * sizeof(double));
    return OPB_OK;
}

int opb_solve_resident(opb_handle* h, int nsolves) {
    int rc = need_device(h); if (rc) return rc;
    if (!h->B || h->ready != opb_handle::FACTORED) return h->fail(OPB_ERR_STATE, "no factor");
    Bundle& B = *h->B;
    for (int k = 0; k < nsolves; k++)
        run_captured(h, h->g_solve, h->mode, [&] {
            launch_permute_in(h->res.p, B.d_perm.p, h->xw.p, B.S.n, h->stream);
            launch_solve(h->dev, B.plan, B.d_sched.p, h->Lval.p, h->Xinv.p, h->xw.p, h->xw2.p, h->uw.p, h->fmode(),
                         h->shard_ctx(), B.d_colowner.p, h->solve_overlap ? &h->side : nullptr, h->stream);
            launch_permute_out_add(h->xw.p, B.d_perm.p, h->b.p, B.S.n, 0, h->stream);
        });
    CK(cudaGetLastError());
    return OPB_OK;
}

int opb_sync_state(opb_handle* h, double* delta_out, int* num_fac_out, int* status_out, double* kkt_err_out) {
    int rc = need_device(h); if (rc) return rc;
    rc = read_state(h); if (rc) return rc;
    if (delta_out) *delta_out = h->h_state.delta;
    if (num_fac_out) *num_fac_out = h->h_state.num_fac;
    if (status_out) *status_out = h->h_state.status;
    if (kkt_err_out) memcpy(kkt_err_out, h->h_state.kkt_err, 6 * sizeof(double));
    return OPB_OK;
}

}  // extern "C"

// ---- L1 compatibility path: arbitrary CSC matrix -----------------------------
namespace {
__global__ void csc_gather_kernel(const double* __restrict__ user, const int64_t* __restrict__ src,
                                  double* __restrict__ Mval, int64_t nnzM) {
    int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < nnzM) { int64_t p = src[e]; Mval[e] = p >= 0 ? user[p] : 0.0; }
}
template <class T>
int64_t copy_out(const std::vector<T>& v, int64_t* out, int64_t cap) {
    int64_t n = (int64_t)v.size();
    if (out) for (int64_t k = 0; k < std::min(n, cap); k++) out[k] = (int64_t)v[k];
    return n;
}
}  // namespace

extern "C" {

int opb_ls_factor_csc(opb_handle* h, int64_t dim, const int64_t* cp, const int64_t* ri, const double* nz,
                      int base, int mode, int64_t n_pos, int64_t m_neg, int* inertia_ok) {
    if (!h) return OPB_ERR_INVALID;
    if (!cp || !ri || !nz || dim <= 0 || (base != 0 && base != 1)) return h->fail(OPB_ERR_INVALID, "bad arguments");
    if (mode != OPB_MODE_CHOLESKY && mode != OPB_MODE_LDLT) return h->fail(OPB_ERR_INVALID, "bad mode");
    if (mode == OPB_MODE_CHOLESKY && m_neg != 0) return h->fail(OPB_ERR_INVALID, "Cholesky requires m == 0 (julia.jl:30)");
    if (h->sharded()) return h->fail(OPB_ERR_INVALID, "opb_ls_factor_csc is not available on a sharded handle");
    int rc = need_device(h); if (rc) return rc;
    const int64_t nnz = cp[dim] - base;
    uint64_t h1 = pattern_hash(dim, cp, ri, nnz) + (uint64_t)base;
    std::string key = "C|" + cache_key(h, h1, 0);
    std::shared_ptr<Bundle> B = cache_get(key);
    if (B && (B->S.n != dim || B->schur || !B->same_pattern(0, dim, cp, ri, base))) B.reset();
    if (B) {
        // the same analysis as last time (one ls_factor! per delta attempt): buffers and CUDA graphs stay
        const bool same = h->B == B;
        h->B = B; h->cached_hit = true;
        if (!same) { rc = alloc_numeric(h); if (rc) return rc; }
    } else {
        B = std::make_shared<Bundle>();
        B->schur = false;
        std::string err;
        if (!build_csc_pattern(dim, cp, ri, base, B->Mp, B->Mi, B->src, err)) return h->fail(OPB_ERR_INVALID, err);
        B->keep_pattern(0, dim, cp, ri, base);
        rc = finish_structure(h, B, key); if (rc) return rc;
    }
    cudaStream_t st = h->stream;
    CK(h->userval.alloc(nnz));
    if (nnz) CK(cudaMemcpyAsync(h->userval.p, nz, nnz * sizeof(double), cudaMemcpyHostToDevice, st));
    const int64_t nnzM = h->B->Mp[dim];
    csc_gather_kernel<<<(unsigned)((nnzM + 255) / 256), 256, 0, st>>>(h->userval.p, h->B->d_src.p, h->Mval.p, nnzM);
    count_launch();
    h->ready = opb_handle::SYSTEM_FORMED;
    int ok = 0;
    rc = single_factor(h, 0.0, mode, &ok); if (rc) return rc;
    if (mode == OPB_MODE_LDLT && ok) {
        const DeltaState& s = h->h_state;
        // inertia_status (linear_system_solvers.jl:48-91) on the classified pivots
        ok = (s.n_bad == 0 && s.n_pos == n_pos && s.n_neg == m_neg) ? 1 : 0;
    }
    if (inertia_ok) *inertia_ok = ok;
    return OPB_OK;
}

int opb_ls_solve(opb_handle* h, const double* rhs, double* sol) {
    int rc = need_device(h); if (rc) return rc;
    if (!h->B || h->ready != opb_handle::FACTORED) return h->fail(OPB_ERR_STATE, "ls_solve before ls_factor!");
    Bundle& B = *h->B;
    const int n = B.S.n;
    cudaStream_t st = h->stream;
    CK(cudaMemcpyAsync(h->res.p, rhs, n * sizeof(double), cudaMemcpyHostToDevice, st));
    launch_permute_in(h->res.p, B.d_perm.p, h->xw.p, n, st);
    launch_solve(h->dev, B.plan, B.d_sched.p, h->Lval.p, h->Xinv.p, h->xw.p, h->xw2.p, h->uw.p, h->fmode(),
                         h->shard_ctx(), B.d_colowner.p, h->solve_overlap ? &h->side : nullptr, st);
    launch_permute_out_add(h->xw.p, B.d_perm.p, h->b.p, n, 0, st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(sol, h->b.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return OPB_OK;
}

// eval_diag_J_T_J (utils/eval.jl:89-100), the building block of compute_schur_diag
// (kkt_system_solver.jl:296-300): out[i] = sum_j J[j,i]^2 * diag_vals[j]
int opb_eval_diag_JtDJ(opb_handle* h, int64_t n, int64_t m, const int64_t* Jp, const int64_t* Ji,
                       const double* Jx, int base, const double* diag_vals, double* out) {
    int rc = need_device(h); if (rc) return rc;
    if (!Jp || !out || n <= 0 || m < 0 || (base != 0 && base != 1)) return h->fail(OPB_ERR_INVALID, "bad arguments");
    const int64_t nnz = Jp[n] - base;
    if (nnz < 0 || (nnz > 0 && (!Ji || !Jx || !diag_vals))) return h->fail(OPB_ERR_INVALID, "bad arguments");
    cudaStream_t st = h->stream;
    CK(h->scr_i0.alloc(n + 1)); CK(h->scr_i1.alloc(nnz)); CK(h->scr_d0.alloc(nnz)); CK(h->scr_d1.alloc(m)); CK(h->scr_d2.alloc(n));
    CK(cudaMemcpyAsync(h->scr_i0.p, Jp, (n + 1) * sizeof(int64_t), cudaMemcpyHostToDevice, st));
    if (nnz) {
        CK(cudaMemcpyAsync(h->scr_i1.p, Ji, nnz * sizeof(int64_t), cudaMemcpyHostToDevice, st));
        CK(cudaMemcpyAsync(h->scr_d0.p, Jx, nnz * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    if (m) CK(cudaMemcpyAsync(h->scr_d1.p, diag_vals, m * sizeof(double), cudaMemcpyHostToDevice, st));
    launch_diag_JtDJ(h->scr_i0.p, h->scr_i1.p, h->scr_d0.p, h->scr_d1.p, h->scr_d2.p, (int)n, base, st);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, h->scr_d2.p, n * sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    return OPB_OK;
}

// ---- introspection -----------------------------------------------------------
int opb_get_info(opb_handle* h, const char* key, double* out) {
    if (!h || !key || !out) return OPB_ERR_INVALID;
    if (!h->B) return h->fail(OPB_ERR_STATE, "no structure");
    const Bundle& B = *h->B;
    const Symbolic& S = B.S;
    std::string k(key);
    if (k == "n") *out = S.n;
    else if (k == "m") *out = B.P.m;
    else if (k == "nnzJ") *out = (double)B.P.nnzJ;
    else if (k == "nnzH") *out = (double)B.P.nnzH;
    else if (k == "nnzM") *out = (double)B.Mp[S.n];
    else if (k == "npairs") *out = B.schur ? (double)B.P.pair_ptr.back() : 0.0;
    else if (k == "nnzL") *out = (double)S.nnzL;
    else if (k == "nnzL_true") *out = (double)S.nnzL_true;
    else if (k == "flops") *out = S.flops;
    else if (k == "nsuper") *out = S.nsuper;
    else if (k == "nlevels") *out = S.nlevels;
    else if (k == "max_front") *out = S.max_front;
    else if (k == "cb_total") *out = (double)S.cb_total;
    else if (k == "n_tiny") *out = B.n_tiny;
    else if (k == "n_small") *out = B.n_small;
    else if (k == "n_big") *out = B.n_big;
    else if (k == "symbolic_cached") *out = h->cached_hit ? 1 : 0;
    else if (k == "device_bytes") *out = (double)g_device_bytes.load();
    else if (k == "loop_graph_active") { *out = h->g_loop.exec ? 1 : 0; h->err = "loop graph: " + h->loop_diag; }
    else if (k == "occ_small_tiles") *out = g_occ_small_tiles;
    else if (k == "sum_rows") *out = (double)S.rowidx.size();
    else if (k == "x_total") *out = (double)B.x_total;
    else if (k == "n_trtri") { int c = 0; for (const TrtriPlan& T : B.trtri) c += T.count; *out = c; }
    else if (k == "shard_rank") *out = B.rank;
    else if (k == "shard_world") *out = B.world;
    else if (k == "shard_load") *out = B.world > 1 ? B.shard.load[B.rank] : S.flops;
    else if (k == "shard_top_flops") *out = B.world > 1 ? B.shard.top_flops : 0.0;
    else if (k == "shard_barriers") { int c = 0; for (const LevelPlan& L : B.plan) if (L.barrier_mask) c += L.barrier_before + 2 * L.split; *out = c; }
    else if (k == "shard_split") { int c = 0; if (B.world > 1) for (char f : B.shard.split) c += f; *out = c; }
    else if (k == "shard_mirrored") { int c = 0; for (const LevelPlan& L : B.plan) c += L.pullcb_count; *out = c; }
    else if (k == "shard_helped") { int c = 0; for (const LevelPlan& L : B.plan) c += L.help_count; *out = c; }
    else if (k.rfind("t_", 0) == 0) {
        // host seconds of the phases of the last analysis of this structure (0 when the phase did not run)
        *out = 0;
        for (const auto& kv : B.timing) if (k == "t_" + kv.first) *out += kv.second;
        for (const auto& kv : S.timing) if (k == "t_" + kv.first) *out += kv.second;
    }
    else return h->fail(OPB_ERR_INVALID, "unknown info key " + k);
    return OPB_OK;
}

int64_t opb_get_symbolic(opb_handle* h, const char* name, int64_t* out, int64_t cap) {
    if (!h || !name) return OPB_ERR_INVALID;
    if (!h->B) return h->fail(OPB_ERR_STATE, "no structure");
    const Bundle& B = *h->B;
    const Symbolic& S = B.S;
    std::string k(name);
    if (k == "perm") return copy_out(S.perm, out, cap);
    if (k == "sfirst") return copy_out(S.sfirst, out, cap);
    if (k == "sparent") return copy_out(S.sparent, out, cap);
    if (k == "rowptr") return copy_out(S.rowptr, out, cap);
    if (k == "rowidx") return copy_out(S.rowidx, out, cap);
    if (k == "rel") return copy_out(S.rel, out, cap);
    if (k == "Loff") return copy_out(S.Loff, out, cap);
    if (k == "CBoff") return copy_out(S.CBoff, out, cap);
    if (k == "level") return copy_out(S.level, out, cap);
    if (k == "amap") return copy_out(S.amap, out, cap);
    if (k == "dpos") return copy_out(S.dpos, out, cap);
    if (k == "Mp") return copy_out(B.Mp, out, cap);
    if (k == "Mi") return copy_out(B.Mi, out, cap);
    if (k == "src") return copy_out(B.src, out, cap);
    if (k == "pair_ptr") return copy_out(B.P.pair_ptr, out, cap);
    if (k == "pairA") return copy_out(B.P.pairA, out, cap);
    if (k == "pairB") return copy_out(B.P.pairB, out, cap);
    if (k == "hmap") return copy_out(B.P.hmap, out, cap);
    if (k == "gptr") return copy_out(S.gptr, out, cap);
    if (k == "gsrc") return copy_out(S.gsrc, out, cap);
    if (k == "gch") return copy_out(S.gch, out, cap);
    if (k == "tcut_ptr") return copy_out(S.tcut_ptr, out, cap);
    if (k == "tcut") return copy_out(S.tcut, out, cap);
    if (k == "owner") { if (B.world > 1) return copy_out(B.shard.owner, out, cap); return copy_out(std::vector<int>(S.nsuper, 0), out, cap); }
    if (k == "level_mask") {      // per level: the ranks THIS rank synchronises with (bit mask, 0 = none)
        std::vector<int> t(S.nlevels, 0);
        if (B.world > 1) for (int l = 0; l < S.nlevels; l++) t[l] = (int)B.shard.level_mask[(size_t)l * B.world + B.rank];
        return copy_out(t, out, cap);
    }
    if (k == "split") { std::vector<int> t(S.nsuper, 0); if (B.world > 1) for (int q = 0; q < S.nsuper; q++) t[q] = B.shard.split[q]; return copy_out(t, out, cap); }
    if (k == "range_a") { if (B.world > 1) return copy_out(B.shard.ra, out, cap); return copy_out(std::vector<int>(S.nsuper, 0), out, cap); }
    if (k == "range_b") { if (B.world > 1) return copy_out(B.shard.rb, out, cap); return copy_out(std::vector<int>(S.nsuper, 1), out, cap); }
    if (k == "top") { std::vector<int> t(S.nsuper, 0); if (B.world > 1) for (int q = 0; q < S.nsuper; q++) t[q] = B.shard.top[q]; return copy_out(t, out, cap); }
    return h->fail(OPB_ERR_INVALID, "unknown symbolic array " + k);
}

int opb_get_L_values(opb_handle* h, double* out, int64_t cap) {
    int rc = need_device(h); if (rc) return rc;
    if (!h->B || h->ready != opb_handle::FACTORED) return h->fail(OPB_ERR_STATE, "no factor");
    int64_t cnt = std::min<int64_t>(cap, h->B->S.nnzL);
    CK(cudaMemcpyAsync(out, h->Lval.p, cnt * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return OPB_OK;
}

}  // extern "C"
