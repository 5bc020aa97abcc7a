// Symbolic analysis (host, once per sparsity pattern).  See symbolic.h.
//
// Stages:
//   1. build_schur_pattern : pattern of tril(J'DJ + H) + full diagonal, and the
//      structure-fixed gather map used by the assembly kernel (replaces the two
//      generic sparse*sparse products of eval.jl:85-87 / schur.jl:55).
//   2. ordering            : nested dissection by BFS level structures with
//      minimum-degree leaves (fill-reducing; CHOLMOD uses AMD/METIS here).
//   3. etree, postorder, column counts (skeleton algorithm), supernodes with
//      relaxed amalgamation, supernodal row structures.
//   4. maps: M_L entry -> L panel slot, child update block -> parent front.
#include "symbolic.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <system_error>
#include <thread>
#include <map>
#include <mutex>
#include <cstring>
#include <cstdio>
#include <cstdlib>
#include <numeric>

// METIS nested dissection from the static library that ships with the CUDA toolkit
// (/usr/local/cuda/lib64/libmetis_static.a, the one cuSOLVER's csrmetisnd uses; idx_t is 64-bit).
// Host-side symbolic analysis only.
extern "C" {
int METIS_NodeND(int64_t* nvtxs, int64_t* xadj, int64_t* adjncy, int64_t* vwgt, int64_t* options,
                 int64_t* perm, int64_t* iperm);
int METIS_SetDefaultOptions(int64_t* options);
}

namespace opb {

uint64_t pattern_hash(int64_t n, const int64_t* p, const int64_t* i, int64_t nnz) {
    uint64_t h = 1469598103934665603ull ^ (uint64_t)n;
    auto mix = [&](uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); };
    for (int64_t k = 0; k <= n; k++) mix((uint64_t)p[k]);
    for (int64_t k = 0; k < nnz; k++) mix((uint64_t)i[k]);
    return h;
}

// ---------------------------------------------------------------------------
// 1. Schur pattern + gather map

// ---------------------------------------------------------------------------
// Host threads for the one-off analysis.  run_chunks(n, fn) calls fn(chunk, begin, end) for a fixed
// partition of [0, n) -- fixed by n and the thread count only, and every use below writes disjoint
// outputs per index, so results do not depend on the schedule.  Falls back to the calling thread
// when no thread can be started.
// ---------------------------------------------------------------------------
namespace {

int host_threads() {
    static const int t = [] {
        unsigned hc = std::thread::hardware_concurrency();
        if (const char* e = getenv("OPB_HOST_THREADS")) { int v = atoi(e); if (v > 0) hc = (unsigned)v; }
        return (int)std::min(16u, std::max(1u, hc));
    }();
    return t;
}

template <class F>
void run_chunks(int64_t n, int64_t min_per_chunk, F&& fn) {
    int nt = (int)std::min<int64_t>(host_threads(), std::max<int64_t>(1, n / std::max<int64_t>(1, min_per_chunk)));
    if (nt <= 1) { fn(0, (int64_t)0, n); return; }
    std::vector<std::thread> th;
    th.reserve(nt - 1);
    auto bound = [&](int c) { return n * c / nt; };
    int started = 1;
    for (int c = 1; c < nt; c++) {
        try { th.emplace_back([&fn, c, &bound] { fn(c, bound(c), bound(c + 1)); }); started++; }
        catch (const std::system_error&) { break; }
    }
    fn(0, bound(0), bound(1));
    for (int c = started; c < nt; c++) fn(c, bound(c), bound(c + 1));   // chunks that got no thread
    for (auto& t : th) t.join();
}

// chunk boundaries that balance a prefix-sum weight (ptr[n] total) instead of the index count
template <class F>
void run_chunks_weighted(int64_t n, const int64_t* ptr, int64_t min_weight, F&& fn) {
    const int64_t total = ptr[n] - ptr[0];
    int nt = (int)std::min<int64_t>(host_threads(), std::max<int64_t>(1, total / std::max<int64_t>(1, min_weight)));
    if (nt <= 1) { fn(0, (int64_t)0, n); return; }
    std::vector<int64_t> b(nt + 1, n);
    b[0] = 0;
    for (int c = 1; c < nt; c++)
        b[c] = std::lower_bound(ptr, ptr + n + 1, ptr[0] + total * c / nt) - ptr;
    for (int c = 1; c <= nt; c++) b[c] = std::max(b[c], b[c - 1]);
    b[nt] = n;
    std::vector<std::thread> th;
    int started = 1;
    for (int c = 1; c < nt; c++) {
        try { th.emplace_back([&fn, c, &b] { fn(c, b[c], b[c + 1]); }); started++; }
        catch (const std::system_error&) { break; }
    }
    fn(0, b[0], b[1]);
    for (int c = started; c < nt; c++) fn(c, b[c], b[c + 1]);
    for (auto& t : th) t.join();
}

}  // namespace

// ---------------------------------------------------------------------------
bool build_schur_pattern(int64_t n64, int64_t m64, const int64_t* Jp_in, const int64_t* Ji_in,
                         const int64_t* Hp_in, const int64_t* Hi_in, int base,
                         SchurPattern& P, std::string& err) {
    if (n64 <= 0 || n64 > 2000000000ll || m64 < 0 || m64 > 2000000000ll) { err = "bad dimensions"; return false; }
    const int n = (int)n64, m = (int)m64;
    P.n = n; P.m = m;
    const int64_t nnzJ = Jp_in[n] - base, nnzH = Hp_in[n] - base;
    if (nnzJ < 0 || nnzH < 0 || nnzJ > 2000000000ll || nnzH > 2000000000ll) { err = "bad colptr"; return false; }
    P.nnzJ = nnzJ; P.nnzH = nnzH;
    P.Jp.resize(n + 1);
    for (int j = 0; j <= n; j++) P.Jp[j] = Jp_in[j] - base;
    P.Jrow.resize(nnzJ);
    for (int j = 0; j < n; j++) {
        if (P.Jp[j + 1] < P.Jp[j]) { err = "J colptr not monotone"; return false; }
        for (int64_t p = P.Jp[j]; p < P.Jp[j + 1]; p++) {
            int64_t k = Ji_in[p] - base;
            if (k < 0 || k >= m) { err = "J row index out of range"; return false; }
            if (p > P.Jp[j] && k <= P.Jrow[p - 1]) { err = "J row indices not sorted/unique"; return false; }
            P.Jrow[p] = (int)k;
        }
    }
    // CSR view of J (columns ascending within each row)
    P.Rp.assign(m + 1, 0);
    for (int64_t p = 0; p < nnzJ; p++) P.Rp[P.Jrow[p] + 1]++;
    for (int k = 0; k < m; k++) P.Rp[k + 1] += P.Rp[k];
    P.Rcol.resize(nnzJ); P.Rpos.resize(nnzJ);
    {
        std::vector<int64_t> nxt(P.Rp.begin(), P.Rp.end() - 1);
        for (int j = 0; j < n; j++)
            for (int64_t p = P.Jp[j]; p < P.Jp[j + 1]; p++) {
                int64_t q = nxt[P.Jrow[p]]++;
                P.Rcol[q] = j; P.Rpos[q] = (int)p;
            }
    }
    // H (lower) checks + symmetric CSR view
    std::vector<int64_t> Hp(n + 1);
    for (int j = 0; j <= n; j++) Hp[j] = Hp_in[j] - base;
    std::vector<int> Hrow(nnzH);
    P.Sp.assign(n + 1, 0);
    for (int j = 0; j < n; j++)
        for (int64_t p = Hp[j]; p < Hp[j + 1]; p++) {
            int64_t i = Hi_in[p] - base;
            if (i < j || i >= n) { err = "H must be lower triangular with in-range rows"; return false; }
            if (p > Hp[j] && i <= Hrow[p - 1]) { err = "H row indices not sorted/unique"; return false; }
            Hrow[p] = (int)i;
            P.Sp[i + 1]++;
            if (i != j) P.Sp[j + 1]++;
        }
    for (int i = 0; i < n; i++) P.Sp[i + 1] += P.Sp[i];
    P.Scol.resize(P.Sp[n]); P.Spos.resize(P.Sp[n]);
    {
        // row i of H_sym: first the lower part (cols j <= i, ascending j), then the
        // mirrored part (cols > i).  Two sweeps keep columns ascending in each row.
        std::vector<int64_t> nxt(P.Sp.begin(), P.Sp.end() - 1);
        for (int j = 0; j < n; j++)
            for (int64_t p = Hp[j]; p < Hp[j + 1]; p++) {  // entry (i,j), i >= j: row i gets col j
                int i = Hrow[p];
                int64_t q = nxt[i]++;
                P.Scol[q] = j; P.Spos[q] = (int)p;
            }
        for (int j = 0; j < n; j++)
            for (int64_t p = Hp[j]; p < Hp[j + 1]; p++) {  // mirrored: row j gets col i (i > j)
                int i = Hrow[p];
                if (i == j) continue;
                int64_t q = nxt[j]++;
                P.Scol[q] = i; P.Spos[q] = (int)p;
            }
    }
    // pattern of tril(J'DJ) U H U diag, column by column (column ranges on host threads, each with
    // its own marker array; the pieces are concatenated in column order)
    P.Mp.assign(n + 1, 0);
    P.Mi.clear();
    {
        const int NTMAX = 16;
        std::vector<int> piece[NTMAX];
        int64_t piece_begin[NTMAX + 1];
        for (int c = 0; c <= NTMAX; c++) piece_begin[c] = -1;
        int nchunks = 0;
        std::mutex mu;
        run_chunks(n, 20000, [&](int c, int64_t jb, int64_t je) {
            std::vector<int> mark(n, -1), list;
            std::vector<int>& out = piece[c];
            for (int j = (int)jb; j < (int)je; j++) {
                list.clear();
                mark[j] = j; list.push_back(j);
                for (int64_t p = P.Jp[j]; p < P.Jp[j + 1]; p++) {
                    int k = P.Jrow[p];
                    for (int64_t q = P.Rp[k + 1] - 1; q >= P.Rp[k]; q--) {
                        int i = P.Rcol[q];
                        if (i < j) break;
                        if (mark[i] != j) { mark[i] = j; list.push_back(i); }
                    }
                }
                for (int64_t p = Hp[j]; p < Hp[j + 1]; p++) {
                    int i = Hrow[p];
                    if (mark[i] != j) { mark[i] = j; list.push_back(i); }
                }
                std::sort(list.begin(), list.end());
                out.insert(out.end(), list.begin(), list.end());
                P.Mp[j + 1] = (int64_t)list.size();           // counts; prefix sum below
            }
            std::lock_guard<std::mutex> g(mu);
            piece_begin[c] = jb;
            nchunks = std::max(nchunks, c + 1);
        });
        for (int j = 0; j < n; j++) P.Mp[j + 1] += P.Mp[j];
        if (P.Mp[n] > 2000000000ll) { err = "Schur complement too large for int32 indexing"; return false; }
        P.Mi.resize(P.Mp[n]);
        run_chunks(nchunks, 1, [&](int, int64_t cb, int64_t ce) {
            for (int64_t c = cb; c < ce; c++)
                if (!piece[c].empty()) memcpy(P.Mi.data() + P.Mp[piece_begin[c]], piece[c].data(), piece[c].size() * sizeof(int));
        });
    }
    const int64_t nnzM = P.Mp[n];
    // pass 2: count pairs, hmap (a column only touches its own entries)
    P.pair_ptr.assign(nnzM + 1, 0);
    P.hmap.assign(nnzM, -1);
    run_chunks_weighted(n, P.Mp.data(), 200000, [&](int, int64_t jb, int64_t je) {
        std::vector<int> where(n, -1);
        for (int j = (int)jb; j < (int)je; j++) {
            for (int64_t e = P.Mp[j]; e < P.Mp[j + 1]; e++) where[P.Mi[e]] = (int)(e - P.Mp[j]);
            const int64_t e0 = P.Mp[j];
            for (int64_t p = P.Jp[j]; p < P.Jp[j + 1]; p++) {
                int k = P.Jrow[p];
                for (int64_t q = P.Rp[k + 1] - 1; q >= P.Rp[k]; q--) {
                    int i = P.Rcol[q];
                    if (i < j) break;
                    P.pair_ptr[e0 + where[i] + 1]++;
                }
            }
            for (int64_t p = Hp[j]; p < Hp[j + 1]; p++) P.hmap[e0 + where[Hrow[p]]] = (int)p;
        }
    });
    for (int64_t e = 0; e < nnzM; e++) P.pair_ptr[e + 1] += P.pair_ptr[e];
    const int64_t npairs = P.pair_ptr[nnzM];
    if (npairs > 2000000000ll) { err = "too many J'DJ products for int32 indexing"; return false; }
    P.pairA.resize(npairs); P.pairB.resize(npairs);
    // pass 3: fill, k ascending inside each entry (CSC rows of J are ascending)
    run_chunks_weighted(n, P.Mp.data(), 200000, [&](int, int64_t jb, int64_t je) {
        std::vector<int> where(n, -1);
        std::vector<int64_t> nxt;
        for (int j = (int)jb; j < (int)je; j++) {
            const int64_t e0 = P.Mp[j], ne = P.Mp[j + 1] - e0;
            nxt.assign(P.pair_ptr.begin() + e0, P.pair_ptr.begin() + e0 + ne);
            for (int64_t e = e0; e < e0 + ne; e++) where[P.Mi[e]] = (int)(e - e0);
            for (int64_t p = P.Jp[j]; p < P.Jp[j + 1]; p++) {
                int k = P.Jrow[p];
                for (int64_t q = P.Rp[k + 1] - 1; q >= P.Rp[k]; q--) {
                    int i = P.Rcol[q];
                    if (i < j) break;
                    int64_t t = nxt[where[i]]++;
                    P.pairA[t] = P.Rpos[q];   // J[k,i]  (scaled by sigma first)
                    P.pairB[t] = (int)p;      // J[k,j]
                }
            }
        }
    });
    return true;
}

bool build_csc_pattern(int64_t n64, const int64_t* Ap, const int64_t* Ai, int base,
                       std::vector<int64_t>& Mp, std::vector<int>& Mi, std::vector<int64_t>& src,
                       std::string& err) {
    if (n64 <= 0 || n64 > 2000000000ll) { err = "bad dimension"; return false; }
    const int n = (int)n64;
    Mp.assign(n + 1, 0); Mi.clear(); src.clear();
    std::vector<std::pair<int, int64_t>> col;
    for (int j = 0; j < n; j++) {
        col.clear();
        bool have_diag = false;
        for (int64_t p = Ap[j] - base; p < Ap[j + 1] - base; p++) {
            int64_t i = Ai[p] - base;
            if (i < 0 || i >= n) { err = "row index out of range"; return false; }
            if (i < j) continue;  // upper triangle is never read (test/linear_system_solvers.jl:74-84)
            if (i == j) have_diag = true;
            col.emplace_back((int)i, p);
        }
        if (!have_diag) col.emplace_back(j, (int64_t)-1);
        std::sort(col.begin(), col.end());
        for (size_t t = 0; t < col.size(); t++) {
            if (t && col[t].first == col[t - 1].first) { err = "duplicate entry"; return false; }
            Mi.push_back(col[t].first); src.push_back(col[t].second);
        }
        Mp[j + 1] = (int64_t)Mi.size();
    }
    return true;
}

// ---------------------------------------------------------------------------
// 2. Ordering
// ---------------------------------------------------------------------------
namespace {

struct Graph {
    int n;
    std::vector<int64_t> xadj;
    std::vector<int> adj;
};

void build_graph(int n, const std::vector<int64_t>& Mp, const std::vector<int>& Mi, Graph& G) {
    G.n = n;
    G.xadj.assign(n + 1, 0);
    for (int j = 0; j < n; j++)
        for (int64_t p = Mp[j]; p < Mp[j + 1]; p++) {
            int i = Mi[p];
            if (i == j) continue;
            G.xadj[i + 1]++; G.xadj[j + 1]++;
        }
    for (int i = 0; i < n; i++) G.xadj[i + 1] += G.xadj[i];
    G.adj.resize(G.xadj[n]);
    std::vector<int64_t> nxt(G.xadj.begin(), G.xadj.end() - 1);
    for (int j = 0; j < n; j++)
        for (int64_t p = Mp[j]; p < Mp[j + 1]; p++) {
            int i = Mi[p];
            if (i == j) continue;
            G.adj[nxt[i]++] = j; G.adj[nxt[j]++] = i;
        }
}

// Exact minimum degree on a small vertex set using bitset adjacency (elimination
// graph model).  verts: global ids; reg[v]==rid marks membership.
void md_small(const Graph& G, const std::vector<int>& verts, const int* reg, int rid,
              std::vector<int>& local, int* out) {
    const int k = (int)verts.size();
    if (k <= 2) { for (int t = 0; t < k; t++) out[t] = verts[t]; return; }
    const int W = (k + 63) / 64;
    for (int t = 0; t < k; t++) local[verts[t]] = t;
    std::vector<uint64_t> A((size_t)k * W, 0), alive(W, 0);
    for (int t = 0; t < k; t++) {
        alive[t >> 6] |= 1ull << (t & 63);
        int v = verts[t];
        for (int64_t p = G.xadj[v]; p < G.xadj[v + 1]; p++) {
            int u = G.adj[p];
            if (__atomic_load_n(&reg[u], __ATOMIC_RELAXED) != rid) continue;
            int lu = local[u];
            A[(size_t)t * W + (lu >> 6)] |= 1ull << (lu & 63);
        }
    }
    std::vector<int> deg(k);
    for (int t = 0; t < k; t++) {
        int d = 0;
        for (int w = 0; w < W; w++) d += __builtin_popcountll(A[(size_t)t * W + w]);
        deg[t] = d;
    }
    std::vector<char> dead(k, 0);
    for (int step = 0; step < k; step++) {
        int best = -1, bd = 1 << 30;
        for (int t = 0; t < k; t++)
            if (!dead[t] && deg[t] < bd) { bd = deg[t]; best = t; }
        out[step] = verts[best];
        dead[best] = 1;
        alive[best >> 6] &= ~(1ull << (best & 63));
        uint64_t* Ab = &A[(size_t)best * W];
        for (int w = 0; w < W; w++) Ab[w] &= alive[w];
        for (int w = 0; w < W; w++) {
            uint64_t bits = Ab[w];
            while (bits) {
                int u = (w << 6) + __builtin_ctzll(bits);
                bits &= bits - 1;
                uint64_t* Au = &A[(size_t)u * W];
                int d = 0;
                for (int x = 0; x < W; x++) { Au[x] = (Au[x] | Ab[x]) & alive[x]; }
                Au[u >> 6] &= ~(1ull << (u & 63));
                for (int x = 0; x < W; x++) d += __builtin_popcountll(Au[x]);
                deg[u] = d;
            }
        }
    }
}

// The two halves of a dissection are independent, so the upper levels of the recursion run on
// several host threads.  Per-vertex arrays (reg, lvl, local) are shared: a task only writes the
// entries of its own vertices; it may READ reg[] of a neighbour that a sibling task is relabelling,
// but only to compare it with its own (unique) region id, so any value it sees gives the same
// answer -- those accesses are relaxed atomics.  The result does not depend on the schedule.
struct NDWork {
    const Graph* G;
    int* reg;                  // region id of each vertex (shared)
    int* lvl;                  // BFS level scratch (shared, own vertices only)
    std::vector<int>* local_v; // md_small scratch, indexed by vertex (shared, own vertices only)
    std::atomic<int>* next_rid;
    std::atomic<int>* live;    // host threads of this ordering that are running (forking stops at the cap)
    std::vector<int> queue;    // per task
    int leaf;
    double balance;            // a separator level must leave at least this fraction on either side
    int* perm;                 // output, new -> old (disjoint ranges per task)
};
inline int reg_of(const NDWork& W, int v) { return __atomic_load_n(&W.reg[v], __ATOMIC_RELAXED); }
inline void set_reg(NDWork& W, int v, int r) { __atomic_store_n(&W.reg[v], r, __ATOMIC_RELAXED); }

// BFS inside region rid from root; fills W.queue (order) and W.lvl; returns #levels.
int bfs(NDWork& W, int root, int rid, std::vector<int>& lptr) {
    const Graph& G = *W.G;
    W.queue.clear(); lptr.clear();
    W.queue.push_back(root); W.lvl[root] = 0;
    lptr.push_back(0);
    size_t head = 0;
    int cur = 0;
    // lvl is reset by the caller via stamp trick: we use lvl = -1 for unvisited in region
    while (head < W.queue.size()) {
        int v = W.queue[head];
        if (W.lvl[v] != cur) { cur = W.lvl[v]; lptr.push_back((int)head); }
        head++;
        for (int64_t p = G.xadj[v]; p < G.xadj[v + 1]; p++) {
            int u = G.adj[p];
            if (reg_of(W, u) != rid || W.lvl[u] >= 0) continue;
            W.lvl[u] = cur + 1;
            W.queue.push_back(u);
        }
    }
    lptr.push_back((int)W.queue.size());
    return (int)lptr.size() - 1;
}

void order_fallback(NDWork& W, std::vector<int>& verts, int rid, int offset) {
    if ((int)verts.size() <= 2048) {
        md_small(*W.G, verts, W.reg, rid, *W.local_v, W.perm + offset);
    } else {
        std::sort(verts.begin(), verts.end());
        for (size_t t = 0; t < verts.size(); t++) W.perm[offset + t] = verts[t];
    }
}

// A dissection forks a host thread for one half while fewer than 2 x host_threads() are running (the
// halves are unbalanced -- 30/70 is common -- so a fixed fork depth leaves one long serial task).
constexpr int ND_PAR_MIN = 20000;      // ... when both halves have at least this many vertices

void nd_rec(NDWork& W, std::vector<int>& verts, int offset, int depth) {
    const Graph& G = *W.G;
    const int k = (int)verts.size();
    if (k == 0) return;
    const int rid = W.next_rid->fetch_add(1);
    for (int v : verts) { set_reg(W, v, rid); W.lvl[v] = -1; }
    if (k <= W.leaf || depth > 200) { order_fallback(W, verts, rid, offset); return; }
    // connected components
    std::vector<int> lptr;
    bfs(W, verts[0], rid, lptr);
    if ((int)W.queue.size() < k) {
        // split into components, recurse on each
        std::vector<std::vector<int>> comps;
        comps.emplace_back(W.queue.begin(), W.queue.end());
        for (int v : verts) {
            if (W.lvl[v] >= 0) continue;
            bfs(W, v, rid, lptr);
            comps.emplace_back(W.queue.begin(), W.queue.end());
        }
        std::vector<int>().swap(verts);
        // group tiny components together to avoid deep recursion on dust
        std::vector<int> dust;
        int off = offset;
        for (auto& c : comps) {
            if ((int)c.size() <= W.leaf / 4 + 1) { dust.insert(dust.end(), c.begin(), c.end()); continue; }
            int sz = (int)c.size();
            nd_rec(W, c, off, depth + 1);
            off += sz;
        }
        if (!dust.empty()) {
            // dust: components are independent; order each by md in chunks
            size_t pos = 0;
            while (pos < dust.size()) {
                size_t end = std::min(dust.size(), pos + (size_t)std::max(W.leaf, 64));
                std::vector<int> chunk(dust.begin() + pos, dust.begin() + end);
                const int r2 = W.next_rid->fetch_add(1);
                for (int v : chunk) set_reg(W, v, r2);
                md_small(G, chunk, W.reg, r2, *W.local_v, W.perm + off);
                off += (int)chunk.size();
                pos = end;
            }
        }
        return;
    }
    // pseudo-peripheral root: repeat BFS from a min-degree vertex of the last level
    int root = verts[0];
    int nlev = (int)lptr.size() - 1;
    for (int it = 0; it < 3; it++) {
        int last0 = lptr[nlev - 1], last1 = lptr[nlev];
        int cand = W.queue[last0];
        int64_t cd = G.xadj[cand + 1] - G.xadj[cand];
        for (int t = last0; t < last1; t++) {
            int v = W.queue[t];
            int64_t d = G.xadj[v + 1] - G.xadj[v];
            if (d < cd) { cd = d; cand = v; }
        }
        for (int v : verts) W.lvl[v] = -1;
        std::vector<int> lptr2;
        int root_prev = root;
        root = cand;
        int nlev2 = bfs(W, root, rid, lptr2);
        bool better = nlev2 > nlev;
        lptr.swap(lptr2); nlev = nlev2;
        (void)root_prev;
        if (!better) break;
    }
    if (nlev < 3) { order_fallback(W, verts, rid, offset); return; }
    // choose separator level: smallest level whose split is balanced
    int best = -1; int64_t bestsz = INT64_MAX;
    for (int l = 1; l + 1 < nlev; l++) {
        int left = lptr[l], sz = lptr[l + 1] - lptr[l], right = k - lptr[l + 1];
        if (left < W.balance * k || right < W.balance * k) continue;
        if (sz < bestsz) { bestsz = sz; best = l; }
    }
    if (best < 0) {
        // median level
        for (int l = 1; l + 1 < nlev; l++) if (lptr[l + 1] >= k / 2) { best = l; break; }
        if (best < 0) best = nlev / 2;
        if (best < 1) best = 1;
        if (best > nlev - 2) best = nlev - 2;
    }
    int sz = lptr[best + 1] - lptr[best];
    if (sz > 0.6 * k) { order_fallback(W, verts, rid, offset); return; }
    std::vector<int> left(W.queue.begin(), W.queue.begin() + lptr[best]);
    std::vector<int> right(W.queue.begin() + lptr[best + 1], W.queue.end());
    std::vector<int> sep;
    // thin the separator: level-`best` vertices with no neighbour in level best+1 go left
    for (int t = lptr[best]; t < lptr[best + 1]; t++) {
        int v = W.queue[t];
        bool touches = false;
        for (int64_t p = G.xadj[v]; p < G.xadj[v + 1] && !touches; p++) {
            int u = G.adj[p];
            if (reg_of(W, u) == rid && W.lvl[u] == best + 1) touches = true;
        }
        if (touches) sep.push_back(v); else left.push_back(v);
    }
    std::vector<int>().swap(verts);
    const int nl = (int)left.size(), nr = (int)right.size();
    // separator last
    std::sort(sep.begin(), sep.end());
    for (size_t t = 0; t < sep.size(); t++) { W.perm[offset + nl + nr + t] = sep[t]; set_reg(W, sep[t], 0); }
    bool fork = nl >= ND_PAR_MIN && nr >= ND_PAR_MIN;
    if (fork && W.live->fetch_add(1) >= 2 * host_threads()) { W.live->fetch_sub(1); fork = false; }
    if (fork) {
        NDWork W2 = W;                 // shares the per-vertex arrays, own queue
        W2.queue.clear();
        std::thread th;
        bool forked = true;
        try {
            th = std::thread([&W2, &left, offset, depth] { nd_rec(W2, left, offset, depth + 1); W2.live->fetch_sub(1); });
        } catch (const std::system_error&) {
            forked = false;            // no thread to be had: same work, serially
            W.live->fetch_sub(1);
        }
        if (!forked) nd_rec(W, left, offset, depth + 1);
        nd_rec(W, right, offset + nl, depth + 1);
        if (forked) th.join();
    } else {
        nd_rec(W, left, offset, depth + 1);
        nd_rec(W, right, offset + nl, depth + 1);
    }
}

void nd_order(const Graph& G, int leaf, double balance, std::vector<int>& perm) {
    std::vector<int> reg(G.n, 0), lvl(G.n, -1), local(G.n, 0);
    std::atomic<int> next_rid{1}, live{1};
    NDWork W;
    W.G = &G;
    W.reg = reg.data(); W.lvl = lvl.data(); W.local_v = &local; W.next_rid = &next_rid; W.live = &live;
    W.leaf = std::max(leaf, 4);
    W.balance = std::min(0.45, std::max(0.05, balance));
    perm.resize(G.n);
    W.perm = perm.data();
    std::vector<int> all(G.n);
    std::iota(all.begin(), all.end(), 0);
    nd_rec(W, all, 0, 0);
}

}  // namespace

// ---------------------------------------------------------------------------
// 3+4. etree, postorder, column counts, supernodes, structures, maps
// ---------------------------------------------------------------------------
namespace {

// lower CSC of the permuted matrix (rows unsorted), strictly-lower entries only
void permuted_lower(int n, const std::vector<int64_t>& Mp, const std::vector<int>& Mi,
                    const std::vector<int>& iperm, std::vector<int64_t>& Bp, std::vector<int>& Bi) {
    Bp.assign(n + 1, 0);
    for (int j = 0; j < n; j++)
        for (int64_t p = Mp[j]; p < Mp[j + 1]; p++) {
            int i = Mi[p];
            if (i == j) continue;
            int a = iperm[i], b = iperm[j];
            Bp[std::min(a, b) + 1]++;
        }
    for (int j = 0; j < n; j++) Bp[j + 1] += Bp[j];
    Bi.resize(Bp[n]);
    std::vector<int64_t> nxt(Bp.begin(), Bp.end() - 1);
    for (int j = 0; j < n; j++)
        for (int64_t p = Mp[j]; p < Mp[j + 1]; p++) {
            int i = Mi[p];
            if (i == j) continue;
            int a = iperm[i], b = iperm[j];
            Bi[nxt[std::min(a, b)]++] = std::max(a, b);
        }
}

// transpose of strictly-lower CSC = for each k the list of i < k with B[k,i] != 0
void transpose_lower(int n, const std::vector<int64_t>& Bp, const std::vector<int>& Bi,
                     std::vector<int64_t>& Up, std::vector<int>& Ui) {
    Up.assign(n + 1, 0);
    for (int64_t p = 0; p < Bp[n]; p++) Up[Bi[p] + 1]++;
    for (int j = 0; j < n; j++) Up[j + 1] += Up[j];
    Ui.resize(Up[n]);
    std::vector<int64_t> nxt(Up.begin(), Up.end() - 1);
    for (int j = 0; j < n; j++)
        for (int64_t p = Bp[j]; p < Bp[j + 1]; p++) Ui[nxt[Bi[p]]++] = j;
}

void etree(int n, const std::vector<int64_t>& Up, const std::vector<int>& Ui, std::vector<int>& parent) {
    parent.assign(n, -1);
    std::vector<int> anc(n, -1);
    for (int k = 0; k < n; k++)
        for (int64_t p = Up[k]; p < Up[k + 1]; p++) {
            int i = Ui[p];
            while (i != -1 && i < k) {
                int nx = anc[i];
                anc[i] = k;
                if (nx == -1) parent[i] = k;
                i = nx;
            }
        }
}

// postorder with children sorted so that the heaviest subtree comes last
void postorder(int n, const std::vector<int>& parent, std::vector<int>& post) {
    std::vector<int> size(n, 1);
    for (int j = 0; j < n; j++) if (parent[j] >= 0) size[parent[j]] += size[j];  // parent[j] > j
    std::vector<int> cptr(n + 2, 0), clist(n);
    for (int j = 0; j < n; j++) cptr[(parent[j] < 0 ? n : parent[j]) + 1]++;
    for (int j = 0; j <= n; j++) cptr[j + 1] += cptr[j];
    {
        std::vector<int> nxt(cptr.begin(), cptr.end() - 1);
        for (int j = 0; j < n; j++) clist[nxt[parent[j] < 0 ? n : parent[j]]++] = j;
    }
    for (int v = 0; v <= n; v++)
        std::stable_sort(clist.begin() + cptr[v], clist.begin() + cptr[v + 1],
                         [&](int a, int b) { return size[a] < size[b]; });
    post.clear(); post.reserve(n);
    // iterative DFS; virtual root = n
    std::vector<int> stack, it(n + 1);
    for (int v = 0; v <= n; v++) it[v] = cptr[v];
    stack.push_back(n);
    while (!stack.empty()) {
        int v = stack.back();
        if (it[v] < cptr[v + 1]) { stack.push_back(clist[it[v]++]); }
        else { stack.pop_back(); if (v != n) post.push_back(v); }
    }
}

// column counts (incl. diagonal) for a postordered matrix; Bp/Bi strictly lower CSC
void colcounts(int n, const std::vector<int64_t>& Bp, const std::vector<int>& Bi,
               const std::vector<int>& parent, std::vector<int64_t>& cc) {
    std::vector<int> first(n, -1), maxfirst(n, -1), prevleaf(n, -1), anc(n);
    std::vector<int64_t> delta(n, 0);
    for (int j = 0; j < n; j++) {
        if (first[j] == -1) { first[j] = j; delta[j] = 1; }  // leaf of the etree
        else delta[j] = 0;
        int p = parent[j];
        if (p >= 0 && first[p] == -1) first[p] = first[j];
    }
    for (int j = 0; j < n; j++) anc[j] = j;
    auto find = [&](int v) {
        int r = v;
        while (anc[r] != r) r = anc[r];
        while (anc[v] != r) { int nx = anc[v]; anc[v] = r; v = nx; }
        return r;
    };
    for (int j = 0; j < n; j++) {
        if (parent[j] >= 0) delta[parent[j]]--;
        for (int64_t p = Bp[j]; p < Bp[j + 1]; p++) {
            int i = Bi[p];  // i > j, entry (i,j): j belongs to row subtree of i
            if (first[j] > maxfirst[i]) {
                // j is a leaf of the row subtree of i
                maxfirst[i] = first[j];
                int pl = prevleaf[i];
                prevleaf[i] = j;
                delta[j]++;
                if (pl != -1) { int q = find(pl); delta[q]--; }
            }
        }
        if (parent[j] >= 0) anc[j] = parent[j];
    }
    cc.assign(n, 0);
    for (int j = 0; j < n; j++) {
        cc[j] += delta[j];
        if (parent[j] >= 0) cc[parent[j]] += cc[j];
    }
}

}  // namespace

bool analyze(int n, const std::vector<int64_t>& Mp, const std::vector<int>& Mi,
             const SymOptions& opt, const int64_t* user_perm, Symbolic& S) {
    S = Symbolic();
    S.n = n;
    auto t_last = std::chrono::steady_clock::now();
    auto mark = [&](const char* name) {
        auto t = std::chrono::steady_clock::now();
        S.timing.emplace_back(name, std::chrono::duration<double>(t - t_last).count());
        t_last = t;
    };
    // ---- ordering
    std::vector<int> perm0(n);
    auto metis_order = [&](std::vector<int>& out) -> bool {
        // METIS_NodeND on the adjacency graph of M
        Graph G;
        build_graph(n, Mp, Mi, G);
        std::vector<int64_t> xadj(G.xadj.begin(), G.xadj.end()), adj(G.adj.begin(), G.adj.end());
        std::vector<int64_t> mperm(n), miperm(n), options(64);
        METIS_SetDefaultOptions(options.data());
        int64_t nv = n;
        out.resize(n);
        if (n <= 1) { if (n == 1) out[0] = 0; return true; }
        if (METIS_NodeND(&nv, xadj.data(), adj.data(), nullptr, options.data(), mperm.data(), miperm.data()) != 1) return false;
        for (int k = 0; k < n; k++) out[k] = (int)mperm[k];
        return true;
    };
    auto own_order = [&](std::vector<int>& out) {
        Graph G;
        build_graph(n, Mp, Mi, G);
        nd_order(G, opt.nd_leaf, opt.nd_balance, out);
    };
    // factorisation flops (sum of squared column counts) of a candidate ordering
    auto flops_of = [&](const std::vector<int>& pm) {
        std::vector<int> ip(n), par, post, pm2(n), ip2(n);
        for (int k = 0; k < n; k++) ip[pm[k]] = k;
        std::vector<int64_t> Bp_, Up_, cc_;
        std::vector<int> Bi_, Ui_;
        permuted_lower(n, Mp, Mi, ip, Bp_, Bi_);
        transpose_lower(n, Bp_, Bi_, Up_, Ui_);
        etree(n, Up_, Ui_, par);
        postorder(n, par, post);
        for (int k = 0; k < n; k++) pm2[k] = pm[post[k]];
        for (int k = 0; k < n; k++) ip2[pm2[k]] = k;
        permuted_lower(n, Mp, Mi, ip2, Bp_, Bi_);
        std::vector<int> par2(n);                // the postordered tree is the same tree, relabelled
        for (int k = 0; k < n; k++) ip[post[k]] = k;
        for (int k = 0; k < n; k++) { const int p0 = par[post[k]]; par2[k] = p0 < 0 ? -1 : ip[p0]; }
        colcounts(n, Bp_, Bi_, par2, cc_);
        double f = 0;
        for (int j = 0; j < n; j++) f += (double)cc_[j] * (double)cc_[j];
        return f;
    };
    if (user_perm) {
        std::vector<char> seen(n, 0);
        for (int k = 0; k < n; k++) {
            int64_t v = user_perm[k];
            if (v < 0 || v >= n || seen[v]) { S.error = "user permutation invalid"; return false; }
            seen[v] = 1; perm0[k] = (int)v;
        }
    } else if (opt.ordering == 1) {
        std::iota(perm0.begin(), perm0.end(), 0);
    } else if (opt.ordering == 3) {
        if (!metis_order(perm0)) { S.error = "METIS_NodeND failed"; return false; }
    } else if (opt.ordering == 0) {
        own_order(perm0);
    } else {
        // auto (default): the level-structure nested dissection of this file and, for graphs that
        // METIS orders in seconds, METIS_NodeND; keep the ordering with fewer factorisation flops
        // The two candidates are independent: METIS runs on the calling thread (its random-number
        // state is process-wide, see finish_structure), the own ordering beside it.
        std::vector<int> pm;
        bool metis_ok = false;
        const bool try_metis = n <= opt.metis_max_n;
        std::thread side;
        bool forked = false;
        if (try_metis) {
            try { side = std::thread([&] { own_order(perm0); }); forked = true; }
            catch (const std::system_error&) {}
        }
        if (!forked) own_order(perm0);
        if (try_metis) metis_ok = metis_order(pm);
        if (forked) side.join();
        mark(try_metis ? "order_candidates" : "order_own");
        if (metis_ok) {
            double f_metis = 0, f_own = 0;
            std::thread cmp;
            bool cf = false;
            try { cmp = std::thread([&] { f_own = flops_of(perm0); }); cf = true; }
            catch (const std::system_error&) {}
            if (!cf) f_own = flops_of(perm0);
            f_metis = flops_of(pm);
            if (cf) cmp.join();
            if (f_metis < f_own) perm0.swap(pm);
            mark("order_compare");
        }
    }
    mark("order");
    std::vector<int> iperm0(n);
    for (int k = 0; k < n; k++) iperm0[perm0[k]] = k;
    // ---- etree + postorder on the first permutation
    std::vector<int64_t> Bp, Up;
    std::vector<int> Bi, Ui, parent0, post;
    permuted_lower(n, Mp, Mi, iperm0, Bp, Bi);
    transpose_lower(n, Bp, Bi, Up, Ui);
    etree(n, Up, Ui, parent0);
    postorder(n, parent0, post);
    S.perm.resize(n); S.iperm.resize(n);
    for (int k = 0; k < n; k++) S.perm[k] = perm0[post[k]];
    for (int k = 0; k < n; k++) S.iperm[S.perm[k]] = k;
    // ---- final permuted pattern and counts; a postordering relabels the elimination tree, it does
    //      not change it: parent[k] = position of parent0[post[k]] in the postorder
    permuted_lower(n, Mp, Mi, S.iperm, Bp, Bi);
    std::vector<int> parent(n);
    {
        std::vector<int>& ipost = Ui;          // scratch: the transpose is no longer needed
        ipost.assign(n, 0);
        for (int k = 0; k < n; k++) ipost[post[k]] = k;
        for (int k = 0; k < n; k++) { const int p0 = parent0[post[k]]; parent[k] = p0 < 0 ? -1 : ipost[p0]; }
    }
    std::vector<int64_t> cc;
    colcounts(n, Bp, Bi, parent, cc);
    S.flops = 0; S.nnzL_true = 0;
    for (int j = 0; j < n; j++) { S.flops += (double)cc[j] * (double)cc[j]; S.nnzL_true += cc[j]; }
    mark("etree_counts");
    // ---- fundamental (maximal) supernodes
    std::vector<int> sfirst;  // first column of each supernode
    for (int j = 0; j < n; j++) {
        bool join = j > 0 && parent[j - 1] == j && cc[j - 1] == cc[j] + 1;
        if (!join) sfirst.push_back(j);
    }
    int ns = (int)sfirst.size();
    sfirst.push_back(n);
    std::vector<int> col2s(n);
    for (int s = 0; s < ns; s++) for (int j = sfirst[s]; j < sfirst[s + 1]; j++) col2s[j] = s;
    // ---- relaxed amalgamation (merge a supernode into its parent when it is the
    //      parent's last child and few explicit zeros are introduced)
    std::vector<int64_t> nc(ns), rr(ns), zz(ns, 0);
    std::vector<int> sp(ns), first_of(ns);
    std::vector<char> dead(ns, 0);
    for (int s = 0; s < ns; s++) {
        nc[s] = sfirst[s + 1] - sfirst[s];
        rr[s] = cc[sfirst[s]] - nc[s];
        int last = sfirst[s + 1] - 1;
        sp[s] = parent[last] < 0 ? -1 : col2s[parent[last]];
        first_of[s] = sfirst[s];
    }
    if (opt.relax_enable) {
        for (int s = 0; s < ns; s++) {
            int p = sp[s];
            if (p < 0) continue;
            if (first_of[s] + nc[s] != first_of[p]) continue;  // not the last (contiguous) child
            int64_t ncm = nc[s] + nc[p];
            int64_t newz = nc[s] * (nc[p] + rr[p] - rr[s]);
            int64_t z = zz[s] + zz[p] + newz;
            double total = (double)ncm * (ncm + 1) / 2.0 + (double)ncm * rr[p];
            double frac = total > 0 ? (double)z / total : 0.0;
            bool merge;
            if (ncm <= opt.relax_small || newz == 0) merge = true;
            else if (ncm <= 16) merge = frac < opt.relax_z16;
            else if (ncm <= 32) merge = frac < opt.relax_z32;
            else if (ncm <= 64) merge = frac < opt.relax_z64;
            else merge = frac < opt.relax_zinf;
            if (merge) {
                dead[s] = 1;
                first_of[p] = first_of[s];
                nc[p] = ncm; zz[p] = z;
            }
        }
    }
    S.sfirst.clear();
    for (int s = 0; s < ns; s++) if (!dead[s]) S.sfirst.push_back(first_of[s]);
    S.nsuper = (int)S.sfirst.size();
    S.sfirst.push_back(n);
    S.col2super.resize(n);
    for (int s = 0; s < S.nsuper; s++) for (int j = S.sfirst[s]; j < S.sfirst[s + 1]; j++) S.col2super[j] = s;
    S.sparent.assign(S.nsuper, -1);
    mark("supernodes");
    // ---- supernodal row structures (bottom-up union of children + own columns)
    const int NS = S.nsuper;
    S.child_ptr.assign(NS + 1, 0);
    S.rowptr.assign(NS + 1, 0);
    S.rowidx.clear();
    {
        std::vector<int> mark(n, -1), tmp;
        // children are discovered as structures are built: parent(s) = super(rows(s)[0])
        std::vector<std::vector<int>> kids(NS);
        for (int s = 0; s < NS; s++) {
            const int f = S.sfirst[s], l = S.sfirst[s + 1] - 1;
            tmp.clear();
            for (int j = f; j <= l; j++)
                for (int64_t p = Bp[j]; p < Bp[j + 1]; p++) {
                    int i = Bi[p];
                    if (i > l && mark[i] != s) { mark[i] = s; tmp.push_back(i); }
                }
            for (int c : kids[s])
                for (int64_t p = S.rowptr[c]; p < S.rowptr[c + 1]; p++) {
                    int i = S.rowidx[p];
                    if (i > l && mark[i] != s) { mark[i] = s; tmp.push_back(i); }
                }
            std::sort(tmp.begin(), tmp.end());
            S.rowidx.insert(S.rowidx.end(), tmp.begin(), tmp.end());
            S.rowptr[s + 1] = (int64_t)S.rowidx.size();
            if (!tmp.empty()) {
                int p = S.col2super[tmp[0]];
                S.sparent[s] = p;
                kids[p].push_back(s);
            }
            std::vector<int>().swap(kids[s]);
        }
    }
    for (int s = 0; s < NS; s++) if (S.sparent[s] >= 0) S.child_ptr[S.sparent[s] + 1]++;
    for (int s = 0; s < NS; s++) S.child_ptr[s + 1] += S.child_ptr[s];
    S.child_list.resize(S.child_ptr[NS]);
    {
        std::vector<int> nxt(S.child_ptr.begin(), S.child_ptr.end() - 1);
        for (int s = 0; s < NS; s++) if (S.sparent[s] >= 0) S.child_list[nxt[S.sparent[s]]++] = s;
    }
    mark("row_structures");
    // ---- storage offsets, levels
    S.Loff.assign(NS + 1, 0); S.CBoff.assign(NS + 1, 0);
    S.level.assign(NS, 0);
    S.max_front = 0;
    for (int s = 0; s < NS; s++) {
        int64_t c = S.sfirst[s + 1] - S.sfirst[s], r = S.rowptr[s + 1] - S.rowptr[s];
        S.Loff[s + 1] = S.Loff[s] + panel_ld(c + r) * c;
        S.max_front = std::max<int64_t>(S.max_front, c + r);
        if (S.sparent[s] >= 0) S.level[S.sparent[s]] = std::max(S.level[S.sparent[s]], S.level[s] + 1);
    }
    S.nnzL = S.Loff[NS];
    S.nlevels = 0;
    for (int s = 0; s < NS; s++) S.nlevels = std::max(S.nlevels, S.level[s] + 1);
    S.level_ptr.assign(S.nlevels + 1, 0);
    for (int s = 0; s < NS; s++) S.level_ptr[S.level[s] + 1]++;
    for (int l = 0; l < S.nlevels; l++) S.level_ptr[l + 1] += S.level_ptr[l];
    S.level_list.resize(NS);
    {
        std::vector<int> nxt(S.level_ptr.begin(), S.level_ptr.end() - 1);
        for (int s = 0; s < NS; s++) S.level_list[nxt[S.level[s]]++] = s;
    }
    // ---- storage of the update blocks.  Block s (r x r) is written while level(s) is processed and
    // read while level(parent(s)) is processed, so it is live on the level interval
    // [level(s), level(parent)].  Offsets are handed out level by level from a first-fit free list:
    // before level l starts, every block whose parent sits below l is returned.  (A prefix-sum
    // layout needs sum r^2 = 90 GB for the 100^3 PDE instance; the live set is a fraction of it.)
    {
        std::vector<std::vector<int>> dies(S.nlevels + 1);     // dies[l]: blocks free again when level l starts
        for (int s = 0; s < NS; s++) {
            const int p = S.sparent[s];
            if (p >= 0) dies[S.level[p] + 1].push_back(s);
        }
        std::map<int64_t, int64_t> free_list;                   // offset -> size, coalesced
        int64_t top = 0;
        S.cb_total = 0;
        auto release = [&](int64_t off, int64_t sz) {
            if (sz <= 0) return;
            auto it = free_list.emplace(off, sz).first;
            auto nx = std::next(it);
            if (nx != free_list.end() && it->first + it->second == nx->first) { it->second += nx->second; free_list.erase(nx); }
            if (it != free_list.begin()) {
                auto pv = std::prev(it);
                if (pv->first + pv->second == it->first) { pv->second += it->second; free_list.erase(it); it = pv; }
            }
            if (it->first + it->second == top) { top = it->first; free_list.erase(it); }   // shrink the arena
        };
        std::vector<int64_t> bsize(NS, 0);
        for (int l = 0; l < S.nlevels; l++) {
            for (int s : dies[l]) release(S.CBoff[s], bsize[s]);
            // biggest blocks first: they find (or create) their place before the small ones fragment it
            std::vector<int> order(S.level_list.begin() + S.level_ptr[l], S.level_list.begin() + S.level_ptr[l + 1]);
            auto rows_of = [&](int s) { return (int64_t)(S.rowptr[s + 1] - S.rowptr[s]); };
            std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return rows_of(a) > rows_of(b); });
            int64_t cursor = 0, cursor_sz = -1;
            bool have_cursor = false;
            for (int s : order) {
                const int64_t r = rows_of(s);
                const int64_t sz = (r * r + 1) & ~(int64_t)1;
                bsize[s] = (S.sparent[s] >= 0) ? sz : 0;
                if (S.sparent[s] < 0 || sz == 0) { S.CBoff[s] = 0; bsize[s] = 0; continue; }
                // first fit by address.  Nothing is released inside a level, so for a run of EQUAL sizes
                // every hole below the place where the previous search stopped is still too small: resume
                // there (same result as a search from the start); a smaller size starts over.
                int64_t off = -1;
                if (sz != cursor_sz) { have_cursor = false; cursor_sz = sz; }
                auto it = have_cursor ? free_list.lower_bound(cursor) : free_list.begin();
                for (; it != free_list.end(); ++it)
                    if (it->second >= sz) {
                        off = it->first;
                        const int64_t rest = it->second - sz;
                        free_list.erase(it);
                        if (rest > 0) free_list.emplace(off + sz, rest);
                        cursor = off; have_cursor = true;
                        break;
                    }
                if (off < 0) { cursor = top; have_cursor = true; }     // no hole fits this size: nor will one for the rest of the run
                if (off < 0) { off = top; top += sz; }
                S.CBoff[s] = off;
                S.cb_total = std::max(S.cb_total, off + sz);
            }
        }
        S.CBoff[NS] = S.cb_total;
    }
    mark("storage");
    // ---- relative indices of each update block inside the parent's front
    S.rel.assign(S.rowidx.size(), -1);
    std::atomic<bool> rel_bad{false};
    run_chunks_weighted(NS, S.rowptr.data(), 200000, [&](int, int64_t sb, int64_t se) {
        for (int s = (int)sb; s < (int)se; s++) {
            int p = S.sparent[s];
            if (p < 0) continue;
            const int pf = S.sfirst[p], pl = S.sfirst[p + 1] - 1, pc = pl - pf + 1;
            int64_t q = S.rowptr[p];
            const int64_t qe = S.rowptr[p + 1];
            for (int64_t t = S.rowptr[s]; t < S.rowptr[s + 1]; t++) {
                int x = S.rowidx[t];
                if (x <= pl) { S.rel[t] = x - pf; continue; }
                while (q < qe && S.rowidx[q] < x) q++;
                if (q >= qe || S.rowidx[q] != x) { rel_bad = true; return; }
                S.rel[t] = pc + (int)(q - S.rowptr[p]);
            }
        }
    });
    if (rel_bad) { S.error = "internal: child row missing from parent front"; return false; }
    // ---- forward-solve gather lists (the transpose of `rel`): per destination of every front the
    // update-vector entries of its children, ascending child order
    {
        const int64_t G = (int64_t)S.rowidx.size() + n;
        S.gptr.assign(G + 1, 0);
        for (int s = 0; s < NS; s++) {
            const int p = S.sparent[s];
            if (p < 0) continue;
            const int64_t gb = S.rowptr[p] + S.sfirst[p];
            for (int64_t t = S.rowptr[s]; t < S.rowptr[s + 1]; t++) S.gptr[gb + S.rel[t] + 1]++;
        }
        for (int64_t g = 0; g < G; g++) S.gptr[g + 1] += S.gptr[g];
        S.gsrc.assign(S.gptr[G], 0);
        S.gch.assign(S.gptr[G], 0);
        std::vector<int64_t> nxt(S.gptr.begin(), S.gptr.end() - 1);
        for (int s = 0; s < NS; s++) {          // ascending s = ascending child order inside every parent
            const int p = S.sparent[s];
            if (p < 0) continue;
            const int64_t gb = S.rowptr[p] + S.sfirst[p];
            for (int64_t t = S.rowptr[s]; t < S.rowptr[s + 1]; t++) {
                const int64_t e = nxt[gb + S.rel[t]]++;
                S.gsrc[e] = t;
                S.gch[e] = s;
            }
        }
    }
    mark("rel_gather");
    // ---- tile cuts of the update blocks inside the parents' update blocks (front_cb_kernel)
    S.tcut_ptr.assign(NS + 1, 0);
    for (int s = 0; s < NS; s++) {
        const int p = S.sparent[s];
        int cnt = 0;
        if (p >= 0) {
            const int pc = S.sfirst[p + 1] - S.sfirst[p], pN = pc + (int)(S.rowptr[p + 1] - S.rowptr[p]);
            cnt = (pN - (pc & ~1) + CB_TILE - 1) / CB_TILE + 2;     // one spare cut: a 128-column tile ends at an even cut
        }
        S.tcut_ptr[s + 1] = S.tcut_ptr[s] + cnt;
    }
    S.tcut.assign(S.tcut_ptr[NS], 0);
    for (int s = 0; s < NS; s++) {
        const int p = S.sparent[s];
        if (p < 0) continue;
        const int ce = (S.sfirst[p + 1] - S.sfirst[p]) & ~1;
        const int* r0 = S.rel.data() + S.rowptr[s];
        const int* r1 = S.rel.data() + S.rowptr[s + 1];
        const int cnt = S.tcut_ptr[s + 1] - S.tcut_ptr[s];
        for (int k = 0; k < cnt; k++)
            S.tcut[S.tcut_ptr[s] + k] = (int)(std::lower_bound(r0, r1, ce + k * CB_TILE) - r0);
    }
    mark("tile_cuts");
    // ---- map M_L entries into the L panels
    S.amap.assign(Mp[n], -1);
    S.dpos.assign(n, -1);
    std::atomic<bool> amap_bad{false};
    run_chunks_weighted(n, Mp.data(), 200000, [&](int, int64_t jb, int64_t je) {
        for (int j = (int)jb; j < (int)je; j++)
            for (int64_t e = Mp[j]; e < Mp[j + 1]; e++) {
                int i = Mi[e];
                int a = S.iperm[i], b = S.iperm[j];
                int row = std::max(a, b), col = std::min(a, b);
                int s = S.col2super[col];
                const int f = S.sfirst[s], l = S.sfirst[s + 1] - 1;
                const int64_t c = l - f + 1, r = S.rowptr[s + 1] - S.rowptr[s], ld = panel_ld(c + r);
                int64_t lrow;
                if (row <= l) lrow = row - f;
                else {
                    const int* b0 = S.rowidx.data() + S.rowptr[s];
                    const int* b1 = S.rowidx.data() + S.rowptr[s + 1];
                    const int* it = std::lower_bound(b0, b1, row);
                    if (it == b1 || *it != row) { amap_bad = true; return; }
                    lrow = c + (it - b0);
                }
                int64_t off = S.Loff[s] + lrow + (int64_t)(col - f) * ld;
                S.amap[e] = off;
                if (i == j) S.dpos[i] = off;
            }
    });
    if (amap_bad) { S.error = "internal: entry missing from supernode structure"; return false; }
    mark("amap");
    return true;
}

void shard_map(const Symbolic& S, int world, double split_flops, ShardMap& out) {
    const int NS = S.nsuper;
    out = ShardMap();
    out.world = std::max(1, world);
    out.owner.assign(NS, 0);
    out.top.assign(NS, 0);
    out.level_barrier.assign(std::max(1, S.nlevels), 0);
    out.level_split.assign(std::max(1, S.nlevels), 0);
    out.split.assign(NS, 0);
    out.ra.assign(NS, 0); out.rb.assign(NS, 1);
    out.load.assign(out.world, 0.0);
    // factorisation flops of every supernode and of the subtree below it (children have smaller
    // indices: the supernodes are numbered in postorder)
    std::vector<double> w(NS), W(NS);
    std::vector<int> fd(NS);             // first descendant: subtree(s) = [fd[s], s]
    for (int s = 0; s < NS; s++) {
        const double c = S.sfirst[s + 1] - S.sfirst[s], N = c + (double)(S.rowptr[s + 1] - S.rowptr[s]);
        // sum_{j<c} (N-j)^2
        w[s] = c * N * N - N * c * (c - 1) + (c - 1) * c * (2 * c - 1) / 6.0;
        W[s] = w[s];
        fd[s] = s;
    }
    for (int s = 0; s < NS; s++) {
        const int p = S.sparent[s];
        if (p >= 0) { W[p] += W[s]; fd[p] = std::min(fd[p], fd[s]); }
    }
    if (out.world > 1) {
        double total_all = 0;
        struct Task { std::vector<int> roots; int a, b; };
        std::vector<Task> stack;
        {
            Task t; t.a = 0; t.b = out.world;
            for (int s = 0; s < NS; s++) if (S.sparent[s] < 0) { t.roots.push_back(s); total_all += W[s]; }
            stack.push_back(std::move(t));
        }
        int expansions = 0;
        const int max_expansions = 64 * out.world;
        while (!stack.empty()) {
            Task T = std::move(stack.back());
            stack.pop_back();
            if (T.b - T.a == 1) {
                for (int r : T.roots) for (int s = fd[r]; s <= r; s++) out.owner[s] = T.a;
                continue;
            }
            const int mid = T.a + (T.b - T.a) / 2;
            const double f1 = (double)(mid - T.a) / (T.b - T.a), f2 = 1.0 - f1;
            std::vector<int> bin1, bin2;
            for (;;) {
                std::sort(T.roots.begin(), T.roots.end(), [&](int x, int y) { return W[x] != W[y] ? W[x] > W[y] : x < y; });
                double total = 0, l1 = 0, l2 = 0;
                for (int r : T.roots) total += W[r];
                bin1.clear(); bin2.clear();
                for (int r : T.roots) {
                    if (l1 / f1 <= l2 / f2) { bin1.push_back(r); l1 += W[r]; } else { bin2.push_back(r); l2 += W[r]; }
                }
                const bool splittable = T.roots.size() >= 2;
                const double imbalance = total > 0 ? std::max(l1 / f1, l2 / f2) / total : 1.0;
                if (splittable && imbalance <= 1.10) break;
                // expand the heaviest root that still has children
                int pick = -1;
                for (size_t k = 0; k < T.roots.size(); k++) {
                    const int r = T.roots[k];
                    if (S.child_ptr[r + 1] > S.child_ptr[r]) { pick = (int)k; break; }
                }
                if (pick < 0 || expansions >= max_expansions || W[T.roots[pick]] < 1e-3 * total_all) break;
                const int hnode = T.roots[pick];
                expansions++;
                out.owner[hnode] = T.a;
                out.top[hnode] = 1;
                out.ra[hnode] = T.a; out.rb[hnode] = T.b;
                T.roots.erase(T.roots.begin() + pick);
                for (int k = S.child_ptr[hnode]; k < S.child_ptr[hnode + 1]; k++) T.roots.push_back(S.child_list[k]);
            }
            Task t1, t2;
            t1.a = T.a; t1.b = mid; t1.roots = std::move(bin1);
            t2.a = mid; t2.b = T.b; t2.roots = std::move(bin2);
            stack.push_back(std::move(t1));
            stack.push_back(std::move(t2));
        }
    }
    for (int s = 0; s < NS; s++) {
        // update blocks worth splitting: r^2 c above split_flops (default 2e10: a millisecond of one GPU)
        const double c = S.sfirst[s + 1] - S.sfirst[s], r = (double)(S.rowptr[s + 1] - S.rowptr[s]);
        if (out.top[s] && out.rb[s] - out.ra[s] >= 2 && r > 0 && r * r * c >= split_flops) {
            out.split[s] = 1;
            out.level_split[S.level[s]] = 1;
        } else if (!out.top[s]) { out.ra[s] = out.owner[s]; out.rb[s] = out.owner[s] + 1; }
        if (out.split[s]) {
            const double fcb = r * r * c;
            out.load[out.owner[s]] += w[s] - fcb;
            for (int q = out.ra[s]; q < out.rb[s]; q++) out.load[q] += fcb / (out.rb[s] - out.ra[s]);
        } else
        out.load[out.owner[s]] += w[s];
        if (out.top[s]) out.top_flops += w[s];
        const int p = S.sparent[s];
        if (p >= 0 && out.owner[p] != out.owner[s]) out.level_barrier[S.level[p]] = 1;
    }
    // synchronisation groups per level (union-find over the ranks)
    const int nl = std::max(1, S.nlevels), Wd = out.world;
    out.level_mask.assign((size_t)nl * Wd, 0u);
    std::vector<int> uf((size_t)nl * Wd);
    for (int l = 0; l < nl; l++) for (int r = 0; r < Wd; r++) uf[(size_t)l * Wd + r] = r;
    auto find = [&](int l, int r) { int* u = uf.data() + (size_t)l * Wd; while (u[r] != r) { u[r] = u[u[r]]; r = u[r]; } return r; };
    auto unite = [&](int l, int a, int b) { a = find(l, a); b = find(l, b); if (a != b) uf[(size_t)l * Wd + std::max(a, b)] = std::min(a, b); };
    for (int s = 0; s < NS; s++) {
        const int p = S.sparent[s];
        if (p >= 0 && out.owner[p] != out.owner[s]) unite(S.level[p], out.owner[p], out.owner[s]);
        if (out.split[s]) for (int q = out.ra[s] + 1; q < out.rb[s]; q++) unite(S.level[s], out.ra[s], q);
    }
    for (int l = 0; l < nl; l++) {
        std::vector<unsigned> comp(Wd, 0u);
        for (int r = 0; r < Wd; r++) comp[find(l, r)] |= 1u << r;
        for (int r = 0; r < Wd; r++) {
            const unsigned m = comp[find(l, r)];
            out.level_mask[(size_t)l * Wd + r] = (m & (m - 1)) ? m : 0u;      // a component of one rank needs no barrier
        }
    }
}

}  // namespace opb
