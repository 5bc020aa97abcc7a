// Host-side symbolic analysis for the B200 KKT path: runs once per sparsity
// pattern and is reused across IPM iterations (the reference redoes CHOLMOD's
// analyze on every ls_factor! call, linear_system_solvers/julia.jl:34, because
// linear_solver_recycle=false, parameters.jl:38).
#pragma once
#include <cstdint>
#include <string>
#include <utility>
#include <vector>

namespace opb {

// Leading dimension of a supernode panel with N = c + r rows: padded to an even
// count so that every panel column starts on a 16-byte boundary (TMA bulk copies).
inline int64_t panel_ld(int64_t N) { return (N + 1) & ~(int64_t)1; }
constexpr int CB_TILE = 64;    // granularity of the tile-cut table (kernels_dense.cu: tiles of 64 / 128 rows, 128 columns)

struct SymOptions {
    int nd_leaf = 96;          // nested dissection stops at parts of this size
    double nd_balance = 0.40;  // a separator level must leave at least this fraction of the part on either side
    int ordering = 4;          // 4 = auto (best of 0 and 3), 0 = level-structure nested dissection +
                               // min-degree leaves, 1 = natural, 3 = METIS_NodeND; a user permutation wins
    int metis_max_n = 400000;  // auto: try METIS only up to this many variables
    double relax_small = 8;    // always merge a last child when merged width <= this
    double relax_z16 = 0.8, relax_z32 = 0.3, relax_z64 = 0.1, relax_zinf = 0.05;
    int relax_enable = 1;
    double shard_split_flops = 2e10;   // sharded instance: update blocks of top fronts with at least this many flops (r^2 c) are split over the ranks of their range
};

// Pattern of M_L = tril(J' D J + H) with full diagonal, and the gather map
// that assembles it (schur.jl:55: Q = J_T * Diagonal(y./s) * J + H).
struct SchurPattern {
    int n = 0, m = 0;
    int64_t nnzJ = 0, nnzH = 0;
    std::vector<int64_t> Mp;     // n+1
    std::vector<int> Mi;         // nnzM, sorted in each column, first entry = diagonal
    std::vector<int64_t> pair_ptr;   // nnzM+1
    std::vector<int> pairA, pairB;   // positions in J.nzval: A -> J[k,row], B -> J[k,col]
    std::vector<int> hmap;       // nnzM: position in H.nzval or -1
    std::vector<int> Jrow;       // nnzJ: 0-based row of every J entry (CSC order)
    std::vector<int64_t> Jp;     // n+1, 0-based
    // CSR view of J for J*x products
    std::vector<int64_t> Rp;     // m+1
    std::vector<int> Rcol, Rpos; // nnzJ
    // symmetric (both triangles) CSR view of H for H_sym*v
    std::vector<int64_t> Sp;     // n+1
    std::vector<int> Scol, Spos; // 2*nnzH - ndiag
};

struct Symbolic {
    int n = 0;
    std::vector<int> perm, iperm;        // perm[new] = old
    int nsuper = 0;
    std::vector<int> sfirst;             // nsuper+1, permuted column ranges
    std::vector<int> sparent;            // supernodal elimination tree (-1 = root)
    std::vector<int64_t> rowptr;         // nsuper+1
    std::vector<int> rowidx;             // below-diagonal rows of each supernode (permuted, sorted)
    std::vector<int> rel;                // same shape as rowidx: position in the parent's front
    std::vector<int64_t> Loff;           // nsuper+1: panel (c+r) x c, column-major, ld = panel_ld(c+r)
    std::vector<int64_t> CBoff;          // nsuper+1: update block r x r, column-major, ld = r
    std::vector<int> level;              // height above the leaves
    int nlevels = 0;
    std::vector<int> level_ptr, level_list;  // supernodes grouped by level
    std::vector<int> child_ptr, child_list;  // children of each supernode (ascending)
    // forward-solve gather lists: destination d (0 <= d < c+r) of supernode s sums the entries
    // gsrc[gptr[g] .. gptr[g+1]) of the update-vector storage, g = rowptr[s] + sfirst[s] + d, in
    // ascending child order (gch = the child each entry belongs to)
    std::vector<int64_t> gptr, gsrc;
    std::vector<int> gch;
    // tile cuts of every update block inside its parent's update block: supernode s (rows rel[.])
    // crosses the parent's CB_TILE-row tile boundary k (front position (c_p & ~1) + k * CB_TILE) at
    // row position tcut[tcut_ptr[s] + k], k = 0 .. tiles(parent) + 1; lets the update-block kernel find
    // the child entries of a tile with four loads instead of four binary searches
    std::vector<int> tcut_ptr, tcut;
    std::vector<int64_t> amap;           // per M_L entry: destination offset in L storage
    std::vector<int64_t> dpos;           // per original variable: offset of its diagonal in L
    std::vector<int> col2super;          // permuted column -> supernode
    int64_t nnzL = 0;                    // sum panel_ld(c+r)*c
    int64_t nnzL_true = 0;               // entries of the trapezoids (lower part only)
    int64_t cb_total = 0;
    double flops = 0;                    // sum_j colcount_j^2
    int max_front = 0;
    std::string error;
    std::vector<std::pair<std::string, double>> timing;   // host seconds per analysis phase (info keys t_<phase>)
};

// index_base: 0 or 1.  Returns false and sets err on invalid input.
bool build_schur_pattern(int64_t n, int64_t m, const int64_t* Jp, const int64_t* Ji,
                         const int64_t* Hp, const int64_t* Hi, int index_base,
                         SchurPattern& out, std::string& err);

// Pattern from a user CSC matrix (L1-compat path, julia.jl:21-97): only entries
// with row >= col are read; the full diagonal is always present.  src[e] is the
// position in the caller's nzval or -1 for an inserted diagonal.
bool build_csc_pattern(int64_t n, const int64_t* Ap, const int64_t* Ai, int index_base,
                       std::vector<int64_t>& Mp, std::vector<int>& Mi, std::vector<int64_t>& src,
                       std::string& err);

bool analyze(int n, const std::vector<int64_t>& Mp, const std::vector<int>& Mi,
             const SymOptions& opt, const int64_t* user_perm, Symbolic& S);

// Mapping of the supernodal elimination tree onto `world` GPUs: proportional
// ("subtree-to-subcube") mapping by factorisation flops.  A rank range that cannot be split
// evenly over the current subtree roots expands its heaviest root: that supernode becomes a
// `top` supernode (it has descendants on other ranks) owned by the first rank of the range, and
// its children become roots.  Everything below the final roots is private to one rank.
struct ShardMap {
    int world = 1;
    std::vector<int> owner;              // per supernode: rank that factorises and solves it
    std::vector<char> top;               // per supernode: descendants live on more than one rank
    std::vector<char> level_barrier;     // per level: some supernode of the level has a child on another rank
    // Top supernodes whose update block is formed by ALL ranks of the range the supernode was expanded
    // in (`split`): rank range [ra, rb), tile t of the update block belongs to rank ra + t mod (rb - ra).
    // The owner factorises the panel; the helpers pull it over NVLink and store their tiles into the
    // owner's update-block arena.  level_split: the level has such a supernode (every rank runs the
    // two extra barriers of that level).
    // level_mask[l * world + r]: the ranks rank r synchronises with at level l (bit mask incl. r itself, 0 = no
    // barrier for r): the connected component of r under "owner of a supernode of the level <-> owner of one
    // of its children" and "ranks of a split front's range".  Ranks whose subtrees are private at a level do
    // not wait for anybody there.
    std::vector<unsigned> level_mask;
    std::vector<char> split, level_split;
    std::vector<int> ra, rb;
    std::vector<double> load;            // per rank: flops of the supernodes it owns
    double top_flops = 0;                // flops of the top supernodes
};
void shard_map(const Symbolic& S, int world, double split_flops, ShardMap& out);

uint64_t pattern_hash(int64_t n, const int64_t* p, const int64_t* i, int64_t nnz);

}  // namespace opb
