// ThreadSanitizer harness for the threaded symbolic analysis (host only, no CUDA):
//   g++ -O1 -g -std=c++17 -fsanitize=thread -pthread -I onephase.jl_b200/csrc tools/symbolic_tsan.cpp \
//       onephase.jl_b200/csrc/symbolic.cpp /usr/local/cuda/lib64/libmetis_static.a -o /tmp/symbolic_tsan
//   /tmp/symbolic_tsan 40 4      # grid size N (n = N^3), ordering (0 = own dissection, 4 = auto: own beside METIS)
// 3-D grid, J = [7-point stencil rows; identity rows], H = I.  Round 2: 40^3 with orderings 0 and 4, 64^3 with 0 -- no reports.
#include "symbolic.h"
#include <algorithm>
#include <cstdio>
#include <cstdlib>
using namespace opb;
int main(int argc, char** argv) {
    int N = argc > 1 ? atoi(argv[1]) : 40;
    int ord = argc > 2 ? atoi(argv[2]) : 0;
    int64_t n = (int64_t)N * N * N, m = 2 * n;
    // J in CSC: column j has the stencil rows of its neighbours (rows 0..n-1) and the identity row n+j
    std::vector<int64_t> Jp(n + 1, 0), Ji;
    auto id = [&](int x, int y, int z) { return ((int64_t)z * N + y) * N + x; };
    for (int z = 0; z < N; z++) for (int y = 0; y < N; y++) for (int x = 0; x < N; x++) {
        int64_t j = id(x, y, z);
        std::vector<int64_t> rows;
        rows.push_back(j);
        if (x > 0) rows.push_back(id(x - 1, y, z)); if (x + 1 < N) rows.push_back(id(x + 1, y, z));
        if (y > 0) rows.push_back(id(x, y - 1, z)); if (y + 1 < N) rows.push_back(id(x, y + 1, z));
        if (z > 0) rows.push_back(id(x, y, z - 1)); if (z + 1 < N) rows.push_back(id(x, y, z + 1));
        rows.push_back(n + j);
        std::sort(rows.begin(), rows.end());
        for (auto r : rows) Ji.push_back(r);
        Jp[j + 1] = (int64_t)Ji.size();
    }
    std::vector<int64_t> Hp(n + 1), Hi(n);
    for (int64_t j = 0; j <= n; j++) Hp[j] = j;
    for (int64_t j = 0; j < n; j++) Hi[j] = j;
    SchurPattern P; std::string err;
    if (!build_schur_pattern(n, m, Jp.data(), Ji.data(), Hp.data(), Hi.data(), 0, P, err)) { printf("pattern: %s\n", err.c_str()); return 1; }
    SymOptions opt; opt.ordering = ord;
    Symbolic S;
    if (!analyze((int)n, P.Mp, P.Mi, opt, nullptr, S)) { printf("analyze: %s\n", S.error.c_str()); return 1; }
    ShardMap sm; shard_map(S, 4, opt.shard_split_flops, sm);
    printf("n=%lld nnzM=%lld flops=%.4e nsuper=%d\n", (long long)n, (long long)P.Mp[n], S.flops, S.nsuper);
    return 0;
}
