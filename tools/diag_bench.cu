// Standalone timing harness for the diagonal-block kernel (development tool, not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I onephase.jl_b200/csrc tools/diag_bench.cu -o tools/diag_bench
#include <cstdio>
#include <vector>
#include <cmath>
#include "kernels_dense.cu"
namespace opb { std::atomic<long long> g_launches{0};
void launch_big_extend_add(const DevSym&, const LevelPlan&, const int*, double*, double*, DeltaState*, cudaStream_t) {} }
using namespace opb;

__global__ void __launch_bounds__(PT) prof_kernel(double* A, int b, int lda, long long* stamps, int* fail) {
    extern __shared__ double D[];
    __shared__ int s_fail;
    const int tid = threadIdx.x;
    double* invbuf = D + LDD * WB;
    double* Xs = invbuf + INVBUF;
    double* base = A + (size_t)blockIdx.x * lda * b;
    long long* st = stamps + blockIdx.x * 64;
    if (tid == 0) st[0] = clock64();
    for (int idx = tid; idx < b * b; idx += PT) { const int i = idx % b, j = idx / b; D[i + j * LDD] = (i >= j) ? base[i + (size_t)j * lda] : 0.0; }
    __syncthreads();
    if (tid == 0) st[1] = clock64();
    bool ok = panel_chol_smem<true>(D, LDD, b, b, invbuf, Xs, &s_fail, st + 2);
    if (!ok) { if (tid == 0) *fail = 1; return; }
    if (tid == 0) st[40] = clock64();
    for (int idx = tid; idx < b * b; idx += PT) { const int i = idx % b, j = idx / b; if (i >= j) base[i + (size_t)j * lda] = D[i + j * LDD]; }
    __syncthreads();
    if (tid == 0) st[41] = clock64();
}

int main() {
    const int b = 128, lda = 128, nblk = 148;
    std::vector<double> h((size_t)nblk * b * lda);
    for (int q = 0; q < nblk; q++) for (int j = 0; j < b; j++) for (int i = 0; i < b; i++)
        h[(size_t)q * b * lda + i + (size_t)j * lda] = (i == j) ? 200.0 : 1.0 / (1.0 + abs(i - j));
    double* d; long long* st; int* fail;
    cudaMalloc(&d, h.size() * 8); cudaMalloc(&st, nblk * 64 * 8); cudaMalloc(&fail, 4); cudaMemset(fail, 0, 4);
    size_t smem = (size_t)(LDD * WB + INVBUF + XS_BLOCKS * INVBUF) * 8;
    cudaFuncSetAttribute(prof_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    for (int rep = 0; rep < 3; rep++) {
        cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        prof_kernel<<<nblk, PT, smem>>>(d, b, lda, st, fail);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> hs(64); cudaMemcpy(hs.data(), st, 64 * 8, cudaMemcpyDeviceToHost);
        int hf; cudaMemcpy(&hf, fail, 4, cudaMemcpyDeviceToHost);
        printf("rep %d: %.1f us (err %s, fail %d)\n", rep, ms * 1e3, cudaGetErrorString(cudaGetLastError()), hf);
        printf("  load %lld\n", hs[1] - hs[0]);
        long long* p = hs.data() + 2;
        for (int kb = 0; kb < 4; kb++) printf("  kb%d: potrf+inv %lld  trsm %lld  update %lld\n", kb, p[3*kb+1]-p[3*kb], p[3*kb+2]-p[3*kb+1], (kb<3? p[3*kb+3]: p[12]) - p[3*kb+2]);
        printf("  offdiag X %lld   store %lld   total %lld clks\n", p[13] - p[12], hs[41] - hs[40], hs[41] - hs[0]);
    }
    return 0;
}
