// Standalone throughput / correctness harness for the FP64 tile engine of kernels_dense.cu
// (development tool, not part of the library):
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I onephase.jl_b200/csrc tools/gemm_bench.cu -o tools/gemm_bench
// Runs C(lower tiles) = A * A' for an N x K column-major A at several K and prints TFLOP/s.
#include <cmath>
#include <cstdio>
#include <vector>
#include "kernels_dense.cu"
namespace opb { std::atomic<long long> g_launches{0}; }
using namespace opb;

__global__ void __launch_bounds__(GEMM_THREADS)
syrk_bench_kernel(const double* __restrict__ A, int lda, int N, int K, double* __restrict__ C, int ldc) {
    extern __shared__ __align__(16) unsigned char smraw[];
    GemmSmem& sm = *reinterpret_cast<GemmSmem*>(smraw);
    const long long tp = blockIdx.x;
    int I = (int)((sqrt(8.0 * (double)tp + 1.0) - 1.0) * 0.5);
    while ((long long)I * (I + 1) / 2 > tp) I--;
    while ((long long)(I + 1) * (I + 2) / 2 <= tp) I++;
    const int J = (int)(tp - (long long)I * (I + 1) / 2);
    const int ri = I * BM, rj = J * BN;
    double acc[8][4][2];
    gemm_mainloop<false>(sm, A + ri, lda, min(BM, N - ri), A + rj, lda, min(BN, N - rj), K, acc);
    if (!gemm_compute_warp()) return;
#pragma unroll
    for (int nt = 0; nt < 4; nt++)
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const int k = rj + acc_col(nt, e);
            if (k >= N) continue;
#pragma unroll
            for (int mt = 0; mt < 8; mt++) {
                const int i = ri + acc_row(mt);
                if (i < N && i >= k) C[i + (size_t)k * ldc] = acc[mt][nt][e];
            }
        }
}

int main(int argc, char** argv) {
    const int N = argc > 1 ? atoi(argv[1]) : 8192 + 70;      // ragged last tile on purpose
    const int Kmax = 4096 + 6;
    const int lda = (N + 1) & ~1;
    std::vector<double> h((size_t)lda * Kmax);
    unsigned long long seed = 12345;
    for (auto& v : h) { seed = seed * 6364136223846793005ull + 1442695040888963407ull; v = ((seed >> 33) % 2001) / 1000.0 - 1.0; }
    double *dA, *dC;
    cudaMalloc(&dA, h.size() * 8);
    cudaMalloc(&dC, (size_t)lda * N * 8);
    cudaMemcpy(dA, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    cudaFuncSetAttribute(syrk_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(GemmSmem));
    const int nt = (N + BM - 1) / BM;
    const unsigned tiles = (unsigned)((long long)nt * (nt + 1) / 2);
    const int Ks[] = {128, 512, 1024, 4096 + 6};
    for (int K : Ks) {
        cudaMemset(dC, 0, (size_t)lda * N * 8);
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        syrk_bench_kernel<<<tiles, GEMM_THREADS, sizeof(GemmSmem)>>>(dA, lda, N, K, dC, lda);
        cudaDeviceSynchronize();
        float best = 1e30f;
        for (int rep = 0; rep < 3; rep++) {
            cudaEventRecord(e0);
            syrk_bench_kernel<<<tiles, GEMM_THREADS, sizeof(GemmSmem)>>>(dA, lda, N, K, dC, lda);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            best = ms < best ? ms : best;
        }
        // spot check against a host dot product
        double maxerr = 0;
        const int pts[][2] = {{0, 0}, {N - 1, 0}, {N - 1, N - 1}, {N - 3, N - 70}, {4097, 129}, {300, 255}, {N / 2, N / 2 - 1}};
        for (auto& p : pts) {
            double ref = 0;
            for (int k = 0; k < K; k++) ref += h[p[0] + (size_t)k * lda] * h[p[1] + (size_t)k * lda];
            double got; cudaMemcpy(&got, dC + p[0] + (size_t)p[1] * lda, 8, cudaMemcpyDeviceToHost);
            maxerr = fmax(maxerr, fabs(got - ref) / (fabs(ref) + 1e-300));
        }
        const double flops = 2.0 * (double)tiles * BM * BN * K;
        printf("N=%d K=%d tiles=%u: %.3f ms  %.2f TFLOP/s (full-tile flops)  max rel err %.2e  (%s)\n", N, K, tiles, best,
               flops / (best * 1e-3) / 1e12, maxerr, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
