// Micro-benchmark of the tile-GEMM engine through the node-level Linear kernel (development aid, not shipped).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I.. tools/gemm_bench.cu ../lin_kernel.o -o gemm_bench
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "../kernels.h"
using namespace gb;

int main(int argc, char** argv) {
    const int HP = argc > 1 ? atoi(argv[1]) : 196;
    const int tiles_per_sm = argc > 2 ? atoi(argv[2]) : 8;
    const int K2 = argc > 3 ? atoi(argv[3]) : 0;
    const int M = 148 * tiles_per_sm * 128;
    float *A, *W, *O;
    cudaMalloc(&A, (size_t)M * HP * 4); cudaMalloc(&W, (size_t)2 * HP * HP * 4); cudaMalloc(&O, (size_t)M * HP * 4);
    std::vector<float> h((size_t)M * HP);
    for (auto& v : h) v = (rand() % 2000 - 1000) * 1e-3f;
    cudaMemcpy(A, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(W, h.data(), (size_t)2 * HP * HP * 4, cudaMemcpyHostToDevice);
    LinArgs a{}; a.A1 = A; a.lda1 = HP; a.K1 = HP; a.A2 = K2 ? A : nullptr; a.lda2 = HP; a.K2 = K2 ? HP : 0; a.wt = W; a.out = O; a.ldo = HP;
    a.M = M; a.ncb = 1; a.epi = EPI_BIAS; a.res_cb = -1;
    for (int i = 0; i < 3; ++i) launch_lin(HP, a, 0);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 10;
    cudaEventRecord(e0);
    for (int i = 0; i < reps; ++i) launch_lin(HP, a, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= reps;
    const double fl = 2.0 * M * (double)HP * (HP + (K2 ? HP : 0));
    printf("HP=%d M=%d K=%d: %.3f ms  %.1f TFLOP/s  (%s)\n", HP, M, HP + (K2 ? HP : 0), ms, fl / ms * 1e-9, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
