// Stand-alone validation of the tcgen05 3xTF32 tile GEMM (development aid):  C[M x N] = A[M x K] * W[N x K]^T
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I.. tools/tc_gemm_test.cu -o tools/tc_gemm_test
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../tc_common.cuh"

using namespace gb;
using namespace gb::tc;

template <int N, int K>
struct Cfg {
    static constexpr int NA = K / ATOM_K;
    static constexpr int S = 2;
    static constexpr int A_BYTES = 128 * ATOM_ROW_BYTES;       // one half (hi or lo) of an A atom
    static constexpr int W_BYTES = N * ATOM_ROW_BYTES;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
    static constexpr int SMEM = S * STAGE_BYTES + 1024 + 256;
};

template <int N, int K>
__global__ void __launch_bounds__(192, 1) tc_gemm_kernel(const float* __restrict__ A, const float* __restrict__ Wp, float* __restrict__ C, int M) {
    using CF = Cfg<N, K>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + CF::S * CF::STAGE_BYTES);
    uint64_t* full_a = bars; uint64_t* full_w = bars + CF::S; uint64_t* empty = bars + 2 * CF::S;
    uint64_t* d_full = bars + 3 * CF::S; uint64_t* d_empty = d_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < CF::S; ++s) { mbar_init(&full_a[s], 128); mbar_init(&full_w[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(d_full, 1); mbar_init(d_empty, 128);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<256>(tmem_slot);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int n_tiles = M / 128;
    constexpr uint32_t idesc = instr_desc_tf32(N);

    if (warp == 0) {                                   // ---- TMA producer: W atoms ----
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int j = 0; j < CF::NA; ++j, ++it) {
                    const uint32_t s = it % CF::S, r = it / CF::S;
                    if (r > 0) mbar_wait(&empty[s], (r - 1) & 1);
                    mbar_arrive_expect_tx(&full_w[s], 2 * CF::W_BYTES);
                    bulk_g2s(base + s * CF::STAGE_BYTES + 2 * CF::A_BYTES, Wp + (size_t)j * 2 * N * ATOM_K, 2 * CF::W_BYTES, &full_w[s]);
                }
        }
    } else if (warp == 1) {                            // ---- MMA issuer ----
        if (lane == 0) {
            uint32_t it = 0, tc_cnt = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tc_cnt) {
                if (tc_cnt > 0) mbar_wait(d_empty, (tc_cnt - 1) & 1);
                fence_after_sync();
                for (int j = 0; j < CF::NA; ++j, ++it) {
                    const uint32_t s = it % CF::S, r = it / CF::S;
                    mbar_wait(&full_a[s], r & 1);
                    mbar_wait(&full_w[s], r & 1);
                    fence_after_sync();
                    const uint32_t a_hi = smem_u32(base + s * CF::STAGE_BYTES), a_lo = a_hi + CF::A_BYTES;
                    const uint32_t w_hi = a_hi + 2 * CF::A_BYTES, w_lo = w_hi + CF::W_BYTES;
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t ko = kk * 32;
                        mma_tf32(tmem_base, smem_desc(a_lo + ko), smem_desc(w_hi + ko), idesc, (j | kk) != 0);
                        mma_tf32(tmem_base, smem_desc(a_hi + ko), smem_desc(w_lo + ko), idesc, 1);
                        mma_tf32(tmem_base, smem_desc(a_hi + ko), smem_desc(w_hi + ko), idesc, 1);
                    }
                    mma_commit(&empty[s]);
                }
                mma_commit(d_full);
            }
        }
    } else {                                           // ---- workers: build A atoms, then epilogue ----
        const int group = warp & 3;                    // TMEM lane group this warp may access
        const int r = group * 32 + lane;
        uint32_t it = 0, tc_cnt = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tc_cnt) {
            const float* arow = A + ((size_t)tile * 128 + r) * K;
            for (int j = 0; j < CF::NA; ++j, ++it) {
                const uint32_t s = it % CF::S, rr = it / CF::S;
                float4 x[8];
#pragma unroll
                for (int c = 0; c < 8; ++c) x[c] = __ldg(reinterpret_cast<const float4*>(arow + j * ATOM_K + 4 * c));
                if (rr > 0) mbar_wait(&empty[s], (rr - 1) & 1);
                unsigned char* a_hi = base + s * CF::STAGE_BYTES;
#pragma unroll
                for (int c = 0; c < 8; ++c) store_split(a_hi, a_hi + CF::A_BYTES, r, c, x[c]);
                fence_proxy_async();
                mbar_arrive(&full_a[s]);
            }
            mbar_wait(d_full, tc_cnt & 1);
            fence_after_sync();
            float* crow = C + ((size_t)tile * 128 + r) * N;
#pragma unroll 1
            for (int ch = 0; ch < N / 32; ++ch) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(group * 32) << 16) + ch * 32, v);
#pragma unroll
                for (int q = 0; q < 8; ++q) *reinterpret_cast<float4*>(crow + ch * 32 + 4 * q) = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
            }
            fence_before_sync();
            mbar_arrive(d_empty);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc<256>(tmem_base);
}

int main(int argc, char** argv) {
    constexpr int N = 192, K = 192;
    const int tiles_per_sm = argc > 1 ? atoi(argv[1]) : 4;
    const int M = 148 * tiles_per_sm * 128;
    std::vector<float> hA((size_t)M * K), hW((size_t)N * K), hWp((size_t)2 * N * K);
    srand(1);
    for (auto& v : hA) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto& v : hW) v = (rand() / (float)RAND_MAX - 0.5f) * 0.2f;
    for (int j = 0; j < K / 32; ++j)
        for (int n = 0; n < N; ++n)
            for (int c = 0; c < 8; ++c)
                for (int e = 0; e < 4; ++e) {
                    const float w = hW[(size_t)n * K + 32 * j + 4 * c + e];
                    uint32_t u; memcpy(&u, &w, 4); u &= 0xffffe000u; float hi; memcpy(&hi, &u, 4);
                    const size_t off = (size_t)(n / 8) * 256 + (n % 8) * 32 + ((c ^ (n % 8)) * 4) + e;
                    hWp[((size_t)j * 2 + 0) * N * 32 + off] = hi;
                    hWp[((size_t)j * 2 + 1) * N * 32 + off] = w - hi;
                }
    float *dA, *dWp, *dC;
    cudaMalloc(&dA, hA.size() * 4); cudaMalloc(&dWp, hWp.size() * 4); cudaMalloc(&dC, (size_t)M * N * 4);
    cudaMemcpy(dA, hA.data(), hA.size() * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(dWp, hWp.data(), hWp.size() * 4, cudaMemcpyHostToDevice);
    cudaMemset(dC, 0xff, (size_t)M * N * 4);
    auto kern = tc_gemm_kernel<N, K>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<N, K>::SMEM);
    kern<<<148, 192, Cfg<N, K>::SMEM>>>(dA, dWp, dC, M);
    cudaError_t e = cudaDeviceSynchronize();
    printf("launch: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<float> hC((size_t)M * N);
    cudaMemcpy(hC.data(), dC, hC.size() * 4, cudaMemcpyDeviceToHost);
    double max_err = 0, max_err32 = 0, max_ref = 0;
    for (int t = 0; t < 64; ++t) {
        const int m = (int)(((long long)t * 7919 * 131) % M);
        for (int n = 0; n < N; ++n) {
            double ref = 0; float ref32 = 0;
            for (int k = 0; k < K; ++k) { ref += (double)hA[(size_t)m * K + k] * hW[(size_t)n * K + k]; ref32 = fmaf(hA[(size_t)m * K + k], hW[(size_t)n * K + k], ref32); }
            max_err = fmax(max_err, fabs(hC[(size_t)m * N + n] - ref));
            max_err32 = fmax(max_err32, fabs(ref32 - ref));
            max_ref = fmax(max_ref, fabs(ref));
        }
    }
    printf("max|ref|=%.4f  max abs err tcgen05-3xTF32 = %.3e   (plain fp32 fma chain: %.3e)\n", max_ref, max_err, max_err32);
    printf("C[0][0..3] = %f %f %f %f\n", hC[0], hC[1], hC[2], hC[3]);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    for (int i = 0; i < 10; ++i) kern<<<148, 192, Cfg<N, K>::SMEM>>>(dA, dWp, dC, M);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= 10;
    printf("M=%d: %.3f ms -> %.1f TFLOP/s (fp32-equivalent), %.1f TFLOP/s tensor (3 MMAs)\n", M, ms, 2.0 * M * N * K / ms * 1e-9, 6.0 * M * N * K / ms * 1e-9);
    return 0;
}
