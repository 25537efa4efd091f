// Weight gradient of a Linear on the tensor cores:  C[m][n] (+)= sum_e G[e][m] * X[e][n]   (gW = gY^T X, training step).
//
// Both operands are read straight from the row-major activation matrices: for a fixed reduction index e the M (resp. N)
// values are contiguous, i.e. both are "MN-major" tcgen05 operands.  For 32-bit elements that is the SWIZZLE_128B_BASE32B
// layout (probed on B200: plain SWIZZLE_128B reads nothing for MN-major TF32): a 32-wide MN block of 4 consecutive e is one
// 512-byte atom whose 128-byte row k holds G[e0+k][32 blk .. 32 blk + 31], with the 32-byte chunk index XOR-ed with k;
// K groups are SBO = 512 bytes apart, MN blocks LBO apart; one K = 8 MMA step spans two atoms.  16 worker warps load a [16 x H] slab of G and of X per stage with coalesced
// float4 loads, split every value into its TF32 hi / lo parts (error-compensated 3xTF32, as everywhere else) and store
// them at the swizzled positions; one thread issues the MMAs (M = 128 x 2 row blocks, N = H rounded to 16, K = 8 per
// instruction) into two TMEM accumulators.  The reduction dimension is split across CTAs; partial results meet in C with
// fp32 atomics (C is zeroed by the launcher unless accumulating).
#include "tc_common.cuh"
#include "kernels.h"

namespace gb {
using namespace tc;

constexpr int WG_KT = 16;                          // reduction rows per stage (two K = 8 MMA steps)
constexpr int WG_STAGES = 3;
constexpr int WG_BLK = (WG_KT / 8) * 1024;         // bytes of one 32-wide MN block in a stage
constexpr int WG_NBLK = 8;                         // operands padded to 256 columns
constexpr int WG_OP = WG_NBLK * WG_BLK;            // bytes of one operand image (hi or lo) in a stage: 16 KB
constexpr int WG_STAGE = 4 * WG_OP;                // A_hi, A_lo, B_hi, B_lo
constexpr int WG_WORK = 512;
constexpr int WG_THREADS = 64 + WG_WORK;
constexpr int WG_SMEM = WG_STAGES * WG_STAGE + 1024 + 256;

__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)(WG_BLK >> 4) << 16;               // LBO: byte offset between consecutive 32-wide MN blocks
    d |= (uint64_t)(512 >> 4) << 32;                  // SBO: byte offset between consecutive 4-deep K groups
    d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
    d |= (uint64_t)1 << 61;                           // SWIZZLE_128B_BASE32B
    return d;
}
__device__ __forceinline__ uint32_t idesc_tf32_mn(int N) {      // TF32 x TF32 -> F32, A and B MN-major, M = 128
    return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
}

__global__ void __launch_bounds__(WG_THREADS, 1) tc_wgrad_kernel(int K, int M, int N, const float* __restrict__ G, int ldg,
                                                                const float* __restrict__ X, int ldx, float* __restrict__ C, int ldc,
                                                                int kt_per_cta, float* __restrict__ partial, int with_colsum) {
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array: the compiler keeps the address space (LDS / STS
    // instead of generic LD / ST for every staging and operand access)
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(base + WG_STAGES * WG_STAGE);
    uint64_t* empty = full + WG_STAGES;
    uint64_t* d_full = empty + WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_full + 1);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full[s], WG_WORK); mbar_init(&empty[s], 1); }
        mbar_init(d_full, 1);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    // zero the operand images once: padding columns / blocks are never written afterwards
    for (int i = tid; i < WG_STAGES * WG_STAGE / 16; i += blockDim.x) reinterpret_cast<float4*>(base)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    fence_proxy_async();
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int n_kt = (K + WG_KT - 1) / WG_KT;
    const int kt_lo = blockIdx.x * kt_per_cta, kt_hi = min(n_kt, kt_lo + kt_per_cta);
    // with_colsum: a column of ones is appended to X (column N of the operand image), so accumulator column N = sum_e G[e][m],
    // the bias gradient of the same Linear, for free (two-phase path only)
    const int NP = (N + (with_colsum ? 1 : 0) + 15) & ~15, n_mblk = (M + 127) / 128;

    if (warp == 1) {
        if (lane == 0 && kt_lo < kt_hi) {
            const uint32_t idesc = idesc_tf32_mn(NP);
            uint32_t it = 0;
            for (int kt = kt_lo; kt < kt_hi; ++kt, ++it) {
                const uint32_t s = it % WG_STAGES, r = it / WG_STAGES;
                mbar_wait(&full[s], r & 1);
                fence_after_sync();
                const uint32_t a_hi = smem_u32(base + s * WG_STAGE), a_lo = a_hi + WG_OP, b_hi = a_hi + 2 * WG_OP, b_lo = a_hi + 3 * WG_OP;
                for (int ks = 0; ks < WG_KT / 8; ++ks) {
                    const uint32_t acc = (it | ks) != 0;
                    for (int mb = 0; mb < n_mblk; ++mb) {
                        const uint32_t ao = mb * 4 * WG_BLK + ks * 1024, bo = ks * 1024, d = tmem_base + mb * 256;
                        mma_tf32(d, smem_desc_mn(a_lo + ao), smem_desc_mn(b_hi + bo), idesc, acc);
                        mma_tf32(d, smem_desc_mn(a_hi + ao), smem_desc_mn(b_lo + bo), idesc, 1);
                        mma_tf32(d, smem_desc_mn(a_hi + ao), smem_desc_mn(b_hi + bo), idesc, 1);
                    }
                }
                mma_commit(&empty[s]);
            }
            mma_commit(d_full);
        }
    } else if (warp >= 2) {
        const int w = tid - 64;                                   // 0 .. 511
        const int m4 = M >> 2, n4 = N >> 2, per_row = m4 + n4;    // float4 per reduction row: G part then X part
        uint32_t it = 0;
        for (int kt = kt_lo; kt < kt_hi; ++kt, ++it) {
            const uint32_t s = it % WG_STAGES, r = it / WG_STAGES;
            unsigned char* st = base + s * WG_STAGE;
            if (r > 0) mbar_wait(&empty[s], (r - 1) & 1);
            for (int i = w; i < WG_KT * per_row; i += WG_WORK) {
                const int k = i / per_row, c = i - k * per_row;
                const int e = kt * WG_KT + k;
                const bool isx = c >= m4;
                const int col = (isx ? c - m4 : c) << 2;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (e < K) v = __ldg(reinterpret_cast<const float4*>((isx ? X + (size_t)e * ldx : G + (size_t)e * ldg) + col));
                float4 h, l;
                h.x = tf32_hi(v.x); h.y = tf32_hi(v.y); h.z = tf32_hi(v.z); h.w = tf32_hi(v.w);
                l.x = v.x - h.x; l.y = v.y - h.y; l.z = v.z - h.z; l.w = v.w - h.w;
                const uint32_t off = (uint32_t)((col >> 5) * WG_BLK + (k >> 2) * 512 + (k & 3) * 128 + ((((col & 31) >> 3) ^ (k & 3)) << 5) + ((col & 7) << 2));
                unsigned char* op = st + (isx ? 2 * WG_OP : 0);
                *reinterpret_cast<float4*>(op + off) = h;
                *reinterpret_cast<float4*>(op + WG_OP + off) = l;
            }
            if (with_colsum && w < WG_KT) {                       // hi = 1, lo = 0 (pre-zeroed) for the rows that exist
                const int k = w, col = N;
                const float one = (kt * WG_KT + k < K) ? 1.f : 0.f;
                const uint32_t off = (uint32_t)((col >> 5) * WG_BLK + (k >> 2) * 512 + (k & 3) * 128 + ((((col & 31) >> 3) ^ (k & 3)) << 5) + ((col & 7) << 2));
                *reinterpret_cast<float*>(st + 2 * WG_OP + off) = one;
            }
            fence_proxy_async();
            mbar_arrive(&full[s]);
        }
        // ---- epilogue: accumulators -> C (atomics: the reduction dimension is split across CTAs) ----
        if (kt_lo < kt_hi) {
            mbar_wait(d_full, 0);
            fence_after_sync();
            const int group = warp & 3, part = (warp - 2) >> 2;
            const int nchunks = NP >> 4;
            for (int mb = 0; mb < n_mblk; ++mb) {
                const int m = mb * 128 + group * 32 + lane;
                const uint32_t taddr = tmem_base + ((uint32_t)(group * 32) << 16) + mb * 256;
                for (int ch = part; ch < nchunks; ch += 4) {
                    float v[16];
                    tmem_ld16(taddr + ch * 16, v);
                    if (m < M) {
                        if (partial) {                    // deterministic two-phase reduction: this CTA's slab [M][NP]
                            float4* prow = reinterpret_cast<float4*>(partial + ((size_t)blockIdx.x * M + m) * NP + ch * 16);
#pragma unroll
                            for (int j = 0; j < 4; ++j) prow[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                        } else {
                            float* crow = C + (size_t)m * ldc + ch * 16;
#pragma unroll
                            for (int j = 0; j < 16; ++j)
                                if (ch * 16 + j < N) atomicAdd(crow + j, v[j]);
                        }
                    }
                }
            }
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// C[m][n] = (accumulate ? C : 0) + sum over the CTA slabs, fixed order; column N of the slabs (if present) -> colsum[m]
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int n_slabs, int M, int N, int NP, float* __restrict__ C, int ldc,
                                    int accumulate, float* __restrict__ colsum) {
    const int ncol = N + (colsum ? 1 : 0);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M * ncol; i += gridDim.x * blockDim.x) {
        const int m = i / ncol, n = i - m * ncol;
        float s0 = 0.f, s1 = 0.f;
        int g = 0;
        for (; g + 1 < n_slabs; g += 2) {
            s0 += partial[((size_t)g * M + m) * NP + n];
            s1 += partial[((size_t)(g + 1) * M + m) * NP + n];
        }
        if (g < n_slabs) s0 += partial[((size_t)g * M + m) * NP + n];
        if (n == N) { colsum[m] = s0 + s1; continue; }
        float* c = C + (size_t)m * ldc + n;
        *c = (accumulate ? *c : 0.f) + (s0 + s1);
    }
}

size_t wgrad_tc_scratch_bytes(int M, int N) { return (size_t)148 * M * ((N + 1 + 15) & ~15) * sizeof(float); }

// scratch (>= wgrad_tc_scratch_bytes) selects the deterministic two-phase reduction; without it C must hold the value to add
// to (zero or the accumulation target) and the CTAs add their partial sums with fp32 atomics
void launch_wgrad_tc(int K, int M, int N, const float* G, int ldg, const float* X, int ldx, float* C, int ldc, int accumulate,
                     float* scratch, float* colsum, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(tc_wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM);
        configured = true;
    }
    const int n_kt = (K + WG_KT - 1) / WG_KT;
    int grid = n_kt / 8;                                  // at least 128 reduction rows per CTA
    if (grid < 1) grid = 1;
    if (grid > 148) grid = 148;
    const int per = (n_kt + grid - 1) / grid;
    grid = (n_kt + per - 1) / per;
    const int wc = (scratch && colsum && N < 256) ? 1 : 0;
    tc_wgrad_kernel<<<grid, WG_THREADS, WG_SMEM, s>>>(K, M, N, G, ldg, X, ldx, C, ldc, per, scratch, wc);
    if (scratch) {
        const int total = M * (N + wc);
        wgrad_reduce_kernel<<<(total + 255) / 256, 256, 0, s>>>(scratch, grid, M, N, (N + wc + 15) & ~15, C, ldc, accumulate, wc ? colsum : nullptr);
    }
}

}  // namespace gb
