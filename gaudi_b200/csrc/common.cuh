// Shared device helpers for the gaudi_b200 kernels (sm_100a only).
//
// Tile GEMM engine used by every node-level and edge-level kernel:
//   C[128 x HP] (+)= A[128 x K] * Wt[K x HP]
// * A lives in shared memory TRANSPOSED: A_s[k * GB_MS + m]  (k-major, 128 rows + 4 pad) so that a
//   lane reads its 4 consecutive rows with one conflict-free LDS.128.
// * Wt (pre-packed, [K][HP] fp32, L2-resident) is streamed through a GB_STAGES-deep shared-memory
//   ring by ONE producer warp with 1-D bulk TMA copies (cp.async.bulk -> UBLKCP) completing on
//   mbarriers; NW consumer warps each own a CW = HP/NW column slab and all 128 rows
//   (lane l -> rows 4l..4l+3), i.e. 4 x CW FP32 accumulators per thread.
// * FP32 FFMA throughout (the "fp32 mode" of BASELINE.json: tolerance 1e-4 per step).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

#define GB_TM 128        // rows (edges or nodes) per tile
#define GB_MS 132        // smem stride of one k-row of the A tile (floats)
#define GB_KC 16         // k-rows of Wt per pipeline stage
#define GB_STAGES 4
#define GB_FFMA2 1       // packed fma.rn.f32x2 in the tile GEMM (halves FP32 issue slots; same IEEE fma per lane)

namespace gb {

// ----------------------------------------------------------------------------------------------
// math
// ----------------------------------------------------------------------------------------------
// sigmoid(x) = 1 / (1 + 2^(-x log2 e)): one MUFU.EX2 and one MUFU.RCP.  The denominator lies in [1, inf], so the bare
// rcp.approx (1 ulp, like __fdividef) needs none of __fdividef's range checks and fix-ups (an FSETP and predicated
// multiplies per element, which showed up in the SASS of every edge kernel).
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float exp_neg(float x) {          // e^-x; flush-to-zero ex2 (no denormal scaling code around the MUFU)
    float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x * -1.4426950408889634f)); return r;
}
__device__ __forceinline__ float sigmoid_f(float x) { return rcp_approx(1.f + exp_neg(x)); }
__device__ __forceinline__ float silu_f(float x) { return x * rcp_approx(1.f + exp_neg(x)); }
// d/dx [x * sigmoid(x)]
__device__ __forceinline__ float dsilu_f(float x) {
    float s = sigmoid_f(x);
    return s * (1.f + x * (1.f - s));
}
__device__ __forceinline__ void silu_both(float x, float& y, float& dy) {
    float s = sigmoid_f(x);
    y = x * s;
    dy = s * (1.f + x * (1.f - s));
}

// ---- packed FP32 (sm_100 FFMA2 / FMUL2 / FADD2: two IEEE fp32 operations per issue slot) versions of the activation math.
// The worker warps of the edge kernels are issue-bound (clock64 timelines, round 2): halving the FMUL / FADD / FFMA count of
// the element-wise phases is worth more than anything else there.  The MUFU ops stay scalar (there is no packed MUFU).
typedef float2 f2;
__device__ __forceinline__ f2 f2s(float s) { return make_float2(s, s); }
__device__ __forceinline__ f2 lo2(const float4& v) { return make_float2(v.x, v.y); }
__device__ __forceinline__ f2 hi2(const float4& v) { return make_float2(v.z, v.w); }
__device__ __forceinline__ float4 cat2(f2 a, f2 b) { return make_float4(a.x, a.y, b.x, b.y); }
__device__ __forceinline__ f2 mul2(f2 a, f2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ f2 add2(f2 a, f2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ f2 fma2(f2 a, f2 b, f2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ f2 ex2_2(f2 x) {
    f2 r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(x.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(x.y));
    return r;
}
__device__ __forceinline__ f2 rcp_2(f2 x) {
    f2 r;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(x.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(x.y));
    return r;
}
__device__ __forceinline__ f2 sigmoid2(f2 x) { return rcp_2(add2(ex2_2(mul2(x, f2s(-1.4426950408889634f))), f2s(1.f))); }
__device__ __forceinline__ f2 silu2(f2 x) { return mul2(x, sigmoid2(x)); }
__device__ __forceinline__ f2 dsilu2(f2 x) {                    // s * (1 + x * (1 - s))
    const f2 s = sigmoid2(x);
    return mul2(s, fma2(x, fma2(s, f2s(-1.f), f2s(1.f)), f2s(1.f)));
}
__device__ __forceinline__ void silu_both2(f2 x, f2& y, f2& dy) {
    const f2 s = sigmoid2(x);
    y = mul2(x, s);
    dy = mul2(s, fma2(x, fma2(s, f2s(-1.f), f2s(1.f)), f2s(1.f)));
}

// packed FP32 FMA (sm_100+ FFMA2): (d0,d1) += (a,a) * (b0,b1) in one issue slot
__device__ __forceinline__ void ffma2(float& d0, float& d1, float a, float b0, float b1) {
    asm("{\n\t.reg .b64 ra, rb, rc;\n\t"
        "mov.b64 ra, {%2, %2};\n\t"
        "mov.b64 rb, {%3, %4};\n\t"
        "mov.b64 rc, {%0, %1};\n\t"
        "fma.rn.f32x2 rc, ra, rb, rc;\n\t"
        "mov.b64 {%0, %1}, rc;\n\t}"
        : "+f"(d0), "+f"(d1)
        : "f"(a), "f"(b0), "f"(b1));
}

// ----------------------------------------------------------------------------------------------
// mbarrier + bulk-copy PTX
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
#ifdef GB_DEBUG_HANG
// development aid: a wait that gives up after a long spin, records (block, thread, barrier smem offset, parity) in a host-mapped
// buffer (gb_debug_set_hang_buf) and traps, so that a deadlock turns into a readable report instead of a hung GPU
__device__ int* gb_hang_buf;
#endif
#ifdef GB_DEBUG_HANG
__device__ int gb_dbg_tag_unused;
#define GB_TAG(x) , (x)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity, int tag = -1) {
#else
#define GB_TAG(x)
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#endif
    uint32_t done;
#ifdef GB_DEBUG_HANG
    for (long long spin = 0;; ++spin) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
        if (spin > 30000) {
            volatile int* hb = gb_hang_buf;
            if (hb && (threadIdx.x & 31) == 0 && blockIdx.x < 24) {
                const int i = atomicAdd((int*)hb, 1);
                if (i < 1000) { hb[4 + 4 * i] = blockIdx.x; hb[5 + 4 * i] = threadIdx.x; hb[6 + 4 * i] = (int)smem_u32(bar); hb[7 + 4 * i] = (int)parity | (tag << 4); }
                __threadfence_system();
            }
            for (int w = 0; w < 3000000; ++w) __nanosleep(1000);
            __trap();
        }
    }
#endif
    // try_wait suspends the thread until the phase completes or a time limit passes; without the hint operand the limit is short and a
    // waiting warp keeps re-issuing SYNCS + BRA + YIELD (25 % of the executed warp-instructions of the round-2c edge kernels were such
    // polls, competing with the worker warps for issue slots).  GB_MBAR_HINT_NS = 0 restores the plain form.
#ifndef GB_MBAR_HINT_NS
#define GB_MBAR_HINT_NS 1000000
#endif
    do {
#if GB_MBAR_HINT_NS > 0
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity), "r"(GB_MBAR_HINT_NS)
            : "memory");
#else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
#endif
    } while (!done);
}
// 1-D bulk async copy global -> shared (TMA engine, SASS UBLKCP); completes `bytes` on `bar`.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void consumer_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// ----------------------------------------------------------------------------------------------
// weight-chunk pipeline (producer warp -> NW consumer warps)
// ----------------------------------------------------------------------------------------------
template <int HP>
struct WPipe {
    float* ring;        // [GB_STAGES][GB_KC][HP]
    uint64_t* full;     // [GB_STAGES]
    uint64_t* empty;    // [GB_STAGES]
    uint32_t cnt;       // running chunk counter (same sequence on both sides)

    __device__ __forceinline__ void init_side(float* r, uint64_t* f, uint64_t* e) {
        ring = r; full = f; empty = e; cnt = 0;
    }
    // ---- producer (one elected lane) ----
    __device__ __forceinline__ void produce(const float* __restrict__ wt, int K) {
        for (int k0 = 0; k0 < K; k0 += GB_KC) {
            const int rows = min(GB_KC, K - k0);
            const uint32_t slot = cnt % GB_STAGES, round = cnt / GB_STAGES;
            if (round > 0) mbar_wait(&empty[slot], (round - 1) & 1);
            const uint32_t bytes = (uint32_t)rows * HP * 4u;
            mbar_arrive_expect_tx(&full[slot], bytes);
            bulk_g2s(ring + (size_t)slot * GB_KC * HP, wt + (size_t)k0 * HP, bytes, &full[slot]);
            ++cnt;
        }
    }
};

// Consumer side of one GEMM call: acc[4][CW] += A_s[0:K, rows of this lane] * Wt[0:K, slab of this warp].
template <int HP, int NW>
__device__ __forceinline__ void gemm_consume(const float* __restrict__ A_s, int K, float (&acc)[4][HP / NW],
                                             WPipe<HP>& pipe, int warp, int lane) {
    constexpr int CW = HP / NW;
    static_assert(CW % 4 == 0, "column slab must be float4 aligned");
    const float* a_base = A_s + 4 * lane;
    for (int k0 = 0; k0 < K; k0 += GB_KC) {
        const int rows = min(GB_KC, K - k0);
        const uint32_t slot = pipe.cnt % GB_STAGES, round = pipe.cnt / GB_STAGES;
        mbar_wait(&pipe.full[slot], round & 1);
        const float* w_base = pipe.ring + (size_t)slot * GB_KC * HP + warp * CW;
        if (rows == GB_KC) {
#ifndef GB_GEMM_UNROLL
#define GB_GEMM_UNROLL 16
#endif
            constexpr int kUnroll = GB_GEMM_UNROLL;
#pragma unroll kUnroll
            for (int kk = 0; kk < GB_KC; ++kk) {
                const float4 a = *reinterpret_cast<const float4*>(a_base + (k0 + kk) * GB_MS);
                float b[CW];
#pragma unroll
                for (int c = 0; c < CW; c += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(w_base + kk * HP + c);
                    b[c] = t.x; b[c + 1] = t.y; b[c + 2] = t.z; b[c + 3] = t.w;
                }
#ifdef GB_FFMA2
#pragma unroll
                for (int c = 0; c < CW; c += 2) {
                    ffma2(acc[0][c], acc[0][c + 1], a.x, b[c], b[c + 1]);
                    ffma2(acc[1][c], acc[1][c + 1], a.y, b[c], b[c + 1]);
                    ffma2(acc[2][c], acc[2][c + 1], a.z, b[c], b[c + 1]);
                    ffma2(acc[3][c], acc[3][c + 1], a.w, b[c], b[c + 1]);
                }
#else
#pragma unroll
                for (int c = 0; c < CW; ++c) {
                    acc[0][c] = fmaf(a.x, b[c], acc[0][c]);
                    acc[1][c] = fmaf(a.y, b[c], acc[1][c]);
                    acc[2][c] = fmaf(a.z, b[c], acc[2][c]);
                    acc[3][c] = fmaf(a.w, b[c], acc[3][c]);
                }
#endif
            }
        } else {
            for (int kk = 0; kk < rows; ++kk) {
                const float4 a = *reinterpret_cast<const float4*>(a_base + (k0 + kk) * GB_MS);
#pragma unroll
                for (int c = 0; c < CW; c += 4) {
                    const float4 t = *reinterpret_cast<const float4*>(w_base + kk * HP + c);
                    acc[0][c] = fmaf(a.x, t.x, acc[0][c]); acc[0][c + 1] = fmaf(a.x, t.y, acc[0][c + 1]);
                    acc[0][c + 2] = fmaf(a.x, t.z, acc[0][c + 2]); acc[0][c + 3] = fmaf(a.x, t.w, acc[0][c + 3]);
                    acc[1][c] = fmaf(a.y, t.x, acc[1][c]); acc[1][c + 1] = fmaf(a.y, t.y, acc[1][c + 1]);
                    acc[1][c + 2] = fmaf(a.y, t.z, acc[1][c + 2]); acc[1][c + 3] = fmaf(a.y, t.w, acc[1][c + 3]);
                    acc[2][c] = fmaf(a.z, t.x, acc[2][c]); acc[2][c + 1] = fmaf(a.z, t.y, acc[2][c + 1]);
                    acc[2][c + 2] = fmaf(a.z, t.z, acc[2][c + 2]); acc[2][c + 3] = fmaf(a.z, t.w, acc[2][c + 3]);
                    acc[3][c] = fmaf(a.w, t.x, acc[3][c]); acc[3][c + 1] = fmaf(a.w, t.y, acc[3][c + 1]);
                    acc[3][c + 2] = fmaf(a.w, t.z, acc[3][c + 2]); acc[3][c + 3] = fmaf(a.w, t.w, acc[3][c + 3]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(&pipe.empty[slot]);
        ++pipe.cnt;
    }
}

template <int CW>
__device__ __forceinline__ void zero_acc(float (&acc)[4][CW]) {
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < CW; ++c) acc[r][c] = 0.f;
}

// Shared-memory carve-up common to all tile kernels.
template <int HP>
struct SmemLayout {
    static constexpr size_t a_bytes = (size_t)HP * GB_MS * 4;
    static constexpr size_t ring_bytes = (size_t)GB_STAGES * GB_KC * HP * 4;
    static constexpr size_t bar_bytes = 2 * GB_STAGES * 8;
    static constexpr size_t base_bytes = a_bytes + ring_bytes + bar_bytes;
};

// Per-template-config constants
template <int HP> struct TileCfg;
template <> struct TileCfg<64>  { static constexpr int NW = 4; };
template <> struct TileCfg<128> { static constexpr int NW = 8; };
template <> struct TileCfg<192> { static constexpr int NW = 8; };
template <> struct TileCfg<196> { static constexpr int NW = 7; };
template <> struct TileCfg<256> { static constexpr int NW = 8; };

}  // namespace gb
