// Tensor-core (tcgen05, 3xTF32) version of the node-level fused Linear kernel -- same contract as lin_kernel.cu.
//
// Warp roles (one persistent CTA per SM, 576 threads):
//   warp 0      TMA producer: streams the pre-swizzled hi/lo weight K-atoms (one 1-D bulk copy per atom)
//   warp 1      MMA issuer: 3 tcgen05.mma (lo*hi, hi*lo, hi*hi) per 8-wide K step, accumulator [128 x NP] fp32 in TMEM
//   warps 2-17  workers, thread = tile row (four parts of 4 warps, each part covers all 128 TMEM lanes; part p owns the
//               16-column chunks ch == p mod 4): build the A K-atoms (global row -> hi/lo split -> 128B-swizzled smem),
//               then run the epilogue on their chunks read back with tcgen05.ld.
#include "tc_common.cuh"
#include "kernels.h"

#ifndef GB_LIN_NPARTS
#define GB_LIN_NPARTS 4     // worker parts of 4 warps
#define GB_LIN_S 2          // operand ring stages
#define GB_LIN_CTAS 1       // CTAs per SM
#endif

namespace gb {
using namespace tc;

template <int NP>
struct TcLinCfg {
    static constexpr int S = GB_LIN_S;
    static constexpr int A_BYTES = 128 * ATOM_ROW_BYTES;
    static constexpr int W_BYTES = NP * ATOM_ROW_BYTES;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
    static constexpr int SMEM = S * STAGE_BYTES + 1024 + 256;
    static constexpr int TMEM_COLS = 256;
    static constexpr int NPARTS = GB_LIN_NPARTS;                          // worker parts of 4 warps; part p owns 16-column chunks ch == p (mod 4)
    static constexpr int NWORK = 128 * NPARTS;
    static constexpr int THREADS = 64 + NWORK;
};

template <int NP>
__global__ void __launch_bounds__(TcLinCfg<NP>::THREADS, GB_LIN_CTAS) tc_lin_kernel(LinArgs a, const float* __restrict__ wimg, int H) {
    using CF = TcLinCfg<NP>;
    extern __shared__ unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + CF::S * CF::STAGE_BYTES);
    uint64_t* full_a = bars; uint64_t* full_w = bars + CF::S; uint64_t* empty = bars + 2 * CF::S;
    uint64_t* d_full = bars + 3 * CF::S; uint64_t* d_empty = d_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < CF::S; ++s) { mbar_init(&full_a[s], 256); mbar_init(&full_w[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(d_full, 1); mbar_init(d_empty, CF::NWORK);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<CF::TMEM_COLS>(tmem_slot);
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int n_tiles = (a.M + 127) / 128;
    const int cb = blockIdx.y;
    const int na1 = (a.K1 + ATOM_K - 1) / ATOM_K, na2 = (a.K2 + ATOM_K - 1) / ATOM_K, na = na1 + na2;
    const size_t atom_floats = (size_t)2 * NP * ATOM_K;
    const float* wcb = wimg + (size_t)cb * na * atom_floats;
    constexpr uint32_t idesc = instr_desc_tf32(NP);

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x)
                for (int j = 0; j < na; ++j, ++it) {
                    const uint32_t s = it % CF::S, r = it / CF::S;
                    if (r > 0) mbar_wait(&empty[s], (r - 1) & 1);
                    mbar_arrive_expect_tx(&full_w[s], 2 * CF::W_BYTES);
                    bulk_g2s(base + s * CF::STAGE_BYTES + 2 * CF::A_BYTES, wcb + (size_t)j * atom_floats, 2 * CF::W_BYTES, &full_w[s]);
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, tcnt = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcnt) {
                if (tcnt > 0) mbar_wait(d_empty, (tcnt - 1) & 1);
                fence_after_sync();
                for (int j = 0; j < na; ++j, ++it) {
                    const uint32_t s = it % CF::S, r = it / CF::S;
                    const int kvalid = (j < na1 ? a.K1 - j * ATOM_K : a.K2 - (j - na1) * ATOM_K);
                    const int ksteps = kvalid >= ATOM_K ? 4 : (kvalid + 7) / 8;
                    mbar_wait(&full_a[s], r & 1);
                    mbar_wait(&full_w[s], r & 1);
                    fence_after_sync();
                    const uint32_t a_hi = smem_u32(base + s * CF::STAGE_BYTES), a_lo = a_hi + CF::A_BYTES;
                    const uint32_t w_hi = a_hi + 2 * CF::A_BYTES, w_lo = w_hi + CF::W_BYTES;
                    for (int kk = 0; kk < ksteps; ++kk) {
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
    } else {
        const int group = warp & 3, part = (warp - 2) >> 2, half = part & 1;
        const int r = group * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(group * 32) << 16);
        const int nchunks = (H + 15) / 16;
        uint32_t tcnt = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcnt) {
            const int row = tile * 128 + r;
            const bool rvalid = row < a.M;
            const float rs = (rvalid && a.rowscale) ? __ldg(a.rowscale + row) : 1.f;
            // ---- build the A atoms owned by this half ----
            for (int j = part >> 1; j < na; j += CF::NPARTS / 2) {
                const uint32_t it = tcnt * na + j;
                const uint32_t s = it % CF::S, rr = it / CF::S;
                const bool first = j < na1;
                const float* src = first ? a.A1 + (size_t)row * a.lda1 + j * ATOM_K : a.A2 + (size_t)row * a.lda2 + (j - na1) * ATOM_K;
                const int kvalid = first ? a.K1 - j * ATOM_K : a.K2 - (j - na1) * ATOM_K;
                float4 x[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int kc = 16 * half + 4 * c;
                    x[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rvalid && kc < kvalid) {
                        x[c] = __ldg(reinterpret_cast<const float4*>(src + kc));
                        x[c].x *= rs; x[c].y *= rs; x[c].z *= rs; x[c].w *= rs;
                    }
                }
                if (rr > 0) mbar_wait(&empty[s], (rr - 1) & 1);
                unsigned char* a_hi = base + s * CF::STAGE_BYTES;
#pragma unroll
                for (int c = 0; c < 4; ++c) store_split(a_hi, a_hi + CF::A_BYTES, r, 4 * half + c, x[c]);
                fence_proxy_async();
                mbar_arrive(&full_a[s]);
            }
            // ---- epilogue on alternating 16-column chunks ----
            mbar_wait(d_full, tcnt & 1);
            fence_after_sync();
            float mk = 1.f;
            if (rvalid && (a.epi == EPI_RES_MASK || (a.epi == EPI_ADD_RES && a.mask))) mk = __ldg(a.mask + row);
            const bool use_res = a.epi == EPI_RES_MASK || (a.epi == EPI_ADD_RES && (a.res_cb < 0 || cb == a.res_cb));
            for (int ch = part; ch < nchunks; ch += CF::NPARTS) {
                float v[16];
                tmem_ld16(lane_addr + ch * 16, v);
                if (!rvalid) continue;
                const int c0 = ch * 16;
                const size_t col0 = (size_t)cb * H + c0;         // column blocks are H wide in the node tensors
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (c0 + 4 * q >= H) break;
                    float o[4] = {v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]};
                    if (a.bias) {
                        const float4 b = __ldg(reinterpret_cast<const float4*>(a.bias + col0 + 4 * q));
                        o[0] += b.x; o[1] += b.y; o[2] += b.z; o[3] += b.w;
                    }
                    if (a.epi == EPI_SILU) {
                        if (a.out2) *reinterpret_cast<float4*>(a.out2 + (size_t)row * a.ldo2 + col0 + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) o[e] = silu_f(o[e]);
                    } else if (a.epi == EPI_MUL_DSILU) {
                        const float4 p = __ldg(reinterpret_cast<const float4*>(a.aux + (size_t)row * a.ldaux + col0 + 4 * q));
                        o[0] *= dsilu_f(p.x); o[1] *= dsilu_f(p.y); o[2] *= dsilu_f(p.z); o[3] *= dsilu_f(p.w);
                    } else if (use_res) {
                        const float4 rsd = __ldg(reinterpret_cast<const float4*>(a.res + (size_t)row * a.ldr + col0 + 4 * q));
                        if (a.epi == EPI_RES_MASK) { o[0] = (rsd.x + o[0]) * mk; o[1] = (rsd.y + o[1]) * mk; o[2] = (rsd.z + o[2]) * mk; o[3] = (rsd.w + o[3]) * mk; }
                        else { o[0] += rsd.x * mk; o[1] += rsd.y * mk; o[2] += rsd.z * mk; o[3] += rsd.w * mk; }
                    }
                    *reinterpret_cast<float4*>(a.out + (size_t)row * a.ldo + col0 + 4 * q) = make_float4(o[0], o[1], o[2], o[3]);
                }
            }
            fence_before_sync();
            mbar_arrive(d_empty);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc<CF::TMEM_COLS>(tmem_base);
}

template <int NP>
static void launch_t(const LinArgs& a, const float* wimg, int H, cudaStream_t s) {
    using CF = TcLinCfg<NP>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(tc_lin_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM);
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_tiles = (a.M + 127) / 128;
    int gx = n_tiles;
    const int cap = max(1, sms * GB_LIN_CTAS / a.ncb);
    if (gx > cap) gx = cap;
    tc_lin_kernel<NP><<<dim3(gx, a.ncb), CF::THREADS, CF::SMEM, s>>>(a, wimg, H);
}

int tc_np(int H) { return H <= 64 ? 64 : (H <= 192 ? 192 : (H <= 208 ? 208 : 256)); }

void launch_lin_tc(int H, const LinArgs& a, const float* wimg, cudaStream_t s) {
    if (a.M <= 0) return;
    switch (tc_np(H)) {
        case 64: launch_t<64>(a, wimg, H, s); break;
        case 192: launch_t<192>(a, wimg, H, s); break;
        case 208: launch_t<208>(a, wimg, H, s); break;
        default: launch_t<256>(a, wimg, H, s); break;
    }
}

// Weight image for the tensor-core path: [atom j][hi | lo][NP rows][32] with the 128B swizzle applied, so that one
// contiguous bulk copy lands a ready-to-use B operand.  value(n, k) = transpose ? W[(k+k_off)*ld + n_off+n]
//                                                                                : W[(n+n_off)*ld + k_off+k]
__global__ void pack_tc_kernel(float* dst, const float* src, int ld, int k_off, int n_off, int Kv, int Nv, int NP, int atoms, int transpose) {
    const int total = atoms * NP * ATOM_K;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int j = idx / (NP * ATOM_K), rem = idx % (NP * ATOM_K);
        const int n = rem / ATOM_K, kk = rem % ATOM_K, k = j * ATOM_K + kk;
        float w = 0.f;
        if (n < Nv && k < Kv) w = transpose ? src[(size_t)(k + k_off) * ld + n_off + n] : src[(size_t)(n + n_off) * ld + k_off + k];
        const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
        const int c = kk >> 2, e = kk & 3;
        const size_t off = (size_t)(n >> 3) * 256 + (n & 7) * 32 + ((c ^ (n & 7)) << 2) + e;
        dst[((size_t)j * 2 + 0) * NP * ATOM_K + off] = hi;
        dst[((size_t)j * 2 + 1) * NP * ATOM_K + off] = w - hi;
    }
}

void launch_pack_tc(float* dst, const float* src, int ld, int k_off, int n_off, int Kv, int Nv, int NP, int atoms, int transpose, cudaStream_t s) {
    const int total = atoms * NP * ATOM_K;
    pack_tc_kernel<<<(total + 255) / 256, 256, 0, s>>>(dst, src, ld, k_off, n_off, Kv, Nv, NP, atoms, transpose);
}

}  // namespace gb
