// Tensor-core (tcgen05, 3xTF32) version of the node-level fused Linear kernel -- same contract as lin_kernel.cu.
//
// Warp roles (one persistent CTA per SM, 576 threads), chosen from a clock64 timeline of the first version (thread = row for
// everything, one accumulator): per 128-row tile the MMAs take 4.7 us but the tile took 13.7 us, because the same threads
// first waited for their A rows (2.3 us), then for the MMAs, then ran a 7 us epilogue whose 16-byte-per-row accesses cost
// 32 L1 tag lookups per warp instruction.  Now the three phases overlap:
//   warp 0       TMA producer: streams the pre-swizzled hi/lo weight K-atoms (one 1-D bulk copy per atom)
//   warp 1       MMA issuer: 3 tcgen05.mma (lo*hi, hi*lo, hi*hi) per 8-wide K step into one of TWO [128 x NP] fp32 accumulators
//   warps 2-9    A builders: 8 consecutive lanes read the 128 bytes of one row of a K-atom (4 full lines per warp load), split
//                hi/lo and store to the 128B-swizzled image; the loads of the next two atoms (also across tiles) are in flight
//                while the current one is stored, so the builders run ahead of the MMA warp, throttled only by the ring
//   warps 10-17  epilogue: two parts of 4 warps (a part covers the 128 TMEM lanes; part p owns the 16-column chunks
//                ch == p mod 2).  tcgen05.ld gives thread = row; the chunk is transposed through a per-warp shared-memory
//                block so that bias / side input / output accesses are 8 rows x 64 contiguous bytes per warp instruction.
//                The epilogue of tile t runs while the MMAs of tile t+1 fill the other accumulator.
#include "tc_common.cuh"
#include "kernels.h"

#ifndef GB_LIN_SPLIT_LD
#define GB_LIN_SPLIT_LD 0     // 1: TMEM read of the next chunk in flight across a chunk step (unsafe: see tc_pred_edge.cu, GB_BWD_SPLIT_LD)
#endif
#ifndef GB_LIN_S
#define GB_LIN_S 2          // operand ring stages
#define GB_LIN_CTAS 1       // CTAs per SM
#endif

namespace gb {
using namespace tc;

// Optional per-tile timeline (make EXTRA=-DGB_TIMELINE; tools/lin_timeline.py): clock64 marks of CTA 0 per role.
#ifdef GB_TIMELINE
__device__ unsigned long long gb_tl_dur[512];             // per CTA: clock64 from kernel entry to exit
__device__ unsigned long long gb_tl[3][1024];
#define TL(role, idx, code) do { if (blockIdx.x == 0 && blockIdx.y == 0 && (idx) < 512) { gb_tl[role][2 * (idx)] = (code); gb_tl[role][2 * (idx) + 1] = clock64(); } } while (0)
#else
#define TL(role, idx, code) do {} while (0)
#endif

template <int NP>
struct TcLinCfg {
    static constexpr int S = GB_LIN_S;
    static constexpr int A_BYTES = 128 * ATOM_ROW_BYTES;
    static constexpr int W_BYTES = NP * ATOM_ROW_BYTES;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
    static constexpr int TMEM_COLS = 512;                 // two accumulators at columns 0 and 256
    static constexpr int NBUILD = 256;                    // warps 2-9
#ifndef GB_LIN_EPARTS
#define GB_LIN_EPARTS 2
#endif
    // epilogue parts of 4 warps (a part covers the 128 TMEM lanes).  Four parts (26 warps, 72 registers) were measured in round 2:
    // 0.098 -> 0.195 ms per launch -- the kernel is bound by the L1 / shared-memory data pipe (93 % busy), not by epilogue warps
    static constexpr int EPARTS = NP > 208 ? 2 : GB_LIN_EPARTS;   // warps 10-17
    static constexpr int NEPI = 128 * EPARTS;
    static constexpr int THREADS = 64 + NBUILD + NEPI;
    static constexpr int STG_PITCH = 20;                  // floats per staged row: conflict-free float4 writes by row
    static constexpr int STG_WARP_BYTES = 32 * STG_PITCH * 4;
    static constexpr int BAR_BYTES = 1024 + 256;
    static constexpr int SMEM = S * STAGE_BYTES + BAR_BYTES + (NEPI / 32) * STG_WARP_BYTES + NP * 4;
    static_assert(SMEM <= 232448, "shared memory budget");
};

template <int NP>
__global__ void __launch_bounds__(TcLinCfg<NP>::THREADS, GB_LIN_CTAS) tc_lin_kernel(LinArgs a, const float* __restrict__ wimg, int H) {
    using CF = TcLinCfg<NP>;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array: the compiler keeps the address space (LDS / STS
    // instead of generic LD / ST for every staging and operand access)
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + CF::S * CF::STAGE_BYTES);
    uint64_t* full_a = bars; uint64_t* full_w = bars + CF::S; uint64_t* empty = bars + 2 * CF::S;
    uint64_t* d_full = bars + 3 * CF::S; uint64_t* d_empty = d_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 2);
    float* stg_all = reinterpret_cast<float*>(base + CF::S * CF::STAGE_BYTES + CF::BAR_BYTES);
    float* bias_s = stg_all + (CF::NEPI / 32) * (CF::STG_WARP_BYTES / 4);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int cb = blockIdx.y;
#ifdef GB_TIMELINE
    const long long tl_t0 = clock64();
#endif
    if (tid == 0) {
        for (int s = 0; s < CF::S; ++s) { mbar_init(&full_a[s], CF::NBUILD); mbar_init(&full_w[s], 1); mbar_init(&empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&d_full[b], 1); mbar_init(&d_empty[b], CF::NEPI); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<CF::TMEM_COLS>(tmem_slot);
    for (int i = tid; i < NP; i += blockDim.x) bias_s[i] = (a.bias && i < H) ? a.bias[(size_t)cb * H + i] : 0.f;
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    const int n_tiles = (a.M + 127) / 128;
    const int na1 = (a.K1 + ATOM_K - 1) / ATOM_K, na2 = (a.K2 + ATOM_K - 1) / ATOM_K, na = na1 + na2;
    const size_t atom_floats = (size_t)2 * NP * ATOM_K;
    const float* wcb = wimg + (size_t)cb * na * atom_floats;
    constexpr uint32_t idesc = instr_desc_tf32(NP), idesc_mix = instr_desc_mix(NP, MIX_BF16);

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
                const uint32_t buf = tcnt & 1, use = tcnt >> 1;
                if (use > 0) mbar_wait(&d_empty[buf], (use - 1) & 1);
                fence_after_sync();
                const uint32_t d_tmem = tmem_base + buf * 256;
                for (int j = 0; j < na; ++j, ++it) {
                    const uint32_t s = it % CF::S, r = it / CF::S;
                    const int kvalid = (j < na1 ? a.K1 - j * ATOM_K : a.K2 - (j - na1) * ATOM_K);
                    const int ksteps = kvalid >= ATOM_K ? 4 : (kvalid + 7) / 8;
                    mbar_wait(&full_a[s], r & 1);
                    TL(1, 2 * it, 1000 + it);
                    mbar_wait(&full_w[s], r & 1);
                    TL(1, 2 * it + 1, 2000 + it);
                    fence_after_sync();
                    const uint32_t a_hi = smem_u32(base + s * CF::STAGE_BYTES), a_lo = a_hi + CF::A_BYTES;
                    const uint32_t w_hi = a_hi + 2 * CF::A_BYTES, w_lo = w_hi + CF::W_BYTES;
                    if (a.mix) {                                   // TF32 product + both correction terms as one bf16 MMA (tc_common.cuh)
                        for (int kk = 0; kk < ksteps; ++kk) {
                            const uint32_t ko = kk * 32;
                            mma_tf32(d_tmem, smem_desc(a_hi + ko), smem_desc(w_hi + ko), idesc, (j | kk) != 0);
                            mma_f16(d_tmem, smem_desc(a_lo + ko), smem_desc(w_lo + ko), idesc_mix, 1);
                        }
                    } else {
                        for (int kk = 0; kk < ksteps; ++kk) {
                            const uint32_t ko = kk * 32;
                            mma_tf32(d_tmem, smem_desc(a_lo + ko), smem_desc(w_hi + ko), idesc, (j | kk) != 0);
                            mma_tf32(d_tmem, smem_desc(a_hi + ko), smem_desc(w_lo + ko), idesc, 1);
                            mma_tf32(d_tmem, smem_desc(a_hi + ko), smem_desc(w_hi + ko), idesc, 1);
                        }
                    }
                    mma_commit(&empty[s]);
                }
                mma_commit(&d_full[buf]);
            }
        }
    } else if (warp < 2 + CF::NBUILD / 32) {
        // ---- A builders: float4 f = bt + 256 i covers row f/8, 16-byte chunk f%8 of the atom; the loads of the next TWO atoms
        //      (across tiles too) are in flight while one atom is split and stored: ~32 KB of reads in flight per SM ----
        const int bt = tid - 64;
        const int my_tiles = blockIdx.x < n_tiles ? (n_tiles - 1 - blockIdx.x) / gridDim.x + 1 : 0;
        const uint32_t total = (uint32_t)my_tiles * na;
        auto load_atom = [&](uint32_t q, float4 (&x)[4]) {
            if (q >= total) return;
            const int tile = blockIdx.x + (q / na) * gridDim.x, j = q % na;
            const bool first = j < na1;
            const int kvalid = first ? a.K1 - j * ATOM_K : a.K2 - (j - na1) * ATOM_K;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int f = bt + 256 * i, kc = (f & 7) << 2;
                const int grow = tile * 128 + (f >> 3);
                x[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (grow < a.M && kc < kvalid) {
                    const float* src = first ? a.A1 + (size_t)grow * a.lda1 + j * ATOM_K : a.A2 + (size_t)grow * a.lda2 + (j - na1) * ATOM_K;
                    x[i] = __ldg(reinterpret_cast<const float4*>(src + kc));
                }
            }
        };
        auto store_atom = [&](uint32_t q, float4 (&x)[4]) {
            if (q >= total) return;
            const uint32_t s = q % CF::S, rr = q / CF::S;
            if (a.rowscale) {
                const int tile = blockIdx.x + (q / na) * gridDim.x;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int grow = tile * 128 + ((bt + 256 * i) >> 3);
                    const float rs = grow < a.M ? __ldg(a.rowscale + grow) : 0.f;
                    x[i].x *= rs; x[i].y *= rs; x[i].z *= rs; x[i].w *= rs;
                }
            }
            if (rr > 0) mbar_wait(&empty[s], (rr - 1) & 1);
            unsigned char* a_hi = base + s * CF::STAGE_BYTES;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int f = bt + 256 * i;
                if (a.mix) store_chunk_mix<MIX_BF16>(a_hi, a_hi + CF::A_BYTES, f >> 3, f & 7, x[i]);
                else store_split(a_hi, a_hi + CF::A_BYTES, f >> 3, f & 7, x[i]);
            }
            fence_proxy_async();
            mbar_arrive(&full_a[s]);
            if (tid == 64) TL(0, q, 200 + q);
        };
        float4 xa[4], xb[4], xc[4];
        load_atom(0, xa);
        load_atom(1, xb);
        for (uint32_t q = 0; q < total; q += 3) {
            load_atom(q + 2, xc); store_atom(q, xa);
            load_atom(q + 3, xa); store_atom(q + 1, xb);
            load_atom(q + 4, xb); store_atom(q + 2, xc);
        }
    } else {
        // ---- epilogue warps ----
        const int group = warp & 3, part = (warp - 2 - CF::NBUILD / 32) >> 2;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(group * 32) << 16);
        const int nchunks = (H + 15) / 16;
        float* stg = stg_all + (warp - 2 - CF::NBUILD / 32) * (CF::STG_WARP_BYTES / 4);
        const int piece = lane & 3, rsub = lane >> 2;
        const bool use_res = a.epi == EPI_RES_MASK || (a.epi == EPI_ADD_RES && (a.res_cb < 0 || cb == a.res_cb));
        uint32_t tcnt = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcnt) {
            const uint32_t buf = tcnt & 1;
            mbar_wait(&d_full[buf], (tcnt >> 1) & 1);
            if (tid == 320) TL(2, 2 * tcnt, 300 + tcnt);
            fence_after_sync();
            // the TMEM read AND the side-input loads (residual / mask / saved pre-activation) of the next chunk are in flight while
            // this one is processed: with the loads issued and consumed inside one chunk step every step exposed a full L2 / HBM
            // latency -- the epilogue then took ~11 us per tile instead of ~6.5 and paced every launch that has a side input
            // (RES_MASK 77 us where the same GEMM without side input takes 45; launch list of round 2e)
            uint32_t vr[16];
            float4 sd_n[4];
            float mk_n[4];
            auto load_side = [&](int ch, float4 (&sd)[4], float (&mk)[4]) {
                const int c0 = ch * 16 + 4 * piece;
                const size_t col = (size_t)cb * H + c0;
                const int row0 = tile * 128 + group * 32 + rsub;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int row = row0 + 8 * i;
                    sd[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    mk[i] = 1.f;
                    if (ch < nchunks && c0 < H && row < a.M) {
                        if (a.epi == EPI_MUL_DSILU) sd[i] = __ldg(reinterpret_cast<const float4*>(a.aux + (size_t)row * a.ldaux + col));
                        else if (use_res) {
                            sd[i] = __ldg(reinterpret_cast<const float4*>(a.res + (size_t)row * a.ldr + col));
                            if (a.epi == EPI_RES_MASK || a.mask) mk[i] = __ldg(a.mask + row);
                        }
                    }
                }
            };
            if (GB_LIN_SPLIT_LD && part < nchunks) tmem_ld16_issue(lane_addr + buf * 256 + part * 16, vr);
            load_side(part, sd_n, mk_n);
            for (int ch = part; ch < nchunks; ch += CF::EPARTS) {
                if (GB_LIN_SPLIT_LD) tmem_ld_wait();
                else tmem_ld16u(lane_addr + buf * 256 + ch * 16, vr);
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    *reinterpret_cast<uint4*>(stg + lane * CF::STG_PITCH + 4 * q) = make_uint4(vr[4 * q], vr[4 * q + 1], vr[4 * q + 2], vr[4 * q + 3]);
                __syncwarp();
                if (GB_LIN_SPLIT_LD && ch + CF::EPARTS < nchunks) tmem_ld16_issue(lane_addr + buf * 256 + (ch + CF::EPARTS) * 16, vr);
                float4 sd[4];
                float mk[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) { sd[i] = sd_n[i]; mk[i] = mk_n[i]; }
                load_side(ch + CF::EPARTS, sd_n, mk_n);
                const int c0 = ch * 16 + 4 * piece;
                if (c0 < H) {
                    const size_t col = (size_t)cb * H + c0;      // column blocks are H wide in the node tensors
                    const float4 b = *reinterpret_cast<const float4*>(bias_s + c0);
                    const int row0 = tile * 128 + group * 32 + rsub;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rl = rsub + 8 * i;
                        const int row = row0 + 8 * i;
                        if (row >= a.M) continue;
                        const float4 acc = *reinterpret_cast<const float4*>(stg + rl * CF::STG_PITCH + 4 * piece);
                        float o[4] = {acc.x + b.x, acc.y + b.y, acc.z + b.z, acc.w + b.w};
                        if (a.epi == EPI_SILU) {
                            if (a.out2) *reinterpret_cast<float4*>(a.out2 + (size_t)row * a.ldo2 + col) = make_float4(o[0], o[1], o[2], o[3]);
#pragma unroll
                            for (int e = 0; e < 4; ++e) o[e] = silu_f(o[e]);
                        } else if (a.epi == EPI_MUL_DSILU) {
                            o[0] *= dsilu_f(sd[i].x); o[1] *= dsilu_f(sd[i].y); o[2] *= dsilu_f(sd[i].z); o[3] *= dsilu_f(sd[i].w);
                        } else if (use_res) {
                            if (a.epi == EPI_RES_MASK) { o[0] = (sd[i].x + o[0]) * mk[i]; o[1] = (sd[i].y + o[1]) * mk[i]; o[2] = (sd[i].z + o[2]) * mk[i]; o[3] = (sd[i].w + o[3]) * mk[i]; }
                            else { o[0] += sd[i].x * mk[i]; o[1] += sd[i].y * mk[i]; o[2] += sd[i].z * mk[i]; o[3] += sd[i].w * mk[i]; }
                        }
                        *reinterpret_cast<float4*>(a.out + (size_t)row * a.ldo + col) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                }
                __syncwarp();
            }
            fence_before_sync();
            mbar_arrive(&d_empty[buf]);
            if (tid == 320) TL(2, 2 * tcnt + 1, 400 + tcnt);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc<CF::TMEM_COLS>(tmem_base);
#ifdef GB_TIMELINE
    if (tid == 0 && blockIdx.y * gridDim.x + blockIdx.x < 512) gb_tl_dur[blockIdx.y * gridDim.x + blockIdx.x] = (unsigned long long)(clock64() - tl_t0);
#endif
}

#ifdef GB_TIMELINE
extern "C" int gb_debug_timeline(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, gb_tl, sizeof(gb_tl)); }
extern "C" int gb_debug_timeline_dur(unsigned long long* out) { return (int)cudaMemcpyFromSymbol(out, gb_tl_dur, sizeof(gb_tl_dur)); }
#endif

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

// Weight image for the tensor-core path: per K-atom j two images of [NP rows][128 bytes] with the 128B swizzle applied, so that
// one contiguous bulk copy lands a ready-to-use B operand.  value(n, k) = transpose ? W[(k+k_off)*ld + n_off+n]
//                                                                                : W[(n+n_off)*ld + k_off+k]
//   fmt 0 (node Linear / training kernels, three TF32 MMAs): [hi | lo] as fp32
//   fmt 1 / 2 (edge kernels, TF32 + one 16-bit MMA, tc_common.cuh): [hi as fp32 | per K step of 8: the 8 hi parts then the 8
//              residuals, as fp16 (1) or bf16 (2)]
__global__ void pack_tc_kernel(float* dst, const float* src, int ld, int k_off, int n_off, int Kv, int Nv, int NP, int atoms, int transpose, int fmt) {
    const int total = atoms * NP * ATOM_K;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int j = idx / (NP * ATOM_K), rem = idx % (NP * ATOM_K);
        const int n = rem / ATOM_K, kk = rem % ATOM_K, k = j * ATOM_K + kk;
        float w = 0.f;
        if (n < Nv && k < Kv) w = transpose ? src[(size_t)(k + k_off) * ld + n_off + n] : src[(size_t)(n + n_off) * ld + k_off + k];
        const float hi = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
        const int c = kk >> 2, e = kk & 3;
        const size_t off = (size_t)(n >> 3) * 256 + (n & 7) * 32 + ((c ^ (n & 7)) << 2) + e;
        float* img_hi = dst + ((size_t)j * 2 + 0) * NP * ATOM_K;
        float* img_2 = dst + ((size_t)j * 2 + 1) * NP * ATOM_K;
        img_hi[off] = hi;
        if (fmt == 0) { img_2[off] = w - hi; continue; }
        // mix image: K step q = kk / 8 occupies the 16-byte chunks 2q (hi parts) and 2q + 1 (residuals) of the row, 8 halfs each
        const int q = kk >> 3, e8 = kk & 7;
        unsigned short* m16 = reinterpret_cast<unsigned short*>(img_2) + (size_t)(n >> 3) * 512 + (n & 7) * 64;
        const float lo = w - hi;
        unsigned short h16, l16;
        if (fmt == 2) { h16 = __bfloat16_as_ushort(__float2bfloat16_rn(hi)); l16 = __bfloat16_as_ushort(__float2bfloat16_rn(lo)); }
        else { h16 = __half_as_ushort(__float2half_rn(hi)); l16 = __half_as_ushort(__float2half_rn(lo)); }
        m16[(((2 * q) ^ (n & 7)) << 3) + e8] = h16;
        m16[(((2 * q + 1) ^ (n & 7)) << 3) + e8] = l16;
    }
}

void launch_pack_tc(float* dst, const float* src, int ld, int k_off, int n_off, int Kv, int Nv, int NP, int atoms, int transpose, cudaStream_t s, int fmt) {
    const int total = atoms * NP * ATOM_K;
    pack_tc_kernel<<<(total + 255) / 256, 256, 0, s>>>(dst, src, ld, k_off, n_off, Kv, Nv, NP, atoms, transpose, fmt);
}

}  // namespace gb
