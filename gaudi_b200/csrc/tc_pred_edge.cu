// Tensor-core (tcgen05, 3xTF32) versions of the predictor E_GCL edge kernels -- same contract and same saved-activation
// layouts as pred_edge.cu (forward: gcl.py:225-279; backward: the hand-written reverse pass replacing autograd).
//
// thread = edge row; 16 worker warps in four parts, part p owns the 16-column chunks ch == p (mod 4) both when building
// A K-atoms (chunk ch = 16-byte chunks 4(ch&1)..+3 of atom ch/2) and in the epilogues, so every atom is built by two
// parts (256 threads).  Two accumulators live in TMEM (columns [0,NP) and [256,256+NP)): the epilogue of GEMM 1 produces the A
// atoms of GEMM 2 on the fly, chunk by chunk, while the MMA warp consumes them.
#include "tc_common.cuh"
#include "kernels.h"

#ifndef GB_PRED_NPARTS
#define GB_PRED_NPARTS 4
#endif

namespace gb {
using namespace tc;

template <int NP>
struct TcPredCfg {
    static constexpr int S = 2;
    static constexpr int A_BYTES = 128 * ATOM_ROW_BYTES;
    static constexpr int W_BYTES = NP * ATOM_ROW_BYTES;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
    static constexpr int NPARTS = GB_PRED_NPARTS;                          // worker parts of 4 warps; part p owns 16-column chunks ch == p (mod 4)
    static constexpr int NWORK = 128 * NPARTS;
    static constexpr int THREADS = 64 + NWORK;
    static constexpr int MAXCH = (NP + 15) / 16;
    static constexpr int MYCH = (MAXCH + NPARTS - 1) / NPARTS;
    static constexpr int EF_STRIDE = 17;
    static constexpr int SCRATCH = 6 * NP * 4 + 4 * NPARTS * 128 * 4 + NPARTS * 128 * EF_STRIDE * 4 + 129 * 4 + 128 * 3 * 4 + 64 + 3 * 132 * 4;
    static constexpr int SMEM = S * STAGE_BYTES + 1024 + 256 + SCRATCH;
    static constexpr int D2_COL = 256;
    // backward only: one extra producer warp streams the saved activations (16-column chunks of the float4 "Q layout",
    // 4 planes x 128 rows x 16 B = 8 KB, contiguous in HBM) through a 4-slot ring that lives in the ef_s scratch area
    static constexpr int BWD_THREADS = THREADS + 32;
    static constexpr int SV_SLOTS = 4;
    static constexpr int SV_SLOT_BYTES = 4 * 128 * 16;
    static_assert(SV_SLOTS * SV_SLOT_BYTES <= NPARTS * 128 * EF_STRIDE * 4, "saved-activation ring must fit the ef scratch");
};

struct SvRing { unsigned char* buf; uint64_t* full; uint64_t* empty; };

// consumer side of the saved-activation ring: wait for chunk q, return this row's first float4 (plane c at +128*c)
__device__ __forceinline__ const float4* sv_acquire(const SvRing& sv, uint32_t q, int r) {
    const uint32_t s = q & 3, rr = q >> 2;
    mbar_wait(&sv.full[s], rr & 1);
    return reinterpret_cast<const float4*>(sv.buf + s * 8192) + r;
}
__device__ __forceinline__ void sv_release(const SvRing& sv, uint32_t q) { mbar_arrive(&sv.empty[q & 3]); }

template <int NPARTS>
__device__ __forceinline__ float psum_parts(const float* red, int r) {
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < NPARTS; ++p) s += red[p * 128 + r];
    return s;
}

__device__ __forceinline__ void nbar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

struct TcPipe {      // shared bookkeeping of the atom ring (same counter sequence in every role)
    unsigned char* base; uint64_t *full_a, *full_w, *empty; int stage_bytes, S;
};

// one GEMM worth of MMAs: na atoms from the ring into accumulator d_tmem
template <int NP>
__device__ __forceinline__ void mma_gemm(const TcPipe& p, uint32_t& it, int na, int H, uint32_t d_tmem) {
    using CF = TcPredCfg<NP>;
    constexpr uint32_t idesc = instr_desc_tf32(NP);
    for (int j = 0; j < na; ++j, ++it) {
        const uint32_t s = it % CF::S, r = it / CF::S;
        const int kvalid = H - j * ATOM_K;
        const int ksteps = kvalid >= ATOM_K ? 4 : (kvalid + 7) / 8;
        mbar_wait(&p.full_a[s], r & 1);
        mbar_wait(&p.full_w[s], r & 1);
        fence_after_sync();
        const uint32_t a_hi = smem_u32(p.base + s * CF::STAGE_BYTES), a_lo = a_hi + CF::A_BYTES;
        const uint32_t w_hi = a_hi + 2 * CF::A_BYTES, w_lo = w_hi + CF::W_BYTES;
        for (int kk = 0; kk < ksteps; ++kk) {
            const uint32_t ko = kk * 32;
            mma_tf32(d_tmem, smem_desc(a_lo + ko), smem_desc(w_hi + ko), idesc, (j | kk) != 0);
            mma_tf32(d_tmem, smem_desc(a_hi + ko), smem_desc(w_lo + ko), idesc, 1);
            mma_tf32(d_tmem, smem_desc(a_hi + ko), smem_desc(w_hi + ko), idesc, 1);
        }
        mma_commit(&p.empty[s]);
    }
}

template <int NP>
__device__ __forceinline__ void tma_gemm(const TcPipe& p, uint32_t& it, int na, const float* wimg) {
    using CF = TcPredCfg<NP>;
    const size_t atom_floats = (size_t)2 * NP * ATOM_K;
    for (int j = 0; j < na; ++j, ++it) {
        const uint32_t s = it % CF::S, r = it / CF::S;
        if (r > 0) mbar_wait(&p.empty[s], (r - 1) & 1);
        mbar_arrive_expect_tx(&p.full_w[s], 2 * CF::W_BYTES);
        bulk_g2s(p.base + s * CF::STAGE_BYTES + 2 * CF::A_BYTES, wimg + (size_t)j * atom_floats, 2 * CF::W_BYTES, &p.full_w[s]);
    }
}

// worker side: publish this thread's 16 columns (4 x float4) of atom `it` (chunk parity h writes 16-byte chunks 4h..4h+3)
template <int NP>
__device__ __forceinline__ void put_chunk(const TcPipe& p, uint32_t it, int r, int half, const float4 (&x)[4]) {
    using CF = TcPredCfg<NP>;
    const uint32_t s = it % CF::S, rr = it / CF::S;
    if (rr > 0) mbar_wait(&p.empty[s], (rr - 1) & 1);
    unsigned char* a_hi = p.base + s * CF::STAGE_BYTES;
#pragma unroll
    for (int c = 0; c < 4; ++c) store_split(a_hi, a_hi + CF::A_BYTES, r, 4 * half + c, x[c]);
    fence_proxy_async();
    mbar_arrive(&p.full_a[s]);
}

// ==================================================================================================================
// forward
// ==================================================================================================================
template <int NP, bool SAVE>
__global__ void __launch_bounds__(TcPredCfg<NP>::THREADS, 1) tc_pred_edge_fwd_kernel(PredEdgeArgs a, const float* __restrict__ w2img,
                                                                   const float* __restrict__ wcimg, int H) {
    using CF = TcPredCfg<NP>;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array: the compiler keeps the address space (LDS / STS
    // instead of generic LD / ST for every staging and operand access)
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + CF::S * CF::STAGE_BYTES);
    TcPipe p{base, bars, bars + CF::S, bars + 2 * CF::S, CF::STAGE_BYTES, CF::S};
    uint64_t* d1_full = bars + 3 * CF::S; uint64_t* d2_full = d1_full + 1; uint64_t* d_empty = d2_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 1);
    float* vec_s = reinterpret_cast<float*>(base + CF::S * CF::STAGE_BYTES + 256);    // [6][NP]: w_r, w_a, b2, att_w, bc, wc_last
    float* red_s = vec_s + 6 * NP;                                                       // [2][2][128]
    float* ef_s = red_s + 4 * CF::NPARTS * 128;                                                       // [2][128][17]
    int* seg_s = reinterpret_cast<int*>(ef_s + CF::NPARTS * 128 * CF::EF_STRIDE);
    float* tr_s = reinterpret_cast<float*>(seg_s + 129);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < CF::S; ++s) { mbar_init(&p.full_a[s], 256); mbar_init(&p.full_w[s], 1); mbar_init(&p.empty[s], 1); }
        mbar_init(d1_full, 1); mbar_init(d2_full, 1); mbar_init(d_empty, CF::NWORK);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    for (int i = tid; i < NP; i += blockDim.x) {
        const bool v = i < H;
        vec_s[i] = v ? a.ext[i] : 0.f; vec_s[NP + i] = v ? a.ext[H + i] : 0.f; vec_s[2 * NP + i] = v ? a.b2[i] : 0.f;
        vec_s[3 * NP + i] = v ? a.att_w[i] : 0.f; vec_s[4 * NP + i] = v ? a.bc[i] : 0.f; vec_s[5 * NP + i] = v ? a.wc_last[i] : 0.f;
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const Graph& g = a.g;
    const int na = (H + ATOM_K - 1) / ATOM_K;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) { tma_gemm<NP>(p, it, na, w2img); tma_gemm<NP>(p, it, na, wcimg); }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, tcnt = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++tcnt) {
                if (tcnt > 0) mbar_wait(d_empty, (tcnt - 1) & 1);
                fence_after_sync();
                mma_gemm<NP>(p, it, na, H, tmem_base);
                mma_commit(d1_full);
                mma_gemm<NP>(p, it, na, H, tmem_base + CF::D2_COL);
                mma_commit(d2_full);
            }
        }
    } else {
        const int group = warp & 3, part = (warp - 2) >> 2, half = part & 1;
        const int r = group * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(group * 32) << 16);
        const int nchunks = (H + 15) / 16;
        float* my_ef = ef_s + part * 128 * CF::EF_STRIDE;
        uint32_t tcnt = 0;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++tcnt) {
            const int4 ti = __ldg(g.tile_info + tile);
            const int node_lo = ti.x, nn = ti.y, e_lo = ti.z, ne = ti.w;
            const bool valid = r < ne;
            int rown = 0, coln = 0; float rad = 0.f, a0 = 0.f;
            {   // the unit difference vector is needed again only for the coordinate update at the end of the tile (part 0):
                // it waits in tr_s instead of in three registers of every worker
                float ux = 0.f, uy = 0.f, uz = 0.f;
                if (valid) {
                    const int e = e_lo + r;
                    rown = g.erow[e]; coln = g.ecol[e];
                    const float dx = a.x[3 * rown] - a.x[3 * coln], dy = a.x[3 * rown + 1] - a.x[3 * coln + 1], dz = a.x[3 * rown + 2] - a.x[3 * coln + 2];
                    rad = dx * dx + dy * dy + dz * dz;
                    const float inv = 1.f / (sqrtf(rad + 1e-8f) + 1.f);
                    ux = dx * inv; uy = dy * inv; uz = dz * inv;
                    const float ex = a.x0[3 * rown] - a.x0[3 * coln], ey = a.x0[3 * rown + 1] - a.x0[3 * coln + 1], ez = a.x0[3 * rown + 2] - a.x0[3 * coln + 2];
                    a0 = ex * ex + ey * ey + ez * ez;
                    if (a.a_edge) a0 = a.a_edge[e];
                }
                if (part == 0) { tr_s[3 * r] = ux; tr_s[3 * r + 1] = uy; tr_s[3 * r + 2] = uz; }
            }
            if (part == 0) for (int i = r; i <= nn; i += 128) seg_s[i] = g.rowptr[node_lo + i] - e_lo;
            const uint32_t it0 = tcnt * 2 * na;
            // ---- GEMM 1 operand: s1 = SiLU(pre1); SiLU'(pre1) is saved ----
            const float* pa_row = a.P + (size_t)rown * (2 * H);
            const float* pb_row = a.P + (size_t)coln * (2 * H) + H;
            for (int j = part >> 1; j < na; j += CF::NPARTS / 2) {
                float4 x[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int k0 = j * ATOM_K + 16 * half + 4 * c;
                    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), dv = v;
                    if (valid && k0 < H) {
                        const float4 pa = __ldg(reinterpret_cast<const float4*>(pa_row + k0));
                        const float4 pb = __ldg(reinterpret_cast<const float4*>(pb_row + k0));
                        const float4 wr = *reinterpret_cast<const float4*>(vec_s + k0);
                        const float4 wa = *reinterpret_cast<const float4*>(vec_s + NP + k0);
                        silu_both(pa.x + pb.x + wr.x * rad + wa.x * a0, v.x, dv.x);
                        silu_both(pa.y + pb.y + wr.y * rad + wa.y * a0, v.y, dv.y);
                        silu_both(pa.z + pb.z + wr.z * rad + wa.z * a0, v.z, dv.z);
                        silu_both(pa.w + pb.w + wr.w * rad + wa.w * a0, v.w, dv.w);
                    }
                    x[c] = v;
                    if (SAVE && k0 < H) *reinterpret_cast<float4*>(a.sv_d1 + (((size_t)tile * (H / 4) + (k0 >> 2)) * 128 + r) * 4) = dv;
                }
                put_chunk<NP>(p, it0 + j, r, half, x);
            }
            // ---- epilogue 1: q = SiLU(pre2), attention gate ----
            mbar_wait(d1_full, tcnt & 1);
            fence_after_sync();
            float q[CF::MYCH][16];
            float psum = 0.f;
#pragma unroll
            for (int ci = 0; ci < CF::MYCH; ++ci) {
                const int ch = part + CF::NPARTS * ci;
                if (ch < nchunks) {
                    tmem_ld16(lane_addr + ch * 16, q[ci]);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const int c0 = ch * 16 + 4 * c4;
                        float pre[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            pre[e] = q[ci][4 * c4 + e] + vec_s[2 * NP + c0 + e];
                            const float v = silu_f(pre[e]);
                            q[ci][4 * c4 + e] = v;
                            psum = fmaf(vec_s[3 * NP + c0 + e], v, psum);
                        }
                        if (SAVE && c0 < H)
                            *reinterpret_cast<float4*>(a.sv_pre2 + (((size_t)tile * (H / 4) + (c0 >> 2)) * 128 + r) * 4) = make_float4(pre[0], pre[1], pre[2], pre[3]);
                    }
                }
            }
            red_s[part * 128 + r] = psum;
            nbar(1, CF::NWORK);
            const float gate = a.attention ? sigmoid_f(psum_parts<CF::NPARTS>(red_s, r) + a.att_b) : 1.f;
            // ---- gated edge feature: segment sums -> agg, and operand atoms of GEMM 2 ----
#pragma unroll
            for (int ci = 0; ci < CF::MYCH; ++ci) {
                const int ch = part + CF::NPARTS * ci;
                if (ch < nchunks) {
                    float4 x[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        x[c] = make_float4(q[ci][4 * c] * gate, q[ci][4 * c + 1] * gate, q[ci][4 * c + 2] * gate, q[ci][4 * c + 3] * gate);
                        if (!valid) x[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                        my_ef[r * CF::EF_STRIDE + 4 * c] = x[c].x; my_ef[r * CF::EF_STRIDE + 4 * c + 1] = x[c].y;
                        my_ef[r * CF::EF_STRIDE + 4 * c + 2] = x[c].z; my_ef[r * CF::EF_STRIDE + 4 * c + 3] = x[c].w;
                    }
                    put_chunk<NP>(p, it0 + na + (ch >> 1), r, half, x);
                    nbar(2 + part, 128);
                    for (int nl = r >> 4; nl < nn; nl += 8) {
                        const int col = r & 15, c = ch * 16 + col;
                        float sum = 0.f;
                        for (int mm = seg_s[nl]; mm < seg_s[nl + 1]; ++mm) sum += my_ef[mm * CF::EF_STRIDE + col];
                        if (c < H) a.agg[(size_t)(node_lo + nl) * H + c] = sum;
                    }
                    nbar(2 + part, 128);
                } else if (ch < 2 * na) {
                    // chunk beyond the hidden width but inside the last atom: publish zeros so the atom completes
                    float4 x[4] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
                    put_chunk<NP>(p, it0 + na + (ch >> 1), r, half, x);
                }
            }
            // ---- epilogue 2: coordinate head ----
            mbar_wait(d2_full, tcnt & 1);
            fence_after_sync();
            float phi_part = 0.f;
#pragma unroll 1
            for (int ch = part; ch < nchunks; ch += CF::NPARTS) {
                float v[16];
                tmem_ld16(lane_addr + CF::D2_COL + ch * 16, v);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int c0 = ch * 16 + 4 * c4;
                    float d3[4];
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        float s3;
                        silu_both(v[4 * c4 + e] + vec_s[4 * NP + c0 + e], s3, d3[e]);
                        phi_part = fmaf(vec_s[5 * NP + c0 + e], s3, phi_part);
                    }
                    if (SAVE && c0 < H)
                        *reinterpret_cast<float4*>(a.sv_d3 + (((size_t)tile * (H / 4) + (c0 >> 2)) * 128 + r) * 4) = make_float4(d3[0], d3[1], d3[2], d3[3]);
                }
            }
            fence_before_sync();
            mbar_arrive(d_empty);
            red_s[CF::NPARTS * 128 + part * 128 + r] = phi_part;
            nbar(1, CF::NWORK);
            if (part == 0) {
                const float phi = psum_parts<CF::NPARTS>(red_s + CF::NPARTS * 128, r);
                const float tau = a.use_tanh ? tanhf(phi) : phi;
                if (SAVE && valid) a.sv_tau[e_lo + r] = tau;
                const float ux = tr_s[3 * r], uy = tr_s[3 * r + 1], uz = tr_s[3 * r + 2];
                if (a.use_tanh) { tr_s[3 * r] = ux * tau * a.coords_range; tr_s[3 * r + 1] = uy * tau * a.coords_range; tr_s[3 * r + 2] = uz * tau * a.coords_range; }
                else { tr_s[3 * r] = ux * tau; tr_s[3 * r + 1] = uy * tau; tr_s[3 * r + 2] = uz * tau; }
            }
            nbar(1, CF::NWORK);
            for (int idx = part * 128 + r; idx < nn * 3; idx += CF::NWORK) {
                const int nl = idx / 3, d = idx - 3 * nl;
                float sum = 0.f;
                for (int mm = seg_s[nl]; mm < seg_s[nl + 1]; ++mm) sum += tr_s[3 * mm + d];
                const int node = node_lo + nl;
                a.x_out[3 * node + d] = (a.x[3 * node + d] + sum) * g.node_mask[node];
            }
            nbar(1, CF::NWORK);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ==================================================================================================================
// backward (input gradient only)
// ==================================================================================================================
template <int NP>
__global__ void __launch_bounds__(TcPredCfg<NP>::BWD_THREADS, 1) tc_pred_edge_bwd_kernel(PredEdgeArgs a, const float* __restrict__ wcimg_nt,
                                                                   const float* __restrict__ w2img_nt, int H) {
    using CF = TcPredCfg<NP>;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array: the compiler keeps the address space (LDS / STS
    // instead of generic LD / ST for every staging and operand access)
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + CF::S * CF::STAGE_BYTES);
    TcPipe p{base, bars, bars + CF::S, bars + 2 * CF::S, CF::STAGE_BYTES, CF::S};
    uint64_t* d1_full = bars + 3 * CF::S; uint64_t* d2_full = d1_full + 1; uint64_t* d_empty = d2_full + 1;
    uint64_t* sv_full = d_empty + 1; uint64_t* sv_empty = sv_full + CF::SV_SLOTS;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(sv_empty + CF::SV_SLOTS);
    float* vec_s = reinterpret_cast<float*>(base + CF::S * CF::STAGE_BYTES + 256);    // [6][NP]: w_r, w_a, b2(unused), att_w, bc(unused), wc_last
    float* red_s = vec_s + 6 * NP;                                                       // [4][128]
    float* ef_s = red_s + 4 * CF::NPARTS * 128;                                          // saved-activation ring (4 x 8 KB)
    int* seg_s = reinterpret_cast<int*>(ef_s + CF::NPARTS * 128 * CF::EF_STRIDE);
    float* geo_s = reinterpret_cast<float*>(seg_s + 132);                                // [6][128] dx, dy, dz, g_u of every tile edge
    const SvRing sv{reinterpret_cast<unsigned char*>(ef_s), sv_full, sv_empty};

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < CF::S; ++s) { mbar_init(&p.full_a[s], 256); mbar_init(&p.full_w[s], 1); mbar_init(&p.empty[s], 1); }
        mbar_init(d1_full, 1); mbar_init(d2_full, 1); mbar_init(d_empty, CF::NWORK);
        for (int s = 0; s < CF::SV_SLOTS; ++s) { mbar_init(&sv_full[s], 1); mbar_init(&sv_empty[s], 128); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    for (int i = tid; i < NP; i += blockDim.x) {
        const bool v = i < H;
        vec_s[i] = v ? a.ext[i] : 0.f; vec_s[NP + i] = v ? a.ext[H + i] : 0.f;
        vec_s[3 * NP + i] = v ? a.att_w[i] : 0.f; vec_s[5 * NP + i] = v ? a.wc_last[i] : 0.f;
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const Graph& g = a.g;
    const int na = (H + ATOM_K - 1) / ATOM_K;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) { tma_gemm<NP>(p, it, na, wcimg_nt); tma_gemm<NP>(p, it, na, w2img_nt); }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, tcnt = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++tcnt) {
                if (tcnt > 0) mbar_wait(d_empty, (tcnt - 1) & 1);
                fence_after_sync();
                mma_gemm<NP>(p, it, na, H, tmem_base);
                mma_commit(d1_full);
                mma_gemm<NP>(p, it, na, H, tmem_base + CF::D2_COL);
                mma_commit(d2_full);
            }
        }
    } else if (warp == 2 + 4 * CF::NPARTS) {
        // saved-activation producer: per tile the chunks of d3 (GEMM-1 operand), pre2 (epilogue 1), pre2 again (GEMM-2
        // operand) and d1 (epilogue 2), in the order the worker parts consume them
        if (lane == 0) {
            const int nchunks = (H + 15) / 16, planes = H / 4;
            uint32_t q = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
#pragma unroll 1
                for (int ph = 0; ph < 4; ++ph) {
                    const float* src = (ph == 0 ? a.sv_d3 : (ph == 3 ? a.sv_d1 : a.sv_pre2)) + (size_t)tile * planes * 512;
                    for (int ch = 0; ch < nchunks; ++ch, ++q) {
                        const uint32_t s = q & 3, rr = q >> 2;
                        const uint32_t bytes = (uint32_t)min(4, planes - 4 * ch) * 2048u;
                        if (rr > 0) mbar_wait(&sv_empty[s], (rr - 1) & 1);
                        mbar_arrive_expect_tx(&sv_full[s], bytes);
                        bulk_g2s(sv.buf + s * CF::SV_SLOT_BYTES, src + (size_t)ch * 2048, bytes, &sv_full[s]);
                    }
                }
            }
        }
    } else {
        const int group = warp & 3, part = (warp - 2) >> 2, half = part & 1;
        const int r = group * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(group * 32) << 16);
        const int nchunks = (H + 15) / 16;
        uint32_t tcnt = 0;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++tcnt) {
            const uint32_t q0 = tcnt * 4 * nchunks;                 // first saved-activation chunk of this tile
            const int node_lo = g.tile_ptr[tile], node_hi = g.tile_ptr[tile + 1];     // (tile_info costs registers here: slower)
            const int nn = node_hi - node_lo;
            const int e_lo = g.rowptr[node_lo], ne = g.rowptr[node_hi] - e_lo;
            const bool valid = r < ne;
            int rown = 0, coln = 0;
            float gphi = 0.f;
            {   // the edge geometry is needed again only at the very end (dL/d(x_i - x_j), part 0): it waits in shared memory
                // instead of occupying seven registers of every worker across both GEMMs
                float dx = 0.f, dy = 0.f, dz = 0.f, gux = 0.f, guy = 0.f, guz = 0.f;
                if (valid) {
                    const int e = e_lo + r;
                    rown = g.erow[e]; coln = g.ecol[e];
                    dx = a.x[3 * rown] - a.x[3 * coln]; dy = a.x[3 * rown + 1] - a.x[3 * coln + 1]; dz = a.x[3 * rown + 2] - a.x[3 * coln + 2];
                    const float nrm = sqrtf(dx * dx + dy * dy + dz * dz + 1e-8f);
                    const float inv = 1.f / (nrm + 1.f);
                    const float mk = g.node_mask[rown];
                    const float gx = a.g_xout[3 * rown] * mk, gy = a.g_xout[3 * rown + 1] * mk, gz = a.g_xout[3 * rown + 2] * mk;
                    const float tau = a.sv_tau[e];
                    const float gdotu = (gx * dx + gy * dy + gz * dz) * inv;
                    if (a.use_tanh) { gphi = gdotu * a.coords_range * (1.f - tau * tau); const float sc = tau * a.coords_range; gux = gx * sc; guy = gy * sc; guz = gz * sc; }
                    else { gphi = gdotu; gux = gx * tau; guy = gy * tau; guz = gz * tau; }
                }
                if (part == 0) {
                    geo_s[r] = dx; geo_s[128 + r] = dy; geo_s[256 + r] = dz;
                    geo_s[384 + r] = gux; geo_s[512 + r] = guy; geo_s[640 + r] = guz;
                }
            }
            if (part == 0) for (int i = r; i <= nn; i += 128) seg_s[i] = g.rowptr[node_lo + i] - e_lo;
            const uint32_t it0 = tcnt * 2 * na;
            // ---- GEMM 1 operand: g_pre3 = g_phi * w_c * SiLU'(pre3) ----
            for (int j = part >> 1; j < na; j += CF::NPARTS / 2) {
                float4 x[4];
                const int ch = 2 * j + half;
                const bool have = ch < nchunks;
                const float4* d3p = have ? sv_acquire(sv, q0 + ch, r) : nullptr;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int k0 = j * ATOM_K + 16 * half + 4 * c;
                    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (k0 < H) {
                        const float4 d3 = d3p[c * 128];
                        const float4 wl = *reinterpret_cast<const float4*>(vec_s + 5 * NP + k0);
                        t = make_float4(gphi * wl.x * d3.x, gphi * wl.y * d3.y, gphi * wl.z * d3.z, gphi * wl.w * d3.w);
                    }
                    x[c] = t;
                }
                if (have) sv_release(sv, q0 + ch);
                put_chunk<NP>(p, it0 + j, r, half, x);
            }
            // ---- epilogue 1: g_ef = coordinate branch + aggregation branch; attention backward ----
            mbar_wait(d1_full, tcnt & 1);
            fence_after_sync();
            // the tile's rows of g_agg (one per row node, shared by its ~9 edges) are copied once, coalesced, into the A half of
            // the operand ring (idle between GEMM 1 and the first operand store of GEMM 2) instead of being gathered from L2
            // with four dependent 16-byte loads per chunk and thread
            constexpr int GA_ROWS = CF::A_BYTES * 2 / (NP * 4);            // rows per ring stage (A hi + A lo)
            const bool ga_staged = nn <= 2 * GA_ROWS;
            if (ga_staged) {
                const int h4 = H >> 2;
                for (int idx = (warp - 2) * 32 + lane; idx < nn * h4; idx += CF::NWORK) {
                    const int nl = idx / h4, k4 = idx - nl * h4;
                    float* dst = reinterpret_cast<float*>(base + (nl / GA_ROWS) * CF::STAGE_BYTES) + (nl % GA_ROWS) * NP + 4 * k4;
                    *reinterpret_cast<float4*>(dst) = __ldg(reinterpret_cast<const float4*>(a.g_agg + (size_t)(node_lo + nl) * a.ld_gagg + 4 * k4));
                }
                nbar(1, CF::NWORK);
            }
            float gef[CF::MYCH][16];
            float plog = 0.f, pdot = 0.f;
            const int rloc = valid ? rown - node_lo : 0;
            const float* ga_row = ga_staged ? reinterpret_cast<const float*>(base + (rloc / GA_ROWS) * CF::STAGE_BYTES) + (rloc % GA_ROWS) * NP
                                            : a.g_agg + (size_t)rown * a.ld_gagg;
#pragma unroll
            for (int ci = 0; ci < CF::MYCH; ++ci) {
                const int ch = part + CF::NPARTS * ci;
                if (ch < nchunks) {
                    tmem_ld16(lane_addr + ch * 16, gef[ci]);
                    const float4* p2p = sv_acquire(sv, q0 + nchunks + ch, r);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const int c0 = ch * 16 + 4 * c4;
                        if (c0 < H) {
                            const float4 ga = *reinterpret_cast<const float4*>(ga_row + c0);
                            const float4 p2 = p2p[c4 * 128];
                            const float gadd[4] = {ga.x, ga.y, ga.z, ga.w};
                            const float pv[4] = {p2.x, p2.y, p2.z, p2.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                const float qv = silu_f(pv[e]);
                                const float gv = gef[ci][4 * c4 + e] + gadd[e];
                                gef[ci][4 * c4 + e] = gv;
                                plog = fmaf(vec_s[3 * NP + c0 + e], qv, plog);
                                pdot = fmaf(gv, qv, pdot);
                            }
                        }
                    }
                    sv_release(sv, q0 + nchunks + ch);
                }
            }
            red_s[part * 128 + r] = plog;
            red_s[CF::NPARTS * 128 + part * 128 + r] = pdot;
            nbar(1, CF::NWORK);
            float gate = 1.f, kap = 0.f;
            if (a.attention) {
                gate = sigmoid_f(psum_parts<CF::NPARTS>(red_s, r) + a.att_b);
                kap = (psum_parts<CF::NPARTS>(red_s + CF::NPARTS * 128, r)) * gate * (1.f - gate);
            }
            // ---- GEMM 2 operand: g_pre2 = (g_ef gate + kappa w_att) SiLU'(pre2) ----
#pragma unroll
            for (int ci = 0; ci < CF::MYCH; ++ci) {
                const int ch = part + CF::NPARTS * ci;
                if (ch < 2 * na) {
                    float4 x[4];
                    const float4* p2p = ch < nchunks ? sv_acquire(sv, q0 + 2 * nchunks + ch, r) : nullptr;
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const int c0 = ch * 16 + 4 * c4;
                        float t[4] = {0.f, 0.f, 0.f, 0.f};
                        if (ch < nchunks && c0 < H && valid) {
                            const float4 p2 = p2p[c4 * 128];
                            const float pv[4] = {p2.x, p2.y, p2.z, p2.w};
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                t[e] = (gef[ci][4 * c4 + e] * gate + kap * vec_s[3 * NP + c0 + e]) * dsilu_f(pv[e]);
                        }
                        x[c4] = make_float4(t[0], t[1], t[2], t[3]);
                    }
                    if (ch < nchunks) sv_release(sv, q0 + 2 * nchunks + ch);
                    put_chunk<NP>(p, it0 + na + (ch >> 1), r, half, x);
                }
            }
            // ---- epilogue 2: g_pre1 = g_s1 * SiLU'(pre1) -> HBM (row-major per edge); radial / attr dots ----
            //      the row sums (g_Pa), column sums (g_Pb) and the coordinate gradient are reduced afterwards by the
            //      node-parallel pred_bwd_reduce_kernel: fixed summation order, no atomics, no per-chunk barriers here
            mbar_wait(d2_full, tcnt & 1);
            fence_after_sync();
            float pr = 0.f, pa = 0.f;
            // g_pre1 rows are written through a per-warp transpose block: thread = row is what TMEM gives, but 16 bytes of 32
            // different rows per store instruction cost 32 L1 tag lookups; transposed, one instruction covers 8 rows x 64 bytes.
            // The block lives in the A half of the operand ring, which is idle between the last MMA of GEMM 2 (d2_full) and the
            // next tile's first operand store (after the barriers below).
            float* stg = reinterpret_cast<float*>(base + ((warp - 2) >> 3) * CF::STAGE_BYTES) + ((warp - 2) & 7) * (32 * 20);
            const int piece = lane & 3, rsub = lane >> 2;
#pragma unroll 1
            for (int ch = part; ch < nchunks; ch += CF::NPARTS) {
                float v[16];
                tmem_ld16(lane_addr + CF::D2_COL + ch * 16, v);
                const float4* d1p = sv_acquire(sv, q0 + 3 * nchunks + ch, r);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int c0 = ch * 16 + 4 * c4;
                    float4 gp = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (c0 < H) {
                        const float4 d1 = d1p[c4 * 128];
                        gp = make_float4(v[4 * c4] * d1.x, v[4 * c4 + 1] * d1.y, v[4 * c4 + 2] * d1.z, v[4 * c4 + 3] * d1.w);
                        const float4 wr = *reinterpret_cast<const float4*>(vec_s + c0);
                        const float4 wa = *reinterpret_cast<const float4*>(vec_s + NP + c0);
                        pr += wr.x * gp.x + wr.y * gp.y + wr.z * gp.z + wr.w * gp.w;
                        pa += wa.x * gp.x + wa.y * gp.y + wa.z * gp.z + wa.w * gp.w;
                    }
                    *reinterpret_cast<float4*>(stg + lane * 20 + 4 * c4) = gp;
                }
                sv_release(sv, q0 + 3 * nchunks + ch);
                __syncwarp();
                const int c = ch * 16 + 4 * piece;
                if (c < H) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rl = group * 32 + rsub + 8 * i;
                        if (rl < ne) *reinterpret_cast<float4*>(a.g_pre1 + (size_t)(e_lo + rl) * H + c) = *reinterpret_cast<const float4*>(stg + (rsub + 8 * i) * 20 + 4 * piece);
                    }
                }
                __syncwarp();
            }
            fence_before_sync();
            mbar_arrive(d_empty);
            red_s[2 * CF::NPARTS * 128 + part * 128 + r] = pr;
            red_s[3 * CF::NPARTS * 128 + part * 128 + r] = pa;
            nbar(1, CF::NWORK);
            if (part == 0 && valid) {
                const float g_r = psum_parts<CF::NPARTS>(red_s + 2 * CF::NPARTS * 128, r);
                const float g_a = psum_parts<CF::NPARTS>(red_s + 3 * CF::NPARTS * 128, r);
                a.g_attr[e_lo + r] += g_a;
                const float dx = geo_s[r], dy = geo_s[128 + r], dz = geo_s[256 + r];
                const float gux = geo_s[384 + r], guy = geo_s[512 + r], guz = geo_s[640 + r];
                const float nrm = sqrtf(dx * dx + dy * dy + dz * dz + 1e-8f);
                const float inv = 1.f / (nrm + 1.f);
                const float k2 = (gux * dx + guy * dy + guz * dz) * inv * inv / nrm;
                float* gd = a.g_d + (size_t)(e_lo + r) * 3;
                gd[0] = 2.f * g_r * dx + gux * inv - k2 * dx;
                gd[1] = 2.f * g_r * dy + guy * inv - k2 * dy;
                gd[2] = 2.f * g_r * dz + guz * inv - k2 * dz;
            }
            nbar(1, CF::NWORK);                                  // red_s / seg_s free for the next tile
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

template <int NP>
static void launch_bwd_t(const PredEdgeArgs& a, const float* wcimg_nt, const float* w2img_nt, int H, cudaStream_t s) {
    using CF = TcPredCfg<NP>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(tc_pred_edge_bwd_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM);
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = a.g.n_tiles < sms ? a.g.n_tiles : sms;
    tc_pred_edge_bwd_kernel<NP><<<grid, CF::BWD_THREADS, CF::SMEM, s>>>(a, wcimg_nt, w2img_nt, H);
}

void launch_pred_edge_bwd_tc(int H, const PredEdgeArgs& a, const float* wcimg_nt, const float* w2img_nt, cudaStream_t s) {
    if (a.g.n_tiles <= 0) return;
    switch (tc_np(H)) {
        case 64: launch_bwd_t<64>(a, wcimg_nt, w2img_nt, H, s); break;
        case 192: launch_bwd_t<192>(a, wcimg_nt, w2img_nt, H, s); break;
        case 208: launch_bwd_t<208>(a, wcimg_nt, w2img_nt, H, s); break;
        default: launch_bwd_t<256>(a, wcimg_nt, w2img_nt, H, s); break;
    }
}

// Node-parallel reductions of the backward (one warp per node, lanes along the feature dimension):
//   g_Pa[i] = sum over the row segment of i,  g_Pb[i] = sum over the edges whose column is i (CSC order),
//   g_x[i]  = g_xout[i]*mask_i + sum_row g_d - sum_col g_d.          Fixed order -> bit-reproducible.
__global__ void pred_bwd_reduce_kernel(PredEdgeArgs a, int H) {
    const Graph& g = a.g;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    const int nq = H >> 2;                                       // float4 columns
    for (int node = warp; node < g.n_nodes; node += nwarps) {
        const int r0 = g.rowptr[node], r1 = g.rowptr[node + 1];
        const int c0 = g.colptr[node], c1 = g.colptr[node + 1];
        for (int q = lane; q < nq; q += 32) {
            float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
            for (int e = r0; e < r1; ++e) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(a.g_pre1 + (size_t)e * H) + q);
                sa.x += v.x; sa.y += v.y; sa.z += v.z; sa.w += v.w;
            }
            for (int p = c0; p < c1; ++p) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(a.g_pre1 + (size_t)__ldg(g.cedge + p) * H) + q);
                sb.x += v.x; sb.y += v.y; sb.z += v.z; sb.w += v.w;
            }
            reinterpret_cast<float4*>(a.g_Pa + (size_t)node * H)[q] = sa;
            reinterpret_cast<float4*>(a.g_Pb + (size_t)node * H)[q] = sb;
        }
        if (lane < 3) {
            float sum = a.g_xout[3 * node + lane] * g.node_mask[node];
            for (int e = r0; e < r1; ++e) sum += a.g_d[(size_t)e * 3 + lane];
            for (int p = c0; p < c1; ++p) sum -= a.g_d[(size_t)g.cedge[p] * 3 + lane];
            a.g_x[3 * node + lane] = sum;
        }
    }
}

void launch_pred_bwd_reduce(int H, const PredEdgeArgs& a, cudaStream_t s) {
    if (a.g.n_nodes <= 0) return;
    const int blocks = min(148 * 8, (a.g.n_nodes + 7) / 8);
    pred_bwd_reduce_kernel<<<blocks, 256, 0, s>>>(a, H);
}

template <int NP>
static void launch_fwd_t(bool save, const PredEdgeArgs& a, const float* w2img, const float* wcimg, int H, cudaStream_t s) {
    using CF = TcPredCfg<NP>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(tc_pred_edge_fwd_kernel<NP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM);
        cudaFuncSetAttribute(tc_pred_edge_fwd_kernel<NP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM);
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = a.g.n_tiles < sms ? a.g.n_tiles : sms;
    if (save) tc_pred_edge_fwd_kernel<NP, true><<<grid, CF::THREADS, CF::SMEM, s>>>(a, w2img, wcimg, H);
    else tc_pred_edge_fwd_kernel<NP, false><<<grid, CF::THREADS, CF::SMEM, s>>>(a, w2img, wcimg, H);
}

void launch_pred_edge_fwd_tc(int H, bool save, const PredEdgeArgs& a, const float* w2img, const float* wcimg, cudaStream_t s) {
    if (a.g.n_tiles <= 0) return;
    switch (tc_np(H)) {
        case 64: launch_fwd_t<64>(save, a, w2img, wcimg, H, s); break;
        case 192: launch_fwd_t<192>(save, a, w2img, wcimg, H, s); break;
        case 208: launch_fwd_t<208>(save, a, w2img, wcimg, H, s); break;
        default: launch_fwd_t<256>(save, a, w2img, wcimg, H, s); break;
    }
}

}  // namespace gb
