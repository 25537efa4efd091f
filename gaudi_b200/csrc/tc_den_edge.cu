// Tensor-core (tcgen05, 3xTF32) version of the denoiser edge kernels -- same contract as den_edge.cu
// (mode 0: GCL edge MLP + attention gate + segment sum, egnn_new.py:42-67; mode 1: EquivariantUpdate, :119-155).
//
// thread = edge row.  Per tile of <=128 compacted edges (complete row segments):
//   workers   build the activation K-atoms  SiLU(Pa[row]+Pb[col]+w_r r+w_d d0)  as hi/lo TF32 halves in swizzled smem
//   warp 1    3 tcgen05.mma per K step into a [128 x NP] fp32 accumulator in TMEM; warp 0 streams the weight atoms (TMA)
//   workers   read the accumulator back (tcgen05.ld): SiLU, row dot with the attention / last coord vector is
//             THREAD-LOCAL (one thread owns one edge), gate, then fixed-order segment sums through a small smem stage.
#include "tc_common.cuh"
#include "kernels.h"

namespace gb {
using namespace tc;

template <int NP>
struct TcEdgeCfg {
    static constexpr int S = 2;
    static constexpr int A_BYTES = 128 * ATOM_ROW_BYTES;
    static constexpr int W_BYTES = NP * ATOM_ROW_BYTES;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
    static constexpr int NPARTS = 4;                         // worker parts of 4 warps; part p owns 16-column chunks ch == p (mod 4)
    static constexpr int NWORK = 128 * NPARTS;
    static constexpr int THREADS = 64 + NWORK;
    static constexpr int MAXCH = (NP + 15) / 16;             // 16-column chunks
    static constexpr int MYCH = (MAXCH + NPARTS - 1) / NPARTS;
    static constexpr int EF_STRIDE = 17;
    static constexpr int SCRATCH = 6 * NP * 4 + NPARTS * 128 * 4 + NPARTS * 128 * EF_STRIDE * 4 + 129 * 4 + 128 * 3 * 4 + 64;
    static constexpr int SMEM = S * STAGE_BYTES + 1024 + 256 + SCRATCH;
    static constexpr int TMEM_COLS = NP <= 64 ? 64 : 256;
};

__device__ __forceinline__ void named_bar(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }

template <int NP, int MODE>
__global__ void __launch_bounds__(TcEdgeCfg<NP>::THREADS, 1) tc_den_edge_kernel(DenEdgeArgs a, const float* __restrict__ wimg, int H) {
    using CF = TcEdgeCfg<NP>;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array: the compiler keeps the address space (LDS / STS
    // instead of generic LD / ST for every staging and operand access)
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + CF::S * CF::STAGE_BYTES);
    uint64_t* full_a = bars; uint64_t* full_w = bars + CF::S; uint64_t* empty = bars + 2 * CF::S;
    uint64_t* d_full = bars + 3 * CF::S; uint64_t* d_empty = d_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d_empty + 1);
    float* vec_s = reinterpret_cast<float*>(base + CF::S * CF::STAGE_BYTES + 256);    // [4][NP]: w_r, w_d, b2, vecw
    float* red_s = vec_s + 6 * NP;                                                       // [2][128]
    float* ef_s = red_s + CF::NPARTS * 128;                                              // [NPARTS][128][17]
    int* seg_s = reinterpret_cast<int*>(ef_s + CF::NPARTS * 128 * CF::EF_STRIDE);        // [129]
    float* tr_s = reinterpret_cast<float*>(seg_s + 129);                                 // [128][3]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < CF::S; ++s) { mbar_init(&full_a[s], 256); mbar_init(&full_w[s], 1); mbar_init(&empty[s], 1); }
        mbar_init(d_full, 1); mbar_init(d_empty, CF::NWORK);
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<CF::TMEM_COLS>(tmem_slot);
    for (int i = tid; i < NP; i += blockDim.x) {
        const bool v = i < H;
        vec_s[i] = v ? a.ext[i] : 0.f; vec_s[NP + i] = v ? a.ext[H + i] : 0.f;
        vec_s[2 * NP + i] = v ? a.b2[i] : 0.f; vec_s[3 * NP + i] = v ? a.vecw[i] : 0.f;
    }
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const Graph& g = a.g;
    const int na = (H + ATOM_K - 1) / ATOM_K;
    const size_t atom_floats = (size_t)2 * NP * ATOM_K;
    constexpr uint32_t idesc = instr_desc_tf32(NP);

    if (warp == 0) {
        if (lane == 0) {
            uint32_t it = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x)
                for (int j = 0; j < na; ++j, ++it) {
                    const uint32_t s = it % CF::S, r = it / CF::S;
                    if (r > 0) mbar_wait(&empty[s], (r - 1) & 1);
                    mbar_arrive_expect_tx(&full_w[s], 2 * CF::W_BYTES);
                    bulk_g2s(base + s * CF::STAGE_BYTES + 2 * CF::A_BYTES, wimg + (size_t)j * atom_floats, 2 * CF::W_BYTES, &full_w[s]);
                }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t it = 0, tcnt = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++tcnt) {
                if (tcnt > 0) mbar_wait(d_empty, (tcnt - 1) & 1);
                fence_after_sync();
                for (int j = 0; j < na; ++j, ++it) {
                    const uint32_t s = it % CF::S, r = it / CF::S;
                    const int kvalid = H - j * ATOM_K;
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
        const int r = group * 32 + lane;                       // tile row == TMEM lane
        const int ht = r;                                      // thread index inside this half (0..127)
        const uint32_t lane_addr = tmem_base + ((uint32_t)(group * 32) << 16);
        const int nchunks = (H + 15) / 16;
        float* my_ef = ef_s + part * 128 * CF::EF_STRIDE;
        uint32_t tcnt = 0;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++tcnt) {
            const int4 ti = __ldg(g.tile_info + tile);
            const int node_lo = ti.x, nn = ti.y, e_lo = ti.z, ne = ti.w;
            const bool valid = r < ne;
            int rown = 0, coln = 0; float rad = 0.f, d0 = 0.f, ux = 0.f, uy = 0.f, uz = 0.f;
            if (valid) {
                const int e = e_lo + r;
                rown = g.erow[e]; coln = g.ecol[e];
                if (a.eattr) { rad = a.eattr[2 * e]; d0 = a.eattr[2 * e + 1]; }
                else {
                    const float dx = a.x[3 * rown] - a.x[3 * coln], dy = a.x[3 * rown + 1] - a.x[3 * coln + 1], dz = a.x[3 * rown + 2] - a.x[3 * coln + 2];
                    rad = dx * dx + dy * dy + dz * dz;
                    const float ex = a.x0[3 * rown] - a.x0[3 * coln], ey = a.x0[3 * rown + 1] - a.x0[3 * coln + 1], ez = a.x0[3 * rown + 2] - a.x0[3 * coln + 2];
                    d0 = ex * ex + ey * ey + ez * ez;
                    if (a.d0_edge) d0 = a.d0_edge[e];
                    if (MODE == 1) { const float inv = 1.f / (sqrtf(rad + 1e-8f) + a.norm_constant); ux = dx * inv; uy = dy * inv; uz = dz * inv; }
                }
                if (MODE == 1 && a.cdiff) { ux = a.cdiff[3 * e]; uy = a.cdiff[3 * e + 1]; uz = a.cdiff[3 * e + 2]; }
            }
            if (part == 0) for (int i = ht; i <= nn; i += 128) seg_s[i] = g.rowptr[node_lo + i] - e_lo;
            // ---- build activation atoms ----
            const float* pa_row = a.P + (size_t)rown * (2 * H);
            const float* pb_row = a.P + (size_t)coln * (2 * H) + H;
            for (int j = part >> 1; j < na; j += 2) {
                const uint32_t it = tcnt * na + j;
                const uint32_t s = it % CF::S, rr = it / CF::S;
                float4 x[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int k0 = j * ATOM_K + 16 * half + 4 * c;
                    x[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid && k0 < H) {
                        const float4 pa = __ldg(reinterpret_cast<const float4*>(pa_row + k0));
                        const float4 pb = __ldg(reinterpret_cast<const float4*>(pb_row + k0));
                        const float4 wr = *reinterpret_cast<const float4*>(vec_s + k0);
                        const float4 wd = *reinterpret_cast<const float4*>(vec_s + NP + k0);
                        x[c].x = silu_f(pa.x + pb.x + wr.x * rad + wd.x * d0);
                        x[c].y = silu_f(pa.y + pb.y + wr.y * rad + wd.y * d0);
                        x[c].z = silu_f(pa.z + pb.z + wr.z * rad + wd.z * d0);
                        x[c].w = silu_f(pa.w + pb.w + wr.w * rad + wd.w * d0);
                    }
                }
                if (rr > 0) mbar_wait(&empty[s], (rr - 1) & 1);
                unsigned char* a_hi = base + s * CF::STAGE_BYTES;
#pragma unroll
                for (int c = 0; c < 4; ++c) store_split(a_hi, a_hi + CF::A_BYTES, r, 4 * half + c, x[c]);
                fence_proxy_async();
                mbar_arrive(&full_a[s]);
            }
            // ---- epilogue ----
            mbar_wait(d_full, tcnt & 1);
            fence_after_sync();
            float m[CF::MYCH][16];
            float psum = 0.f;
#pragma unroll
            for (int ci = 0; ci < CF::MYCH; ++ci) {
                const int ch = part + CF::NPARTS * ci;
                if (ch < nchunks) {
                    tmem_ld16(lane_addr + ch * 16, m[ci]);
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        const int c = ch * 16 + q;
                        const float v = silu_f(m[ci][q] + vec_s[2 * NP + c]);     // padded columns: acc = 0, bias = 0 -> 0
                        m[ci][q] = v;
                        psum = fmaf(vec_s[3 * NP + c], v, psum);
                    }
                }
            }
            fence_before_sync();
            mbar_arrive(d_empty);                               // accumulator is in registers: the next tile's MMAs may start
            red_s[part * 128 + r] = psum;
            named_bar(1, CF::NWORK);
            const float dot = red_s[r] + red_s[128 + r] + red_s[256 + r] + red_s[384 + r];
            if (MODE == 0) {
                const float gate = a.attention ? sigmoid_f(dot + a.att_b) : 1.f;
#pragma unroll
                for (int ci = 0; ci < CF::MYCH; ++ci) {
                    const int ch = part + CF::NPARTS * ci;
                    if (ch < nchunks) {                          // uniform across the half
#pragma unroll
                        for (int q = 0; q < 16; ++q) my_ef[r * CF::EF_STRIDE + q] = m[ci][q] * gate;
                        named_bar(2 + part, 128);
                        for (int nl = ht >> 4; nl < nn; nl += 8) {
                            const int col = ht & 15, c = ch * 16 + col;
                            float sum = 0.f;
                            for (int mm = seg_s[nl]; mm < seg_s[nl + 1]; ++mm) sum += my_ef[mm * CF::EF_STRIDE + col];
                            if (c < H) a.agg[(size_t)(node_lo + nl) * H + c] = sum / a.normf;
                        }
                        named_bar(2 + part, 128);
                    }
                }
            } else {
                if (part == 0) {
                    float sc;
                    if (a.use_tanh) {
                        const float th = tanhf(dot);
                        tr_s[3 * r] = ux * th * a.coords_range; tr_s[3 * r + 1] = uy * th * a.coords_range; tr_s[3 * r + 2] = uz * th * a.coords_range;
                    } else {
                        sc = dot;
                        tr_s[3 * r] = ux * sc; tr_s[3 * r + 1] = uy * sc; tr_s[3 * r + 2] = uz * sc;
                    }
                }
                named_bar(1, CF::NWORK);
                const int wt = part * 128 + ht;
                for (int idx = wt; idx < nn * 3; idx += CF::NWORK) {
                    const int nl = idx / 3, d = idx - 3 * nl;
                    float sum = 0.f;
                    for (int mm = seg_s[nl]; mm < seg_s[nl + 1]; ++mm) sum += tr_s[3 * mm + d];
                    const int node = node_lo + nl;
                    a.x_out[3 * node + d] = (a.x[3 * node + d] + sum / a.normf) * g.node_mask[node];
                }
            }
            named_bar(1, CF::NWORK);                                  // scratch (seg_s, red_s, tr_s) free for the next tile
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc<CF::TMEM_COLS>(tmem_base);
}

template <int NP>
static void launch_t(int mode, const DenEdgeArgs& a, const float* wimg, int H, cudaStream_t s) {
    using CF = TcEdgeCfg<NP>;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(tc_den_edge_kernel<NP, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM);
        cudaFuncSetAttribute(tc_den_edge_kernel<NP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, CF::SMEM);
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = a.g.n_tiles < sms ? a.g.n_tiles : sms;
    if (mode == 0) tc_den_edge_kernel<NP, 0><<<grid, CF::THREADS, CF::SMEM, s>>>(a, wimg, H);
    else tc_den_edge_kernel<NP, 1><<<grid, CF::THREADS, CF::SMEM, s>>>(a, wimg, H);
}

void launch_den_edge_tc(int H, int mode, const DenEdgeArgs& a, const float* wimg, cudaStream_t s) {
    if (a.g.n_tiles <= 0) return;
    switch (tc_np(H)) {
        case 64: launch_t<64>(mode, a, wimg, H, s); break;
        case 192: launch_t<192>(mode, a, wimg, H, s); break;
        case 208: launch_t<208>(mode, a, wimg, H, s); break;
        default: launch_t<256>(mode, a, wimg, H, s); break;
    }
}

}  // namespace gb
