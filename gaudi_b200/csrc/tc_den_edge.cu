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

// Optional clock64 timeline of CTA 0 (make EXTRA=-DGB_TIMELINE; tools/edge_timeline.py): role 0 = worker part 0 / lane 0,
// 1 = worker part 2 / lane 0, 2 = MMA lane, 3 = aux warp 0, 4 = aux warp 1.  Each record = (code, clock).
#ifdef GB_TIMELINE
__device__ unsigned long long gb_tl_den[5][2048];
__device__ unsigned int gb_tl_den_n[5];
#define TLD(role, code) do { if (blockIdx.x == 0) { unsigned int i_ = gb_tl_den_n[role]; if (i_ < 1023) { gb_tl_den[role][2 * i_] = (code); gb_tl_den[role][2 * i_ + 1] = clock64(); gb_tl_den_n[role] = i_ + 1; } } } while (0)
#else
#define TLD(role, code) do {} while (0)
#endif

template <int NP>
struct TcEdgeCfg {
#ifndef GB_DEN_ACC_STRIDE
#define GB_DEN_ACC_STRIDE 256
#endif
    static constexpr int ACC_STRIDE = NP <= 64 ? 64 : GB_DEN_ACC_STRIDE;   // two accumulators: the MMAs of tile k+1 run during the epilogue of tile k
    static constexpr int TMEM_COLS = NP <= 64 ? 128 : 512;
#ifndef GB_DEN_AT
#define GB_DEN_AT 1
#endif
    // activation operand in tensor memory when 64 columns are free behind each accumulator (hidden 192: 2 x (192 + 64) = 512)
    static constexpr bool AT = GB_DEN_AT && ACC_STRIDE - NP >= 64;
    using R = Rings<NP, MIX_BF16, (NP > 208 ? 2 : 3), 2, AT>;
    static constexpr int NPARTS = 4;                         // worker parts of 4 warps; part p owns 16-column chunks ch == p (mod 4)
    static constexpr int NWORK = 128 * NPARTS;
    // two auxiliary warps stage the P rows of every K-atom (even / odd atoms); the first one also prepares the edge geometry
    // of its next tile before that tile's first atom, i.e. while the workers are still in the previous tile's epilogue.
    // (20 warps: register allocation is per 4 warps, a 21st warp would cap every thread at 80 registers)
    static constexpr int AUX_WARP = 2 + NWORK / 32;
    static constexpr int THREADS = 64 + NWORK + 64;
    static constexpr int MAXCH = (NP + 15) / 16;             // 16-column chunks
    static constexpr int MYCH = (MAXCH + NPARTS - 1) / NPARTS;
    static constexpr int EF_STRIDE = 17;                     // (a 20-float pitch with 16-byte stores and two buffers per part = one barrier per chunk: no gain)
    static constexpr int GEO_NF = 7;                         // P-stage row of the row node / of the col node, radial, d0, unit vector (3)
    static constexpr int GEO_WORDS = geo_words(GEO_NF);
    static constexpr int BAR_BYTES = 256;
    static constexpr int NGEO = 3;                           // tile k builds while tile k-1 is in its epilogue and k+1 is being prepared
    static constexpr int SCRATCH = 4 * NP * 4 + 2 * NPARTS * 128 * 4 + NPARTS * 128 * EF_STRIDE * 4 + NGEO * GEO_WORDS * 4 + 2 * 128 * 3 * 4 + PS_BYTES + 64;
    static constexpr int SMEM = R::BYTES + 1024 + BAR_BYTES + SCRATCH;
    static_assert(SMEM <= 232448, "shared memory budget");
};

template <int NP, int MODE>
__global__ void __launch_bounds__(TcEdgeCfg<NP>::THREADS, 1) tc_den_edge_kernel(DenEdgeArgs a, const float* __restrict__ wimg, int H) {
    using CF = TcEdgeCfg<NP>;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array: the compiler keeps the address space (LDS / STS
    // instead of generic LD / ST for every staging and operand access)
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + CF::R::BYTES);
    typename CF::R rg; rg.carve(base, bars);
    uint64_t* d_full = bars + CF::R::NBARS; uint64_t* d_empty = d_full + 2;
    uint64_t* geo_full = d_empty + 2; uint64_t* geo_empty = geo_full + CF::NGEO;
    uint64_t* ps_full = geo_empty + CF::NGEO; uint64_t* ps_empty = ps_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ps_empty + 4);
    float* vec_s = reinterpret_cast<float*>(base + CF::R::BYTES + CF::BAR_BYTES);     // [4][NP]: w_r, w_d, b2, vecw
    float* red_s = vec_s + 4 * NP;                                                       // [2][NPARTS][128] (tile parity)
    float* ef_s = red_s + 2 * CF::NPARTS * 128;                                          // [NPARTS][128][17]
    int* geo_s = reinterpret_cast<int*>(ef_s + CF::NPARTS * 128 * CF::EF_STRIDE);        // [NGEO][GEO_WORDS]
    float* tr_s = reinterpret_cast<float*>(geo_s + CF::NGEO * CF::GEO_WORDS);            // [2][128][3] (tile parity)
    const PStage ps{tr_s + 2 * 128 * 3, ps_full, ps_empty, a.g.ps_rows <= PS_ROWS / 2 ? 2 : 1};  // 4 x 48 or 2 x 96 rows of 36 floats

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        rg.init(256);
        for (int b = 0; b < 2; ++b) { mbar_init(&d_full[b], 1); mbar_init(&d_empty[b], CF::NWORK); }
        for (int b = 0; b < CF::NGEO; ++b) { mbar_init(&geo_full[b], 1); mbar_init(&geo_empty[b], CF::NWORK); }
        ps.init();
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
    if (CF::AT) { rg.set_tmem(0, tmem_base + NP); rg.set_tmem(1, tmem_base + CF::ACC_STRIDE + NP); }
    const Graph& g = a.g;
    const int na = (H + ATOM_K - 1) / ATOM_K;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t wq = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) rg.tma_gemm(wq, na, wimg);
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t gq = 0, wq = 0, tcnt = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++tcnt) {
                const uint32_t ab = tcnt & 1, use = tcnt >> 1;
                if (use > 0) mbar_wait(&d_empty[ab], (use - 1) & 1);
                fence_after_sync();
                TLD(2, 1);
                rg.mma_gemm(gq, wq, na, H, tmem_base + ab * CF::ACC_STRIDE);
                mma_commit(&d_full[ab]);
                TLD(2, 2);
            }
        }
    } else if (warp >= CF::AUX_WARP) {
        // ---- auxiliary warps.  Both stage P rows (warp 0: even atoms, warp 1: odd atoms of the CTA's atom sequence); warp 1 also
        //      keeps the NEXT tile's edge list in registers while this tile's atoms are loaded and writes that tile's geometry block
        //      (header, row-segment table, per-edge P-stage rows / radial / d0 / unit vector) before the workers get there ----
        const int ldw = warp - CF::AUX_WARP;
        auto geo_emit = [&](const TileMeta& m, uint32_t tc) {
            const uint32_t gb_ = tc % CF::NGEO, use = tc / CF::NGEO;
            if (use > 0) mbar_wait(&geo_empty[gb_], (use - 1) & 1);
            int* gi = geo_s + gb_ * CF::GEO_WORDS;
            float* gf = reinterpret_cast<float*>(gi);
            if (lane == 0) { gi[0] = m.node_lo; gi[1] = m.nn; gi[2] = m.e_lo; gi[3] = m.ne; }
            for (int i = lane; i <= m.nn; i += 32) gi[4 + i] = __ldg(g.rowptr + m.node_lo + i) - m.e_lo;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = lane + 32 * q;
                int prow = 0, pcol = 0; float rad = 0.f, d0 = 0.f, ux = 0.f, uy = 0.f, uz = 0.f;
                if (r < m.ne) {
                    const int e = m.e_lo + r, rown = m.row[q], coln = m.col[q];
                    prow = rown - m.node_lo; pcol = m.nn + coln - m.cn_lo;
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
                gi[GEO_HDR + r] = prow; gi[GEO_HDR + 128 + r] = pcol;
                gf[GEO_HDR + 256 + r] = rad; gf[GEO_HDR + 384 + r] = d0;
                if (MODE == 1) { gf[GEO_HDR + 512 + r] = ux; gf[GEO_HDR + 640 + r] = uy; gf[GEO_HDR + 768 + r] = uz; }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&geo_full[gb_]);
        };
        TileMeta cur, nxt;
        int tile = blockIdx.x;
        if (tile < g.n_tiles) {
            tile_meta_load(cur, g, tile, lane, ldw == 1);
            if (ldw == 1) geo_emit(cur, 0);
        }
        for (uint32_t tcnt = 0; tile < g.n_tiles; tile += gridDim.x, ++tcnt) {
            const int ntile = tile + gridDim.x;
            if (ntile < g.n_tiles) tile_meta_load(nxt, g, ntile, lane, ldw == 1);
            for (int j = ldw; j < na; j += 2) {
                pstage_load_atom(ps, tcnt, j, na, H, a.P, cur.node_lo, cur.nn, cur.cn_lo, cur.ncn, lane);
                if (lane == 0) TLD(3 + ldw, 20 + j);
            }
            if (ntile < g.n_tiles) {
                if (ldw == 1) { geo_emit(nxt, tcnt + 1); if (lane == 0) TLD(4, 70); }
                cur = nxt;
            }
        }
    } else {
        const int group = warp & 3, part = (warp - 2) >> 2, half = part & 1;
        const int r = group * 32 + lane;                       // tile row == TMEM lane
        const int ht = r;                                      // thread index inside this part (0..127)
        const uint32_t lane_addr = tmem_base + ((uint32_t)(group * 32) << 16);
        const int nchunks = (H + 15) / 16;
        float* my_ef = ef_s + part * 128 * CF::EF_STRIDE;
        const int tlr = (lane == 0 && group == 0) ? (part == 0 ? 0 : (part == 2 ? 1 : -1)) : -1;
        (void)tlr;
        // ---- operand build of tile number k (the CTA's k-th tile): SiLU(Pa[row] + Pb[col] + w_r r + w_d d0) as hi/lo atoms ----
        auto build = [&](uint32_t k) {
            const uint32_t gb_ = k % CF::NGEO;
            if (tlr >= 0) TLD(tlr, 10);
            mbar_wait(&geo_full[gb_], (k / CF::NGEO) & 1);
            if (tlr >= 0) TLD(tlr, 11);
            const int* gi = geo_s + gb_ * CF::GEO_WORDS;
            const float* gf = reinterpret_cast<const float*>(gi);
            const bool valid = r < gi[3];
            const int pa_off = gi[GEO_HDR + r] * PS_PITCH + 16 * half, pb_off = gi[GEO_HDR + 128 + r] * PS_PITCH + 16 * half;
            const float rad = gf[GEO_HDR + 256 + r], d0 = gf[GEO_HDR + 384 + r];
            for (int j = part >> 1; j < na; j += 2) {
                const float* pst = ps.acquire(k, j, na);
                if (tlr >= 0) TLD(tlr, 20 + j);
                float4 x[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int k0 = j * ATOM_K + 16 * half + 4 * c;
                    x[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (valid && k0 < H) {
                        const float4 pa = *reinterpret_cast<const float4*>(pst + pa_off + 4 * c);
                        const float4 pb = *reinterpret_cast<const float4*>(pst + pb_off + 4 * c);
                        const float4 wr = *reinterpret_cast<const float4*>(vec_s + k0);
                        const float4 wd = *reinterpret_cast<const float4*>(vec_s + NP + k0);
                        x[c].x = silu_f(pa.x + pb.x + wr.x * rad + wd.x * d0);
                        x[c].y = silu_f(pa.y + pb.y + wr.y * rad + wd.y * d0);
                        x[c].z = silu_f(pa.z + pb.z + wr.z * rad + wd.z * d0);
                        x[c].w = silu_f(pa.w + pb.w + wr.w * rad + wd.w * d0);
                    }
                }
                ps.release(k, j, na);                            // P slice consumed (the loads above fed the arithmetic)
                if (tlr >= 0) TLD(tlr, 30 + j);
                if (CF::AT) rg.put_chunk_t(k, j, na, (uint32_t)(group * 32) << 16, half, x);
                else rg.put_chunk(k, j, na, r, half, x);
                if (tlr >= 0) TLD(tlr, 40 + j);
            }
        };
        // Software pipeline over the CTA's tiles: build(0); then per tile k: build(k+1), epilogue(k) -- the tensor pipe works on
        // tile k+1 (second accumulator) while the workers run the epilogue of tile k.
        uint32_t k = 0;
        int tile = blockIdx.x;
        if (tile < g.n_tiles) build(0);
        for (; tile < g.n_tiles; tile += gridDim.x, ++k) {
            if (tile + (int)gridDim.x < g.n_tiles) build(k + 1);
            // ---- epilogue of tile k ----
            const uint32_t gb_ = k % CF::NGEO, ab = k & 1;
            const int* gi = geo_s + gb_ * CF::GEO_WORDS;
            const float* gf = reinterpret_cast<const float*>(gi);
            const int* seg_s = gi + 4;
            const int node_lo = gi[0], nn = gi[1];
            float* red = red_s + ab * CF::NPARTS * 128;
            mbar_wait(&d_full[ab], (k >> 1) & 1);
            if (tlr >= 0) TLD(tlr, 50);
            fence_after_sync();
            float m[CF::MYCH][16];
            float psum = 0.f;
#pragma unroll
            for (int ci = 0; ci < CF::MYCH; ++ci) {
                const int ch = part + CF::NPARTS * ci;
                if (ch < nchunks) {
                    tmem_ld16(lane_addr + ab * CF::ACC_STRIDE + ch * 16, m[ci]);
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
            mbar_arrive(&d_empty[ab]);                          // accumulator is in registers: the MMAs of tile k+2 may overwrite it
            // red / tr are double-buffered by tile parity: a warp can be one tile ahead of another one (it passed the barriers of
            // tile k with it), never two, because it needs that warp's arrival at the barriers of tile k+1
            red[part * 128 + r] = psum;
            if (tlr >= 0) TLD(tlr, 51);
            bar_named(BAR_QUAD + group, 128);                   // the four warps of this TMEM lane quadrant, one per part
            if (tlr >= 0) TLD(tlr, 52);
            const float dot = red[r] + red[128 + r] + red[256 + r] + red[384 + r];
            if (MODE == 0) {
                const float gate = a.attention ? sigmoid_f(dot + a.att_b) : 1.f;
#pragma unroll
                for (int ci = 0; ci < CF::MYCH; ++ci) {
                    const int ch = part + CF::NPARTS * ci;
                    if (ch < nchunks) {                          // uniform across the part
#pragma unroll
                        for (int q = 0; q < 16; ++q) my_ef[r * CF::EF_STRIDE + q] = m[ci][q] * gate;
                        bar_named(BAR_PART + part, 128);
                        for (int nl = ht >> 4; nl < nn; nl += 8) {
                            const int col = ht & 15, c = ch * 16 + col;
                            float sum = 0.f;
                            for (int mm = seg_s[nl]; mm < seg_s[nl + 1]; ++mm) sum += my_ef[mm * CF::EF_STRIDE + col];   // (a 4-way unrolled form with batched loads measured 4-6 % slower)
                            if (c < H) a.agg[(size_t)(node_lo + nl) * H + c] = sum / a.normf;
                        }
                        bar_named(BAR_PART + part, 128);
                    }
                }
            } else {
                float* tr = tr_s + ab * 128 * 3;
                if (part == 0) {
                    const float ux = gf[GEO_HDR + 512 + r], uy = gf[GEO_HDR + 640 + r], uz = gf[GEO_HDR + 768 + r];
                    const float sc = a.use_tanh ? tanhf(dot) * a.coords_range : dot;
                    tr[3 * r] = ux * sc; tr[3 * r + 1] = uy * sc; tr[3 * r + 2] = uz * sc;
                }
                bar_named(BAR_WORKERS, CF::NWORK);
                const int wt = part * 128 + ht;
                for (int idx = wt; idx < nn * 3; idx += CF::NWORK) {
                    const int nl = idx / 3, d = idx - 3 * nl;
                    float sum = 0.f;
                    for (int mm = seg_s[nl]; mm < seg_s[nl + 1]; ++mm) sum += tr[3 * mm + d];
                    const int node = node_lo + nl;
                    a.x_out[3 * node + d] = (a.x[3 * node + d] + sum / a.normf) * g.node_mask[node];
                }
            }
            mbar_arrive(&geo_empty[gb_]);                        // header / segment table of tile k no longer needed
            if (tlr >= 0) TLD(tlr, 60);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc<CF::TMEM_COLS>(tmem_base);
}

#ifdef GB_TIMELINE
extern "C" int gb_debug_timeline_den(unsigned long long* out, unsigned int* n, int reset) {
    if (reset) { unsigned int z[5] = {0, 0, 0, 0, 0}; return (int)cudaMemcpyToSymbol(gb_tl_den_n, z, sizeof(z)); }
    cudaMemcpyFromSymbol(n, gb_tl_den_n, sizeof(gb_tl_den_n));
    return (int)cudaMemcpyFromSymbol(out, gb_tl_den, sizeof(gb_tl_den));
}
#endif

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
