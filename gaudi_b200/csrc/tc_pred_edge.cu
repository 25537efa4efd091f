// Tensor-core (tcgen05, 3xTF32) versions of the predictor E_GCL edge kernels -- same contract and same saved-activation
// layouts as pred_edge.cu (forward: gcl.py:225-279; backward: the hand-written reverse pass replacing autograd).
//
// thread = edge row; 16 worker warps in four parts, part p owns the 16-column chunks ch == p (mod 4) both when building
// A K-atoms (chunk ch = 16-byte chunks 4(ch&1)..+3 of atom ch/2) and in the epilogues, so every atom is built by two
// parts (256 threads).  Two accumulators live in TMEM (columns [0,NP) and [256,256+NP)): the epilogue of GEMM 1 produces the A
// atoms of GEMM 2 on the fly, chunk by chunk, while the MMA warp consumes them.
#include <type_traits>
#include "tc_common.cuh"
#include "kernels.h"

#ifndef GB_PRED_NPARTS
#define GB_PRED_NPARTS 4
#endif
#ifndef GB_BWD_SW
#define GB_BWD_SW 2         // backward: weight half-atom slots
#endif
#ifndef GB_BWD_SVS
#define GB_BWD_SVS 6        // backward: saved-activation ring slots (8 KB each).  5 -> 6 slots: -4 % (the stall attribution of the r2f
                            // profile had 17 % of the samples in svq_acquire); sharing a slot between two 4 KB derivative chunks: no further gain
#endif
#ifndef GB_SV_D
#define GB_SV_D SV_FX       // 16-bit code of the saved SiLU derivatives (tc_common.cuh)
#endif
#ifndef GB_FWD_AT
#define GB_FWD_AT 1         // forward: activation operands in tensor memory (half-atom stages)
#endif
#ifndef GB_BWD_AT
#define GB_BWD_AT 0         // backward: shared-memory activation ring (see TcPredCfg)
#endif
#ifndef GB_BWD_SPLIT_LD
// backward: 1 = the TMEM read of the next chunk is issued one chunk ahead and waited for at the top of the next step (rounds 2a-2e).
// That form is UNSAFE and off: between the issue and the wait the destination registers are in flight, but the compiler does not
// know it and may copy or spill them (it did once register pressure rose in round 2f: 2-7 of 24 000 molecules per run came out
// 1e-5 .. 1e-4 off, different ones each run -- found by the replica test).  Loading and waiting together is also 6 % FASTER here.
#define GB_BWD_SPLIT_LD 0
#endif
#ifndef GB_BWD_GA_TMA
#define GB_BWD_GA_TMA 1     // backward: the tile's g_agg rows arrive by bulk TMA one tile ahead (two blocks of 16 rows) instead of a staging loop
#endif
#ifndef GB_BWD_SA
#define GB_BWD_SA (GB_BWD_GA_TMA ? 2 : 3)         // backward: activation (A operand) ring stages (the third one gave 1 %; the g_agg blocks need its room)
#endif

namespace gb {
using namespace tc;

// Optional clock64 timeline of CTA 0 (make EXTRA=-DGB_TIMELINE; tools/edge_timeline.py): role 0 = worker part 0, 1 = worker part 3
// (lane 0 of the first quadrant), 2 = MMA lane.  Each record = (code, clock).
#ifdef GB_TIMELINE
__device__ unsigned long long gb_tl_pred[5][2048];
__device__ unsigned int gb_tl_pred_n[5];
#define TLP(role, code) do { if (blockIdx.x == 0) { unsigned int i_ = gb_tl_pred_n[role]; if (i_ < 1023) { gb_tl_pred[role][2 * i_] = (code); gb_tl_pred[role][2 * i_ + 1] = clock64(); gb_tl_pred_n[role] = i_ + 1; } } } while (0)
#define TLW(code) do { if (tlr >= 0) TLP(tlr, code); } while (0)
#else
#define TLP(role, code) do {} while (0)
#define TLW(code) do {} while (0)
#endif

template <int NP, bool AT_WANTED>
struct TcPredCfg {
    // bf16 correction terms in both directions: fp16 would be 3x more accurate for O(1) activations, but it has no range for the
    // gradients of the backward and it overflows on the forward too (random-init trajectories reach |x| ~ 1e3, i.e. squared
    // distances and pre-activations beyond 65504: the 1000-step chain test diverged with fp16)
    // activation operands in tensor memory (RingsH, tc_common.cuh) when the two accumulators leave 96 TMEM columns: NP <= 208.
    // Hidden 256 (2 x 256 = 512 columns) keeps the shared-memory activation ring.  Measured at the bench shape: forward 1.113 ->
    // 1.074 ms; backward 1.143 -> 1.210 ms -- its operand builds are fast enough to be paced by the tensor pipe, and three
    // half-atom stages let the four worker parts run only 1.5 atoms ahead instead of 3, so the backward keeps the shared-memory
    // ring (GB_BWD_AT=1 selects the tested tensor-memory form).
    static constexpr bool AT = AT_WANTED && 2 * NP + 96 <= 512;
    using RS = Rings<NP, MIX_BF16>;
    using RSB = Rings<NP, MIX_BF16, (NP > 208 ? 2 : GB_BWD_SW), (NP > 208 ? 2 : GB_BWD_SA)>;    // NP = 256: two + two is what fits
    using RH = RingsH<NP, MIX_BF16>;
    using R = typename std::conditional<AT, RH, RS>::type;
    using RB = typename std::conditional<AT, RH, RSB>::type;
    static constexpr int NPARTS = GB_PRED_NPARTS;                          // worker parts of 4 warps; part p owns 16-column chunks ch == p (mod 4)
    static constexpr int NWORK = 128 * NPARTS;
    static constexpr int THREADS = 64 + NWORK;
    static constexpr int MAXCH = (NP + 15) / 16;
    static constexpr int MYCH = (MAXCH + NPARTS - 1) / NPARTS;
    static constexpr int EF_STRIDE = 17;
    static constexpr int BAR_BYTES = 512;                                  // up to 64 mbarriers + the TMEM address slot
    static constexpr int D2_COL = AT ? NP + RH::TMEM_COLS : 256;             // accumulator 2 (AT: behind the three activation stages)
    // backward only: one extra producer warp streams the saved activations (16-column chunks of the plane layouts of
    // tc_common.cuh: 2 planes x 128 rows x 16 B = 4 KB for the 16-bit derivative codes, 4 planes = 8 KB for the fp32 pre2,
    // contiguous in HBM) through a ring of 8 KB slots (more than one chunk per worker part in flight: with one slot per part the
    // HBM latency of every chunk was exposed -- ~1.5 us on each of the 13 chunk steps of a part and tile)
    static constexpr int SV_SLOT_BYTES = 4 * SV_PLANE_BYTES;
};

// saved-activation ring.  `round[s]` = number of the round (q / slots) whose chunk the producer last started to load into slot s.
// A consumer checks it before waiting on full[s]: the occupants of a slot alternate between two worker parts when the slot count
// is not a multiple of the part count, so a part can reach its wait for round r while the producer has not even started round
// r-1 of that slot -- and an mbarrier parity wait cannot tell "round r" from "round r-2" (it would pass on stale data and then
// release the slot a second time; this was a real, timing-dependent deadlock with 6 slots and 4 parts).
struct SvRing { unsigned char* buf; uint64_t* full; uint64_t* empty; volatile uint32_t* round; };

template <int NPARTS>
__device__ __forceinline__ float psum_parts(const float* red, int r) {
    float s = 0.f;
#pragma unroll
    for (int p = 0; p < NPARTS; ++p) s += red[p * 128 + r];
    return s;
}

// ==================================================================================================================
// forward
// ==================================================================================================================
// Warps: 0 = weight TMA, 1 = MMA issue, 2..17 = workers, 18 / 19 = auxiliary (staged P rows of the even / odd K-atoms of GEMM 1;
// warp 19 also prepares the next tile's edge geometry).  Worker schedule, software-pipelined over the CTA's tiles:
//     build1(0);  for k: epilogue1(k) [= operand of GEMM 2(k)], build1(k+1), epilogue2(k)
// so that GEMM 1 of tile k+1 runs on the tensor pipe while the workers are in epilogue 2 of tile k, and GEMM 2 of tile k while they
// build tile k+1 (accumulator 1 is free again once epilogue 1 has read it, accumulator 2 once epilogue 2 has).
template <int NP>
struct TcFwdCfg : TcPredCfg<NP, GB_FWD_AT != 0> {
    using B = TcPredCfg<NP, GB_FWD_AT != 0>;
    static constexpr int AUX_WARP = 2 + 4 * B::NPARTS;
    static constexpr int THREADS = B::THREADS + 64;
    static constexpr int GEO_NF = 6;                                       // P-stage rows (row | col << 16), radial, edge_attr, unit vector (3)
    static constexpr int GEO_NODE = geo_words(GEO_NF);                     // then per row node (up to 128): x (3), node mask -- read on the tile's tail
    static constexpr int GEO_WORDS = GEO_NODE + 4 * 128;
    static constexpr int NGEO = 3;
    static constexpr int SCRATCH = 6 * NP * 4 + 2 * B::NPARTS * 128 * 4 + B::NPARTS * 128 * B::EF_STRIDE * 4 + NGEO * GEO_WORDS * 4 + 128 * 3 * 4 + PS_BYTES + 64;
    static constexpr int SMEM = B::R::BYTES + 1024 + B::BAR_BYTES + SCRATCH;
    static_assert(SMEM <= 232448, "shared memory budget (forward)");
};

template <int NP, bool SAVE>
__global__ void __launch_bounds__(TcFwdCfg<NP>::THREADS, 1) tc_pred_edge_fwd_kernel(PredEdgeArgs a, const float* __restrict__ w2img,
                                                                   const float* __restrict__ wcimg, int H) {
    using CF = TcFwdCfg<NP>;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array: the compiler keeps the address space (LDS / STS
    // instead of generic LD / ST for every staging and operand access)
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + CF::R::BYTES);
    typename CF::R rg;
    if constexpr (CF::AT) rg.carve(base, bars, reinterpret_cast<volatile uint32_t*>(base + CF::R::BYTES + 448));   // built[3], inside the barrier block
    else rg.carve(base, bars);
    uint64_t* d1_full = bars + CF::R::NBARS; uint64_t* d2_full = d1_full + 1; uint64_t* d1_empty = d2_full + 1; uint64_t* d2_empty = d1_empty + 1;
    uint64_t* geo_full = d2_empty + 1; uint64_t* geo_empty = geo_full + CF::NGEO;
    uint64_t* ps_full = geo_empty + CF::NGEO; uint64_t* ps_empty = ps_full + 4;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ps_empty + 4);
    float* vec_s = reinterpret_cast<float*>(base + CF::R::BYTES + CF::BAR_BYTES);     // [6][NP]: w_r, w_a, b2, att_w, bc, wc_last
    float* red_s = vec_s + 6 * NP;                                                       // [2][NPARTS][128]: gate logits | coordinate head
    float* ef_s = red_s + 2 * CF::NPARTS * 128;                                          // [NPARTS][128][17]
    int* geo_s = reinterpret_cast<int*>(ef_s + CF::NPARTS * 128 * CF::EF_STRIDE);        // [NGEO][GEO_WORDS]
    float* tr_s = reinterpret_cast<float*>(geo_s + CF::NGEO * CF::GEO_WORDS);            // [128][3]
    const PStage ps{tr_s + 128 * 3, ps_full, ps_empty, a.g.ps_rows <= PS_ROWS / 2 ? 2 : 1};

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        if constexpr (CF::AT) rg.init(); else rg.init(256);
        mbar_init(d1_full, 1); mbar_init(d2_full, 1); mbar_init(d1_empty, CF::NWORK); mbar_init(d2_empty, CF::NWORK);
        for (int b = 0; b < CF::NGEO; ++b) { mbar_init(&geo_full[b], 1); mbar_init(&geo_empty[b], CF::NWORK); }
        ps.init();
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
    if constexpr (CF::AT) rg.a_tm0 = tmem_base + NP;
    const Graph& g = a.g;
    const int na = (H + ATOM_K - 1) / ATOM_K;
    const int npl = sv_planes(H);                                // 8-column planes of the 16-bit saved activations
    (void)npl;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t wq = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) { rg.tma_gemm(wq, na, w2img); rg.tma_gemm(wq, na, wcimg); }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            uint32_t gq = 0, wq = 0, tcnt = 0;
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++tcnt) {
                if (tcnt > 0) mbar_wait(d1_empty, (tcnt - 1) & 1);
                fence_after_sync();
                TLP(2, 1);
                rg.mma_gemm(gq, wq, na, H, tmem_base);
                mma_commit(d1_full);
                TLP(2, 2);
                if (tcnt > 0) mbar_wait(d2_empty, (tcnt - 1) & 1);
                fence_after_sync();
                TLP(2, 3);
                rg.mma_gemm(gq, wq, na, H, tmem_base + CF::D2_COL);
                mma_commit(d2_full);
                TLP(2, 4);
            }
        }
    } else if (warp >= CF::AUX_WARP) {
        const int ldw = warp - CF::AUX_WARP;
        auto geo_emit = [&](const TileMeta& m, uint32_t tc) {
            const uint32_t gb_ = tc % CF::NGEO, use = tc / CF::NGEO;
            if (use > 0) mbar_wait(&geo_empty[gb_], (use - 1) & 1);
            int* gi = geo_s + gb_ * CF::GEO_WORDS;
            float* gf = reinterpret_cast<float*>(gi);
            if (lane == 0) { gi[0] = m.node_lo; gi[1] = m.nn; gi[2] = m.e_lo; gi[3] = m.ne; }
            for (int i = lane; i <= m.nn; i += 32) gi[4 + i] = __ldg(g.rowptr + m.node_lo + i) - m.e_lo;
            for (int i = lane; i < m.nn; i += 32) {
                const int node = m.node_lo + i;
                gf[CF::GEO_NODE + 4 * i] = a.x[3 * node]; gf[CF::GEO_NODE + 4 * i + 1] = a.x[3 * node + 1]; gf[CF::GEO_NODE + 4 * i + 2] = a.x[3 * node + 2];
                gf[CF::GEO_NODE + 4 * i + 3] = g.node_mask[node];
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = lane + 32 * q;
                int prc = 0; float rad = 0.f, a0 = 0.f, ux = 0.f, uy = 0.f, uz = 0.f;
                if (r < m.ne) {
                    const int e = m.e_lo + r, rown = m.row[q], coln = m.col[q];
                    prc = (rown - m.node_lo) | ((m.nn + coln - m.cn_lo) << 16);
                    const float dx = a.x[3 * rown] - a.x[3 * coln], dy = a.x[3 * rown + 1] - a.x[3 * coln + 1], dz = a.x[3 * rown + 2] - a.x[3 * coln + 2];
                    rad = dx * dx + dy * dy + dz * dz;
                    const float inv = 1.f / (sqrtf(rad + 1e-8f) + 1.f);
                    ux = dx * inv; uy = dy * inv; uz = dz * inv;
                    const float ex = a.x0[3 * rown] - a.x0[3 * coln], ey = a.x0[3 * rown + 1] - a.x0[3 * coln + 1], ez = a.x0[3 * rown + 2] - a.x0[3 * coln + 2];
                    a0 = ex * ex + ey * ey + ez * ez;
                    if (a.a_edge) a0 = a.a_edge[e];
                }
                gi[GEO_HDR + r] = prc; gf[GEO_HDR + 128 + r] = rad; gf[GEO_HDR + 256 + r] = a0;
                gf[GEO_HDR + 384 + r] = ux; gf[GEO_HDR + 512 + r] = uy; gf[GEO_HDR + 640 + r] = uz;
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
            for (int j = ldw; j < na; j += 2) pstage_load_atom(ps, tcnt, j, na, H, a.P, cur.node_lo, cur.nn, cur.cn_lo, cur.ncn, lane);
            if (ntile < g.n_tiles) {
                if (ldw == 1) geo_emit(nxt, tcnt + 1);
                cur = nxt;
            }
        }
    } else {
        const int group = warp & 3, part = (warp - 2) >> 2, half = part & 1;
        const int r = group * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(group * 32) << 16);
        const int nchunks = (H + 15) / 16, nfull = H >> 4;       // all chunks / chunks whose 16 columns are all real
        const uint32_t lane_off = (uint32_t)(group * 32) << 16;
        const bool leader = group == 0 && lane == 0;
        (void)lane_off; (void)leader;
#ifndef GB_FWD_ROT
#define GB_FWD_ROT 2
#endif
        // chunk ownership of the epilogues: part p owns the chunks ch == pv (mod 4).  In the GEMM-1 operand build the extra (7th) atom
        // belongs to parts 0 / 1 (the pairs are tied to the staged P rows); rotating the epilogues' ownership by two gives the extra 13th
        // chunk to part 2, so that no part is the slow one in every phase.  (Only with the operands in tensor memory: the shared-memory
        // ring ties a chunk's atom to a pair.)
        const int pv = CF::AT ? ((part + GB_FWD_ROT) & (CF::NPARTS - 1)) : part;
        float* my_ef = ef_s + part * 128 * CF::EF_STRIDE;
        float* red1 = red_s; float* red2 = red_s + CF::NPARTS * 128;
        const int tlr = (lane == 0 && group == 0) ? (part == 0 ? 0 : (part == 3 ? 1 : -1)) : -1;
        (void)tlr;
        // ---- GEMM 1 operand of the CTA's k-th tile: s1 = SiLU(pre1) from the staged P rows; SiLU'(pre1) is saved ----
        auto build1 = [&](uint32_t k, int tile) {
            const uint32_t gb_ = k % CF::NGEO;
            TLW(10);
            mbar_wait(&geo_full[gb_], (k / CF::NGEO) & 1);
            TLW(11);
            const int* gi = geo_s + gb_ * CF::GEO_WORDS;
            const float* gf = reinterpret_cast<const float*>(gi);
            const bool valid = r < gi[3];
            const int prc = gi[GEO_HDR + r];
            const int pa_off = (prc & 0xffff) * PS_PITCH + 16 * half, pb_off = (prc >> 16) * PS_PITCH + 16 * half;
            const float rad = gf[GEO_HDR + 128 + r], a0 = gf[GEO_HDR + 256 + r];
            for (int j = part >> 1; j < na; j += CF::NPARTS / 2) {
                const float* pst = ps.acquire(k, j, na);
                float4 x[4];
#pragma unroll
                for (int pp = 0; pp < 2; ++pp) {                    // one 8-column plane of the saved SiLU'(pre1) per pass
                    float dv8[8];
#pragma unroll
                    for (int c2 = 0; c2 < 2; ++c2) {
                        const int c = 2 * pp + c2;
                        const int k0 = j * ATOM_K + 16 * half + 4 * c;
                        float4 v = make_float4(0.f, 0.f, 0.f, 0.f), dv = v;
                        if (valid && k0 < H) {
                            const float4 pa = *reinterpret_cast<const float4*>(pst + pa_off + 4 * c);
                            const float4 pb = *reinterpret_cast<const float4*>(pst + pb_off + 4 * c);
                            const float4 wr = *reinterpret_cast<const float4*>(vec_s + k0);
                            const float4 wa = *reinterpret_cast<const float4*>(vec_s + NP + k0);
                            silu_both(pa.x + pb.x + wr.x * rad + wa.x * a0, v.x, dv.x);
                            silu_both(pa.y + pb.y + wr.y * rad + wa.y * a0, v.y, dv.y);
                            silu_both(pa.z + pb.z + wr.z * rad + wa.z * a0, v.z, dv.z);
                            silu_both(pa.w + pb.w + wr.w * rad + wa.w * a0, v.w, dv.w);
                        }
                        x[c] = v;
                        dv8[4 * c2] = dv.x; dv8[4 * c2 + 1] = dv.y; dv8[4 * c2 + 2] = dv.z; dv8[4 * c2 + 3] = dv.w;
                    }
                    const int plane = 4 * j + 2 * half + pp;
                    if (SAVE && plane < npl) sv_store8<GB_SV_D>(a.sv_d1, tile, npl, plane, r, dv8);
                }
                ps.release(k, j, na);
                TLW(20 + j);
                if constexpr (CF::AT) { if (2 * j + half < nchunks) rg.put(2 * k, 2 * j + half, nchunks, lane_off, leader, x); }
                else rg.put_chunk(2 * k, j, na, r, half, x);
                TLW(30 + j);
            }
        };
        uint32_t k = 0;
        int tile = blockIdx.x;
        if (tile < g.n_tiles) build1(0, tile);
        for (; tile < g.n_tiles; tile += gridDim.x, ++k) {
            const uint32_t gb_ = k % CF::NGEO;
            const int* gi = geo_s + gb_ * CF::GEO_WORDS;
            const float* gf = reinterpret_cast<const float*>(gi);
            const int* seg_s = gi + 4;
            const int node_lo = gi[0], nn = gi[1], e_lo = gi[2], ne = gi[3];
            const bool valid = r < ne;
            // ---- epilogue 1: q = SiLU(pre2), attention gate ----
            TLW(40);
            mbar_wait(d1_full, k & 1);
            TLW(41);
            fence_after_sync();
            float q[CF::MYCH][16];
            float psum = 0.f;
            // Two code paths per chunk: a chunk whose 16 columns are all real runs without per-column tests; the last chunk of a
            // hidden width that is not a multiple of 16 (4 real columns of 16 at H = 196) only touches its real 4-column groups.
            // The part that owns that extra chunk is the critical path of every chunk phase (4 chunks against 3), and runtime tests
            // inside ALL chunks cost more than they save (+7 % when tried: they break the scheduling of the unrolled loops).
            auto e1_chunk = [&](auto tail_t, float (&qr)[16], int ch) {
                constexpr bool TAIL = decltype(tail_t)::value;
                tmem_ld16(lane_addr + ch * 16, qr);
#pragma unroll
                for (int c4 = 0; c4 < 4; ++c4) {
                    const int c0 = ch * 16 + 4 * c4;
                    if (!TAIL || c0 < H) {
                        float pre[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            pre[e] = qr[4 * c4 + e] + vec_s[2 * NP + c0 + e];
                            const float v = silu_f(pre[e]);
                            qr[4 * c4 + e] = v;
                            psum = fmaf(vec_s[3 * NP + c0 + e], v, psum);
                        }
                        if (SAVE)                    // pre2 stays fp32: [tile][k / 4][128 rows][4]
                            *reinterpret_cast<float4*>(a.sv_pre2 + (((size_t)tile * (H / 4) + (c0 >> 2)) * 128 + r) * 4) = make_float4(pre[0], pre[1], pre[2], pre[3]);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e) qr[4 * c4 + e] = 0.f;
                    }
                }
            };
#pragma unroll
            for (int ci = 0; ci < CF::MYCH; ++ci) {
                const int ch = pv + CF::NPARTS * ci;
                if (ch < nfull) e1_chunk(std::false_type{}, q[ci], ch);
                else if (ch < nchunks) e1_chunk(std::true_type{}, q[ci], ch);
            }
            fence_before_sync();
            mbar_arrive(d1_empty);                              // accumulator 1 is in registers: GEMM 1 of the next tile may start
            // red1 / red2 alternate (gate logits of tile k, coordinate head of tile k, gate logits of tile k+1, ...): a warp is never
            // more than one quadrant barrier ahead of the warps it shares the rows with, so one copy of each is enough
            red1[part * 128 + r] = psum;
            TLW(42);
            bar_named(BAR_QUAD + group, 128);
            TLW(43);
            const float gate = a.attention ? sigmoid_f(psum_parts<CF::NPARTS>(red1, r) + a.att_b) : 1.f;
            // ---- gated edge feature: segment sums -> agg, and operand atoms of GEMM 2 ----
            auto g_chunk = [&](auto tail_t, const float (&qr)[16], int ch) {
                constexpr bool TAIL = decltype(tail_t)::value;
                float4 x[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    x[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!TAIL || ch * 16 + 4 * c < H) {
                        x[c] = make_float4(qr[4 * c] * gate, qr[4 * c + 1] * gate, qr[4 * c + 2] * gate, qr[4 * c + 3] * gate);
                        if (!valid) x[c] = make_float4(0.f, 0.f, 0.f, 0.f);
                        my_ef[r * CF::EF_STRIDE + 4 * c] = x[c].x; my_ef[r * CF::EF_STRIDE + 4 * c + 1] = x[c].y;
                        my_ef[r * CF::EF_STRIDE + 4 * c + 2] = x[c].z; my_ef[r * CF::EF_STRIDE + 4 * c + 3] = x[c].w;
                    }
                }
                if constexpr (CF::AT) rg.put(2 * k + 1, ch, nchunks, lane_off, leader, x);
                else rg.put_chunk(2 * k + 1, ch >> 1, na, r, half, x);
                bar_named(BAR_PART + part, 128);
                if (!TAIL || ch * 16 + (r & 15) < H) {
                    for (int nl = r >> 4; nl < nn; nl += 8) {
                        const int col = r & 15, c = ch * 16 + col;
                        float sum = 0.f;
                        for (int mm = seg_s[nl]; mm < seg_s[nl + 1]; ++mm) sum += my_ef[mm * CF::EF_STRIDE + col];   // (a 4-way unrolled form with batched loads measured 4-6 % slower)
                        a.agg[(size_t)(node_lo + nl) * H + c] = sum;
                    }
                }
                bar_named(BAR_PART + part, 128);
            };
#pragma unroll
            for (int ci = 0; ci < CF::MYCH; ++ci) {
                const int ch = pv + CF::NPARTS * ci;
                if (ch < nfull) g_chunk(std::false_type{}, q[ci], ch);
                else if (ch < nchunks) g_chunk(std::true_type{}, q[ci], ch);
                else if (!CF::AT && ch < 2 * na) {
                    // chunk beyond the hidden width but inside the last atom: publish zeros so the atom completes
                    // (the half-atom stages of the AT form have no such chunk)
                    float4 x[4] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
                    if constexpr (!CF::AT) rg.put_chunk(2 * k + 1, ch >> 1, na, r, half, x);
                }
            }
            // ---- GEMM 1 operand of the next tile (its MMAs overlap epilogue 2 below) ----
            TLW(50);
            if (tile + (int)gridDim.x < g.n_tiles) build1(k + 1, tile + gridDim.x);
            // ---- epilogue 2: coordinate head ----
            TLW(60);
            mbar_wait(d2_full, k & 1);
            TLW(61);
            fence_after_sync();
            float phi_part = 0.f;
            auto e2_chunk = [&](auto tail_t, int ch) {
                constexpr bool TAIL = decltype(tail_t)::value;
                float v[16];
                tmem_ld16(lane_addr + CF::D2_COL + ch * 16, v);
#pragma unroll
                for (int pp = 0; pp < 2; ++pp) {
                    const int c0 = ch * 16 + 8 * pp;
                    if (TAIL && c0 >= H) break;
                    float d3[8];
#pragma unroll
                    for (int hq = 0; hq < 2; ++hq) {
                        if (!TAIL || c0 + 4 * hq < H) {
#pragma unroll
                            for (int e = 4 * hq; e < 4 * hq + 4; ++e) {
                                float s3;
                                silu_both(v[8 * pp + e] + vec_s[4 * NP + c0 + e], s3, d3[e]);
                                phi_part = fmaf(vec_s[5 * NP + c0 + e], s3, phi_part);
                            }
                        } else {
#pragma unroll
                            for (int e = 4 * hq; e < 4 * hq + 4; ++e) d3[e] = 0.f;
                        }
                    }
                    if (SAVE) sv_store8<GB_SV_D>(a.sv_d3, tile, npl, 2 * ch + pp, r, d3);
                }
            };
#pragma unroll 1
            for (int ch = pv; ch < nfull; ch += CF::NPARTS) e2_chunk(std::false_type{}, ch);
            if (nfull < nchunks && pv == (nfull & (CF::NPARTS - 1))) e2_chunk(std::true_type{}, nfull);
            fence_before_sync();
            mbar_arrive(d2_empty);
            red2[part * 128 + r] = phi_part;
            TLW(62);
            bar_named(BAR_QUAD + group, 128);
            TLW(63);
            if (part == 0) {
                const float phi = psum_parts<CF::NPARTS>(red2, r);
                const float tau = a.use_tanh ? tanhf(phi) : phi;
                if (SAVE && valid) a.sv_tau[e_lo + r] = tau;
                const float sc = a.use_tanh ? tau * a.coords_range : tau;
                tr_s[3 * r] = gf[GEO_HDR + 384 + r] * sc; tr_s[3 * r + 1] = gf[GEO_HDR + 512 + r] * sc; tr_s[3 * r + 2] = gf[GEO_HDR + 640 + r] * sc;
            }
            // tr_s is single-buffered: its next writer has to pass build1 of tile k+2, whose ring slots are only released by
            // MMAs that consumed GEMM-2 atoms of tile k+1 from every part, i.e. after every warp has left this tile
            bar_named(BAR_WORKERS, CF::NWORK);
            for (int idx = part * 128 + r; idx < nn * 3; idx += CF::NWORK) {
                const int nl = idx / 3, d = idx - 3 * nl;
                float sum = 0.f;
                for (int mm = seg_s[nl]; mm < seg_s[nl + 1]; ++mm) sum += tr_s[3 * mm + d];
                const int node = node_lo + nl;
                a.x_out[3 * node + d] = (gf[CF::GEO_NODE + 4 * nl + d] + sum) * gf[CF::GEO_NODE + 4 * nl + 3];     // x and mask came with the geometry block
            }
            mbar_arrive(&geo_empty[gb_]);
            TLW(70);
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

// ==================================================================================================================
// backward (input gradient only)
// ==================================================================================================================
// Warps: 0 = weight TMA, 1 = MMA issue, 2..17 = workers, 18 = saved-activation TMA, 19 = edge geometry of the next tiles.
// Worker schedule, software-pipelined over the CTA's tiles (same idea as the forward kernel):
//     build1(0);  for k: epilogue1(k), build2(k) [operand of GEMM 2(k)], build1(k+1) [operand of GEMM 1(k+1)], epilogue2(k)
// GEMM 1 of tile k+1 runs while the workers are in epilogue 2 of tile k; accumulator 1 (which also parks g_ef between epilogue 1
// and build2) is released after build2(k), accumulator 2 after epilogue2(k).
template <int NP>
struct TcBwdCfg : TcPredCfg<NP, GB_BWD_AT != 0> {
    using B = TcPredCfg<NP, GB_BWD_AT != 0>;
    static constexpr int SV_WARP = 2 + 4 * B::NPARTS, GEO_WARP = SV_WARP + 1;
    static constexpr int THREADS = B::THREADS + 64;
    static constexpr int GEO_NF = 9;                                       // local row node, g_phi, d (3), g_u (3), g_attr so far
    static constexpr int GEO_WORDS = geo_words(GEO_NF);
    static constexpr int SV_SLOTS = B::AT ? 7 : (NP > 208 ? 6 : GB_BWD_SVS);  // saved-activation ring: slots of 8 KB
    static constexpr int STG_WARP_FLOATS = 32 * 8;                         // per-warp transpose block of epilogue 2: 32 rows x 8 columns
    static constexpr int GA_SCRATCH_ROWS = B::AT ? 32 : 0;                 // AT form: staged g_agg rows of a tile (else they borrow the A ring)
    // g_agg rows of the coming tile by bulk TMA (SV warp) into one of two blocks: tiles with more row nodes read g_agg from L2
    static constexpr bool GA_TMA = GB_BWD_GA_TMA && !B::AT && NP <= 208;
#ifndef GB_GA_TMA_ROWS
#define GB_GA_TMA_ROWS 15          // (15 rows of NP floats per block leave room for the sixth 8 KB slot of the saved-activation ring)
#endif
    static constexpr int GA_TMA_ROWS = GB_GA_TMA_ROWS;
    static constexpr int GA_BLOCK_BYTES = GA_TMA_ROWS * NP * 4;
    static constexpr int SCRATCH = 6 * NP * 4 + 2 * B::NPARTS * 128 * 4 + SV_SLOTS * B::SV_SLOT_BYTES + (B::NWORK / 32) * STG_WARP_FLOATS * 4 +
                                   2 * GEO_WORDS * 4 + GA_SCRATCH_ROWS * NP * 4 + (GA_TMA ? 2 * GA_BLOCK_BYTES : 0) + 64;
    static constexpr int SMEM = B::RB::BYTES + 1024 + B::BAR_BYTES + SCRATCH;
    static_assert(SMEM <= 232448, "shared memory budget (backward)");
};

template <int SLOTS>
__device__ __forceinline__ const uint4* svq_acquire(const SvRing& sv, uint32_t q, int r) {
    const uint32_t s = q % SLOTS, rr = q / SLOTS;
#ifdef GB_DEBUG_HANG
    for (long long spin = 0; sv.round[s] != rr; ++spin) {
        if (spin > 3000000) {
            volatile int* hb = gb_hang_buf;
            if (hb && (threadIdx.x & 31) == 0 && blockIdx.x < 24) {
                const int i = atomicAdd((int*)hb, 1);
                if (i < 1000) { hb[4 + 4 * i] = blockIdx.x; hb[5 + 4 * i] = threadIdx.x; hb[6 + 4 * i] = 0x1000000 + (int)q; hb[7 + 4 * i] = (int)sv.round[s]; }
                __threadfence_system();
            }
            for (int w = 0; w < 3000000; ++w) __nanosleep(1000);
            __trap();
        }
    }
#else
    while (sv.round[s] != rr) { }
#endif
    mbar_wait(&sv.full[s], rr & 1);
    return reinterpret_cast<const uint4*>(sv.buf + s * (4 * SV_PLANE_BYTES)) + r;
}
template <int SLOTS>
__device__ __forceinline__ void svq_release(const SvRing& sv, uint32_t q) { mbar_arrive(&sv.empty[q % SLOTS]); }

template <int NP>
__global__ void __launch_bounds__(TcBwdCfg<NP>::THREADS, 1) tc_pred_edge_bwd_kernel(PredEdgeArgs a, const float* __restrict__ wcimg_nt,
                                                                   const float* __restrict__ w2img_nt, int H) {
    using CF = TcBwdCfg<NP>;
    constexpr int SVS = CF::SV_SLOTS;
    extern __shared__ unsigned char smem_raw[];
    // 1024-byte alignment by pointer arithmetic on the __shared__ array: the compiler keeps the address space (LDS / STS
    // instead of generic LD / ST for every staging and operand access)
    unsigned char* base = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* bars = reinterpret_cast<uint64_t*>(base + CF::RB::BYTES);
    typename CF::RB rg;
    if constexpr (CF::AT) rg.carve(base, bars, reinterpret_cast<volatile uint32_t*>(base + CF::RB::BYTES + 448));   // built[3], inside the barrier block
    else rg.carve(base, bars);
    uint64_t* d1_full = bars + CF::RB::NBARS; uint64_t* d2_full = d1_full + 1; uint64_t* d1_empty = d2_full + 1; uint64_t* d2_empty = d1_empty + 1;
    uint64_t* sv_full = d2_empty + 1; uint64_t* sv_empty = sv_full + SVS;
    uint64_t* geo_full = sv_empty + SVS; uint64_t* geo_empty = geo_full + 2;
    uint64_t* ga_full = geo_empty + 2; uint64_t* ga_empty = ga_full + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ga_empty + 2);
    volatile uint32_t* sv_round = reinterpret_cast<volatile uint32_t*>(base + CF::RB::BYTES + 384);     // [SV_SLOTS], inside the barrier block
    float* vec_s = reinterpret_cast<float*>(base + CF::RB::BYTES + CF::BAR_BYTES);     // [6][NP]: w_r, w_a, -, att_w, -, wc_last
    float* red_s = vec_s + 6 * NP;                                                       // [2][NPARTS][128]
    unsigned char* sv_buf = reinterpret_cast<unsigned char*>(red_s + 2 * CF::NPARTS * 128);   // saved-activation ring
    float* stg_s = reinterpret_cast<float*>(sv_buf + SVS * CF::SV_SLOT_BYTES);           // [16 warps][32][8]
    int* geo_s = reinterpret_cast<int*>(stg_s + (CF::NWORK / 32) * CF::STG_WARP_FLOATS);  // [2][GEO_WORDS]
    const SvRing sv{sv_buf, sv_full, sv_empty, sv_round};
    unsigned char* ga_scratch;
    if constexpr (CF::AT || CF::GA_TMA) ga_scratch = reinterpret_cast<unsigned char*>(geo_s + 2 * CF::GEO_WORDS);
    else ga_scratch = rg.a_base;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        if constexpr (CF::AT) rg.init(); else rg.init(256);
        mbar_init(d1_full, 1); mbar_init(d2_full, 1); mbar_init(d1_empty, CF::NWORK); mbar_init(d2_empty, CF::NWORK);
        for (int s = 0; s < SVS; ++s) { mbar_init(&sv_full[s], 1); mbar_init(&sv_empty[s], 128); sv_round[s] = 0xffffffffu; }
        for (int b = 0; b < 2; ++b) { mbar_init(&geo_full[b], 1); mbar_init(&geo_empty[b], CF::NWORK); }
        for (int b = 0; b < 2; ++b) { mbar_init(&ga_full[b], 1); mbar_init(&ga_empty[b], CF::NWORK); }
        mbar_fence_init();
    }
    if (warp == 1) tmem_alloc<512>(tmem_slot);
    for (int i = tid; i < NP; i += blockDim.x) {
        const bool v = i < H;
        vec_s[i] = v ? a.ext[i] : 0.f; vec_s[NP + i] = v ? a.ext[H + i] : 0.f;
        vec_s[3 * NP + i] = v ? a.att_w[i] : 0.f; vec_s[5 * NP + i] = v ? a.wc_last[i] : 0.f;
    }
    // (the last 16-column chunk of a tile may hold one plane only: sv_load8 reads the missing plane as zeros instead of the slot's
    // previous occupant, so a NaN saved for another tile can never leak into this one through a zero-padded weight)
    fence_before_sync();
    __syncthreads();
    fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    if constexpr (CF::AT) rg.a_tm0 = tmem_base + NP;
    const Graph& g = a.g;
    const int na = (H + ATOM_K - 1) / ATOM_K;
    const int my_tiles = (int)blockIdx.x < g.n_tiles ? (g.n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const int npl = sv_planes(H);                                // 8-column planes of the 16-bit saved activations

    if (warp == 0) {
        if (lane == 0 && my_tiles > 0) {
            uint32_t wq = 0;
            rg.tma_gemm(wq, na, wcimg_nt);
            for (int k = 0; k < my_tiles; ++k) { rg.tma_gemm(wq, na, w2img_nt); if (k + 1 < my_tiles) rg.tma_gemm(wq, na, wcimg_nt); }
        }
    } else if (warp == 1) {
        if (lane == 0 && my_tiles > 0) {
            uint32_t gq = 0, wq = 0;
            TLP(2, 1);
            rg.mma_gemm(gq, wq, na, H, tmem_base);                       // GEMM 1 of the first tile
            mma_commit(d1_full);
            for (int k = 0; k < my_tiles; ++k) {
                if (k > 0) mbar_wait(d2_empty, (k - 1) & 1);
                fence_after_sync();
                TLP(2, 3);
                rg.mma_gemm(gq, wq, na, H, tmem_base + CF::D2_COL);      // GEMM 2 of tile k
                mma_commit(d2_full);
                TLP(2, 4);
                if (k + 1 < my_tiles) {
                    mbar_wait(d1_empty, k & 1);
                    fence_after_sync();
                    TLP(2, 1);
                    rg.mma_gemm(gq, wq, na, H, tmem_base);               // GEMM 1 of tile k+1
                    mma_commit(d1_full);
                    TLP(2, 2);
                }
            }
        }
    } else if (warp == CF::SV_WARP) {
        // saved-activation producer, in the order the worker parts consume the 16-column chunks:
        //   d3(0);  for k: pre2(k) [epilogue 1], pre2(k) [GEMM-2 operand], d3(k+1) [GEMM-1 operand of the next tile], d1(k) [epilogue 2]
        if (lane == 0 && my_tiles > 0) {
            const int nchunks = (H + 15) / 16, planes = sv_planes(H);
            uint32_t q = 0;
            auto stream = [&](const float* src, bool p32) {     // 16-bit derivative planes (8 columns each) or fp32 pre2 planes (4 columns each)
                for (int ch = 0; ch < nchunks; ++ch, ++q) {
                    const uint32_t s = q % SVS, rr = q / SVS;
                    const uint32_t bytes = (uint32_t)(p32 ? min(4, H / 4 - 4 * ch) : min(2, planes - 2 * ch)) * (uint32_t)SV_PLANE_BYTES;
                    const size_t off = (size_t)ch * (p32 ? 4 : 2) * (SV_PLANE_BYTES / 4);
                    if (rr > 0) mbar_wait(&sv_empty[s], (rr - 1) & 1);
                    sv_round[s] = rr;                        // full[s] is in phase rr from here on (see SvRing)
                    mbar_arrive_expect_tx(&sv_full[s], bytes);
                    bulk_g2s(sv.buf + s * CF::SV_SLOT_BYTES, src + off, bytes, &sv_full[s]);
                }
            };
            // g_agg rows (one per row node, shared by its ~9 edges) of the CTA's kk-th tile -> block kk & 1, one bulk copy per row
            auto ga_load = [&](int tile_, uint32_t kk) {
                if constexpr (CF::GA_TMA) {
                    const uint32_t b = kk & 1, use = kk >> 1;
                    if (use > 0) mbar_wait(&ga_empty[b], (use - 1) & 1);
                    const int4 ti = __ldg(g.tile_info + tile_);
                    if (ti.y <= CF::GA_TMA_ROWS) {
                        mbar_arrive_expect_tx(&ga_full[b], (uint32_t)(ti.y * H * 4));
                        for (int nl = 0; nl < ti.y; ++nl)
                            bulk_g2s(ga_scratch + b * CF::GA_BLOCK_BYTES + nl * NP * 4, a.g_agg + (size_t)(ti.x + nl) * a.ld_gagg, (uint32_t)(H * 4), &ga_full[b]);
                    } else {
                        mbar_arrive(&ga_full[b]);                // too many row nodes: the workers read g_agg from L2
                    }
                }
            };
            const size_t tstride = (size_t)planes * (SV_PLANE_BYTES / 4);          // floats per tile: derivative codes
            const size_t tstride_p = (size_t)(H / 4) * (SV_PLANE_BYTES / 4);       //                  pre2
            ga_load(blockIdx.x, 0);
            stream(a.sv_d3 + (size_t)blockIdx.x * tstride, false);
            int tile = blockIdx.x;
            for (int k = 0; k < my_tiles; ++k, tile += gridDim.x) {
                stream(a.sv_pre2 + (size_t)tile * tstride_p, true);
                if (k + 1 < my_tiles) ga_load(tile + gridDim.x, (uint32_t)k + 1);
                stream(a.sv_pre2 + (size_t)tile * tstride_p, true);
                if (k + 1 < my_tiles) stream(a.sv_d3 + (size_t)(tile + gridDim.x) * tstride, false);
                stream(a.sv_d1 + (size_t)tile * tstride, false);
            }
        }
    } else if (warp == CF::GEO_WARP) {
        // ---- geometry warp: header and per-edge backward geometry of the coming tiles (its dependent load chain
        //      tile_info -> erow/ecol -> x / g_xout / tau used to stall all worker warps at every tile start) ----
        auto geo_emit = [&](const TileMeta& m, uint32_t tc) {
            const uint32_t gb_ = tc & 1, use = tc >> 1;
            if (use > 0) mbar_wait(&geo_empty[gb_], (use - 1) & 1);
            int* gi = geo_s + gb_ * CF::GEO_WORDS;
            float* gf = reinterpret_cast<float*>(gi);
            if (lane == 0) { gi[0] = m.node_lo; gi[1] = m.nn; gi[2] = m.e_lo; gi[3] = m.ne; }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int r = lane + 32 * q;
                int rloc = 0;
                float gphi = 0.f, dx = 0.f, dy = 0.f, dz = 0.f, gux = 0.f, guy = 0.f, guz = 0.f, gat = 0.f;
                if (r < m.ne) {
                    const int e = m.e_lo + r, rown = m.row[q], coln = m.col[q];
                    gat = a.g_attr[e];                       // sum of the later layers (every edge is updated once per launch, by its own tile)
                    rloc = rown - m.node_lo;
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
                gi[GEO_HDR + r] = rloc; gf[GEO_HDR + 128 + r] = gphi;
                gf[GEO_HDR + 256 + r] = dx; gf[GEO_HDR + 384 + r] = dy; gf[GEO_HDR + 512 + r] = dz;
                gf[GEO_HDR + 640 + r] = gux; gf[GEO_HDR + 768 + r] = guy; gf[GEO_HDR + 896 + r] = guz;
                gf[GEO_HDR + 1024 + r] = gat;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&geo_full[gb_]);
        };
        TileMeta cur;
        uint32_t tcnt = 0;
        for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x, ++tcnt) {
            tile_meta_load(cur, g, tile, lane, true);
            geo_emit(cur, tcnt);                                 // blocks on geo_empty: at most two tiles ahead of the workers
        }
    } else {
        const int group = warp & 3, part = (warp - 2) >> 2, half = part & 1;
        const int r = group * 32 + lane;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(group * 32) << 16);
        const int nchunks = (H + 15) / 16, nfull = H >> 4;       // all chunks / chunks whose 16 columns are all real
#ifndef GB_BWD_ROT
#define GB_BWD_ROT 0          // (2 = extra chunk of the epilogues on part 2 as in the forward kernel: no measurable gain here)
#endif
        // chunk ownership of the two EPILOGUES: ch == pv (mod 4).  The operand builds are tied to the part pairs of the activation ring
        // (the extra chunk is part 0's there); the epilogues only read tensor memory and the saved-activation ring, so their extra chunk
        // can go to part 2 (see the forward kernel)
        const int pv = (part + GB_BWD_ROT) & (CF::NPARTS - 1);
        const int tlr = (lane == 0 && group == 0) ? (part == 0 ? 0 : (part == 3 ? 1 : -1)) : -1;
        (void)tlr;
        const uint32_t lane_off = (uint32_t)(group * 32) << 16;
        const bool leader = group == 0 && lane == 0;
        (void)lane_off; (void)leader;
        float* stg = stg_s + (warp - 2) * CF::STG_WARP_FLOATS;
        uint32_t sq = 0;                                             // first saved-activation chunk of the current stream
        // ---- GEMM 1 operand of the CTA's k-th tile: g_pre3 = g_phi * w_c * SiLU'(pre3) ----
        // (no per-element bounds checks: w_c is zero-padded beyond H, rows beyond the tile's edges have g_phi = 0)
        auto build1 = [&](uint32_t k) {
            const uint32_t gb_ = k & 1;
            TLW(10);
            mbar_wait(&geo_full[gb_], (k >> 1) & 1);
            TLW(11);
            const f2 gphi2 = f2s(reinterpret_cast<const float*>(geo_s + gb_ * CF::GEO_WORDS)[GEO_HDR + 128 + r]);
            for (int j = part >> 1; j < na; j += CF::NPARTS / 2) {
                f2 x[8];
                const int ch = 2 * j + half;
                if (ch < nchunks) {
                    const uint4* d3p = svq_acquire<SVS>(sv, sq + ch, r);
                    TLW(100 + j);
#pragma unroll
                    for (int pp = 0; pp < 2; ++pp) {
                        f2 d3[4];
                        sv_load8<GB_SV_D>(d3p, pp, 2 * ch + pp < npl, d3);
#pragma unroll
                        for (int c2 = 0; c2 < 2; ++c2) {
                            const int c = 2 * pp + c2;
                            const float4 wl = *reinterpret_cast<const float4*>(vec_s + 5 * NP + ch * 16 + 4 * c);
                            x[2 * c] = mul2(mul2(gphi2, lo2(wl)), d3[2 * c2]);
                            x[2 * c + 1] = mul2(mul2(gphi2, hi2(wl)), d3[2 * c2 + 1]);
                        }
                    }
                    svq_release<SVS>(sv, sq + ch);
                } else {
#pragma unroll
                    for (int c = 0; c < 8; ++c) x[c] = f2s(0.f);
                }
                TLW(110 + j);
                if constexpr (CF::AT) { if (ch < nchunks) rg.put(2 * k, ch, nchunks, lane_off, leader, x); }
                else rg.put_chunk2(2 * k, j, na, r, half, x);
                TLW(120 + j);
            }
            sq += nchunks;
        };
        if (my_tiles > 0) build1(0);
        int tile = blockIdx.x;
        for (uint32_t k = 0; k < (uint32_t)my_tiles; ++k, tile += gridDim.x) {
            const uint32_t gb_ = k & 1;
            const int* gi = geo_s + gb_ * CF::GEO_WORDS;
            const float* gf = reinterpret_cast<const float*>(gi);
            const int node_lo = gi[0], nn = gi[1], e_lo = gi[2], ne = gi[3];
            const bool valid = r < ne;
            // ---- epilogue 1: g_ef = coordinate branch + aggregation branch; attention backward ----
            TLW(40);
            mbar_wait(d1_full, k & 1);
            TLW(41);
            fence_after_sync();
            // the tile's rows of g_agg (one per row node, shared by its ~9 edges) are copied once, coalesced, into the A ring: every
            // operand atom stored so far (GEMM 1 of this tile was the last) has been consumed, and nothing is stored before build2
            // (AT form: there is no activation ring in shared memory; a dedicated block of GA_SCRATCH_ROWS rows takes its place)
            // (GA_TMA form: the rows were copied by the SV warp's bulk TMA one tile ahead into block k & 1 -- no staging loop here)
            constexpr int GA_STAGE_BYTES = CF::GA_TMA ? CF::GA_BLOCK_BYTES : (CF::AT ? CF::GA_SCRATCH_ROWS * NP * 4 : CF::RSB::A_STAGE);
            constexpr int GA_ROWS = GA_STAGE_BYTES / (NP * 4);             // rows per ring stage (A hi + A lo)
            unsigned char* ga_base = ga_scratch + (CF::GA_TMA ? (k & 1) * CF::GA_BLOCK_BYTES : 0);
            const bool ga_staged = nn <= (CF::AT || CF::GA_TMA ? 1 : CF::RSB::SA) * GA_ROWS;
            if constexpr (CF::GA_TMA) mbar_wait(&ga_full[k & 1], (k >> 1) & 1);
            if (!CF::GA_TMA && ga_staged) {
                const int h4 = H >> 2;
                for (int idx = (warp - 2) * 32 + lane; idx < nn * h4; idx += CF::NWORK) {
                    const int nl = idx / h4, k4 = idx - nl * h4;
                    float* dst = reinterpret_cast<float*>(ga_base + (nl / GA_ROWS) * GA_STAGE_BYTES) + (nl % GA_ROWS) * NP + 4 * k4;
                    *reinterpret_cast<float4*>(dst) = __ldg(reinterpret_cast<const float4*>(a.g_agg + (size_t)(node_lo + nl) * a.ld_gagg + 4 * k4));
                }
            }
            TLW(42);
            bar_named(BAR_WORKERS, CF::NWORK);                     // staged rows visible; also orders red_s against the previous tile
            TLW(43);
            f2 plog2 = f2s(0.f), pdot2 = f2s(0.f);
            const int rloc = gi[GEO_HDR + r];
            const float* ga_row = ga_staged ? reinterpret_cast<const float*>(ga_base + (rloc / GA_ROWS) * GA_STAGE_BYTES) + (rloc % GA_ROWS) * NP
                                            : a.g_agg + (size_t)(node_lo + rloc) * a.ld_gagg;
            // chunk-streaming: g_ef = accumulator 1 + g_agg[row] goes BACK to tensor memory (same columns) instead of waiting in 64
            // registers for the row-wide attention reduction; the operand build of GEMM 2 reads it again.  The TMEM read of the
            // next chunk is in flight while the current one is processed.
            // (two code paths per chunk, as in the forward kernel: chunks whose 16 columns are all real carry no per-column tests; the
            //  last, partly filled chunk of a hidden width that is not a multiple of 16 only touches its real 4-column groups)
            {
                float v[16];
                if (GB_BWD_SPLIT_LD && pv < nchunks) tmem_ld16_issue_f(lane_addr + pv * 16, v);
                auto e1_chunk = [&](auto tail_t, int ch) {
                    constexpr bool TAIL = decltype(tail_t)::value;
                    if (GB_BWD_SPLIT_LD) tmem_ld_wait16(v); else tmem_ld16(lane_addr + ch * 16, v);
                    TLW(200 + ch);
                    const float4* p2p = reinterpret_cast<const float4*>(svq_acquire<SVS>(sv, sq + ch, r));
                    TLW(220 + ch);
                    float w[16];
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const int c0 = ch * 16 + 4 * c4;
                        if (!TAIL || c0 < H) {                      // (the staged g_agg copy only holds the H real columns; planes beyond H
                            const float4 ga = *reinterpret_cast<const float4*>(ga_row + c0);       //  would be the slot's previous occupant)
                            const float4 p2 = p2p[c4 * 128];
                            const float4 wq = *reinterpret_cast<const float4*>(vec_s + 3 * NP + c0);
                            const f2 qa = silu2(lo2(p2)), qb = silu2(hi2(p2));
                            const f2 ga_ = add2(make_float2(v[4 * c4], v[4 * c4 + 1]), lo2(ga));
                            const f2 gb2 = add2(make_float2(v[4 * c4 + 2], v[4 * c4 + 3]), hi2(ga));
                            w[4 * c4] = ga_.x; w[4 * c4 + 1] = ga_.y; w[4 * c4 + 2] = gb2.x; w[4 * c4 + 3] = gb2.y;
                            plog2 = fma2(lo2(wq), qa, plog2); plog2 = fma2(hi2(wq), qb, plog2);
                            pdot2 = fma2(ga_, qa, pdot2); pdot2 = fma2(gb2, qb, pdot2);
                        } else {
                            w[4 * c4] = 0.f; w[4 * c4 + 1] = 0.f; w[4 * c4 + 2] = 0.f; w[4 * c4 + 3] = 0.f;
                        }
                    }
                    svq_release<SVS>(sv, sq + ch);
                    if (GB_BWD_SPLIT_LD && ch + CF::NPARTS < nchunks) tmem_ld16_issue_f(lane_addr + (ch + CF::NPARTS) * 16, v);
                    tmem_st16(lane_addr + ch * 16, w);
                };
#pragma unroll 1
                for (int ch = pv; ch < nfull; ch += CF::NPARTS) e1_chunk(std::false_type{}, ch);
                if (nfull < nchunks && pv == (nfull & (CF::NPARTS - 1))) e1_chunk(std::true_type{}, nfull);
                tmem_st_wait();
            }
            if constexpr (CF::GA_TMA) mbar_arrive(&ga_empty[k & 1]);      // this tile's g_agg block may be overwritten (two tiles from now)
            sq += nchunks;
            red_s[part * 128 + r] = plog2.x + plog2.y;
            red_s[CF::NPARTS * 128 + part * 128 + r] = pdot2.x + pdot2.y;
            TLW(44);
            bar_named(BAR_WORKERS, CF::NWORK);                     // also: nobody reads the staged g_agg rows in the A ring any more
            TLW(45);
            float gate = 1.f, kap = 0.f;
            if (a.attention) {
                gate = sigmoid_f(psum_parts<CF::NPARTS>(red_s, r) + a.att_b);
                kap = (psum_parts<CF::NPARTS>(red_s + CF::NPARTS * 128, r)) * gate * (1.f - gate);
            }
            // ---- GEMM 2 operand: g_pre2 = (g_ef gate + kappa w_att) SiLU'(pre2) ----
            {
                const f2 gate2 = f2s(gate), kap2 = f2s(kap);
                float v[16];
                if (GB_BWD_SPLIT_LD && part < nchunks) tmem_ld16_issue_f(lane_addr + part * 16, v);
                auto b2_chunk = [&](auto tail_t, int ch) {
                    constexpr bool TAIL = decltype(tail_t)::value;
                    f2 x[8];
                    if (GB_BWD_SPLIT_LD) tmem_ld_wait16(v); else tmem_ld16(lane_addr + ch * 16, v);
                    const float4* p2p = reinterpret_cast<const float4*>(svq_acquire<SVS>(sv, sq + ch, r));
                    TLW(300 + ch);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const int c0 = ch * 16 + 4 * c4;
                        if (!TAIL || c0 < H) {
                            const float4 p2 = p2p[c4 * 128];
                            const float4 wq = *reinterpret_cast<const float4*>(vec_s + 3 * NP + c0);
                            const f2 ta = fma2(kap2, lo2(wq), mul2(make_float2(v[4 * c4], v[4 * c4 + 1]), gate2));
                            const f2 tb = fma2(kap2, hi2(wq), mul2(make_float2(v[4 * c4 + 2], v[4 * c4 + 3]), gate2));
                            x[2 * c4] = mul2(ta, dsilu2(lo2(p2)));
                            x[2 * c4 + 1] = mul2(tb, dsilu2(hi2(p2)));
                        } else {
                            x[2 * c4] = f2s(0.f); x[2 * c4 + 1] = f2s(0.f);
                        }
                    }
                    svq_release<SVS>(sv, sq + ch);
                    if (GB_BWD_SPLIT_LD && ch + CF::NPARTS < nchunks) tmem_ld16_issue_f(lane_addr + (ch + CF::NPARTS) * 16, v);
                    TLW(320 + ch);
                    if constexpr (CF::AT) rg.put(2 * k + 1, ch, nchunks, lane_off, leader, x);
                    else rg.put_chunk2(2 * k + 1, ch >> 1, na, r, half, x);
                    TLW(340 + ch);
                };
#pragma unroll 1
                for (int ch = part; ch < nfull; ch += CF::NPARTS) b2_chunk(std::false_type{}, ch);
                if (nfull < nchunks && part == (nfull & (CF::NPARTS - 1))) b2_chunk(std::true_type{}, nfull);
                if constexpr (!CF::AT) {         // chunks beyond the hidden width but inside the last atom: zeros, so that the atom completes
                    for (int ch = nchunks; ch < 2 * na; ++ch)
                        if ((ch & (CF::NPARTS - 1)) == part) {
                            f2 x[8];
#pragma unroll
                            for (int c = 0; c < 8; ++c) x[c] = f2s(0.f);
                            rg.put_chunk2(2 * k + 1, ch >> 1, na, r, half, x);
                        }
                }
            }
            sq += nchunks;
            fence_before_sync();
            mbar_arrive(d1_empty);                                 // accumulator 1 (and g_ef parked in it) consumed: GEMM 1 of tile k+1 may start
            // ---- GEMM 1 operand of the next tile (its MMAs overlap epilogue 2 below) ----
            TLW(50);
            if (k + 1 < (uint32_t)my_tiles) build1(k + 1);
            // ---- epilogue 2: g_pre1 = g_s1 * SiLU'(pre1) -> HBM (row-major per edge); radial / attr dots ----
            //      the row sums (g_Pa), column sums (g_Pb) and the coordinate gradient are reduced afterwards by the
            //      node-parallel pred_bwd_reduce_kernel: fixed summation order, no atomics, no per-chunk barriers here
            TLW(60);
            mbar_wait(d2_full, k & 1);
            TLW(61);
            fence_after_sync();
            f2 pr2 = f2s(0.f), pa2 = f2s(0.f);
            // g_pre1 rows are written through a per-warp transpose block: thread = row is what TMEM gives, but 16 bytes of 32
            // different rows per store instruction cost 32 L1 tag lookups; transposed, one instruction covers 16 rows x 32 bytes
            // (full sectors).  The block holds 32 rows x 8 columns (16-byte chunk c of row l at position c ^ ((l >> 2) & 1):
            // conflict-free both ways); a 16-column chunk goes through it in two halves.
            {
                const int piece = lane & 1, rsub = lane >> 1;
                const int wsw = (lane >> 2) & 1;
                float v[16];
                if (GB_BWD_SPLIT_LD && pv < nchunks) tmem_ld16_issue_f(lane_addr + CF::D2_COL + pv * 16, v);
                auto e2_chunk = [&](auto tail_t, int ch) {
                    constexpr bool TAIL = decltype(tail_t)::value;
                    if (GB_BWD_SPLIT_LD) tmem_ld_wait16(v); else tmem_ld16(lane_addr + CF::D2_COL + ch * 16, v);
                    TLW(400 + ch);
                    const uint4* d1p = svq_acquire<SVS>(sv, sq + ch, r);
                    TLW(420 + ch);
                    f2 gp[8], d1v[8];
                    sv_load8<GB_SV_D>(d1p, 0, true, d1v);
                    sv_load8<GB_SV_D>(d1p, 1, !TAIL || 2 * ch + 1 < npl, d1v + 4);
#pragma unroll
                    for (int c4 = 0; c4 < 4; ++c4) {
                        const int c0 = ch * 16 + 4 * c4;
                        const float4 wr = *reinterpret_cast<const float4*>(vec_s + c0);
                        const float4 wa = *reinterpret_cast<const float4*>(vec_s + NP + c0);
                        gp[2 * c4] = mul2(make_float2(v[4 * c4], v[4 * c4 + 1]), d1v[2 * c4]);
                        gp[2 * c4 + 1] = mul2(make_float2(v[4 * c4 + 2], v[4 * c4 + 3]), d1v[2 * c4 + 1]);
                        pr2 = fma2(lo2(wr), gp[2 * c4], pr2); pr2 = fma2(hi2(wr), gp[2 * c4 + 1], pr2);
                        pa2 = fma2(lo2(wa), gp[2 * c4], pa2); pa2 = fma2(hi2(wa), gp[2 * c4 + 1], pa2);
                    }
                    svq_release<SVS>(sv, sq + ch);
                    if (GB_BWD_SPLIT_LD && ch + CF::NPARTS < nchunks) tmem_ld16_issue_f(lane_addr + CF::D2_COL + (ch + CF::NPARTS) * 16, v);   // next chunk in flight during the stores
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        if (TAIL && ch * 16 + 8 * hh >= H) break;      // (uniform: the partly filled chunk's halves beyond the hidden width)
                        *reinterpret_cast<float4*>(stg + lane * 8 + 4 * (0 ^ wsw)) = cat2(gp[4 * hh], gp[4 * hh + 1]);
                        *reinterpret_cast<float4*>(stg + lane * 8 + 4 * (1 ^ wsw)) = cat2(gp[4 * hh + 2], gp[4 * hh + 3]);
                        __syncwarp();
                        const int c = ch * 16 + 8 * hh + 4 * piece;
                        if (!TAIL || c < H) {
#pragma unroll
                            for (int i = 0; i < 2; ++i) {
                                const int rw = rsub + 16 * i, rl = group * 32 + rw;
                                if (rl < ne) *reinterpret_cast<float4*>(a.g_pre1 + (size_t)(e_lo + rl) * H + c) = *reinterpret_cast<const float4*>(stg + rw * 8 + 4 * (piece ^ ((rw >> 2) & 1)));
                            }
                        }
                        __syncwarp();
                    }
                };
#pragma unroll 1
                for (int ch = pv; ch < nfull; ch += CF::NPARTS) e2_chunk(std::false_type{}, ch);
                if (nfull < nchunks && pv == (nfull & (CF::NPARTS - 1))) e2_chunk(std::true_type{}, nfull);
            }
            sq += nchunks;
            fence_before_sync();
            mbar_arrive(d2_empty);
            red_s[part * 128 + r] = pr2.x + pr2.y;
            red_s[CF::NPARTS * 128 + part * 128 + r] = pa2.x + pa2.y;
            TLW(62);
            bar_named(BAR_WORKERS, CF::NWORK);
            TLW(63);
            if (part == 0 && valid) {
                const float g_r = psum_parts<CF::NPARTS>(red_s, r);
                const float g_a = psum_parts<CF::NPARTS>(red_s + CF::NPARTS * 128, r);
                a.g_attr[e_lo + r] = gf[GEO_HDR + 1024 + r] + g_a;         // (the old value came with the geometry block: no load on the tile's tail)
                const float dx = gf[GEO_HDR + 256 + r], dy = gf[GEO_HDR + 384 + r], dz = gf[GEO_HDR + 512 + r];
                const float gux = gf[GEO_HDR + 640 + r], guy = gf[GEO_HDR + 768 + r], guz = gf[GEO_HDR + 896 + r];
                const float nrm = sqrtf(dx * dx + dy * dy + dz * dz + 1e-8f);
                const float inv = 1.f / (nrm + 1.f);
                const float k2 = (gux * dx + guy * dy + guz * dz) * inv * inv / nrm;
                float* gd = a.g_d + (size_t)(e_lo + r) * 3;
                gd[0] = 2.f * g_r * dx + gux * inv - k2 * dx;
                gd[1] = 2.f * g_r * dy + guy * inv - k2 * dy;
                gd[2] = 2.f * g_r * dz + guz * inv - k2 * dz;
            }
            mbar_arrive(&geo_empty[gb_]);
            TLW(70);
            // red_s: its next writers (epilogue 1 of tile k+1) first pass the worker barrier at the top of that epilogue, which the
            // part-0 warps reach only after the reads above
        }
    }
    fence_before_sync();
    __syncthreads();
    if (warp == 1) tmem_dealloc<512>(tmem_base);
}

#ifdef GB_DEBUG_HANG
extern "C" int gb_debug_set_hang_buf(int* dev_ptr) { return (int)cudaMemcpyToSymbol(gb_hang_buf, &dev_ptr, sizeof(dev_ptr)); }
#endif

#ifdef GB_TIMELINE
extern "C" int gb_debug_timeline_pred(unsigned long long* out, unsigned int* n, int reset) {
    if (reset) { unsigned int z[5] = {0, 0, 0, 0, 0}; return (int)cudaMemcpyToSymbol(gb_tl_pred_n, z, sizeof(z)); }
    cudaMemcpyFromSymbol(n, gb_tl_pred_n, sizeof(gb_tl_pred_n));
    return (int)cudaMemcpyFromSymbol(out, gb_tl_pred, sizeof(gb_tl_pred));
}
#endif

template <int NP>
static void launch_bwd_t(const PredEdgeArgs& a, const float* wcimg_nt, const float* w2img_nt, int H, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(tc_pred_edge_bwd_kernel<NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcBwdCfg<NP>::SMEM);
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = a.g.n_tiles < sms ? a.g.n_tiles : sms;
    tc_pred_edge_bwd_kernel<NP><<<grid, TcBwdCfg<NP>::THREADS, TcBwdCfg<NP>::SMEM, s>>>(a, wcimg_nt, w2img_nt, H);
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

// Node-parallel reductions of the backward:
//   g_Pa[i] = sum over the row segment of i,  g_Pb[i] = sum over the edges whose column is i (CSC order),
//   g_x[i]  = g_xout[i]*mask_i + sum_row g_d - sum_col g_d.          Fixed order -> bit-reproducible.
// One thread per (node, float4 column): a block of 256 threads covers 256 / (H / 4) nodes (5 at H = 196: 245 busy threads; the
// first version gave a whole warp to a node and ran its second pass over the 49 float4 columns with 17 of 32 lanes).  The row and
// column loops issue four independent 16-byte loads before they add, in the same order as a plain loop.
__global__ void __launch_bounds__(256) pred_bwd_reduce_kernel(PredEdgeArgs a, int H) {
    const Graph& g = a.g;
    const int nq = H >> 2;                                       // float4 columns
    const int npb = 256 / nq;                                    // nodes per block and pass
    const int nl = threadIdx.x / nq, q = threadIdx.x - nl * nq;
    for (int base = blockIdx.x * npb; base < g.n_nodes; base += gridDim.x * npb) {
        const int node = base + nl;
        if (nl < npb && node < g.n_nodes) {
            const int r0 = g.rowptr[node], r1 = g.rowptr[node + 1];
            const int c0 = g.colptr[node], c1 = g.colptr[node + 1];
            const float4* src = reinterpret_cast<const float4*>(a.g_pre1) + q;
            float4 sa = make_float4(0.f, 0.f, 0.f, 0.f), sb = sa;
            int e = r0;
            for (; e + 4 <= r1; e += 4) {
                float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = __ldg(src + (size_t)(e + u) * nq);
#pragma unroll
                for (int u = 0; u < 4; ++u) { sa.x += v[u].x; sa.y += v[u].y; sa.z += v[u].z; sa.w += v[u].w; }
            }
            for (; e < r1; ++e) { const float4 v = __ldg(src + (size_t)e * nq); sa.x += v.x; sa.y += v.y; sa.z += v.z; sa.w += v.w; }
            int p = c0;
            for (; p + 4 <= c1; p += 4) {
                int ce[4]; float4 v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) ce[u] = __ldg(g.cedge + p + u);
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = __ldg(src + (size_t)ce[u] * nq);
#pragma unroll
                for (int u = 0; u < 4; ++u) { sb.x += v[u].x; sb.y += v[u].y; sb.z += v[u].z; sb.w += v[u].w; }
            }
            for (; p < c1; ++p) { const float4 v = __ldg(src + (size_t)__ldg(g.cedge + p) * nq); sb.x += v.x; sb.y += v.y; sb.z += v.z; sb.w += v.w; }
            reinterpret_cast<float4*>(a.g_Pa + (size_t)node * H)[q] = sa;
            reinterpret_cast<float4*>(a.g_Pb + (size_t)node * H)[q] = sb;
            if (q < 3) {                                         // coordinate gradient: three threads of the node
                float sum = a.g_xout[3 * node + q] * g.node_mask[node];
                for (int ee = r0; ee < r1; ++ee) sum += a.g_d[(size_t)ee * 3 + q];
                for (int pp = c0; pp < c1; ++pp) sum -= a.g_d[(size_t)g.cedge[pp] * 3 + q];
                a.g_x[3 * node + q] = sum;
            }
        }
    }
}

void launch_pred_bwd_reduce(int H, const PredEdgeArgs& a, cudaStream_t s) {
    if (a.g.n_nodes <= 0) return;
    const int npb = 256 / (H >> 2);
    const int blocks = min(148 * 16, (a.g.n_nodes + npb - 1) / npb);
    pred_bwd_reduce_kernel<<<blocks, 256, 0, s>>>(a, H);
}

template <int NP>
static void launch_fwd_t(bool save, const PredEdgeArgs& a, const float* w2img, const float* wcimg, int H, cudaStream_t s) {
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(tc_pred_edge_fwd_kernel<NP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcFwdCfg<NP>::SMEM);
        cudaFuncSetAttribute(tc_pred_edge_fwd_kernel<NP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcFwdCfg<NP>::SMEM);
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = a.g.n_tiles < sms ? a.g.n_tiles : sms;
    if (save) tc_pred_edge_fwd_kernel<NP, true><<<grid, TcFwdCfg<NP>::THREADS, TcFwdCfg<NP>::SMEM, s>>>(a, w2img, wcimg, H);
    else tc_pred_edge_fwd_kernel<NP, false><<<grid, TcFwdCfg<NP>::THREADS, TcFwdCfg<NP>::SMEM, s>>>(a, w2img, wcimg, H);
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
