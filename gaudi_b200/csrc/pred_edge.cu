// Fused per-edge kernels of one predictor E_GCL layer (edm/egnn_predictor/gcl.py:225-316) and the
// hand-written reverse pass that replaces torch.autograd through it (en_diffusion.py:899-903).
//
// forward, per tile of <=128 compacted edges (row segments complete):
//   pre1 = Pa[row]+Pb[col]+w_r r+w_a a          s1 = SiLU(pre1)           (a = |x0_i-x0_j|^2, models.py:452)
//   pre2 = s1 W2^T + b2      q = SiLU(pre2)     gate = sigmoid(w_att.q+b)  ef = q*gate      (gcl.py:225-238)
//   agg_i = sum_j ef                                                                        (gcl.py:242)
//   pre3 = ef Wc^T + bc      s3 = SiLU(pre3)    tau = tanh(w_c.s3)                          (gcl.py:208-216)
//   x_i' = (x_i + sum_j u_ij tau range) mask_i                                              (gcl.py:256-279,305)
// saved for the gradient pass (tile-blocked so that both passes read/write coalesced):
//   d1 = SiLU'(pre1)  layout [tile][k/4][m][4]      pre2, d3 = SiLU'(pre3)  layout [tile][k][m]     tau [edge]
//
// backward, same tiling, two GEMMs with the un-transposed weights:
//   g_phi = (g_x'_i mask_i . u) range (1-tau^2)
//   g_ef  = (g_phi w_c * d3) Wc  + g_agg_i
//   g_q   = g_ef gate + (g_ef.q) gate(1-gate) w_att ;  g_pre2 = g_q SiLU'(pre2)
//   g_pre1 = (g_pre2 W2) * d1
//   g_Pa_i = sum_j g_pre1 (row segments, plain stores)   g_Pb_j = sum_i g_pre1 (column scatter, RED)
//   g_r = w_r.g_pre1   g_a = w_a.g_pre1 (accumulated per edge over layers)
//   g_x  += d(r,u)/dx terms (row and column)
#include "common.cuh"
#include "kernels.h"

namespace gb {

template <int HP>
struct PredSmem {
    static constexpr int NW = TileCfg<HP>::NW;
    float* A_s; float* ring; uint64_t* full; uint64_t* empty;
    float* vec_s;    // [6][HP]: w_r, w_attr, b2, att_w, bc, wc_last
    float* red_s;    // [2][NW][128]
    float* f_s;      // [16][128] per-edge float scratch
    int* row_s; int* col_s; int* seg_s;
    __device__ __forceinline__ void carve(unsigned char* raw) {
        A_s = reinterpret_cast<float*>(raw);
        ring = A_s + HP * GB_MS;
        full = reinterpret_cast<uint64_t*>(ring + GB_STAGES * GB_KC * HP);
        empty = full + GB_STAGES;
        vec_s = reinterpret_cast<float*>(empty + GB_STAGES);
        red_s = vec_s + 6 * HP;
        f_s = red_s + 2 * NW * GB_TM;
        row_s = reinterpret_cast<int*>(f_s + 16 * GB_TM);
        col_s = row_s + GB_TM;
        seg_s = col_s + GB_TM;
    }
};

template <int HP>
__device__ __forceinline__ void load_vecs(const PredEdgeArgs& a, float* vec_s, int tid, int NT) {
    for (int i = tid; i < HP; i += NT) {
        vec_s[i] = a.ext[i]; vec_s[HP + i] = a.ext[HP + i]; vec_s[2 * HP + i] = a.b2[i];
        vec_s[3 * HP + i] = a.att_w[i]; vec_s[4 * HP + i] = a.bc[i]; vec_s[5 * HP + i] = a.wc_last[i];
    }
}

// row-segment sums of the [c][m] tile in A_s -> dst[node][c] (lane = 8 columns x 4 nodes: conflict-free for odd segments)
template <int HP, int NW>
__device__ __forceinline__ void segsum_rows(const float* A_s, const int* seg_s, int nn, int node_lo, float* dst, int ld,
                                            int warp, int lane) {
    const int ncg = (HP + 7) / 8, nng = (nn + 3) / 4;
    for (int item = warp; item < ncg * nng; item += NW) {
        const int cg = item % ncg, ng = item / ncg;
        const int c = cg * 8 + (lane & 7), nl = ng * 4 + (lane >> 3);
        if (c < HP && nl < nn) {
            float sum = 0.f;
            for (int m = seg_s[nl]; m < seg_s[nl + 1]; ++m) sum += A_s[c * GB_MS + m];
            dst[(size_t)(node_lo + nl) * ld + c] = sum;
        }
    }
}

// =================================================================================================
// forward
// =================================================================================================
template <int HP, bool SAVE>
__global__ void __launch_bounds__((TileCfg<HP>::NW + 1) * 32, 1) pred_edge_fwd_kernel(PredEdgeArgs a) {
    constexpr int NW = TileCfg<HP>::NW;
    constexpr int CW = HP / NW;
    constexpr int NT = NW * 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PredSmem<HP> S;
    S.carve(smem_raw);
    float* r_s = S.f_s; float* a0_s = S.f_s + GB_TM; float* u_s = S.f_s + 2 * GB_TM;   // u: [128][3]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < GB_STAGES; ++i) { mbar_init(&S.full[i], 1); mbar_init(&S.empty[i], NW); }
        mbar_fence_init();
    }
    __syncthreads();
    WPipe<HP> pipe;
    pipe.init_side(S.ring, S.full, S.empty);
    const Graph& g = a.g;
    if (warp == NW) {
        if (lane == 0)
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) { pipe.produce(a.wt2, HP); pipe.produce(a.wtc, HP); }
        return;
    }
    load_vecs<HP>(a, S.vec_s, tid, NT);

    for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        const int node_lo = g.tile_ptr[tile], node_hi = g.tile_ptr[tile + 1];
        const int nn = node_hi - node_lo;
        const int e_lo = g.rowptr[node_lo], ne = g.rowptr[node_hi] - e_lo;
        consumer_bar(NT);
        for (int m = tid; m < GB_TM; m += NT) {
            int row = 0, col = 0; float r = 0.f, a0 = 0.f, ux = 0.f, uy = 0.f, uz = 0.f;
            if (m < ne) {
                const int e = e_lo + m;
                row = g.erow[e]; col = g.ecol[e];
                const float dx = a.x[3 * row] - a.x[3 * col], dy = a.x[3 * row + 1] - a.x[3 * col + 1], dz = a.x[3 * row + 2] - a.x[3 * col + 2];
                r = dx * dx + dy * dy + dz * dz;                                      // coord2radial, gcl.py:308-316
                const float inv = 1.f / (sqrtf(r + 1e-8f) + 1.f);
                ux = dx * inv; uy = dy * inv; uz = dz * inv;
                const float ex = a.x0[3 * row] - a.x0[3 * col], ey = a.x0[3 * row + 1] - a.x0[3 * col + 1], ez = a.x0[3 * row + 2] - a.x0[3 * col + 2];
                a0 = ex * ex + ey * ey + ez * ez;
                if (a.a_edge) a0 = a.a_edge[e];
            }
            S.row_s[m] = row; S.col_s[m] = col; r_s[m] = r; a0_s[m] = a0;
            u_s[3 * m] = ux; u_s[3 * m + 1] = uy; u_s[3 * m + 2] = uz;
        }
        for (int i = tid; i <= nn; i += NT) S.seg_s[i] = g.rowptr[node_lo + i] - e_lo;
        consumer_bar(NT);
        // ---- build s1 (and save SiLU'(pre1)) ------------------------------------------------------------
        for (int idx = tid; idx < GB_TM * (HP / 4); idx += NT) {
            const int m = idx & (GB_TM - 1), kq = idx >> 7;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f), dv = v;
            if (m < ne) {
                const float4 pa = __ldg(reinterpret_cast<const float4*>(a.P + (size_t)S.row_s[m] * (2 * HP) + 4 * kq));
                const float4 pb = __ldg(reinterpret_cast<const float4*>(a.P + (size_t)S.col_s[m] * (2 * HP) + HP + 4 * kq));
                const float4 wr = *reinterpret_cast<const float4*>(S.vec_s + 4 * kq);
                const float4 wa = *reinterpret_cast<const float4*>(S.vec_s + HP + 4 * kq);
                const float r = r_s[m], a0 = a0_s[m];
                silu_both(pa.x + pb.x + wr.x * r + wa.x * a0, v.x, dv.x);
                silu_both(pa.y + pb.y + wr.y * r + wa.y * a0, v.y, dv.y);
                silu_both(pa.z + pb.z + wr.z * r + wa.z * a0, v.z, dv.z);
                silu_both(pa.w + pb.w + wr.w * r + wa.w * a0, v.w, dv.w);
            }
            float* d = S.A_s + (4 * kq) * GB_MS + m;
            d[0] = v.x; d[GB_MS] = v.y; d[2 * GB_MS] = v.z; d[3 * GB_MS] = v.w;
            if (SAVE) *reinterpret_cast<float4*>(a.sv_d1 + (((size_t)tile * (HP / 4) + kq) * GB_TM + m) * 4) = dv;
        }
        consumer_bar(NT);
        float acc[4][CW];
        zero_acc<CW>(acc);
        gemm_consume<HP, NW>(S.A_s, HP, acc, pipe, warp, lane);
        // ---- epilogue 1: q, attention gate ------------------------------------------------------------------
        {
            float part[4] = {0.f, 0.f, 0.f, 0.f};
            const float* b2 = S.vec_s + 2 * HP + warp * CW;
            const float* aw = S.vec_s + 3 * HP + warp * CW;
#pragma unroll
            for (int c = 0; c < CW; ++c) {
                const float bb = b2[c], ww = aw[c];
                float p[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    p[r] = acc[r][c] + bb;
                    const float v = silu_f(p[r]);
                    acc[r][c] = v;
                    part[r] = fmaf(ww, v, part[r]);
                }
                if (SAVE) *reinterpret_cast<float4*>(a.sv_pre2 + ((size_t)tile * HP + warp * CW + c) * GB_TM + 4 * lane) = make_float4(p[0], p[1], p[2], p[3]);
            }
            *reinterpret_cast<float4*>(S.red_s + warp * GB_TM + 4 * lane) = make_float4(part[0], part[1], part[2], part[3]);
        }
        consumer_bar(NT);
        {
            float gate[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float sdot = a.att_b;
                for (int w = 0; w < NW; ++w) sdot += S.red_s[w * GB_TM + 4 * lane + r];
                gate[r] = a.attention ? sigmoid_f(sdot) : 1.f;
            }
#pragma unroll
            for (int c = 0; c < CW; ++c)
                *reinterpret_cast<float4*>(S.A_s + (warp * CW + c) * GB_MS + 4 * lane) =
                    make_float4(acc[0][c] * gate[0], acc[1][c] * gate[1], acc[2][c] * gate[2], acc[3][c] * gate[3]);
        }
        consumer_bar(NT);
        segsum_rows<HP, NW>(S.A_s, S.seg_s, nn, node_lo, a.agg, HP, warp, lane);
        // ---- GEMM 2 on the gated edge feature ------------------------------------------------------------------
        zero_acc<CW>(acc);
        gemm_consume<HP, NW>(S.A_s, HP, acc, pipe, warp, lane);
        {
            float part[4] = {0.f, 0.f, 0.f, 0.f};
            const float* bc = S.vec_s + 4 * HP + warp * CW;
            const float* wl = S.vec_s + 5 * HP + warp * CW;
#pragma unroll
            for (int c = 0; c < CW; ++c) {
                const float bb = bc[c], ww = wl[c];
                float dv[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    float v;
                    silu_both(acc[r][c] + bb, v, dv[r]);
                    part[r] = fmaf(ww, v, part[r]);
                }
                if (SAVE) *reinterpret_cast<float4*>(a.sv_d3 + ((size_t)tile * HP + warp * CW + c) * GB_TM + 4 * lane) = make_float4(dv[0], dv[1], dv[2], dv[3]);
            }
            *reinterpret_cast<float4*>(S.red_s + warp * GB_TM + 4 * lane) = make_float4(part[0], part[1], part[2], part[3]);
        }
        consumer_bar(NT);
        if (tid < GB_TM) {
            const int m = tid;
            float phi = 0.f;
            for (int w = 0; w < NW; ++w) phi += S.red_s[w * GB_TM + m];
            const float tau = a.use_tanh ? tanhf(phi) : phi;
            if (SAVE && m < ne) a.sv_tau[e_lo + m] = tau;
            if (a.use_tanh) {
                u_s[3 * m] = u_s[3 * m] * tau * a.coords_range;
                u_s[3 * m + 1] = u_s[3 * m + 1] * tau * a.coords_range;
                u_s[3 * m + 2] = u_s[3 * m + 2] * tau * a.coords_range;
            } else {
                u_s[3 * m] *= tau; u_s[3 * m + 1] *= tau; u_s[3 * m + 2] *= tau;
            }
        }
        consumer_bar(NT);
        for (int idx = tid; idx < nn * 3; idx += NT) {
            const int nl = idx / 3, d = idx - 3 * nl;
            float sum = 0.f;
            for (int m = S.seg_s[nl]; m < S.seg_s[nl + 1]; ++m) sum += u_s[3 * m + d];
            const int node = node_lo + nl;
            a.x_out[3 * node + d] = (a.x[3 * node + d] + sum) * g.node_mask[node];
        }
    }
}

// =================================================================================================
// backward (input gradient only: no weight gradients)
// =================================================================================================
template <int HP>
__global__ void __launch_bounds__((TileCfg<HP>::NW + 1) * 32, 1) pred_edge_bwd_kernel(PredEdgeArgs a) {
    constexpr int NW = TileCfg<HP>::NW;
    constexpr int CW = HP / NW;
    constexpr int NT = NW * 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    PredSmem<HP> S;
    S.carve(smem_raw);
    float* gphi_s = S.f_s;                 // [128]
    float* nrm_s = S.f_s + GB_TM;          // [128]
    float* d_s = S.f_s + 2 * GB_TM;        // [128][3]
    float* gu_s = S.f_s + 5 * GB_TM;       // [128][3]
    float* gd_s = S.f_s + 8 * GB_TM;       // [128][3]
    float* gate_s = S.f_s + 11 * GB_TM;    // [128]
    float* kap_s = S.f_s + 12 * GB_TM;     // [128]

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < GB_STAGES; ++i) { mbar_init(&S.full[i], 1); mbar_init(&S.empty[i], NW); }
        mbar_fence_init();
    }
    __syncthreads();
    WPipe<HP> pipe;
    pipe.init_side(S.ring, S.full, S.empty);
    const Graph& g = a.g;
    if (warp == NW) {
        if (lane == 0)
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) { pipe.produce(a.wc_nt, HP); pipe.produce(a.w2_nt, HP); }
        return;
    }
    load_vecs<HP>(a, S.vec_s, tid, NT);

    for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        const int node_lo = g.tile_ptr[tile], node_hi = g.tile_ptr[tile + 1];
        const int nn = node_hi - node_lo;
        const int e_lo = g.rowptr[node_lo], ne = g.rowptr[node_hi] - e_lo;
        consumer_bar(NT);
        for (int m = tid; m < GB_TM; m += NT) {
            int row = 0, col = 0; float gphi = 0.f, nrm = 1.f, dx = 0.f, dy = 0.f, dz = 0.f, gux = 0.f, guy = 0.f, guz = 0.f;
            if (m < ne) {
                const int e = e_lo + m;
                row = g.erow[e]; col = g.ecol[e];
                dx = a.x[3 * row] - a.x[3 * col]; dy = a.x[3 * row + 1] - a.x[3 * col + 1]; dz = a.x[3 * row + 2] - a.x[3 * col + 2];
                const float r = dx * dx + dy * dy + dz * dz;
                nrm = sqrtf(r + 1e-8f);
                const float inv = 1.f / (nrm + 1.f);
                const float mk = g.node_mask[row];
                const float gx = a.g_xout[3 * row] * mk, gy = a.g_xout[3 * row + 1] * mk, gz = a.g_xout[3 * row + 2] * mk;
                const float tau = a.sv_tau[e];
                const float gdotu = (gx * dx + gy * dy + gz * dz) * inv;
                if (a.use_tanh) {
                    gphi = gdotu * a.coords_range * (1.f - tau * tau);
                    const float sc = tau * a.coords_range;
                    gux = gx * sc; guy = gy * sc; guz = gz * sc;
                } else {
                    gphi = gdotu;
                    gux = gx * tau; guy = gy * tau; guz = gz * tau;
                }
            }
            S.row_s[m] = row; S.col_s[m] = col; gphi_s[m] = gphi; nrm_s[m] = nrm;
            d_s[3 * m] = dx; d_s[3 * m + 1] = dy; d_s[3 * m + 2] = dz;
            gu_s[3 * m] = gux; gu_s[3 * m + 1] = guy; gu_s[3 * m + 2] = guz;
        }
        for (int i = tid; i <= nn; i += NT) S.seg_s[i] = g.rowptr[node_lo + i] - e_lo;
        consumer_bar(NT);
        // ---- A = g_pre3 = g_phi * w_c * SiLU'(pre3) ------------------------------------------------------------
        for (int idx = tid; idx < HP * (GB_TM / 4); idx += NT) {
            const int mq = idx & 31, k = idx >> 5;
            const float4 d3 = __ldg(reinterpret_cast<const float4*>(a.sv_d3 + ((size_t)tile * HP + k) * GB_TM + 4 * mq));
            const float4 gp = *reinterpret_cast<const float4*>(gphi_s + 4 * mq);
            const float w = S.vec_s[5 * HP + k];
            *reinterpret_cast<float4*>(S.A_s + k * GB_MS + 4 * mq) = make_float4(gp.x * w * d3.x, gp.y * w * d3.y, gp.z * w * d3.z, gp.w * w * d3.w);
        }
        consumer_bar(NT);
        float acc[4][CW];
        zero_acc<CW>(acc);
        gemm_consume<HP, NW>(S.A_s, HP, acc, pipe, warp, lane);     // acc = g_ef (coordinate branch)
        // ---- epilogue 1: add aggregation branch, attention backward ------------------------------------------------
        const float* pre2_base = a.sv_pre2 + ((size_t)tile * HP + warp * CW) * GB_TM + 4 * lane;
        {
            const float* ga_row[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) ga_row[r] = a.g_agg + (size_t)S.row_s[4 * lane + r] * a.ld_gagg + warp * CW;
            float plog[4] = {0.f, 0.f, 0.f, 0.f}, pdot[4] = {0.f, 0.f, 0.f, 0.f};
            const float* aw = S.vec_s + 3 * HP + warp * CW;
#pragma unroll
            for (int c = 0; c < CW; c += 4) {
                float4 gg[4];
#pragma unroll
                for (int r = 0; r < 4; ++r) gg[r] = __ldg(reinterpret_cast<const float4*>(ga_row[r] + c));
#pragma unroll
                for (int cc = 0; cc < 4; ++cc) {
                    const float4 p2 = __ldg(reinterpret_cast<const float4*>(pre2_base + (size_t)(c + cc) * GB_TM));
                    const float pv[4] = {p2.x, p2.y, p2.z, p2.w};
                    const float ww = aw[c + cc];
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float gadd = cc == 0 ? gg[r].x : cc == 1 ? gg[r].y : cc == 2 ? gg[r].z : gg[r].w;
                        const float gef = acc[r][c + cc] + gadd;
                        acc[r][c + cc] = gef;
                        const float q = silu_f(pv[r]);
                        plog[r] = fmaf(ww, q, plog[r]);
                        pdot[r] = fmaf(gef, q, pdot[r]);
                    }
                }
            }
            *reinterpret_cast<float4*>(S.red_s + warp * GB_TM + 4 * lane) = make_float4(plog[0], plog[1], plog[2], plog[3]);
            *reinterpret_cast<float4*>(S.red_s + (NW + warp) * GB_TM + 4 * lane) = make_float4(pdot[0], pdot[1], pdot[2], pdot[3]);
        }
        consumer_bar(NT);
        {
            float gate[4], kap[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float sl = a.att_b, sd = 0.f;
                for (int w = 0; w < NW; ++w) { sl += S.red_s[w * GB_TM + 4 * lane + r]; sd += S.red_s[(NW + w) * GB_TM + 4 * lane + r]; }
                if (a.attention) { gate[r] = sigmoid_f(sl); kap[r] = sd * gate[r] * (1.f - gate[r]); }
                else { gate[r] = 1.f; kap[r] = 0.f; }
            }
            const float* aw = S.vec_s + 3 * HP + warp * CW;
#pragma unroll
            for (int c = 0; c < CW; ++c) {
                const float4 p2 = __ldg(reinterpret_cast<const float4*>(pre2_base + (size_t)c * GB_TM));
                const float ww = aw[c];
                float4 o;
                o.x = (acc[0][c] * gate[0] + kap[0] * ww) * dsilu_f(p2.x);
                o.y = (acc[1][c] * gate[1] + kap[1] * ww) * dsilu_f(p2.y);
                o.z = (acc[2][c] * gate[2] + kap[2] * ww) * dsilu_f(p2.z);
                o.w = (acc[3][c] * gate[3] + kap[3] * ww) * dsilu_f(p2.w);
                *reinterpret_cast<float4*>(S.A_s + (warp * CW + c) * GB_MS + 4 * lane) = o;      // g_pre2
            }
        }
        consumer_bar(NT);
        zero_acc<CW>(acc);
        gemm_consume<HP, NW>(S.A_s, HP, acc, pipe, warp, lane);     // acc = g_s1
        // ---- epilogue 2: g_pre1 = g_s1 * SiLU'(pre1); radial / attr dots ------------------------------------------------
        {
            float pr[4] = {0.f, 0.f, 0.f, 0.f}, pa[4] = {0.f, 0.f, 0.f, 0.f};
            const float* wr = S.vec_s + warp * CW;
            const float* wa = S.vec_s + HP + warp * CW;
#pragma unroll
            for (int c = 0; c < CW; c += 4) {
                const int kq = (warp * CW + c) >> 2;
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float4 d1 = __ldg(reinterpret_cast<const float4*>(a.sv_d1 + (((size_t)tile * (HP / 4) + kq) * GB_TM + 4 * lane + r) * 4));
                    acc[r][c] *= d1.x; acc[r][c + 1] *= d1.y; acc[r][c + 2] *= d1.z; acc[r][c + 3] *= d1.w;
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        pr[r] = fmaf(wr[c + cc], acc[r][c + cc], pr[r]);
                        pa[r] = fmaf(wa[c + cc], acc[r][c + cc], pa[r]);
                    }
                }
            }
            *reinterpret_cast<float4*>(S.red_s + warp * GB_TM + 4 * lane) = make_float4(pr[0], pr[1], pr[2], pr[3]);
            *reinterpret_cast<float4*>(S.red_s + (NW + warp) * GB_TM + 4 * lane) = make_float4(pa[0], pa[1], pa[2], pa[3]);
        }
        consumer_bar(NT);                                            // GEMM 2 done everywhere: A_s reusable
#pragma unroll
        for (int c = 0; c < CW; ++c)
            *reinterpret_cast<float4*>(S.A_s + (warp * CW + c) * GB_MS + 4 * lane) = make_float4(acc[0][c], acc[1][c], acc[2][c], acc[3][c]);
        if (tid < GB_TM) {
            const int m = tid;
            float g_r = 0.f, g_a = 0.f;
            for (int w = 0; w < NW; ++w) { g_r += S.red_s[w * GB_TM + m]; g_a += S.red_s[(NW + w) * GB_TM + m]; }
            float gdx = 0.f, gdy = 0.f, gdz = 0.f;
            if (m < ne) {
                a.g_attr[e_lo + m] += g_a;
                const float dx = d_s[3 * m], dy = d_s[3 * m + 1], dz = d_s[3 * m + 2];
                const float gux = gu_s[3 * m], guy = gu_s[3 * m + 1], guz = gu_s[3 * m + 2];
                const float nrm = nrm_s[m], inv = 1.f / (nrm + 1.f);
                const float k2 = (gux * dx + guy * dy + guz * dz) * inv * inv / nrm;
                gdx = 2.f * g_r * dx + gux * inv - k2 * dx;
                gdy = 2.f * g_r * dy + guy * inv - k2 * dy;
                gdz = 2.f * g_r * dz + guz * inv - k2 * dz;
            }
            gd_s[3 * m] = gdx; gd_s[3 * m + 1] = gdy; gd_s[3 * m + 2] = gdz;
        }
        consumer_bar(NT);
        // rows: g_Pa (plain stores); columns: g_Pb (reduction over the tile, then one RED per (node, c))
        segsum_rows<HP, NW>(S.A_s, S.seg_s, nn, node_lo, a.g_Pa, HP, warp, lane);
        {
            const int t0 = g.tc_ptr[tile], t1 = g.tc_ptr[tile + 1];
            const int ncg = (HP + 31) / 32;
            for (int item = warp; item < (t1 - t0) * ncg; item += NW) {
                const int ti = t0 + item / ncg, c = (item % ncg) * 32 + lane;
                const int p0 = g.tc_start[ti], p1 = g.tc_start[ti + 1];
                if (c < HP) {
                    float sum = 0.f;
                    for (int p = p0; p < p1; ++p) sum += S.A_s[c * GB_MS + g.cperm[p]];
                    atomicAdd(a.g_Pb + (size_t)g.tc_node[ti] * HP + c, sum);
                }
            }
        }
        for (int idx = tid; idx < nn * 3; idx += NT) {
            const int nl = idx / 3, d = idx - 3 * nl;
            const int node = node_lo + nl;
            float sum = a.g_xout[3 * node + d] * g.node_mask[node];
            for (int m = S.seg_s[nl]; m < S.seg_s[nl + 1]; ++m) sum += gd_s[3 * m + d];
            atomicAdd(a.g_x + 3 * node + d, sum);
        }
        for (int idx = tid; idx < ne * 3; idx += NT) {
            const int m = idx / 3, d = idx - 3 * m;
            atomicAdd(a.g_x + 3 * S.col_s[m] + d, -gd_s[idx]);
        }
    }
}

template <int HP>
static void launch_fwd_t(bool save, const PredEdgeArgs& a, cudaStream_t s) {
    constexpr int NW = TileCfg<HP>::NW;
    const size_t smem = tile_kernel_smem_bytes(HP);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(pred_edge_fwd_kernel<HP, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(pred_edge_fwd_kernel<HP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(pred_edge_bwd_kernel<HP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = a.g.n_tiles < sms ? a.g.n_tiles : sms;
    if (save) pred_edge_fwd_kernel<HP, true><<<grid, (NW + 1) * 32, smem, s>>>(a);
    else pred_edge_fwd_kernel<HP, false><<<grid, (NW + 1) * 32, smem, s>>>(a);
}

template <int HP>
static void launch_bwd_t(const PredEdgeArgs& a, cudaStream_t s) {
    constexpr int NW = TileCfg<HP>::NW;
    const size_t smem = tile_kernel_smem_bytes(HP);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(pred_edge_bwd_kernel<HP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = a.g.n_tiles < sms ? a.g.n_tiles : sms;
    pred_edge_bwd_kernel<HP><<<grid, (NW + 1) * 32, smem, s>>>(a);
}

void launch_pred_edge_fwd(int HP, bool save, const PredEdgeArgs& a, cudaStream_t s) {
    if (a.g.n_tiles <= 0) return;
    switch (HP) {
        case 64: launch_fwd_t<64>(save, a, s); break;
        case 192: launch_fwd_t<192>(save, a, s); break;
        case 196: launch_fwd_t<196>(save, a, s); break;
        case 256: launch_fwd_t<256>(save, a, s); break;
        default: break;
    }
}

void launch_pred_edge_bwd(int HP, const PredEdgeArgs& a, cudaStream_t s) {
    if (a.g.n_tiles <= 0) return;
    switch (HP) {
        case 64: launch_bwd_t<64>(a, s); break;
        case 192: launch_bwd_t<192>(a, s); break;
        case 196: launch_bwd_t<196>(a, s); break;
        case 256: launch_bwd_t<256>(a, s); break;
        default: break;
    }
}

}  // namespace gb
