// Fused per-edge kernels of one denoiser EquivariantBlock (edm/egnn/egnn_new.py:214-235):
//   mode 0  GCL.edge_model + attention gate + unsorted_segment_sum      (egnn_new.py:42-67, 403-414)
//   mode 1  EquivariantUpdate.coord_model + tanh*range + segment sum      (egnn_new.py:119-155)
// One CTA owns a tile of <=128 compacted edges whose row segments are complete, so the segment sums are
// plain shared-memory reductions (no atomics, deterministic).  Per tile:
//   build   A[k][m] = SiLU( Pa[row_m][k] + Pb[col_m][k] + w_r[k] r_m + w_d[k] d0_m )        (CUDA cores, MUFU)
//   GEMM    acc     = A^T-tile x W2^T                                                        (FP32 FFMA, TMA-fed ring)
//   mode 0  m = SiLU(acc+b2); gate = sigmoid(w_a.m + b_a); agg_i = sum_j m*gate
//   mode 1  phi = w7 . SiLU(acc+b6); x_i' = (x_i + sum_j u_ij tanh(phi) range) mask_i
#include "common.cuh"
#include "kernels.h"

namespace gb {

template <int HP, int MODE>
__global__ void __launch_bounds__((TileCfg<HP>::NW + 1) * 32, 1) den_edge_kernel(DenEdgeArgs a) {
    constexpr int NW = TileCfg<HP>::NW;
    constexpr int CW = HP / NW;
    constexpr int NT = NW * 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* A_s = reinterpret_cast<float*>(smem_raw);
    float* ring = A_s + HP * GB_MS;
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + GB_STAGES * GB_KC * HP);
    uint64_t* empty = full + GB_STAGES;
    float* vec_s = reinterpret_cast<float*>(empty + GB_STAGES);   // [4][HP]: w_r, w_d, b2, vecw
    float* red_s = vec_s + 4 * HP;                                 // [NW][128]
    float* r_s = red_s + NW * GB_TM;                               // [128]
    float* d0_s = r_s + GB_TM;                                     // [128]
    float* u_s = d0_s + GB_TM;                                     // [128][3]
    int* row_s = reinterpret_cast<int*>(u_s + 3 * GB_TM);          // [128]
    int* col_s = row_s + GB_TM;                                    // [128]
    int* seg_s = col_s + GB_TM;                                    // [129] tile-local segment starts

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < GB_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NW); }
        mbar_fence_init();
    }
    __syncthreads();
    WPipe<HP> pipe;
    pipe.init_side(ring, full, empty);
    const Graph& g = a.g;

    if (warp == NW) {
        if (lane == 0)
            for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) pipe.produce(a.wt2, HP);
        return;
    }
    for (int i = tid; i < HP; i += NT) {
        vec_s[i] = a.ext[i]; vec_s[HP + i] = a.ext[HP + i]; vec_s[2 * HP + i] = a.b2[i]; vec_s[3 * HP + i] = a.vecw[i];
    }

    for (int tile = blockIdx.x; tile < g.n_tiles; tile += gridDim.x) {
        const int node_lo = g.tile_ptr[tile], node_hi = g.tile_ptr[tile + 1];
        const int nn = node_hi - node_lo;
        const int e_lo = g.rowptr[node_lo], ne = g.rowptr[node_hi] - e_lo;
        consumer_bar(NT);                                   // previous tile fully consumed
        // ---- per-edge scalars -------------------------------------------------------------------
        for (int m = tid; m < GB_TM; m += NT) {
            int row = 0, col = 0; float r = 0.f, d0 = 0.f, ux = 0.f, uy = 0.f, uz = 0.f;
            if (m < ne) {
                const int e = e_lo + m;
                row = g.erow[e]; col = g.ecol[e];
                if (a.eattr) { r = a.eattr[2 * e]; d0 = a.eattr[2 * e + 1]; }
                else {
                    const float dx = a.x[3 * row] - a.x[3 * col], dy = a.x[3 * row + 1] - a.x[3 * col + 1], dz = a.x[3 * row + 2] - a.x[3 * col + 2];
                    r = dx * dx + dy * dy + dz * dz;                                  // coord2diff, egnn_new.py:394-400
                    const float ex = a.x0[3 * row] - a.x0[3 * col], ey = a.x0[3 * row + 1] - a.x0[3 * col + 1], ez = a.x0[3 * row + 2] - a.x0[3 * col + 2];
                    d0 = ex * ex + ey * ey + ez * ez;
                    if (a.d0_edge) d0 = a.d0_edge[e];
                    if (MODE == 1) {
                        const float inv = 1.f / (sqrtf(r + 1e-8f) + a.norm_constant);
                        ux = dx * inv; uy = dy * inv; uz = dz * inv;
                    }
                }
                if (MODE == 1 && a.cdiff) { ux = a.cdiff[3 * e]; uy = a.cdiff[3 * e + 1]; uz = a.cdiff[3 * e + 2]; }
            }
            row_s[m] = row; col_s[m] = col; r_s[m] = r; d0_s[m] = d0;
            if (MODE == 1) { u_s[3 * m] = ux; u_s[3 * m + 1] = uy; u_s[3 * m + 2] = uz; }
        }
        for (int i = tid; i <= nn; i += NT) seg_s[i] = g.rowptr[node_lo + i] - e_lo;
        consumer_bar(NT);
        // ---- build the activation tile ------------------------------------------------------------
        for (int idx = tid; idx < GB_TM * (HP / 4); idx += NT) {
            const int m = idx & (GB_TM - 1), kq = idx >> 7;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < ne) {
                const float4 pa = __ldg(reinterpret_cast<const float4*>(a.P + (size_t)row_s[m] * (2 * HP) + 4 * kq));
                const float4 pb = __ldg(reinterpret_cast<const float4*>(a.P + (size_t)col_s[m] * (2 * HP) + HP + 4 * kq));
                const float4 wr = *reinterpret_cast<const float4*>(vec_s + 4 * kq);
                const float4 wd = *reinterpret_cast<const float4*>(vec_s + HP + 4 * kq);
                const float r = r_s[m], d0 = d0_s[m];
                v.x = silu_f(pa.x + pb.x + wr.x * r + wd.x * d0);
                v.y = silu_f(pa.y + pb.y + wr.y * r + wd.y * d0);
                v.z = silu_f(pa.z + pb.z + wr.z * r + wd.z * d0);
                v.w = silu_f(pa.w + pb.w + wr.w * r + wd.w * d0);
            }
            float* d = A_s + (4 * kq) * GB_MS + m;
            d[0] = v.x; d[GB_MS] = v.y; d[2 * GB_MS] = v.z; d[3 * GB_MS] = v.w;
        }
        consumer_bar(NT);
        // ---- GEMM ------------------------------------------------------------------------------------
        float acc[4][CW];
        zero_acc<CW>(acc);
        gemm_consume<HP, NW>(A_s, HP, acc, pipe, warp, lane);
        // ---- epilogue: activation + row dot with vecw ---------------------------------------------------
        {
            float part[4] = {0.f, 0.f, 0.f, 0.f};
            const float* b2 = vec_s + 2 * HP + warp * CW;
            const float* vw = vec_s + 3 * HP + warp * CW;
#pragma unroll
            for (int c = 0; c < CW; ++c) {
                const float bb = b2[c], ww = vw[c];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const float v = silu_f(acc[r][c] + bb);
                    acc[r][c] = v;
                    part[r] = fmaf(ww, v, part[r]);
                }
            }
            *reinterpret_cast<float4*>(red_s + warp * GB_TM + 4 * lane) = make_float4(part[0], part[1], part[2], part[3]);
        }
        consumer_bar(NT);   // all warps finished the GEMM (A_s reusable) and published their partial dots
        if (MODE == 0) {
            float gate[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float sdot = a.att_b;
                for (int w = 0; w < NW; ++w) sdot += red_s[w * GB_TM + 4 * lane + r];
                gate[r] = a.attention ? sigmoid_f(sdot) : 1.f;
            }
#pragma unroll
            for (int c = 0; c < CW; ++c)
                *reinterpret_cast<float4*>(A_s + (warp * CW + c) * GB_MS + 4 * lane) =
                    make_float4(acc[0][c] * gate[0], acc[1][c] * gate[1], acc[2][c] * gate[2], acc[3][c] * gate[3]);
            consumer_bar(NT);
            // segment sums: lane = (column within group of 8, node within group of 4)
            const int ncg = (HP + 7) / 8, nng = (nn + 3) / 4;
            for (int item = warp; item < ncg * nng; item += NW) {
                const int cg = item % ncg, ng = item / ncg;
                const int c = cg * 8 + (lane & 7), nl = ng * 4 + (lane >> 3);
                if (c < HP && nl < nn) {
                    const int m0 = seg_s[nl], m1 = seg_s[nl + 1];
                    float sum = 0.f;
                    for (int m = m0; m < m1; ++m) sum += A_s[c * GB_MS + m];
                    a.agg[(size_t)(node_lo + nl) * HP + c] = sum / a.normf;
                }
            }
        } else {
            if (tid < GB_TM) {
                const int m = tid;
                float phi = 0.f;
                for (int w = 0; w < NW; ++w) phi += red_s[w * GB_TM + m];
                const float sc = a.use_tanh ? tanhf(phi) * a.coords_range : phi;
                if (a.use_tanh) {          // (coord_diff * tanh(phi)) * range, egnn_new.py:123-127
                    const float th = tanhf(phi);
                    u_s[3 * m] = u_s[3 * m] * th * a.coords_range;
                    u_s[3 * m + 1] = u_s[3 * m + 1] * th * a.coords_range;
                    u_s[3 * m + 2] = u_s[3 * m + 2] * th * a.coords_range;
                } else {
                    u_s[3 * m] *= sc; u_s[3 * m + 1] *= sc; u_s[3 * m + 2] *= sc;
                }
            }
            consumer_bar(NT);
            for (int idx = tid; idx < nn * 3; idx += NT) {
                const int nl = idx / 3, d = idx - 3 * nl;
                const int m0 = seg_s[nl], m1 = seg_s[nl + 1];
                float sum = 0.f;
                for (int m = m0; m < m1; ++m) sum += u_s[3 * m + d];
                const int node = node_lo + nl;
                a.x_out[3 * node + d] = (a.x[3 * node + d] + sum / a.normf) * g.node_mask[node];
            }
        }
    }
}

template <int HP>
static void launch_den_edge_t(int mode, const DenEdgeArgs& a, cudaStream_t s) {
    constexpr int NW = TileCfg<HP>::NW;
    const size_t smem = tile_kernel_smem_bytes(HP);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(den_edge_kernel<HP, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(den_edge_kernel<HP, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int grid = a.g.n_tiles < sms ? a.g.n_tiles : sms;
    if (mode == 0) den_edge_kernel<HP, 0><<<grid, (NW + 1) * 32, smem, s>>>(a);
    else den_edge_kernel<HP, 1><<<grid, (NW + 1) * 32, smem, s>>>(a);
}

void launch_den_edge(int HP, int mode, const DenEdgeArgs& a, cudaStream_t s) {
    if (a.g.n_tiles <= 0) return;
    switch (HP) {
        case 64: launch_den_edge_t<64>(mode, a, s); break;
        case 192: launch_den_edge_t<192>(mode, a, s); break;
        case 196: launch_den_edge_t<196>(mode, a, s); break;
        case 256: launch_den_edge_t<256>(mode, a, s); break;
        default: break;
    }
}

}  // namespace gb
