// Node-level fused Linear kernels (node MLPs, first-layer factorisation P = h W1a|W1b, backward dgrads).
// Reference ops replaced: nn.Linear inside GCL.node_model (edm/egnn/egnn_new.py:59-73),
// E_GCL.node_model (edm/egnn_predictor/gcl.py:240-250) and the h[row]/h[col] halves of the first
// edge/coord Linear, which are applied per NODE here instead of per EDGE (exact algebraic factorisation:
// W[h_i,h_j,e] = W_a h_i + W_b h_j + W_e e).
#include "common.cuh"
#include "kernels.h"

namespace gb {

template <int HP>
__global__ void __launch_bounds__((TileCfg<HP>::NW + 1) * 32, 1) lin_kernel(LinArgs a) {
    constexpr int NW = TileCfg<HP>::NW;
    constexpr int CW = HP / NW;
    constexpr int NT = NW * 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* A_s = reinterpret_cast<float*>(smem_raw);
    float* ring = A_s + HP * GB_MS;
    uint64_t* full = reinterpret_cast<uint64_t*>(ring + GB_STAGES * GB_KC * HP);
    uint64_t* empty = full + GB_STAGES;

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int i = 0; i < GB_STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NW); }
        mbar_fence_init();
    }
    __syncthreads();

    const int cb = blockIdx.y;
    const int Ktot = a.K1 + a.K2;
    const float* wt = a.wt + (size_t)cb * Ktot * HP;
    const int n_tiles = (a.M + GB_TM - 1) / GB_TM;
    WPipe<HP> pipe;
    pipe.init_side(ring, full, empty);

    if (warp == NW) {                       // ---- producer warp ----
        if (lane == 0) {
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                pipe.produce(wt, a.K1);
                if (a.K2 > 0) pipe.produce(wt + (size_t)a.K1 * HP, a.K2);
            }
        }
        return;
    }

    const float* bias = a.bias ? a.bias + cb * HP + warp * CW : nullptr;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int row0 = tile * GB_TM;
        float acc[4][CW];
        zero_acc<CW>(acc);
#pragma unroll 1
        for (int part = 0; part < 2; ++part) {
            const float* A = part ? a.A2 : a.A1;
            const int lda = part ? a.lda2 : a.lda1;
            const int K = part ? a.K2 : a.K1;
            if (K == 0) break;
            consumer_bar(NT);               // previous readers of A_s are done
            for (int idx = tid; idx < GB_TM * (K / 4); idx += NT) {
                const int m = idx & (GB_TM - 1), kq = idx >> 7;
                const int row = row0 + m;
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (row < a.M) {
                    v = __ldg(reinterpret_cast<const float4*>(A + (size_t)row * lda + 4 * kq));
                    if (a.rowscale) { const float sc = __ldg(a.rowscale + row); v.x *= sc; v.y *= sc; v.z *= sc; v.w *= sc; }
                }
                float* d = A_s + (4 * kq) * GB_MS + m;
                d[0] = v.x; d[GB_MS] = v.y; d[2 * GB_MS] = v.z; d[3 * GB_MS] = v.w;
            }
            consumer_bar(NT);
            gemm_consume<HP, NW>(A_s, K, acc, pipe, warp, lane);
        }
        // ---- epilogue straight from registers ----
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int row = row0 + 4 * lane + r;
            if (row >= a.M) continue;
            const size_t col0 = (size_t)cb * HP + warp * CW;
            float mk = 1.f;
            if (a.epi == EPI_RES_MASK) mk = __ldg(a.mask + row);
            else if (a.epi == EPI_ADD_RES && a.mask) mk = __ldg(a.mask + row);
#pragma unroll
            for (int c = 0; c < CW; c += 4) {
                float v[4] = {acc[r][c], acc[r][c + 1], acc[r][c + 2], acc[r][c + 3]};
                if (bias) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(bias + c));
                    v[0] += b.x; v[1] += b.y; v[2] += b.z; v[3] += b.w;
                }
                if (a.epi == EPI_SILU) {
                    if (a.out2) *reinterpret_cast<float4*>(a.out2 + (size_t)row * a.ldo2 + col0 + c) = make_float4(v[0], v[1], v[2], v[3]);
#pragma unroll
                    for (int q = 0; q < 4; ++q) v[q] = silu_f(v[q]);
                } else if (a.epi == EPI_RES_MASK) {
                    const float4 rs = __ldg(reinterpret_cast<const float4*>(a.res + (size_t)row * a.ldr + col0 + c));
                    v[0] = (rs.x + v[0]) * mk; v[1] = (rs.y + v[1]) * mk; v[2] = (rs.z + v[2]) * mk; v[3] = (rs.w + v[3]) * mk;
                } else if (a.epi == EPI_MUL_DSILU) {
                    const float4 p = __ldg(reinterpret_cast<const float4*>(a.aux + (size_t)row * a.ldaux + col0 + c));
                    v[0] *= dsilu_f(p.x); v[1] *= dsilu_f(p.y); v[2] *= dsilu_f(p.z); v[3] *= dsilu_f(p.w);
                } else if (a.epi == EPI_ADD_RES && (a.res_cb < 0 || cb == a.res_cb)) {
                    const float4 rs = __ldg(reinterpret_cast<const float4*>(a.res + (size_t)row * a.ldr + col0 + c));
                    v[0] += rs.x * mk; v[1] += rs.y * mk; v[2] += rs.z * mk; v[3] += rs.w * mk;
                }
                *reinterpret_cast<float4*>(a.out + (size_t)row * a.ldo + col0 + c) = make_float4(v[0], v[1], v[2], v[3]);
            }
        }
    }
}

size_t tile_kernel_smem_bytes(int HP) {
    // A tile + weight ring + barriers + per-kernel scratch (vectors, reductions, per-edge scalars)
    size_t base = (size_t)HP * GB_MS * 4 + (size_t)GB_STAGES * GB_KC * HP * 4 + 2 * GB_STAGES * 8;
    size_t scratch = (size_t)8 * HP * 4 + (size_t)3 * 8 * GB_TM * 4 + (size_t)GB_TM * 4 * 16 + 1024;
    return base + scratch;
}

template <int HP>
static void launch_lin_t(const LinArgs& a, cudaStream_t s) {
    constexpr int NW = TileCfg<HP>::NW;
    const size_t smem = SmemLayout<HP>::base_bytes;
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(lin_kernel<HP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        configured = true;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_tiles = (a.M + GB_TM - 1) / GB_TM;
    int gx = n_tiles;
    const int cap = max(1, sms / a.ncb);
    if (gx > cap) gx = cap;
    dim3 grid(gx, a.ncb);
    lin_kernel<HP><<<grid, (NW + 1) * 32, smem, s>>>(a);
}

void launch_lin(int HP, const LinArgs& a, cudaStream_t s) {
    if (a.M <= 0) return;
    switch (HP) {
        case 64: launch_lin_t<64>(a, s); break;
        case 192: launch_lin_t<192>(a, s); break;
        case 196: launch_lin_t<196>(a, s); break;
        case 256: launch_lin_t<256>(a, s); break;
        default: break;
    }
}

// ------------------------------------------------------------------------------------------------
// tiny-K / tiny-N heads
// ------------------------------------------------------------------------------------------------
__global__ void embed_in_kernel(EmbedInArgs a) {
    // one thread per (node, 4 output features), one 16-byte store each (HBM-write-bound: n_nodes x HP x 4 B); K = F+1 <= 16.
    // HP is a multiple of 4 and so is H for every supported width, but the bounds are checked per element anyway.
    const int F = a.D - 3, q4 = a.HP >> 2;
    const int total = a.n_nodes * q4;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int node = idx / q4, c0 = (idx - node * q4) << 2;
        const float mk = a.node_mask[node];
        const float* zr = a.z + (size_t)node * a.D;
        const float t = a.t_per_mol ? a.t_ptr[node / a.N] : a.t_ptr[0];
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = c0 + e;
            v[e] = 0.f;
            if (c < a.H) {
                const float* wr = a.w + (size_t)c * (F + 1);
                float acc = a.b[c];
                for (int k = 0; k < F; ++k) acc = fmaf(zr[3 + k] * mk, wr[k], acc);
                v[e] = fmaf(t, wr[F], acc);
            }
        }
        *reinterpret_cast<float4*>(a.h + (size_t)node * a.HP + c0) = make_float4(v[0], v[1], v[2], v[3]);
        if (c0 == 0) { a.x[(size_t)node * 3] = zr[0] * mk; a.x[(size_t)node * 3 + 1] = zr[1] * mk; a.x[(size_t)node * 3 + 2] = zr[2] * mk; }
    }
}

void launch_embed_in(const EmbedInArgs& a, cudaStream_t s) {
    const long long total = (long long)a.n_nodes * (a.HP >> 2);
    int blocks = (int)min((long long)148 * 16, (total + 255) / 256);
    embed_in_kernel<<<blocks, 256, 0, s>>>(a);
}

__global__ void embed_plain_kernel(const float* h_in, int K, const float* w, const float* b, int n_nodes, int H, int HP, float* h) {
    const long long total = (long long)n_nodes * HP;
    for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
        const int node = (int)(idx / HP), c = (int)(idx % HP);
        float v = 0.f;
        if (c < H) {
            v = b[c];
            for (int k = 0; k < K; ++k) v = fmaf(h_in[(size_t)node * K + k], w[(size_t)c * K + k], v);
        }
        h[idx] = v;
    }
}

void launch_embed_plain(const float* h_in, int K, const float* w, const float* b, int n_nodes, int H, int HP, float* h, cudaStream_t s) {
    const long long total = (long long)n_nodes * HP;
    embed_plain_kernel<<<(int)min((long long)148 * 16, (total + 255) / 256), 256, 0, s>>>(h_in, K, w, b, n_nodes, H, HP, h);
}

__global__ void embed_out_kernel(EmbedOutArgs a) {
    // one warp per node: the row of h is read once (lane l keeps columns l, l+32, ...; H <= 256), then n_out <= 16 dot products
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int node = warp; node < a.n_nodes; node += nwarps) {
        const float* hr = a.h + (size_t)node * a.HP;
        const float mk = a.node_mask[node];
        float hv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int k = lane + 32 * i; hv[i] = k < a.H ? hr[k] : 0.f; }
        for (int o = 0; o < a.n_out; ++o) {
            const float* wr = a.w + (size_t)o * a.H;
            float p = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) { const int k = lane + 32 * i; if (k < a.H) p = fmaf(hv[i], wr[k], p); }
#pragma unroll
            for (int off = 16; off; off >>= 1) p += __shfl_xor_sync(0xffffffffu, p, off);
            if (lane == 0) a.out[(size_t)node * a.ldo + o] = (p + a.b[o]) * mk;
        }
    }
}

void launch_embed_out(const EmbedOutArgs& a, cudaStream_t s) {
    int blocks = min(148 * 8, (a.n_nodes + 7) / 8);
    if (blocks < 1) blocks = 1;
    embed_out_kernel<<<blocks, 256, 0, s>>>(a);
}

// dst[k*HP + n] = W[(n+n_off)*ld + k_off + k]  (transpose)   or   W[(k+k_off)*ld + n_off + n]
__global__ void pack_kernel(float* dst, const float* src, int ld, int k_off, int n_off, int Kv, int Nv, int Kp,
                            int HP, int transpose) {
    const int total = Kp * HP;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int k = idx / HP, n = idx % HP;
        float v = 0.f;
        if (k < Kv && n < Nv) v = transpose ? src[(size_t)(n + n_off) * ld + k_off + k] : src[(size_t)(k + k_off) * ld + n_off + n];
        dst[idx] = v;
    }
}

void launch_pack(float* dst, const float* src, int ld, int k_off, int n_off, int Kv, int Nv, int Kp, int HP,
                 int transpose, cudaStream_t s) {
    const int total = Kp * HP;
    pack_kernel<<<(total + 255) / 256, 256, 0, s>>>(dst, src, ld, k_off, n_off, Kv, Nv, Kp, HP, transpose);
}

}  // namespace gb
