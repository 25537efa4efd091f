// HBM-bound per-molecule kernels of the reverse-diffusion step (one warp per molecule):
//   den_finish   EGNN_dynamics tail: vel=(x_out-x)*mask, NaN->0, remove_mean_with_mask, concat  (edm/egnn/models.py:116-152)
//   step_pre     mu = zt/alpha_ts - c*eps ; zs = mu + sigma*noise [; CoM projection]             (en_diffusion.py:831-851, 888-897)
//   step_guide   grad clip to norm 10, CoM removal, zs -= sigma*grad, CoM removal, nan_to_num    (en_diffusion.py:905-934)
//   decode       sample_p_xh_given_z0: x, one-hot(argmax)                                         (en_diffusion.py:533-560)
// plus the Philox noise source (masked, centre-of-gravity-free noise: utils.py:116-125,146-149) and the tiny head /
// embedding backward kernels of the predictor input gradient.
#include "common.cuh"
#include "kernels.h"
#include <float.h>

namespace gb {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int off = 16; off; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int off = 16; off; off >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, off));
    return v;
}
// max of non-negative floats (NaN compares as huge -> surfaces as a failed invariant, like the reference's assert)
__device__ __forceinline__ void atomic_max_pos(float* addr, float v) { atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v)); }
__device__ __forceinline__ float nan_to_num_f(float v) {     // torch.nan_to_num(v, 0.)
    if (isnan(v)) return 0.f;
    if (isinf(v)) return v > 0.f ? FLT_MAX : -FLT_MAX;
    return v;
}

// ---- Philox4x32-10 ------------------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0; c[1] = lo1; c[2] = n2; c[3] = lo0;
}
// standard normal for element `idx` of draw `draw` under `seed` (Box-Muller on two 24-bit uniforms)
__device__ __forceinline__ float philox_normal(unsigned long long seed, unsigned long long draw, uint32_t idx) {
    uint32_t c[4] = {idx, (uint32_t)draw, (uint32_t)(draw >> 32), 0x9E3779B9u};
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
#pragma unroll
    for (int i = 0; i < 10; ++i) { philox_round(c, k0, k1); k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    const float u1 = ((c[0] >> 8) + 1) * (1.0f / 16777216.0f);   // (0,1]
    const float u2 = (c[1] >> 8) * (1.0f / 16777216.0f);         // [0,1)
    return sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
}

// masked, x-centred noise value for element (b, i, d); needs the warp to own molecule b
__device__ __forceinline__ void molecule_noise(float* buf /*[N*D] smem or local via callback*/, const float* nm, int b, int N, int D,
                                               float std, unsigned long long seed, unsigned long long draw, int lane) {
    float sx = 0.f, sy = 0.f, sz = 0.f, cnt = 0.f;
    for (int e = lane; e < N * D; e += 32) {
        const int i = e / D, d = e - i * D;
        const float mk = nm[b * N + i];
        const float v = philox_normal(seed, draw, (uint32_t)(b * N * D + e)) * std * mk;
        buf[e] = v;
        if (d == 0) { sx += v; cnt += mk; } else if (d == 1) sy += v; else if (d == 2) sz += v;
    }
    sx = warp_sum(sx); sy = warp_sum(sy); sz = warp_sum(sz); cnt = fmaxf(warp_sum(cnt), 1.f);
    __syncwarp();
    for (int e = lane; e < N * D; e += 32) {
        const int i = e / D, d = e - i * D;
        if (d < 3) buf[e] -= ((d == 0 ? sx : d == 1 ? sy : sz) / cnt) * nm[b * N + i];
    }
    __syncwarp();
}

#define GB_MOL_SMEM_MAX 1024   // floats of per-warp scratch (N*D <= 1024)

__global__ void noise_kernel(float* out, const float* nm, int B, int N, int D, float std, unsigned long long seed,
                             unsigned long long draw) {
    extern __shared__ float sm[];
    const int wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float* buf = sm + warp * N * D;
    for (int b = blockIdx.x * wpb + warp; b < B; b += gridDim.x * wpb) {
        molecule_noise(buf, nm, b, N, D, std, seed, draw, lane);
        for (int e = lane; e < N * D; e += 32) out[(size_t)b * N * D + e] = buf[e];
        __syncwarp();
    }
}

void launch_noise(float* out, const float* nm, int B, int N, int D, float std, unsigned long long seed,
                  unsigned long long draw, cudaStream_t s) {
    const int wpb = 4;
    int blocks = (B + wpb - 1) / wpb; if (blocks > 148 * 8) blocks = 148 * 8;
    noise_kernel<<<blocks, wpb * 32, wpb * N * D * sizeof(float), s>>>(out, nm, B, N, D, std, seed, draw);
}

// ---- denoiser tail ----------------------------------------------------------------------------------------
__global__ void den_finish_kernel(DenFinishArgs a) {
    const int wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int D = 3 + a.F;
    float mzx = 0.f, mzc = 0.f, mex = 0.f, mec = 0.f, mvz = 0.f;
    float* stats = a.stats ? a.stats + (a.stats_step ? 8 * (*a.stats_step) : 0) : nullptr;
    for (int b = blockIdx.x * wpb + warp; b < a.B; b += gridDim.x * wpb) {
        float s[3] = {0.f, 0.f, 0.f}, zs[3] = {0.f, 0.f, 0.f}, cnt = 0.f;
        for (int i = lane; i < a.N; i += 32) {
            const int node = b * a.N + i;
            const float mk = a.node_mask[node];
            cnt += mk;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float v = (a.x_fin[3 * node + d] - a.x_in[3 * node + d]) * mk;
                if (isnan(v)) v = 0.f;
                s[d] += v;
                if (a.stats) {
                    const float z = a.zt[(size_t)node * D + d];
                    zs[d] += z; mzx = fmaxf(mzx, fabsf(z)); mvz = fmaxf(mvz, fabsf(z * (1.f - mk)));
                }
            }
        }
        cnt = fmaxf(warp_sum(cnt), 1.f);
#pragma unroll
        for (int d = 0; d < 3; ++d) { s[d] = warp_sum(s[d]); zs[d] = warp_sum(zs[d]); mzc = fmaxf(mzc, fabsf(zs[d])); }
        float es[3] = {0.f, 0.f, 0.f};
        for (int i = lane; i < a.N; i += 32) {
            const int node = b * a.N + i;
            const float mk = a.node_mask[node];
            float* out = a.eps + (size_t)node * D;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                float v = (a.x_fin[3 * node + d] - a.x_in[3 * node + d]) * mk;
                if (isnan(v)) v = 0.f;
                v = v - (s[d] / cnt) * mk;
                if (a.scrub_all) v = nan_to_num_f(v);
                out[d] = v;
                es[d] += v; mex = fmaxf(mex, fabsf(v));
            }
            for (int k = 0; k < a.F; ++k) {
                float v = a.h_out[(size_t)node * a.ld_h + k];
                if (a.scrub_all) v = nan_to_num_f(v);
                out[3 + k] = v;
            }
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) mec = fmaxf(mec, fabsf(warp_sum(es[d])));
    }
    if (a.stats) {
        mzx = warp_max(mzx); mex = warp_max(mex); mvz = warp_max(mvz);
        if (lane == 0) {
            atomic_max_pos(stats + 0, mzx); atomic_max_pos(stats + 1, mzc);
            atomic_max_pos(stats + 2, mex); atomic_max_pos(stats + 3, mec); atomic_max_pos(stats + 4, mvz);
        }
    }
}

static inline int mol_blocks(int B, int wpb) { int b = (B + wpb - 1) / wpb; return b > 148 * 8 ? 148 * 8 : (b < 1 ? 1 : b); }

void launch_den_finish(const DenFinishArgs& a, cudaStream_t s) { den_finish_kernel<<<mol_blocks(a.B, 8), 256, 0, s>>>(a); }

// ---- z_s = mu + sigma * noise ------------------------------------------------------------------------------
__global__ void step_pre_kernel(StepArgs a) {
    extern __shared__ float sm[];
    const int wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ND = a.N * a.D;
    float* buf = sm + warp * ND;
    const float alpha = a.coef[0], ceps = a.coef[1], sigma = a.coef[2];
    const unsigned long long draw = a.draw_ptr ? *a.draw_ptr : a.draw;
    const float* noise = a.noise ? a.noise + (a.draw_ptr ? draw * a.noise_stride : 0) : nullptr;
    for (int b = blockIdx.x * wpb + warp; b < a.B; b += gridDim.x * wpb) {
        const size_t base = (size_t)b * ND;
        if (noise) { for (int e = lane; e < ND; e += 32) buf[e] = noise[base + e]; __syncwarp(); }
        else molecule_noise(buf, a.node_mask, b, a.N, a.D, a.noise_std, a.seed, draw, lane);
        float s[3] = {0.f, 0.f, 0.f}, cnt = 0.f;
        for (int e = lane; e < ND; e += 32) {
            const float mu = __fsub_rn(__fdiv_rn(a.zt[base + e], alpha), __fmul_rn(ceps, a.eps[base + e]));
            const float v = __fadd_rn(mu, __fmul_rn(sigma, buf[e]));
            buf[e] = v;
            const int i = e / a.D, d = e - i * a.D;
            if (d < 3) s[d] += v;
            if (d == 0) cnt += a.node_mask[b * a.N + i];
        }
        if (a.project) {
            cnt = fmaxf(warp_sum(cnt), 1.f);
#pragma unroll
            for (int d = 0; d < 3; ++d) s[d] = warp_sum(s[d]) / cnt;
        }
        __syncwarp();
        for (int e = lane; e < ND; e += 32) {
            float v = buf[e];
            if (a.project) {
                const int i = e / a.D, d = e - i * a.D;
                if (d < 3) v = v - s[d] * a.node_mask[b * a.N + i];
            }
            a.zs[base + e] = v;
        }
        __syncwarp();
    }
}

void launch_step_pre(const StepArgs& a, cudaStream_t s) {
    const int wpb = 4;
    step_pre_kernel<<<mol_blocks(a.B, wpb), wpb * 32, wpb * a.N * a.D * sizeof(float), s>>>(a);
}

// ---- guidance update ------------------------------------------------------------------------------------------
__global__ void step_guide_kernel(GuideArgs a) {
    extern __shared__ float sm[];
    const int wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ND = a.N * a.D;
    float* buf = sm + warp * ND;
    const float sigma = a.coef[2];
    for (int b = blockIdx.x * wpb + warp; b < a.B; b += gridDim.x * wpb) {
        const size_t base = (size_t)b * ND;
        float n2 = 0.f, cnt = 0.f;
        for (int e = lane; e < ND; e += 32) {
            const float gq = a.grad[base + e];
            buf[e] = gq; n2 += gq * gq;
            if (e % a.D == 0) cnt += a.node_mask[b * a.N + e / a.D];
        }
        n2 = warp_sum(n2); cnt = fmaxf(warp_sum(cnt), 1.f);
        // en_diffusion.py:905-909.  torch.clamp(x, max=1) propagates a NaN norm (the whole molecule then becomes NaN and is zeroed
        // by the final nan_to_num); fminf alone would silently drop it
        const float craw = a.max_norm / (sqrtf(n2) + 1e-6f);
        const float coef = (craw != craw) ? craw : fminf(craw, 1.0f);
        float s[3] = {0.f, 0.f, 0.f};
        __syncwarp();
        for (int e = lane; e < ND; e += 32) {
            const float gq = buf[e] * coef;
            buf[e] = gq;
            const int d = e % a.D;
            if (d < 3) s[d] += gq;
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) s[d] = warp_sum(s[d]) / cnt;
        float t[3] = {0.f, 0.f, 0.f};
        __syncwarp();
        for (int e = lane; e < ND; e += 32) {
            const int i = e / a.D, d = e - i * a.D;
            float gq = buf[e];
            if (d < 3) gq = gq - s[d] * a.node_mask[b * a.N + i];                    // :911-919
            const float v = __fsub_rn(a.zs_pre[base + e], __fmul_rn(sigma, gq));     // :920
            buf[e] = v;
            if (d < 3) t[d] += v;
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) t[d] = warp_sum(t[d]) / cnt;
        __syncwarp();
        for (int e = lane; e < ND; e += 32) {
            const int i = e / a.D, d = e - i * a.D;
            float v = buf[e];
            if (d < 3) v = v - t[d] * a.node_mask[b * a.N + i];                      // :923-931
            a.zs[base + e] = nan_to_num_f(v);                                        // :933-934
        }
        __syncwarp();
    }
}

void launch_step_guide(const GuideArgs& a, cudaStream_t s) {
    const int wpb = 4;
    step_guide_kernel<<<mol_blocks(a.B, wpb), wpb * 32, wpb * a.N * a.D * sizeof(float), s>>>(a);
}

// ---- final decode ----------------------------------------------------------------------------------------------
__global__ void decode_kernel(DecodeArgs a) {
    extern __shared__ float sm[];
    const int wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ND = a.N * a.D, F = a.D - 3;
    float* buf = sm + warp * ND;
    const float sigma0 = a.coef[0], alpha0 = a.coef[1], sigma_x = a.coef[2];
    const float inv_alpha = __fdiv_rn(1.0f, alpha0);
    float cog = 0.f;
    for (int b = blockIdx.x * wpb + warp; b < a.B; b += gridDim.x * wpb) {
        const size_t base = (size_t)b * ND;
        if (a.noise) { for (int e = lane; e < ND; e += 32) buf[e] = a.noise[base + e]; __syncwarp(); }
        else molecule_noise(buf, a.node_mask, b, a.N, a.D, 1.0f, a.seed, a.draw, lane);
        float s[3] = {0.f, 0.f, 0.f};
        for (int i = lane; i < a.N; i += 32) {
            const int node = b * a.N + i;
            const float mk = a.node_mask[node];
            const float* z = a.z0 + (size_t)node * a.D;
            const float* ep = a.eps + (size_t)node * a.D;
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                const float mu = __fmul_rn(inv_alpha, __fsub_rn(z[d], __fmul_rn(sigma0, ep[d])));   // :501
                const float v = __fmul_rn(__fadd_rn(mu, __fmul_rn(sigma_x, buf[i * a.D + d])), a.norm_x);
                a.x[3 * node + d] = v;
                s[d] += v;
            }
            int best = 0; float bv = -INFINITY;
            for (int k = 0; k < F; ++k) {
                const float v = __fmul_rn(__fadd_rn(__fmul_rn(z[3 + k], a.norm_h), a.bias_h), mk);
                if (k == 0 || v > bv) { bv = v; best = k; }                     // first maximum wins, like torch.argmax
            }
            for (int k = 0; k < F; ++k) a.one_hot[(size_t)node * F + k] = (k == best ? 1.f : 0.f) * mk;
        }
#pragma unroll
        for (int d = 0; d < 3; ++d) cog = fmaxf(cog, fabsf(warp_sum(s[d])));
        __syncwarp();
    }
    if (lane == 0 && a.cog_max) atomic_max_pos(a.cog_max, cog);
}

void launch_decode(const DecodeArgs& a, cudaStream_t s) {
    const int wpb = 4;
    decode_kernel<<<mol_blocks(a.B, wpb), wpb * 32, wpb * a.N * a.D * sizeof(float), s>>>(a);
}

__global__ void cog_fix_kernel(float* x, const float* nm, const float* cog_max, float thresh, int B, int N) {
    if (!(*cog_max > thresh)) return;                           // en_diffusion.py:1000-1006 / 1059-1065
    const int wpb = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int b = blockIdx.x * wpb + warp; b < B; b += gridDim.x * wpb) {
        float s[3] = {0.f, 0.f, 0.f}, cnt = 0.f;
        for (int i = lane; i < N; i += 32) {
            cnt += nm[b * N + i];
            for (int d = 0; d < 3; ++d) s[d] += x[3 * (b * N + i) + d];
        }
        cnt = fmaxf(warp_sum(cnt), 1.f);
        for (int d = 0; d < 3; ++d) s[d] = warp_sum(s[d]) / cnt;
        for (int i = lane; i < N; i += 32)
            for (int d = 0; d < 3; ++d) x[3 * (b * N + i) + d] -= s[d] * nm[b * N + i];
    }
}

void launch_cog_fix(float* x, const float* nm, const float* cog_max, float thresh, int B, int N, cudaStream_t s) {
    cog_fix_kernel<<<mol_blocks(B, 8), 256, 0, s>>>(x, nm, cog_max, thresh, B, N);
}

// ---- predictor head: mean over padded nodes, and its backward ------------------------------------------------------
__global__ void pool_mean_kernel(const float* hout, int B, int N, int n_out, float* pred) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * n_out) return;
    const int b = idx / n_out, o = idx - b * n_out;
    float s = 0.f;
    for (int i = 0; i < N; ++i) s += hout[(size_t)(b * N + i) * n_out + o];
    pred[idx] = s / (float)N;                                   // .mean(1) over PADDED N, models.py:457
}

void launch_pool_mean(const float* hout, const float*, int B, int N, int n_out, float* pred, cudaStream_t s) {
    pool_mean_kernel<<<(B * n_out + 255) / 256, 256, 0, s>>>(hout, B, N, n_out, pred);
}

__global__ void head_bwd_kernel(HeadBwdArgs a) {
    // one thread per (node, 4 columns), one 16-byte store each
    const int q4 = a.HP >> 2, total = a.B * a.N * q4;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
        const int node = idx / q4, c0 = (idx - node * q4) << 2;
        const int b = node / a.N;
        const float sc = a.node_mask[node] / (float)a.N;
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int c = c0 + e;
            float acc = 0.f;
            if (c < a.H)
                for (int o = 0; o < a.n_out; ++o) acc = fmaf(a.g_pred[b * a.n_out + o], a.w[(size_t)o * a.H + c], acc);
            v[e] = c < a.H ? acc * sc : 0.f;
        }
        *reinterpret_cast<float4*>(a.gh + (size_t)node * a.HP + c0) = make_float4(v[0], v[1], v[2], v[3]);
    }
}

void launch_head_bwd(const HeadBwdArgs& a, cudaStream_t s) {
    const long long total = (long long)a.B * a.N * (a.HP >> 2);
    head_bwd_kernel<<<(int)min((long long)148 * 16, (total + 255) / 256), 256, 0, s>>>(a);
}

// dz from dL/dh0 (embedding), dL/dx0 and the accumulated edge-attribute gradient (a_ij = |x0_i - x0_j|^2 feeds every
// layer): node-parallel over the CSR (row) and CSC (column) views, fixed summation order, no atomics.
__global__ void in_bwd_kernel(InBwdArgs a) {
    const Graph& g = a.g;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int node = warp; node < g.n_nodes; node += nwarps) {
        const float mk = g.node_mask[node];
        float* out = a.gz + (size_t)node * a.D;
        if (lane < 3) {
            float sum = a.gx0[3 * node + lane];
            const float xi = a.x0[3 * node + lane];
            for (int e = g.rowptr[node]; e < g.rowptr[node + 1]; ++e)
                sum += 2.f * a.g_attr[e] * (xi - a.x0[3 * g.ecol[e] + lane]);
            for (int p = g.colptr[node]; p < g.colptr[node + 1]; ++p) {
                const int e = g.cedge[p];
                sum -= 2.f * a.g_attr[e] * (a.x0[3 * g.erow[e] + lane] - xi);
            }
            out[lane] = sum * mk;
        }
        const float* gh = a.gh0 + (size_t)node * a.HP;
        for (int k = 0; k < a.F; ++k) {
            float p = 0.f;
            for (int c = lane; c < a.H; c += 32) p = fmaf(gh[c], a.w_in[(size_t)c * (a.F + 1) + k], p);
            p = warp_sum(p);
            if (lane == 0) out[3 + k] = p * mk;
        }
    }
}

void launch_in_bwd(const InBwdArgs& a, cudaStream_t s) {
    in_bwd_kernel<<<min(148 * 8, (a.g.n_nodes + 7) / 8), 256, 0, s>>>(a);
}

}  // namespace gb
