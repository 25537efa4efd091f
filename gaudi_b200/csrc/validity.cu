// Geometric validity of generated ring graphs: the per-molecule triple Python loops of
// positions2adj (utils/helpers.py:172-196), check_stability / check_angels3 / check_angels4 (analyze/analyze.py:19-100) and
// find_triplets_quads / angel3 / angel4 (analyze/analyze.py:234-318) as one thread per molecule.  Everything here is small
// integer / bit-set work on <= 16 rings plus a handful of fp32 angles; the batch is the parallel axis.
#include "../../include/gaudi_b200.h"
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace gb {

constexpr int MAXR = 16;          // rings per molecule (cata <= 11, hetro <= 10)
constexpr int MAXRANGE = 4;       // angle ranges per ring symbol

struct ValidityTables {
    const float* pair_lo;   // [T*T] lo*(1-tol) as fp32, +inf where the pair has no entry
    const float* pair_hi;   // [T*T] hi*(1+tol)
    const float* a3_lo;     // [T*MAXRANGE]
    const float* a3_hi;
    const int* a3_cnt;      // [T] number of ranges, -1: symbol missing from the table (the reference raises KeyError)
    float min_dist;         // min lo * (1-tol)
    float a4_hi;            // a4["180"]*(1-tol)
    float a4_lo;            // a4["0"]*(1+tol)
    int check_a4;           // cata only (analyze.py:40-41)
    int n_types;
    int orientation_type;   // -1: no orientation nodes (cata)
};

// Rounding model of the reference's CPU ops (probed against torch 2.11 / AVX-512, tests/test_validity.py pins it):
// torch.dot and torch.sum accumulate the rounded products left to right; torch.norm / linalg.norm contract to FMAs.
__device__ __forceinline__ float norm3(float a, float b, float c) {
    return sqrtf(__fmaf_rn(c, c, __fmaf_rn(b, b, __fmul_rn(a, a))));
}

__device__ __forceinline__ float angel3_dev(const float* p0, const float* p1, const float* p2) {
    // analyze.py:234-240: rad2deg(acos(v1.v2 / (|v1| |v2|)))
    const float ax = p0[0] - p1[0], ay = p0[1] - p1[1], az = p0[2] - p1[2];
    const float bx = p2[0] - p1[0], by = p2[1] - p1[1], bz = p2[2] - p1[2];
    const float dot = __fadd_rn(__fadd_rn(__fmul_rn(ax, bx), __fmul_rn(ay, by)), __fmul_rn(az, bz));
    const float na = norm3(ax, ay, az), nb = norm3(bx, by, bz);
    const float a = acosf(__fdiv_rn(dot, __fmul_rn(na, nb))) * 57.295779513082320876798154814105f;
    return a >= 0.f ? a : a + 360.f;
}

__device__ __forceinline__ float dot3(const float* a, const float* b) {
    return __fadd_rn(__fadd_rn(__fmul_rn(a[0], b[0]), __fmul_rn(a[1], b[1])), __fmul_rn(a[2], b[2]));
}

__device__ __forceinline__ float angel4_dev(const float* p0, const float* p1, const float* p2, const float* p3) {
    // analyze.py:243-273 (Praxeolitic dihedral), |degrees|
    float b0[3], b1[3], b2[3], v[3], w[3], c[3];
    for (int d = 0; d < 3; ++d) { b0[d] = -1.0f * (p1[d] - p0[d]); b1[d] = p2[d] - p1[d]; b2[d] = p3[d] - p2[d]; }
    const float n1 = norm3(b1[0], b1[1], b1[2]);
    for (int d = 0; d < 3; ++d) b1[d] = __fdiv_rn(b1[d], n1);
    const float d0 = dot3(b0, b1), d2 = dot3(b2, b1);
    for (int d = 0; d < 3; ++d) { v[d] = __fsub_rn(b0[d], __fmul_rn(d0, b1[d])); w[d] = __fsub_rn(b2[d], __fmul_rn(d2, b1[d])); }
    c[0] = __fsub_rn(__fmul_rn(b1[1], v[2]), __fmul_rn(b1[2], v[1]));
    c[1] = __fsub_rn(__fmul_rn(b1[2], v[0]), __fmul_rn(b1[0], v[2]));
    c[2] = __fsub_rn(__fmul_rn(b1[0], v[1]), __fmul_rn(b1[1], v[0]));
    const float x = dot3(v, w), y = dot3(c, w);
    return fabsf(atan2f(y, x) * 57.295779513082320876798154814105f);
}

// flags[b][0..4] = orientation_nodes, dist_stable, connected, angels3, angels4; [5] = all five; [6] = error bits
// (1: no ring nodes, 2: more than MAXR rings, 4: angle table has no entry for a centre ring symbol), [7] = number of rings
__global__ void stability_kernel(const float* __restrict__ x, const int* __restrict__ ring_type, const float* __restrict__ node_mask,
                                 int B, int N, ValidityTables t, unsigned char* __restrict__ flags) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    unsigned char* out = flags + (size_t)b * 8;
    for (int i = 0; i < 8; ++i) out[i] = 0;
    out[0] = 1;                                             // results start as {orientation: True, others False}
    int idx[2 * MAXR], n_tot = 0;
    for (int i = 0; i < N; ++i)
        if (node_mask[(size_t)b * N + i] != 0.f) { if (n_tot < 2 * MAXR) idx[n_tot] = i; ++n_tot; }
    if (n_tot > 2 * MAXR) { out[6] = 2; return; }
    int n = n_tot;
    if (t.orientation_type >= 0) {                          // analyze.py:62-73
        n = n_tot / 2;
        bool ok = n_tot - n > 0;
        for (int i = n; i < n_tot; ++i) ok = ok && ring_type[(size_t)b * N + idx[i]] == t.orientation_type;
        for (int i = 0; i < n; ++i) ok = ok && ring_type[(size_t)b * N + idx[i]] != t.orientation_type;
        if (!ok) { out[0] = 0; out[7] = (unsigned char)n; return; }
    }
    out[7] = (unsigned char)n;
    if (n == 0) { out[6] = 1; return; }
    if (n > MAXR) { out[6] = 2; return; }
    float p[MAXR][3];
    int rt[MAXR];
    for (int i = 0; i < n; ++i) {
        for (int d = 0; d < 3; ++d) p[i][d] = x[((size_t)b * N + idx[i]) * 3 + d];
        rt[i] = ring_type[(size_t)b * N + idx[i]];
    }
    // ---- positions2adj + minimum-distance check (helpers.py:164-196, analyze.py:80-84) ----
    uint32_t adj[MAXR];
    for (int i = 0; i < n; ++i) adj[i] = 0;
    bool too_close = false;
    for (int i = 0; i < n; ++i)
        for (int j = i + 1; j < n; ++j) {
            const float dx = p[i][0] - p[j][0], dy = p[i][1] - p[j][1], dz = p[i][2] - p[j][2];
            const float dist = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
            if (dist < t.min_dist) too_close = true;
            const int k = rt[i] * t.n_types + rt[j];
            if (t.pair_lo[k] < dist && dist < t.pair_hi[k]) { adj[i] |= 1u << j; adj[j] |= 1u << i; }
        }
    if (too_close) return;
    out[1] = 1;
    // ---- connectivity: BFS from ring 0 with neighbours in ascending order (nx.bfs_edges), recording the tree edges ----
    int queue[MAXR], parent_of[MAXR], qh = 0, qt = 0;
    uint32_t seen = 1u;
    queue[qt++] = 0; parent_of[0] = -1;
    while (qh < qt) {
        const int u = queue[qh++];
        uint32_t nb = adj[u] & ~seen;
        while (nb) { const int v = __ffs(nb) - 1; nb &= nb - 1; seen |= 1u << v; parent_of[v] = u; queue[qt++] = v; }
    }
    if (qt != n) return;
    out[2] = 1;
    // ---- triplets around every BFS tree edge (analyze.py:282-291), as a set: trip[centre][a] has bit b (a < b) ----
    uint32_t trip[MAXR][MAXR];
    for (int c = 0; c < n; ++c) for (int a = 0; a < n; ++a) trip[c][a] = 0;
    for (int k = 1; k < qt; ++k) {
        const int n2 = queue[k], n1 = parent_of[n2];
        uint32_t nb = adj[n1] & ~(1u << n2);
        while (nb) { const int n3 = __ffs(nb) - 1; nb &= nb - 1; trip[n1][min(n2, n3)] |= 1u << max(n2, n3); }
        nb = adj[n2] & ~(1u << n1);
        while (nb) { const int n3 = __ffs(nb) - 1; nb &= nb - 1; trip[n2][min(n1, n3)] |= 1u << max(n1, n3); }
    }
    bool a3_ok = true, a4_ok = true, missing = false;
    for (int c = 0; c < n; ++c)
        for (int a = 0; a < n; ++a) {
            uint32_t bits = trip[c][a];
            while (bits) {
                const int bb = __ffs(bits) - 1; bits &= bits - 1;
                const float ang = angel3_dev(p[a], p[c], p[bb]);
                const int cnt = t.a3_cnt[rt[c]];
                if (cnt < 0) missing = true;
                bool in_any = false;
                for (int q = 0; q < cnt; ++q) in_any = in_any || (t.a3_lo[rt[c] * MAXRANGE + q] <= ang && ang <= t.a3_hi[rt[c] * MAXRANGE + q]);
                a3_ok = a3_ok && in_any;
                if (!t.check_a4) continue;
                if (170.f < ang && ang < 190.f) continue;               // only angular triplets seed quads (analyze.py:297-311)
                uint32_t nb = adj[a] & ~((1u << c) | (1u << bb));
                while (nb) {
                    const int n4 = __ffs(nb) - 1; nb &= nb - 1;
                    const float lin = angel3_dev(p[n4], p[a], p[c]);
                    if (175.f < lin && lin < 185.f) continue;
                    // quad (n4, a, c, bb), canonical orientation: first < last
                    const float d4 = n4 < bb ? angel4_dev(p[n4], p[a], p[c], p[bb]) : angel4_dev(p[bb], p[c], p[a], p[n4]);
                    a4_ok = a4_ok && (t.a4_hi <= d4 || d4 <= t.a4_lo);
                }
                nb = adj[bb] & ~((1u << a) | (1u << c));
                while (nb) {
                    const int n4 = __ffs(nb) - 1; nb &= nb - 1;
                    const float lin = angel3_dev(p[c], p[bb], p[n4]);
                    if (175.f < lin && lin < 185.f) continue;
                    const float d4 = a < n4 ? angel4_dev(p[a], p[c], p[bb], p[n4]) : angel4_dev(p[n4], p[bb], p[c], p[a]);
                    a4_ok = a4_ok && (t.a4_hi <= d4 || d4 <= t.a4_lo);
                }
            }
        }
    if (missing) out[6] |= 4;
    out[3] = a3_ok;
    out[4] = a4_ok;
    out[5] = out[0] && out[1] && out[2] && out[3] && out[4];
}

// positions2adj on the full padded tensors (helpers.py:172-196): dist [B,N,N], adj [B,N,N] fp32 0/1
__global__ void positions2adj_kernel(const float* __restrict__ x, const int* __restrict__ ring_type, int B, int N, ValidityTables t,
                                     float* __restrict__ dist, float* __restrict__ adj) {
    const size_t total = (size_t)B * N * N;
    for (size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
        const int j = (int)(e % N), i = (int)((e / N) % N), b = (int)(e / ((size_t)N * N));
        const float* pi = x + ((size_t)b * N + i) * 3;
        const float* pj = x + ((size_t)b * N + j) * 3;
        const float dx = pi[0] - pj[0], dy = pi[1] - pj[1], dz = pi[2] - pj[2];
        const float d = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        dist[e] = d;
        float a = 0.f;
        if (i != j) {
            const int lo = min(i, j), hi = max(i, j);            // the reference looks the pair up as (type_i, type_j) with i < j
            const int k = ring_type[(size_t)b * N + lo] * t.n_types + ring_type[(size_t)b * N + hi];
            if (t.pair_lo[k] < d && d < t.pair_hi[k]) a = 1.f;
        }
        adj[e] = a;
    }
}

}  // namespace gb

using namespace gb;
extern int gb_train_fail(const char* what);
extern void gb_train_launched(int n);

static ValidityTables make_tables(const float* pair_lo, const float* pair_hi, const float* a3_lo, const float* a3_hi, const int* a3_cnt,
                                  float min_dist, float a4_hi, float a4_lo, int check_a4, int n_types, int orientation_type) {
    ValidityTables t;
    t.pair_lo = pair_lo; t.pair_hi = pair_hi; t.a3_lo = a3_lo; t.a3_hi = a3_hi; t.a3_cnt = a3_cnt;
    t.min_dist = min_dist; t.a4_hi = a4_hi; t.a4_lo = a4_lo; t.check_a4 = check_a4; t.n_types = n_types;
    t.orientation_type = orientation_type;
    return t;
}

extern "C" int gb_check_stability(const float* x, const int* ring_type, const float* node_mask, int B, int N, int n_types,
                                  int orientation_type, const float* pair_lo, const float* pair_hi, float min_dist, const float* a3_lo,
                                  const float* a3_hi, const int* a3_cnt, float a4_hi, float a4_lo, int check_a4, unsigned char* flags,
                                  void* stream) {
    if (B <= 0) return 0;
    ValidityTables t = make_tables(pair_lo, pair_hi, a3_lo, a3_hi, a3_cnt, min_dist, a4_hi, a4_lo, check_a4, n_types, orientation_type);
    stability_kernel<<<(B + 127) / 128, 128, 0, (cudaStream_t)stream>>>(x, ring_type, node_mask, B, N, t, flags);
    if (cudaPeekAtLastError() != cudaSuccess) { cudaGetLastError(); return gb_train_fail("check_stability"); }
    gb_train_launched(1);
    return 0;
}

extern "C" int gb_positions2adj(const float* x, const int* ring_type, int B, int N, int n_types, const float* pair_lo, const float* pair_hi,
                                float* dist, float* adj, void* stream) {
    if (B <= 0 || N <= 0) return 0;
    ValidityTables t = make_tables(pair_lo, pair_hi, nullptr, nullptr, nullptr, 0.f, 0.f, 0.f, 0, n_types, -1);
    const size_t total = (size_t)B * N * N;
    const int blocks = (int)((total + 255) / 256 > 148 * 16 ? 148 * 16 : (total + 255) / 256);
    positions2adj_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, ring_type, B, N, t, dist, adj);
    if (cudaPeekAtLastError() != cudaSuccess) { cudaGetLastError(); return gb_train_fail("positions2adj"); }
    gb_train_launched(1);
    return 0;
}
