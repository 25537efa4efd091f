// C-ABI of gaudi_b200 (see include/gaudi_b200.h): weight packing, workspace carving and the launch sequences
// of the denoiser forward, the predictor forward / input gradient and the reverse-diffusion loop.
#include "../../include/gaudi_b200.h"
#include "kernels.h"

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include <string>
#include <vector>

using namespace gb;

// ------------------------------------------------------------------------------------------------------
// error handling / launch counting
// ------------------------------------------------------------------------------------------------------
static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};   // process-wide: autograd runs backward ops on its own thread

static int fail(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}
#define GB_CUDA(x)                                                                      \
    do {                                                                                \
        cudaError_t e_ = (x);                                                           \
        if (e_ != cudaSuccess) return fail("%s:%d %s: %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); \
    } while (0)
#define GB_LAUNCHED(n) (g_launches += (n))
// train_ops.cu reports launch failures through the same thread-local error string and counts its launches here
int gb_train_fail(const char* what) { return fail("%s: launch failed", what); }
void gb_train_launched(int n) { g_launches += n; }
static int check_launch(const char* what) {
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) { cudaGetLastError(); return fail("%s: %s", what, cudaGetErrorString(e)); }
    return 0;
}

extern "C" int gb_abi_version(void) { return 1; }
extern "C" const char* gb_last_error(void) { return g_err.c_str(); }
extern "C" long long gb_launch_count(int reset) {
    long long v = g_launches;
    if (reset) g_launches.store(0);
    return v;
}

// ------------------------------------------------------------------------------------------------------
// handles
// ------------------------------------------------------------------------------------------------------
struct EdgeMlpW {           // first Linear factorised per node (2 column blocks), second Linear packed^T
    size_t l1_wt, l1_b, l1_ext, l2_wt, l2_b;
    size_t l1_nt, l2_nt;    // un-transposed copies for the input-gradient pass (predictor only)
    size_t l1_tc, l2_tc, l1_nt_tc, l2_nt_tc;   // tensor-core images of the same blocks
};
struct NodeMlpW { size_t l1_wt, l1_b, l2_wt, l2_b, l1_nt, l2_nt, l1_tc, l2_tc, l1_nt_tc, l2_nt_tc; };
struct DenGcl { EdgeMlpW e; NodeMlpW n; size_t att_w; float att_b; };
struct DenEquiv { EdgeMlpW c; size_t last_w; };
struct PredLayer { EdgeMlpW e; NodeMlpW n; size_t att_w; float att_b; size_t c_wt, c_b, c_nt, c_last, c_tc, c_nt_tc; };

struct gb_net {
    int kind;               // 0 denoiser, 1 predictor
    int F, H, HP, L, n_sub, out_nf, attention, use_tanh;
    float coords_range, norm_constant, normf;
    int tc_lin = 0, tc_den = 0, tc_pred = 0;   // which kernel families run on tcgen05 (GAUDI_B200_GEMM)
    int lin_fmt = 2;        // image format of the node-Linear weights: 2 = hi + bf16 mix (2 MMAs per K step), 0 = hi | lo (3xTF32; GAUDI_B200_LIN_MIX=0)
    float* buf = nullptr;
    size_t n_floats = 0;
    size_t emb_w, emb_b, out_w, out_b;      // raw (unpacked) small heads
    std::vector<DenGcl> gcl;                // L * n_sub
    std::vector<DenEquiv> eq;               // L
    std::vector<PredLayer> pl;              // L
    const float* p(size_t off) const { return buf + off; }
};

struct gb_graph { Graph g; };

static int pad_hidden(int H) {
    if (H <= 64) return 64;
    if (H <= 192) return 192;
    if (H <= 196) return 196;
    if (H <= 256) return 256;
    return -1;
}

namespace {
struct Packer {
    gb_net* net; cudaStream_t s; size_t cursor = 0; bool dry = true;
    size_t take(size_t n) { size_t o = cursor; cursor += (n + 63) & ~size_t(63); return o; }
    // Linear weight w [out][ld] -> block [Kp][HP] at dst; transpose: k = input column (k_off..), n = output row (n_off..)
    void pack_at(size_t dst, const float* w, int ld, int k_off, int n_off, int Kv, int Nv, int Kp, int transpose) {
        if (!dry) { launch_pack(net->buf + dst, w, ld, k_off, n_off, Kv, Nv, Kp, net->HP, transpose, s); GB_LAUNCHED(1); }
    }
    size_t block(const float* w, int ld, int k_off, int n_off, int Kv, int Nv, int Kp, int transpose) {
        size_t o = take((size_t)Kp * net->HP);
        pack_at(o, w, ld, k_off, n_off, Kv, Nv, Kp, transpose);
        return o;
    }
    // two blocks stored back to back (column blocks of one Linear, or a K-concatenated weight)
    size_t block2(const float* w, int ld, int k0, int n0, int k1, int n1, int Kv, int Nv, int Kp, int transpose) {
        const size_t sz = (size_t)Kp * net->HP;
        size_t o = take(2 * sz);
        pack_at(o, w, ld, k0, n0, Kv, Nv, Kp, transpose);
        pack_at(o + sz, w, ld, k1, n1, Kv, Nv, Kp, transpose);
        return o;
    }
    // tensor-core image of ONE [Nv x Kv] block: atoms(Kv) x [hi|lo][NP][32]
    size_t tc_size(int K) const { return (size_t)((K + 31) / 32) * 2 * tc_np(net->HP) * 32; }
    void tc_at(size_t dst, const float* w, int ld, int k_off, int n_off, int Kv, int Nv, int Kp, int transpose, int fmt = 0) {
        if (!dry) { launch_pack_tc(net->buf + dst, w, ld, k_off, n_off, Kv, Nv, tc_np(net->HP), (Kp + 31) / 32, transpose, s, fmt); GB_LAUNCHED(1); }
    }
    // transpose==1: forward use (k = input column, n = output row); 0: dgrad use (k = output row, n = input column)
    // fmt: image format (launch_pack_tc): 0 for the node-Linear kernel, 1 (fp16 mix) / 2 (bf16 mix) for the edge kernels
    size_t tc_block(const float* w, int ld, int k_off, int n_off, int Kv, int Nv, int transpose, int fmt = 0) {
        size_t o = take(tc_size(net->HP));
        tc_at(o, w, ld, transpose ? k_off : k_off, n_off, Kv, Nv, net->HP, transpose ? 0 : 1, fmt);
        return o;
    }
    size_t tc_block2(const float* w, int ld, int k0, int n0, int k1, int n1, int Kv, int Nv, int transpose) {
        const size_t sz = tc_size(net->HP);
        size_t o = take(2 * sz);
        tc_at(o, w, ld, k0, n0, Kv, Nv, net->HP, transpose ? 0 : 1, net->lin_fmt);
        tc_at(o + sz, w, ld, k1, n1, Kv, Nv, net->HP, transpose ? 0 : 1, net->lin_fmt);
        return o;
    }
    size_t vec(const float* v, int n, int np) {      // zero-padded copy of a vector
        size_t o = take(np);
        if (!dry) {
            cudaMemsetAsync(net->buf + o, 0, (size_t)np * 4, s);
            if (v && n > 0) cudaMemcpyAsync(net->buf + o, v, (size_t)n * 4, cudaMemcpyDeviceToDevice, s);
        }
        return o;
    }
    // first edge Linear  W [H][2H+2] (+bias): P blocks (W_a^T with bias folded | W_b^T), ext rows (2), nt copies
    void edge_l1(EdgeMlpW& e, const float* w, const float* b, bool want_nt) {
        const int H = net->H, HP = net->HP, ld = 2 * H + 2;
        e.l1_wt = block2(w, ld, 0, 0, H, 0, H, H, HP, 1);      // column blocks W_a^T | W_b^T
        e.l1_tc = tc_block2(w, ld, 0, 0, H, 0, H, H, 1);
        e.l1_b = vec(b, H, 2 * HP);                            // (b | 0)
        e.l1_ext = take(2 * HP);
        if (!dry) {
            cudaMemsetAsync(net->buf + e.l1_ext, 0, (size_t)2 * HP * 4, s);
            // columns 2H and 2H+1 of W, strided gather
            cudaMemcpy2DAsync(net->buf + e.l1_ext, 4, w + 2 * H, (size_t)ld * 4, 4, H, cudaMemcpyDeviceToDevice, s);
            cudaMemcpy2DAsync(net->buf + e.l1_ext + HP, 4, w + 2 * H + 1, (size_t)ld * 4, 4, H, cudaMemcpyDeviceToDevice, s);
        }
        e.l1_nt = 0;
        if (want_nt) {                                          // [2HP][HP]: rows = output index, W_a then W_b
            e.l1_nt = block2(w, ld, 0, 0, 0, H, H, H, HP, 0);
            e.l1_nt_tc = tc_block2(w, ld, 0, 0, 0, H, H, H, 0);
        }
    }
    // edge = true: the image feeds an edge kernel (TF32 + bf16 correction terms, tc_common.cuh)
    void square(size_t& wt, size_t& bias, size_t* nt, const float* w, const float* b, size_t* tcw = nullptr, size_t* tcnt = nullptr, bool edge = false) {
        const int H = net->H, HP = net->HP;
        wt = block(w, H, 0, 0, H, H, HP, 1);
        bias = vec(b, H, HP);
        if (nt) *nt = block(w, H, 0, 0, H, H, HP, 0);
        if (tcw) *tcw = tc_block(w, H, 0, 0, H, H, 1, edge ? 2 : net->lin_fmt);
        if (tcnt) *tcnt = tc_block(w, H, 0, 0, H, H, 0, edge ? 2 : net->lin_fmt);
    }
    void node_mlp(NodeMlpW& n, const float* w1, const float* b1, const float* w2, const float* b2, bool want_nt) {
        const int H = net->H, HP = net->HP;
        n.l1_wt = block2(w1, 2 * H, 0, 0, H, 0, H, H, HP, 1);   // rows k<HP: h part, rows HP..2HP: agg part
        n.l1_tc = tc_block2(w1, 2 * H, 0, 0, H, 0, H, H, 1);
        n.l1_b = vec(b1, H, HP);
        n.l1_nt = 0; n.l2_nt = 0;
        if (want_nt) {                                           // two column blocks [HP][HP]: W[:, :H], W[:, H:]
            n.l1_nt = block2(w1, 2 * H, 0, 0, 0, H, H, H, HP, 0);
            n.l1_nt_tc = tc_block2(w1, 2 * H, 0, 0, 0, H, H, H, 0);
        }
        square(n.l2_wt, n.l2_b, want_nt ? &n.l2_nt : nullptr, w2, b2, &n.l2_tc, want_nt ? &n.l2_nt_tc : nullptr);
    }
};

int read_scalar(const float* dev, float* out, cudaStream_t s) {
    GB_CUDA(cudaMemcpyAsync(out, dev, 4, cudaMemcpyDeviceToHost, s));
    GB_CUDA(cudaStreamSynchronize(s));
    return 0;
}
}  // namespace

static void parse_gemm_mode(gb_net* net) {
    // GAUDI_B200_GEMM = "tc" (default: every converted kernel family on tcgen05) | "fp32" | comma list of lin,den,pred
    const char* e = getenv("GAUDI_B200_GEMM");
    std::string m = e ? e : "tc";
    const bool all = (m == "tc" || m == "all");
    net->tc_lin = all || m.find("lin") != std::string::npos;
    net->tc_den = all || m.find("den") != std::string::npos;
    net->tc_pred = all || m.find("pred") != std::string::npos;
    if (m == "fp32") net->tc_lin = net->tc_den = net->tc_pred = 0;
    const char* lm = getenv("GAUDI_B200_LIN_MIX");
    net->lin_fmt = (lm && lm[0] == '0') ? 0 : 2;
}

static void run_lin(const gb_net* n, LinArgs& a, cudaStream_t s) {
    a.mix = n->lin_fmt == 2;
    if (n->tc_lin && a.wt_tc) launch_lin_tc(n->HP, a, a.wt_tc, s);
    else launch_lin(n->HP, a, s);
}

static int build_net(gb_net* net, const float* const* P, int n_params, cudaStream_t s) {
    parse_gemm_mode(net);
    const int H = net->H;
    for (int pass = 0; pass < 2; ++pass) {
        Packer pk{net, s};
        pk.dry = (pass == 0);
        int i = 0;
        const int Fin = net->F + 1;
        net->emb_w = pk.vec(P[i], H * Fin, H * Fin); ++i;
        net->emb_b = pk.vec(P[i], H, net->HP); ++i;
        const int n_out = net->kind == 0 ? Fin : net->out_nf;
        net->out_w = pk.vec(P[i], n_out * H, n_out * H); ++i;
        net->out_b = pk.vec(P[i], n_out, n_out); ++i;
        if (net->kind == 0) {
            net->gcl.resize((size_t)net->L * net->n_sub);
            net->eq.resize(net->L);
            for (int b = 0; b < net->L; ++b) {
                for (int q = 0; q < net->n_sub; ++q) {
                    DenGcl& G = net->gcl[(size_t)b * net->n_sub + q];
                    pk.edge_l1(G.e, P[i], P[i + 1], false);
                    pk.square(G.e.l2_wt, G.e.l2_b, nullptr, P[i + 2], P[i + 3], &G.e.l2_tc, nullptr, true);
                    pk.node_mlp(G.n, P[i + 4], P[i + 5], P[i + 6], P[i + 7], false);
                    i += 8;
                    if (net->attention) {
                        G.att_w = pk.vec(P[i], H, net->HP);
                        if (!pk.dry && read_scalar(P[i + 1], &G.att_b, s)) return 1;
                        i += 2;
                    } else { G.att_w = pk.vec(nullptr, 0, net->HP); G.att_b = 0.f; }
                }
                DenEquiv& E = net->eq[b];
                pk.edge_l1(E.c, P[i], P[i + 1], false);
                pk.square(E.c.l2_wt, E.c.l2_b, nullptr, P[i + 2], P[i + 3], &E.c.l2_tc, nullptr, true);
                E.last_w = pk.vec(P[i + 4], H, net->HP);
                i += 5;
            }
        } else {
            net->pl.resize(net->L);
            for (int l = 0; l < net->L; ++l) {
                PredLayer& Lr = net->pl[l];
                pk.edge_l1(Lr.e, P[i], P[i + 1], true);
                pk.square(Lr.e.l2_wt, Lr.e.l2_b, &Lr.e.l2_nt, P[i + 2], P[i + 3], &Lr.e.l2_tc, &Lr.e.l2_nt_tc, true);
                pk.node_mlp(Lr.n, P[i + 4], P[i + 5], P[i + 6], P[i + 7], true);
                pk.square(Lr.c_wt, Lr.c_b, &Lr.c_nt, P[i + 8], P[i + 9], &Lr.c_tc, &Lr.c_nt_tc, true);
                Lr.c_last = pk.vec(P[i + 10], H, net->HP);
                i += 11;
                if (net->attention) {
                    Lr.att_w = pk.vec(P[i], H, net->HP);
                    if (!pk.dry && read_scalar(P[i + 1], &Lr.att_b, s)) return 1;
                    i += 2;
                } else { Lr.att_w = pk.vec(nullptr, 0, net->HP); Lr.att_b = 0.f; }
            }
        }
        if (i != n_params) return fail("expected %d parameter pointers, got %d", i, n_params);
        if (pass == 0) {
            net->n_floats = pk.cursor;
            GB_CUDA(cudaMalloc(&net->buf, net->n_floats * sizeof(float)));
        }
    }
    GB_CUDA(cudaStreamSynchronize(s));
    return check_launch("pack");
}

extern "C" int gb_denoiser_create(gb_net** out, int in_node_nf, int hidden_nf, int n_layers, int inv_sublayers,
                                  int attention, int use_tanh, float coords_range, float norm_constant,
                                  float normalization_factor, const float* const* params, int n_params, void* stream) {
    if (!out || !params) return fail("null argument");
    const int HP = pad_hidden(hidden_nf);
    if (HP < 0) return fail("hidden_nf %d unsupported (max 256)", hidden_nf);
    if (in_node_nf + 1 > 16) return fail("in_node_nf %d too large", in_node_nf);
    gb_net* net = new gb_net();
    net->kind = 0; net->F = in_node_nf; net->H = hidden_nf; net->HP = HP; net->L = n_layers; net->n_sub = inv_sublayers;
    net->out_nf = in_node_nf + 1; net->attention = attention; net->use_tanh = use_tanh;
    net->coords_range = coords_range; net->norm_constant = norm_constant; net->normf = normalization_factor;
    if (build_net(net, params, n_params, (cudaStream_t)stream)) { gb_net_destroy(net); return 1; }
    *out = net;
    return 0;
}

extern "C" int gb_predictor_create(gb_net** out, int in_node_nf, int out_nf, int hidden_nf, int n_layers, int attention,
                                   int use_tanh, float coords_range, const float* const* params, int n_params,
                                   void* stream) {
    if (!out || !params) return fail("null argument");
    const int HP = pad_hidden(hidden_nf);
    if (HP < 0) return fail("hidden_nf %d unsupported (max 256)", hidden_nf);
    if (in_node_nf + 1 > 16 || out_nf > 16) return fail("in_node_nf/out_nf too large");
    gb_net* net = new gb_net();
    net->kind = 1; net->F = in_node_nf; net->H = hidden_nf; net->HP = HP; net->L = n_layers; net->n_sub = 1;
    net->out_nf = out_nf; net->attention = attention; net->use_tanh = use_tanh;
    net->coords_range = coords_range / (float)n_layers;        // edm/egnn_predictor/models.py:515
    net->norm_constant = 1.f; net->normf = 1.f;
    if (build_net(net, params, n_params, (cudaStream_t)stream)) { gb_net_destroy(net); return 1; }
    *out = net;
    return 0;
}

extern "C" int gb_net_destroy(gb_net* net) {
    if (net) { if (net->buf) cudaFree(net->buf); delete net; }
    return 0;
}
extern "C" int gb_net_hidden_padded(const gb_net* net) { return net ? net->HP : -1; }

// Greedy packing of consecutive nodes into tiles of <=128 edges and <=128 nodes; a node's edge segment is
// never split (a node with more than 128 edges is an error).
extern "C" int gb_tile_pack(const int32_t* rowptr, int n_nodes, int32_t* tile_ptr, int* n_tiles_out) {
    int nt = 0, start = 0;
    if (tile_ptr) tile_ptr[0] = 0;
    while (start < n_nodes) {
        int end = start;
        while (end < n_nodes && (end - start) < GB_TM_HOST && rowptr[end + 1] - rowptr[start] <= GB_TM_HOST) ++end;
        if (end == start) return fail("node %d has %d edges (> %d per tile)", start, rowptr[start + 1] - rowptr[start], GB_TM_HOST);
        ++nt;
        if (tile_ptr) tile_ptr[nt] = end;
        start = end;
    }
    *n_tiles_out = nt;
    return 0;
}

extern "C" int gb_stage_rows(void) { return GB_PS_ROWS_HOST; }
extern "C" int gb_tile_pack_graphs(const int32_t* rowptr, int n_nodes, int N, int32_t* tile_ptr, int* n_tiles_out) {
    if (N <= 0) return gb_tile_pack(rowptr, n_nodes, tile_ptr, n_tiles_out);
    int nt = 0, start = 0;
    if (tile_ptr) tile_ptr[0] = 0;
    while (start < n_nodes) {
        int end = start;
        while (end < n_nodes && (end - start) < GB_TM_HOST && rowptr[end + 1] - rowptr[start] <= GB_TM_HOST &&
               (end + 1 - start) + (end / N - start / N + 1) * N <= GB_PS_ROWS_HOST) ++end;
        if (end == start) {
            if (rowptr[start + 1] - rowptr[start] > GB_TM_HOST)
                return fail("node %d has %d edges (> %d per tile)", start, rowptr[start + 1] - rowptr[start], GB_TM_HOST);
            return fail("graphs of %d padded nodes do not fit the %d staged rows of the edge kernels", N, GB_PS_ROWS_HOST);
        }
        ++nt;
        if (tile_ptr) tile_ptr[nt] = end;
        start = end;
    }
    *n_tiles_out = nt;
    return 0;
}

__global__ void tile_info_kernel(const int* __restrict__ tile_ptr, const int* __restrict__ rowptr, int n_tiles, int N, int4* __restrict__ out,
                                 int* __restrict__ bad) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_tiles) return;
    const int lo = tile_ptr[t], hi = tile_ptr[t + 1];
    const int e_lo = rowptr[lo];
    out[t] = make_int4(lo, hi - lo, e_lo, rowptr[hi] - e_lo);
    // contract of the edge kernels: <= 128 edges / nodes per tile, and the staged node-projection rows fit (gb_tile_pack_graphs)
    const int rows = (hi - lo) + ((hi - 1) / N - lo / N + 1) * N;
    if (hi <= lo || hi - lo > GB_TM_HOST || rowptr[hi] - e_lo > GB_TM_HOST || rows > GB_PS_ROWS_HOST) atomicOr(bad, 1);
    atomicMax(bad + 1, rows);
}

extern "C" int gb_graph_create(gb_graph** out, int B, int N, int n_edges, int n_tiles, int n_tc, const int32_t* rowptr,
                               const int32_t* erow, const int32_t* ecol, const int32_t* tile_ptr, const int32_t* tc_ptr,
                               const int32_t* tc_node, const int32_t* tc_start, const int32_t* cperm,
                               const int32_t* colptr, const int32_t* cedge, const float* node_mask) {
    (void)n_tc;
    if (!out) return fail("null argument");
    int ps_rows = 0;
    int4* tinfo = nullptr;                       // the only memory a graph owns: 16 bytes per tile, derived from the caller's arrays
    if (n_tiles > 0) {
        if (N <= 0) return fail("gb_graph_create: N must be positive");
        GB_CUDA(cudaMalloc(&tinfo, ((size_t)n_tiles + 1) * sizeof(int4)));
        int* bad = reinterpret_cast<int*>(tinfo + n_tiles);
        GB_CUDA(cudaMemset(bad, 0, 2 * sizeof(int)));
        tile_info_kernel<<<(n_tiles + 255) / 256, 256>>>(tile_ptr, rowptr, n_tiles, N, tinfo, bad);
        GB_CUDA(cudaDeviceSynchronize());        // not a hot-path call: graphs are created once per mask pair
        int bad_h[2] = {0, 0};
        GB_CUDA(cudaMemcpy(bad_h, bad, 2 * sizeof(int), cudaMemcpyDeviceToHost));
        ps_rows = bad_h[1];
        if (bad_h[0]) { cudaFree(tinfo); return fail("gb_graph_create: tile_ptr violates the tile contract (pack it with gb_tile_pack_graphs)"); }
    }
    gb_graph* g = new gb_graph();
    g->g = Graph{B * N, n_edges, n_tiles, B, N, rowptr, erow, ecol, tile_ptr, tc_ptr, tc_node, tc_start, cperm, colptr, cedge, node_mask, tinfo, ps_rows};
    *out = g;
    return 0;
}
extern "C" int gb_graph_destroy(gb_graph* g) {
    if (g) { if (g->g.tile_info) cudaFree((void*)g->g.tile_info); delete g; }
    return 0;
}

// ------------------------------------------------------------------------------------------------------
// workspace carving
// ------------------------------------------------------------------------------------------------------
namespace {
struct Bump {
    char* base; size_t off = 0; size_t cap;
    Bump(void* b, size_t c) : base((char*)b), cap(c) {}
    template <typename T> T* get(size_t n) {
        size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
        T* p = base ? (T*)(base + off) : nullptr;
        off += bytes;
        return p;
    }
};

struct DenWs { float *h, *h2, *s, *agg, *P, *x0, *xa, *xb, *hout; };
void carve_den(Bump& b, const gb_net* n, const Graph& g, DenWs& w) {
    const size_t nn = g.n_nodes, HP = n->HP;
    w.h = b.get<float>(nn * HP); w.h2 = b.get<float>(nn * HP); w.s = b.get<float>(nn * HP);
    w.agg = b.get<float>(nn * HP); w.P = b.get<float>(nn * 2 * HP);
    w.x0 = b.get<float>(nn * 3); w.xa = b.get<float>(nn * 3); w.xb = b.get<float>(nn * 3);
    w.hout = b.get<float>(nn * (n->F + 1));
}

struct PredWs {
    float *h, *h2, *s, *agg, *P, *hout;
    float* x;        // [(L+1)][nn][3]
    float* pre4;     // [L][nn][HP]
    float *sv_d1, *sv_pre2, *sv_d3, *sv_tau;   // per layer strides below
    size_t sv_stride, sv_stride_p, tau_stride;
    float *gh, *gh2, *gcat, *gpre4, *gPa, *gPb, *gx, *gx2, *gattr, *gpre1, *gd;
};
void carve_pred(Bump& b, const gb_net* n, const Graph& g, bool grad, PredWs& w) {
    const size_t nn = g.n_nodes, HP = n->HP, L = n->L;
    w.h = b.get<float>(nn * HP); w.h2 = b.get<float>(nn * HP); w.s = b.get<float>(nn * HP);
    w.agg = b.get<float>(nn * HP); w.P = b.get<float>(nn * 2 * HP); w.hout = b.get<float>(nn * n->out_nf);
    w.x = b.get<float>((L + 1) * nn * 3);
    // saved activations per layer and tensor, in floats: the tensor-core kernels keep the two SiLU derivatives as 16-bit codes in
    // planes of 8 columns ([tile][(H + 7) / 8][128 rows][8 codes], tc_common.cuh), the FP32 engine as fp32 [tile][HP][128]
    w.sv_stride = n->tc_pred ? (size_t)g.n_tiles * ((n->H + 7) / 8) * (GB_TM_HOST * 16 / 4) : (size_t)g.n_tiles * HP * GB_TM_HOST;
    w.sv_stride_p = (size_t)g.n_tiles * HP * GB_TM_HOST;        // pre2: fp32 in both engines
    w.tau_stride = (size_t)g.n_edges;
    if (grad) {
        w.pre4 = b.get<float>(L * nn * HP);
        w.sv_d1 = b.get<float>(L * w.sv_stride); w.sv_pre2 = b.get<float>(L * w.sv_stride_p);
        w.sv_d3 = b.get<float>(L * w.sv_stride); w.sv_tau = b.get<float>(L * w.tau_stride + 1);
        w.gh = b.get<float>(nn * HP); w.gh2 = b.get<float>(nn * HP); w.gcat = b.get<float>(nn * 2 * HP);
        w.gpre4 = b.get<float>(nn * HP); w.gPa = b.get<float>(nn * HP); w.gPb = b.get<float>(nn * HP);
        w.gx = b.get<float>(nn * 3); w.gx2 = b.get<float>(nn * 3); w.gattr = b.get<float>(g.n_edges + 1);
        w.gpre1 = n->tc_pred ? b.get<float>((size_t)g.n_edges * HP + 4) : nullptr;
        w.gd = n->tc_pred ? b.get<float>((size_t)g.n_edges * 3 + 4) : nullptr;
    } else {
        w.pre4 = w.sv_d1 = w.sv_pre2 = w.sv_d3 = w.sv_tau = nullptr;
        w.gh = w.gh2 = w.gcat = w.gpre4 = w.gPa = w.gPb = w.gx = w.gx2 = w.gattr = w.gpre1 = w.gd = nullptr;
    }
}
}  // namespace

extern "C" size_t gb_denoiser_workspace_bytes(const gb_net* net, const gb_graph* g) {
    Bump b(nullptr, 0); DenWs w; carve_den(b, net, g->g, w); return b.off;
}
extern "C" size_t gb_predictor_workspace_bytes(const gb_net* net, const gb_graph* g, int with_grad) {
    Bump b(nullptr, 0); PredWs w; carve_pred(b, net, g->g, with_grad != 0, w); return b.off;
}

// ------------------------------------------------------------------------------------------------------
// launch sequences
// ------------------------------------------------------------------------------------------------------
static LinArgs lin_base(int M) {
    LinArgs a;
    memset(&a, 0, sizeof(a));
    a.M = M; a.ncb = 1; a.epi = EPI_BIAS; a.res_cb = -1;
    return a;
}

// P = h @ [W_a^T | W_b^T] + (b | 0)
static void lin_P(const gb_net* n, const EdgeMlpW& e, const float* h, float* P, int M, cudaStream_t s) {
    LinArgs a = lin_base(M);
    a.A1 = h; a.lda1 = n->HP; a.K1 = n->HP; a.wt = n->p(e.l1_wt); a.wt_tc = n->p(e.l1_tc); a.bias = n->p(e.l1_b);
    a.out = P; a.ldo = 2 * n->HP; a.ncb = 2;
    run_lin(n, a, s); GB_LAUNCHED(1);
}
// h_new = (h + W2 SiLU(W1 [h, agg] + b1) + b2) * mask ; optionally keeps the pre-activation
static void node_update(const gb_net* n, const NodeMlpW& w, const float* h, const float* agg, float* hid, float* h_new,
                        float* pre_save, const Graph& g, cudaStream_t s) {
    LinArgs a = lin_base(g.n_nodes);
    a.A1 = h; a.lda1 = n->HP; a.K1 = n->HP; a.A2 = agg; a.lda2 = n->HP; a.K2 = n->HP;
    a.wt = n->p(w.l1_wt); a.wt_tc = n->p(w.l1_tc); a.bias = n->p(w.l1_b); a.out = hid; a.ldo = n->HP; a.epi = EPI_SILU;
    a.out2 = pre_save; a.ldo2 = n->HP;
    run_lin(n, a, s);
    LinArgs c = lin_base(g.n_nodes);
    c.A1 = hid; c.lda1 = n->HP; c.K1 = n->HP; c.wt = n->p(w.l2_wt); c.wt_tc = n->p(w.l2_tc); c.bias = n->p(w.l2_b);
    c.out = h_new; c.ldo = n->HP; c.epi = EPI_RES_MASK; c.res = h; c.ldr = n->HP; c.mask = g.node_mask;
    run_lin(n, c, s);
    GB_LAUNCHED(2);
}

static void embed_in(const gb_net* n, const Graph& g, const float* z, const float* t, int t_per_mol, float* h, float* x,
                     cudaStream_t s) {
    EmbedInArgs e{z, 3 + n->F, t, t_per_mol, n->p(n->emb_w), n->p(n->emb_b), g.node_mask, g.n_nodes, g.N, n->H, n->HP, h, x};
    launch_embed_in(e, s); GB_LAUNCHED(1);
}

static int denoiser_forward_impl(const gb_net* n, const Graph& g, const float* z, const float* t, int t_per_mol, float* eps,
                                 int scrub_all, float* stats, const long long* stats_step, void* ws, size_t ws_bytes,
                                 cudaStream_t s) {
    if (n->kind != 0) return fail("not a denoiser handle");
    Bump b(ws, ws_bytes); DenWs w; carve_den(b, n, g, w);
    if (b.off > ws_bytes) return fail("denoiser workspace too small: need %zu bytes, got %zu", b.off, ws_bytes);
    embed_in(n, g, z, t, t_per_mol, w.h, w.x0, s);
    float *h = w.h, *h2 = w.h2; const float* xc = w.x0; float* xn = w.xa;
    lin_P(n, n->gcl[0].e, h, w.P, g.n_nodes, s);
    for (int blk = 0; blk < n->L; ++blk) {
        for (int q = 0; q < n->n_sub; ++q) {
            const DenGcl& G = n->gcl[(size_t)blk * n->n_sub + q];
            DenEdgeArgs a;
            memset(&a, 0, sizeof(a));
            a.g = g; a.P = w.P; a.ext = n->p(G.e.l1_ext); a.wt2 = n->p(G.e.l2_wt); a.b2 = n->p(G.e.l2_b);
            a.vecw = n->p(G.att_w); a.att_b = G.att_b; a.attention = n->attention; a.use_tanh = n->use_tanh;
            a.norm_constant = n->norm_constant; a.normf = n->normf; a.coords_range = n->coords_range;
            a.x = xc; a.x0 = w.x0; a.agg = w.agg;
            if (n->tc_den) launch_den_edge_tc(n->HP, 0, a, n->p(G.e.l2_tc), s); else launch_den_edge(n->HP, 0, a, s);
            GB_LAUNCHED(1);
            node_update(n, G.n, h, w.agg, w.s, h2, nullptr, g, s);
            float* tmp = h; h = h2; h2 = tmp;
            const EdgeMlpW& next = (q + 1 < n->n_sub) ? n->gcl[(size_t)blk * n->n_sub + q + 1].e : n->eq[blk].c;
            lin_P(n, next, h, w.P, g.n_nodes, s);
        }
        const DenEquiv& E = n->eq[blk];
        DenEdgeArgs a;
        memset(&a, 0, sizeof(a));
        a.g = g; a.P = w.P; a.ext = n->p(E.c.l1_ext); a.wt2 = n->p(E.c.l2_wt); a.b2 = n->p(E.c.l2_b);
        a.vecw = n->p(E.last_w); a.attention = 0; a.use_tanh = n->use_tanh;
        a.norm_constant = n->norm_constant; a.normf = n->normf; a.coords_range = n->coords_range;
        a.x = xc; a.x0 = w.x0; a.x_out = xn;
        if (n->tc_den) launch_den_edge_tc(n->HP, 1, a, n->p(E.c.l2_tc), s); else launch_den_edge(n->HP, 1, a, s);
        GB_LAUNCHED(1);
        xc = xn; xn = (xn == w.xa) ? w.xb : w.xa;
        if (blk + 1 < n->L) lin_P(n, n->gcl[(size_t)(blk + 1) * n->n_sub].e, h, w.P, g.n_nodes, s);
    }
    EmbedOutArgs eo{h, n->HP, n->H, n->p(n->out_w), n->p(n->out_b), n->F + 1, g.node_mask, g.n_nodes, w.hout, n->F + 1};
    launch_embed_out(eo, s);
    DenFinishArgs f{xc, w.x0, w.hout, n->F + 1, g.node_mask, g.B, g.N, n->F, eps, scrub_all, stats, z, stats_step};
    launch_den_finish(f, s);
    GB_LAUNCHED(2);
    return check_launch("denoiser_forward");
}

extern "C" int gb_denoiser_forward(const gb_net* net, const gb_graph* g, const float* z, const float* t, int t_per_mol,
                                   float* eps, int scrub_all, float* stats, void* workspace, size_t workspace_bytes,
                                   void* stream) {
    if (!net || !g || !z || !t || !eps || !workspace) return fail("null argument");
    return denoiser_forward_impl(net, g->g, z, t, t_per_mol, eps, scrub_all, stats, nullptr, workspace, workspace_bytes,
                                 (cudaStream_t)stream);
}

static PredEdgeArgs pred_edge_args(const gb_net* n, const PredLayer& Lr, const Graph& g, const PredWs& w, int l, bool grad) {
    PredEdgeArgs a;
    memset(&a, 0, sizeof(a));
    const size_t nn = g.n_nodes;
    a.g = g; a.P = w.P; a.ext = n->p(Lr.e.l1_ext); a.wt2 = n->p(Lr.e.l2_wt); a.b2 = n->p(Lr.e.l2_b);
    a.att_w = n->p(Lr.att_w); a.att_b = Lr.att_b; a.attention = n->attention;
    a.wtc = n->p(Lr.c_wt); a.bc = n->p(Lr.c_b); a.wc_last = n->p(Lr.c_last);
    a.use_tanh = n->use_tanh; a.coords_range = n->coords_range;
    a.x0 = w.x;
    if (grad) {
        a.x = w.x + (size_t)l * nn * 3; a.x_out = w.x + (size_t)(l + 1) * nn * 3;
        a.sv_d1 = w.sv_d1 + l * w.sv_stride; a.sv_pre2 = w.sv_pre2 + l * w.sv_stride_p;
        a.sv_d3 = w.sv_d3 + l * w.sv_stride; a.sv_tau = w.sv_tau + l * w.tau_stride;
    } else {                                  // inference: ping-pong between slots 1 and 2, slot 0 keeps x0
        a.x = l == 0 ? w.x : w.x + (size_t)(1 + ((l - 1) & 1)) * nn * 3;
        a.x_out = w.x + (size_t)(1 + (l & 1)) * nn * 3;
    }
    a.agg = w.agg;
    return a;
}

static int predictor_forward_impl(const gb_net* n, const Graph& g, const float* z, const float* t, int t_per_mol, float* pred,
                                  int save, void* ws, size_t ws_bytes, cudaStream_t s) {
    if (n->kind != 1) return fail("not a predictor handle");
    Bump b(ws, ws_bytes); PredWs w; carve_pred(b, n, g, save != 0, w);
    if (b.off > ws_bytes) return fail("predictor workspace too small: need %zu bytes, got %zu", b.off, ws_bytes);
    const size_t nn = g.n_nodes;
    embed_in(n, g, z, t, t_per_mol, w.h, w.x, s);
    float *h = w.h, *h2 = w.h2;
    lin_P(n, n->pl[0].e, h, w.P, g.n_nodes, s);
    for (int l = 0; l < n->L; ++l) {
        const PredLayer& Lr = n->pl[l];
        PredEdgeArgs a = pred_edge_args(n, Lr, g, w, l, save != 0);
        if (n->tc_pred) launch_pred_edge_fwd_tc(n->HP, save != 0, a, n->p(Lr.e.l2_tc), n->p(Lr.c_tc), s);
        else launch_pred_edge_fwd(n->HP, save != 0, a, s);
        GB_LAUNCHED(1);
        node_update(n, Lr.n, h, w.agg, w.s, h2, save ? w.pre4 + (size_t)l * nn * n->HP : nullptr, g, s);
        float* tmp = h; h = h2; h2 = tmp;
        if (l + 1 < n->L) lin_P(n, n->pl[l + 1].e, h, w.P, g.n_nodes, s);
    }
    EmbedOutArgs eo{h, n->HP, n->H, n->p(n->out_w), n->p(n->out_b), n->out_nf, g.node_mask, g.n_nodes, w.hout, n->out_nf};
    launch_embed_out(eo, s);
    launch_pool_mean(w.hout, nullptr, g.B, g.N, n->out_nf, pred, s);
    GB_LAUNCHED(2);
    return check_launch("predictor_forward");
}

extern "C" int gb_predictor_forward(const gb_net* net, const gb_graph* g, const float* z, const float* t, int t_per_mol,
                                    float* pred, int save_for_grad, void* workspace, size_t workspace_bytes, void* stream) {
    if (!net || !g || !z || !t || !pred || !workspace) return fail("null argument");
    return predictor_forward_impl(net, g->g, z, t, t_per_mol, pred, save_for_grad, workspace, workspace_bytes, (cudaStream_t)stream);
}

__global__ void broadcast_rows_kernel(const float* v, int n, int B, float* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B * n) out[i] = v[i % n];
}

static int predictor_grad_impl(const gb_net* n, const Graph& g, const float* g_pred, int bcast, float* g_z, void* ws,
                               size_t ws_bytes, cudaStream_t s) {
    if (n->kind != 1) return fail("not a predictor handle");
    Bump b(ws, ws_bytes); PredWs w; carve_pred(b, n, g, true, w);
    if (b.off > ws_bytes) return fail("predictor workspace too small: need %zu bytes, got %zu", b.off, ws_bytes);
    const size_t nn = g.n_nodes; const int HP = n->HP;
    const float* gp = g_pred;
    if (bcast) {          // expand the shared [out] vector into gcat scratch (free until the first layer)
        broadcast_rows_kernel<<<(g.B * n->out_nf + 255) / 256, 256, 0, s>>>(g_pred, n->out_nf, g.B, w.gcat);
        gp = w.gcat; GB_LAUNCHED(1);
    }
    HeadBwdArgs hb{gp, n->p(n->out_w), n->out_nf, n->H, HP, g.node_mask, g.B, g.N, w.gh};
    launch_head_bwd(hb, s); GB_LAUNCHED(1);
    float *gh = w.gh, *gh2 = w.gh2, *gx = w.gx, *gx2 = w.gx2;
    cudaMemsetAsync(gx, 0, nn * 3 * sizeof(float), s);      // dL/dx_L = 0: the head only reads h
    cudaMemsetAsync(w.gattr, 0, (size_t)g.n_edges * sizeof(float), s);
    for (int l = n->L - 1; l >= 0; --l) {
        const PredLayer& Lr = n->pl[l];
        // g_pre4 = ((gh*mask) W4) * SiLU'(pre4)
        LinArgs a = lin_base(g.n_nodes);
        a.A1 = gh; a.lda1 = HP; a.K1 = HP; a.rowscale = g.node_mask; a.wt = n->p(Lr.n.l2_nt); a.wt_tc = n->p(Lr.n.l2_nt_tc);
        a.out = w.gpre4; a.ldo = HP; a.epi = EPI_MUL_DSILU; a.aux = w.pre4 + (size_t)l * nn * HP; a.ldaux = HP;
        run_lin(n, a, s);
        // gcat = g_pre4 W3  -> [:, :HP] (+ gh*mask) = dL/dh (direct), [:, HP:] = dL/dagg
        LinArgs c = lin_base(g.n_nodes);
        c.A1 = w.gpre4; c.lda1 = HP; c.K1 = HP; c.wt = n->p(Lr.n.l1_nt); c.wt_tc = n->p(Lr.n.l1_nt_tc); c.ncb = 2; c.out = w.gcat; c.ldo = 2 * HP;
        c.epi = EPI_ADD_RES; c.res = gh; c.ldr = HP; c.mask = g.node_mask; c.res_cb = 0;
        run_lin(n, c, s);
        PredEdgeArgs e = pred_edge_args(n, Lr, g, w, l, true);
        e.w2_nt = n->p(Lr.e.l2_nt); e.wc_nt = n->p(Lr.c_nt);
        e.g_agg = w.gcat + HP; e.ld_gagg = 2 * HP; e.g_xout = gx; e.g_Pa = w.gPa; e.g_Pb = w.gPb; e.g_x = gx2; e.g_attr = w.gattr;
        e.g_pre1 = w.gpre1; e.g_d = w.gd;
        if (n->tc_pred) {                   // tile kernel writes dL/dpre1 and dL/dd per edge; node-parallel kernel reduces them
            launch_pred_edge_bwd_tc(HP, e, n->p(Lr.c_nt_tc), n->p(Lr.e.l2_nt_tc), s);
            launch_pred_bwd_reduce(HP, e, s);
            GB_LAUNCHED(1);
        } else {
            cudaMemsetAsync(w.gPb, 0, nn * HP * sizeof(float), s);
            cudaMemsetAsync(gx2, 0, nn * 3 * sizeof(float), s);
            launch_pred_edge_bwd(HP, e, s);
        }
        // gh_l = gcat[:, :HP] + gPa W1a + gPb W1b
        LinArgs d = lin_base(g.n_nodes);
        d.A1 = w.gPa; d.lda1 = HP; d.K1 = HP; d.A2 = w.gPb; d.lda2 = HP; d.K2 = HP; d.wt = n->p(Lr.e.l1_nt); d.wt_tc = n->p(Lr.e.l1_nt_tc);
        d.out = gh2; d.ldo = HP; d.epi = EPI_ADD_RES; d.res = w.gcat; d.ldr = 2 * HP;
        run_lin(n, d, s);
        GB_LAUNCHED(4);
        float* tmp = gh; gh = gh2; gh2 = tmp;
        tmp = gx; gx = gx2; gx2 = tmp;
    }
    InBwdArgs ib{g, gh, HP, n->H, n->p(n->emb_w), n->F, gx, w.gattr, w.x, 3 + n->F, g_z};
    launch_in_bwd(ib, s); GB_LAUNCHED(1);
    return check_launch("predictor_input_grad");
}

extern "C" int gb_predictor_input_grad(const gb_net* net, const gb_graph* g, const float* g_pred, int g_pred_broadcast,
                                       float* g_z, void* workspace, size_t workspace_bytes, void* stream) {
    if (!net || !g || !g_pred || !g_z || !workspace) return fail("null argument");
    return predictor_grad_impl(net, g->g, g_pred, g_pred_broadcast, g_z, workspace, workspace_bytes, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------------
// sub-module forwards (inference only): GCL / EquivariantUpdate / EquivariantBlock / EGNN / E_GCL / predictor EGNN.
// Node tensors are [n_nodes, H] with H == padded hidden width (hidden_nf in {64,192,196,256}).
// ------------------------------------------------------------------------------------------------------
static int check_layer_api(const gb_net* n, int kind, int layer) {
    if (!n) return fail("null net");
    if (n->kind != kind) return fail("wrong network kind for this call");
    if (n->H != n->HP) return fail("sub-module API needs hidden_nf in {64,192,196,256} (got %d)", n->H);
    if (layer < 0 || layer >= n->L) return fail("layer index %d out of range", layer);
    return 0;
}

static DenEdgeArgs den_args_base(const gb_net* n, const Graph& g) {
    DenEdgeArgs a;
    memset(&a, 0, sizeof(a));
    a.g = g; a.use_tanh = n->use_tanh; a.norm_constant = n->norm_constant; a.normf = n->normf; a.coords_range = n->coords_range;
    return a;
}

static void run_den_gcl(const gb_net* n, const Graph& g, const DenGcl& G, const DenWs& w, const float* h, float* h_out,
                        const float* x, const float* x0, const float* eattr, const float* d0_edge, cudaStream_t s) {
    lin_P(n, G.e, h, w.P, g.n_nodes, s);
    DenEdgeArgs a = den_args_base(n, g);
    a.P = w.P; a.ext = n->p(G.e.l1_ext); a.wt2 = n->p(G.e.l2_wt); a.b2 = n->p(G.e.l2_b); a.vecw = n->p(G.att_w);
    a.att_b = G.att_b; a.attention = n->attention; a.x = x; a.x0 = x0; a.eattr = eattr; a.d0_edge = d0_edge; a.agg = w.agg;
    if (n->tc_den) launch_den_edge_tc(n->HP, 0, a, n->p(G.e.l2_tc), s); else launch_den_edge(n->HP, 0, a, s);
    GB_LAUNCHED(1);
    node_update(n, G.n, h, w.agg, w.s, h_out, nullptr, g, s);
}

static void run_den_equiv(const gb_net* n, const Graph& g, const DenEquiv& E, const DenWs& w, const float* h, const float* x,
                          const float* x0, const float* cdiff, const float* eattr, const float* d0_edge, float* x_out, cudaStream_t s) {
    lin_P(n, E.c, h, w.P, g.n_nodes, s);
    DenEdgeArgs a = den_args_base(n, g);
    a.P = w.P; a.ext = n->p(E.c.l1_ext); a.wt2 = n->p(E.c.l2_wt); a.b2 = n->p(E.c.l2_b); a.vecw = n->p(E.last_w);
    a.x = x; a.x0 = x0; a.cdiff = cdiff; a.eattr = eattr; a.d0_edge = d0_edge; a.x_out = x_out;
    if (n->tc_den) launch_den_edge_tc(n->HP, 1, a, n->p(E.c.l2_tc), s); else launch_den_edge(n->HP, 1, a, s);
    GB_LAUNCHED(1);
}

extern "C" int gb_den_gcl_forward(const gb_net* n, const gb_graph* gg, int block, int sub, const float* h_in, const float* eattr,
                                  float* h_out, void* ws, size_t ws_bytes, void* stream) {
    if (check_layer_api(n, 0, block)) return 1;
    if (sub < 0 || sub >= n->n_sub || !gg || !h_in || !eattr || !h_out) return fail("bad argument");
    Bump b(ws, ws_bytes); DenWs w; carve_den(b, n, gg->g, w);
    if (b.off > ws_bytes) return fail("workspace too small");
    run_den_gcl(n, gg->g, n->gcl[(size_t)block * n->n_sub + sub], w, h_in, h_out, nullptr, nullptr, eattr, nullptr, (cudaStream_t)stream);
    return check_launch("gcl_forward");
}

extern "C" int gb_den_equiv_forward(const gb_net* n, const gb_graph* gg, int block, const float* h_in, const float* x_in,
                                    const float* cdiff, const float* eattr, float* x_out, void* ws, size_t ws_bytes, void* stream) {
    if (check_layer_api(n, 0, block)) return 1;
    if (!gg || !h_in || !x_in || !cdiff || !eattr || !x_out) return fail("bad argument");
    Bump b(ws, ws_bytes); DenWs w; carve_den(b, n, gg->g, w);
    if (b.off > ws_bytes) return fail("workspace too small");
    run_den_equiv(n, gg->g, n->eq[block], w, h_in, x_in, x_in, cdiff, eattr, nullptr, x_out, (cudaStream_t)stream);
    return check_launch("equiv_forward");
}

// EquivariantBlock.forward: h, x -> h', x' with the second edge attribute given per compacted edge
extern "C" int gb_den_block_forward(const gb_net* n, const gb_graph* gg, int block, const float* h_in, const float* x_in,
                                    const float* d0_edge, float* h_out, float* x_out, void* ws, size_t ws_bytes, void* stream) {
    if (check_layer_api(n, 0, block)) return 1;
    if (!gg || !h_in || !x_in || !d0_edge || !h_out || !x_out) return fail("bad argument");
    const Graph& g = gg->g; cudaStream_t s = (cudaStream_t)stream;
    Bump b(ws, ws_bytes); DenWs w; carve_den(b, n, g, w);
    if (b.off > ws_bytes) return fail("workspace too small");
    const float* h = h_in;
    for (int q = 0; q < n->n_sub; ++q) {
        float* dst = (q + 1 == n->n_sub) ? h_out : ((q & 1) ? w.h2 : w.h);
        run_den_gcl(n, g, n->gcl[(size_t)block * n->n_sub + q], w, h, dst, x_in, x_in, nullptr, d0_edge, s);
        h = dst;
    }
    run_den_equiv(n, g, n->eq[block], w, h, x_in, x_in, nullptr, nullptr, d0_edge, x_out, s);
    return check_launch("block_forward");
}

// EGNN.forward of the denoiser: h_in [n, F+1] (time column included), x_in -> h_out [n, F+1], x_out
extern "C" int gb_den_egnn_forward(const gb_net* n, const gb_graph* gg, const float* h_in, const float* x_in, float* h_out,
                                   float* x_out, void* ws, size_t ws_bytes, void* stream) {
    if (check_layer_api(n, 0, 0)) return 1;
    if (!gg || !h_in || !x_in || !h_out || !x_out) return fail("bad argument");
    const Graph& g = gg->g; cudaStream_t s = (cudaStream_t)stream;
    Bump b(ws, ws_bytes); DenWs w; carve_den(b, n, g, w);
    if (b.off > ws_bytes) return fail("workspace too small");
    launch_embed_plain(h_in, n->F + 1, n->p(n->emb_w), n->p(n->emb_b), g.n_nodes, n->H, n->HP, w.h, s); GB_LAUNCHED(1);
    GB_CUDA(cudaMemcpyAsync(w.x0, x_in, (size_t)g.n_nodes * 3 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    float *h = w.h, *h2 = w.h2; const float* xc = w.x0; float* xn = w.xa;
    for (int blk = 0; blk < n->L; ++blk) {
        for (int q = 0; q < n->n_sub; ++q) {
            run_den_gcl(n, g, n->gcl[(size_t)blk * n->n_sub + q], w, h, h2, xc, w.x0, nullptr, nullptr, s);
            float* tmp = h; h = h2; h2 = tmp;
        }
        float* dst = (blk + 1 == n->L) ? x_out : xn;
        run_den_equiv(n, g, n->eq[blk], w, h, xc, w.x0, nullptr, nullptr, nullptr, dst, s);
        xc = dst; xn = (xn == w.xa) ? w.xb : w.xa;
    }
    EmbedOutArgs eo{h, n->HP, n->H, n->p(n->out_w), n->p(n->out_b), n->F + 1, g.node_mask, g.n_nodes, h_out, n->F + 1};
    launch_embed_out(eo, s); GB_LAUNCHED(1);
    return check_launch("egnn_forward");
}

// E_GCL.forward: h, coord, edge_attr (per compacted edge) -> h', coord'
extern "C" int gb_pred_layer_forward(const gb_net* n, const gb_graph* gg, int layer, const float* h_in, const float* x_in,
                                     const float* a_edge, float* h_out, float* x_out, void* ws, size_t ws_bytes, void* stream) {
    if (check_layer_api(n, 1, layer)) return 1;
    if (!gg || !h_in || !x_in || !a_edge || !h_out || !x_out) return fail("bad argument");
    const Graph& g = gg->g; cudaStream_t s = (cudaStream_t)stream;
    Bump b(ws, ws_bytes); PredWs w; carve_pred(b, n, g, false, w);
    if (b.off > ws_bytes) return fail("workspace too small");
    const PredLayer& Lr = n->pl[layer];
    lin_P(n, Lr.e, h_in, w.P, g.n_nodes, s);
    PredEdgeArgs a = pred_edge_args(n, Lr, g, w, 0, false);
    a.x = x_in; a.x0 = x_in; a.a_edge = a_edge; a.x_out = x_out;
    if (n->tc_pred) launch_pred_edge_fwd_tc(n->HP, false, a, n->p(Lr.e.l2_tc), n->p(Lr.c_tc), s);
    else launch_pred_edge_fwd(n->HP, false, a, s);
    GB_LAUNCHED(1);
    node_update(n, Lr.n, h_in, w.agg, w.s, h_out, nullptr, g, s);
    return check_launch("e_gcl_forward");
}

// predictor EGNN.forward: h_in [n, F+1], x_in, edge_attr per compacted edge -> h_out [n, out_nf] (masked), x_out
extern "C" int gb_pred_egnn_forward(const gb_net* n, const gb_graph* gg, const float* h_in, const float* x_in, const float* a_edge,
                                    float* h_out, float* x_out, void* ws, size_t ws_bytes, void* stream) {
    if (check_layer_api(n, 1, 0)) return 1;
    if (!gg || !h_in || !x_in || !a_edge || !h_out || !x_out) return fail("bad argument");
    const Graph& g = gg->g; cudaStream_t s = (cudaStream_t)stream;
    Bump b(ws, ws_bytes); PredWs w; carve_pred(b, n, g, false, w);
    if (b.off > ws_bytes) return fail("workspace too small");
    const size_t nn = g.n_nodes;
    launch_embed_plain(h_in, n->F + 1, n->p(n->emb_w), n->p(n->emb_b), g.n_nodes, n->H, n->HP, w.h, s); GB_LAUNCHED(1);
    GB_CUDA(cudaMemcpyAsync(w.x, x_in, nn * 3 * sizeof(float), cudaMemcpyDeviceToDevice, s));
    float *h = w.h, *h2 = w.h2;
    for (int l = 0; l < n->L; ++l) {
        const PredLayer& Lr = n->pl[l];
        lin_P(n, Lr.e, h, w.P, g.n_nodes, s);
        PredEdgeArgs a = pred_edge_args(n, Lr, g, w, l, false);
        a.a_edge = a_edge;
        if (l + 1 == n->L) a.x_out = x_out;
        if (n->tc_pred) launch_pred_edge_fwd_tc(n->HP, false, a, n->p(Lr.e.l2_tc), n->p(Lr.c_tc), s);
        else launch_pred_edge_fwd(n->HP, false, a, s);
        GB_LAUNCHED(1);
        node_update(n, Lr.n, h, w.agg, w.s, h2, nullptr, g, s);
        float* tmp = h; h = h2; h2 = tmp;
    }
    EmbedOutArgs eo{h, n->HP, n->H, n->p(n->out_w), n->p(n->out_b), n->out_nf, g.node_mask, g.n_nodes, h_out, n->out_nf};
    launch_embed_out(eo, s); GB_LAUNCHED(1);
    return check_launch("pred_egnn_forward");
}

// ------------------------------------------------------------------------------------------------------
// step pieces
// ------------------------------------------------------------------------------------------------------
extern "C" int gb_step_sample(const float* zt, const float* eps, const float* noise, const float* coef,
                              const float* node_mask, int B, int N, int D, unsigned long long seed,
                              unsigned long long draw, int project, float* zs, void* stream) {
    StepArgs a{zt, eps, noise, coef, node_mask, B, N, D, seed, draw, 1.0f, project, zs, nullptr, 0};
    launch_step_pre(a, (cudaStream_t)stream); GB_LAUNCHED(1);
    return check_launch("step_sample");
}
extern "C" int gb_step_guide(const float* zs_pre, const float* grad, const float* coef, const float* node_mask, int B,
                             int N, int D, float max_norm, float* zs, void* stream) {
    GuideArgs a{zs_pre, grad, coef, node_mask, B, N, D, max_norm, zs};
    launch_step_guide(a, (cudaStream_t)stream); GB_LAUNCHED(1);
    return check_launch("step_guide");
}
extern "C" int gb_decode(const float* z0, const float* eps, const float* noise, const float* coef, const float* node_mask,
                         int B, int N, int D, unsigned long long seed, unsigned long long draw, float norm_x, float norm_h,
                         float bias_h, float* x, float* one_hot, float* cog_max, void* stream) {
    DecodeArgs a{z0, eps, noise, coef, node_mask, B, N, D, seed, draw, norm_x, norm_h, bias_h, x, one_hot, cog_max};
    launch_decode(a, (cudaStream_t)stream); GB_LAUNCHED(1);
    return check_launch("decode");
}
extern "C" int gb_cog_fix(float* x, const float* node_mask, const float* cog_max, float thresh, int B, int N, void* stream) {
    launch_cog_fix(x, node_mask, cog_max, thresh, B, N, (cudaStream_t)stream); GB_LAUNCHED(1);
    return check_launch("cog_fix");
}
extern "C" int gb_noise(float* out, const float* node_mask, int B, int N, int D, float std, unsigned long long seed,
                        unsigned long long draw, void* stream) {
    launch_noise(out, node_mask, B, N, D, std, seed, draw, (cudaStream_t)stream); GB_LAUNCHED(1);
    return check_launch("noise");
}

// ------------------------------------------------------------------------------------------------------
// whole loop
// ------------------------------------------------------------------------------------------------------
struct Cursor { float coef[3]; float t; unsigned long long draw; long long s; };

__global__ void cursor_set_kernel(Cursor* c, const float* sched, const float* tvals, int T, long long s) {
    c->s = s; c->coef[0] = sched[3 * s]; c->coef[1] = sched[3 * s + 1]; c->coef[2] = sched[3 * s + 2];
    c->t = tvals[s + 1]; c->draw = (unsigned long long)(T - s);
}
__global__ void cursor_next_kernel(Cursor* c, const float* sched, const float* tvals, int T) {
    const long long s = c->s - 1;
    if (s < 0) return;
    c->s = s; c->coef[0] = sched[3 * s]; c->coef[1] = sched[3 * s + 1]; c->coef[2] = sched[3 * s + 2];
    c->t = tvals[s + 1]; c->draw = (unsigned long long)(T - s);
}

namespace {
struct LoopWs { Cursor* cur; float *eps, *zs_pre, *grad, *pred; void* den; size_t den_bytes; void* prd; size_t prd_bytes; };
void carve_loop(Bump& b, const gb_net* den, const gb_net* pred, const Graph& g, LoopWs& w) {
    const size_t nd = (size_t)g.n_nodes * (3 + den->F);
    w.cur = b.get<Cursor>(1);
    w.eps = b.get<float>(nd); w.zs_pre = b.get<float>(nd); w.grad = b.get<float>(nd);
    w.pred = b.get<float>((size_t)g.B * (pred ? pred->out_nf : 1));
    { Bump d(nullptr, 0); DenWs dw; carve_den(d, den, g, dw); w.den_bytes = d.off; }
    w.den = b.get<char>(w.den_bytes);
    w.prd_bytes = 0; w.prd = nullptr;
    if (pred) { Bump p(nullptr, 0); PredWs pw; carve_pred(p, pred, g, true, pw); w.prd_bytes = p.off; w.prd = b.get<char>(w.prd_bytes); }
}
}  // namespace

extern "C" size_t gb_sample_loop_workspace_bytes(const gb_net* den, const gb_net* pred, const gb_graph* g) {
    Bump b(nullptr, 0); LoopWs w; carve_loop(b, den, pred, g->g, w); return b.off;
}

static int one_step(const gb_net* den, const gb_net* pred, const Graph& g, float* z, const LoopWs& w, const float* target_w,
                    const float* noise, unsigned long long seed, float* stats, cudaStream_t s) {
    const int D = 3 + den->F;
    if (denoiser_forward_impl(den, g, z, &w.cur->t, 0, w.eps, pred ? 1 : 0, stats, stats ? &w.cur->s : nullptr, w.den,
                              w.den_bytes, s)) return 1;
    StepArgs a{z, w.eps, noise, w.cur->coef, g.node_mask, g.B, g.N, D, seed, 0ull, 1.0f, pred ? 0 : 1,
               pred ? w.zs_pre : z, &w.cur->draw, (size_t)g.n_nodes * D};
    launch_step_pre(a, s); GB_LAUNCHED(1);
    if (pred) {
        if (predictor_forward_impl(pred, g, w.zs_pre, &w.cur->t, 0, w.pred, 1, w.prd, w.prd_bytes, s)) return 1;
        if (predictor_grad_impl(pred, g, target_w, 1, w.grad, w.prd, w.prd_bytes, s)) return 1;
        GuideArgs gd{w.zs_pre, w.grad, w.cur->coef, g.node_mask, g.B, g.N, D, 10.0f, z};
        launch_step_guide(gd, s); GB_LAUNCHED(1);
    }
    return check_launch("sample step");
}

extern "C" int gb_sample_loop(const gb_net* den, const gb_net* pred, const gb_graph* gg, float* z, int T, int s_hi, int s_lo,
                              const float* sched, const float* tvals, const float* target_w, const float* noise,
                              unsigned long long seed, float* stats, void* workspace, size_t workspace_bytes, int use_graph,
                              void* stream) {
    if (!den || !gg || !z || !sched || !tvals || !workspace) return fail("null argument");
    if (pred && !target_w) return fail("guided loop needs target_w");
    if (s_hi > T || s_lo < 0 || s_lo >= s_hi) return fail("bad step range [%d,%d) for T=%d", s_lo, s_hi, T);
    const Graph& g = gg->g;
    cudaStream_t s = (cudaStream_t)stream;
    Bump b(workspace, workspace_bytes); LoopWs w; carve_loop(b, den, pred, g, w);
    if (b.off > workspace_bytes) return fail("loop workspace too small: need %zu bytes, got %zu", b.off, workspace_bytes);
    const int n_steps = s_hi - s_lo;
    cursor_set_kernel<<<1, 1, 0, s>>>(w.cur, sched, tvals, T, (long long)(s_hi - 1)); GB_LAUNCHED(1);
    if (!use_graph || n_steps < 3) {
        for (int i = 0; i < n_steps; ++i) {
            if (one_step(den, pred, g, z, w, target_w, noise, seed, stats, s)) return 1;
            cursor_next_kernel<<<1, 1, 0, s>>>(w.cur, sched, tvals, T); GB_LAUNCHED(1);
        }
        return check_launch("sample_loop");
    }
    // capture ONE step (all per-step scalars are read through the device cursor), replay it n_steps times
    cudaGraph_t graph = nullptr; cudaGraphExec_t exec = nullptr;
    cudaStream_t cs = s;
    bool own_stream = false;
    if (cs == nullptr || cs == cudaStreamLegacy) { GB_CUDA(cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking)); own_stream = true; GB_CUDA(cudaStreamSynchronize(s)); }
    const long long before = g_launches;
    GB_CUDA(cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal));
    int rc = one_step(den, pred, g, z, w, target_w, noise, seed, stats, cs);
    cursor_next_kernel<<<1, 1, 0, cs>>>(w.cur, sched, tvals, T); GB_LAUNCHED(1);
    cudaError_t ce = cudaStreamEndCapture(cs, &graph);
    if (rc || ce != cudaSuccess) { if (graph) cudaGraphDestroy(graph); if (own_stream) cudaStreamDestroy(cs); return rc ? rc : fail("graph capture: %s", cudaGetErrorString(ce)); }
    const long long per_step = g_launches - before;
    GB_CUDA(cudaGraphInstantiate(&exec, graph, 0));
    for (int i = 0; i < n_steps; ++i) GB_CUDA(cudaGraphLaunch(exec, cs));
    g_launches += per_step * (n_steps - 1);
    GB_CUDA(cudaStreamSynchronize(cs));        // the exec object must outlive its launches
    cudaGraphExecDestroy(exec); cudaGraphDestroy(graph);
    if (own_stream) cudaStreamDestroy(cs);
    return check_launch("sample_loop(graph)");
}

// ------------------------------------------------------------------------------------------------------
// measurement aid: re-launch ONE kernel of the path `repeats` times on the workspace left behind by the last
// forward / input-gradient call, so that bench.py can bracket exactly that kernel with CUDA events.
//   which: 0 denoiser GCL edge kernel, 1 denoiser EquivariantUpdate edge kernel (workspace of gb_denoiser_forward)
//          2 predictor edge forward (saving), 3 predictor edge backward, 4 node MLP first Linear (predictor
//          workspace after gb_predictor_forward(save)+gb_predictor_input_grad)
// ------------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------------
// training step: one Linear on the tcgen05 path (weights are re-packed on every call: they change every step)
// ------------------------------------------------------------------------------------------------------
extern "C" size_t gb_linear_scratch_bytes(int N, int K1, int K2) {
    return (size_t)((K1 + 31) / 32 + (K2 + 31) / 32) * 2 * tc_np(N) * 32 * sizeof(float);
}
extern "C" int gb_linear(int M, int N, int K1, int K2, const float* A1, int lda1, const float* A2, int lda2, const float* W,
                         int ldw, int transpose_w, const float* bias, int epi, float* out, float* out2, const float* res_or_aux,
                         const float* mask, void* wimg, size_t wimg_bytes, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (M <= 0) return 0;
    if (N > 256 || (N & 3) || (K1 & 3) || (K2 & 3) || (lda1 & 3) || (K2 && (lda2 & 3)))
        return fail("gb_linear: needs N <= 256 and N, K1, K2, lda multiples of 4 (got N=%d K1=%d K2=%d)", N, K1, K2);
    if (K2 && transpose_w) return fail("gb_linear: a K-concatenated input is only supported with transpose_w = 0");
    if (wimg_bytes < gb_linear_scratch_bytes(N, K1, K2)) return fail("gb_linear: scratch too small");
    const int NP = tc_np(N), na1 = (K1 + 31) / 32, na2 = (K2 + 31) / 32;
    float* img = (float*)wimg;
    launch_pack_tc(img, W, ldw, 0, 0, K1, N, NP, na1, transpose_w, s);
    if (K2) launch_pack_tc(img + (size_t)na1 * 2 * NP * 32, W, ldw, K1, 0, K2, N, NP, na2, 0, s);
    LinArgs a{};
    a.A1 = A1; a.lda1 = lda1; a.K1 = K1; a.A2 = A2; a.lda2 = lda2; a.K2 = K2;
    a.bias = bias; a.out = out; a.ldo = N; a.out2 = out2; a.ldo2 = N; a.M = M; a.ncb = 1; a.res_cb = -1;
    switch (epi) {
        case 0: a.epi = EPI_BIAS; break;
        case 1: a.epi = EPI_SILU; break;
        case 2: a.epi = EPI_RES_MASK; a.res = res_or_aux; a.ldr = N; a.mask = mask; break;
        case 3: a.epi = EPI_MUL_DSILU; a.aux = res_or_aux; a.ldaux = N; break;
        case 4: a.epi = EPI_ADD_RES; a.res = res_or_aux; a.ldr = N; break;
        default: return fail("gb_linear: unknown epilogue %d", epi);
    }
    launch_lin_tc(N, a, img, s);
    GB_LAUNCHED(K2 ? 3 : 2);
    return check_launch("gb_linear");
}

extern "C" size_t gb_wgrad_scratch_bytes(int M, int N) { return wgrad_tc_scratch_bytes(M, N); }
extern "C" int gb_wgrad(int K, int M, int N, const float* G, int ldg, const float* X, int ldx, float* C, int ldc, int accumulate,
                        float* colsum, void* scratch, size_t scratch_bytes, void* stream) {
    cudaStream_t s = (cudaStream_t)stream;
    if (M <= 0 || N <= 0) return 0;
    if (M > 256 || N > 256 || (M & 3) || (N & 3) || (ldg & 3) || (ldx & 3) || ((uintptr_t)G & 15) || ((uintptr_t)X & 15))
        return fail("gb_wgrad: needs M, N <= 256, multiples of 4, and 16-byte aligned rows (got M=%d N=%d ldg=%d ldx=%d)", M, N, ldg, ldx);
    float* sc = (scratch && scratch_bytes >= wgrad_tc_scratch_bytes(M, N) && ((uintptr_t)scratch & 15) == 0) ? (float*)scratch : nullptr;
    if (colsum && (!sc || N >= 256)) return fail("gb_wgrad: the fused column sum needs the scratch (two-phase) path and N < 256");
    if (K <= 0) {
        if (!accumulate) GB_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, M, s));
        if (colsum) GB_CUDA(cudaMemsetAsync(colsum, 0, (size_t)M * 4, s));
        return 0;
    }
    if (!sc && !accumulate) GB_CUDA(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)N * 4, M, s));
    launch_wgrad_tc(K, M, N, G, ldg, X, ldx, C, ldc, accumulate, sc, colsum, s);
    GB_LAUNCHED(sc ? 2 : 1);
    return check_launch("gb_wgrad");
}

extern "C" int gb_profile_kernel(const gb_net* n, const gb_graph* gg, int which, int layer, void* ws, size_t ws_bytes,
                                 int repeats, void* stream) {
    if (!n || !gg || !ws) return fail("null argument");
    const Graph& g = gg->g;
    cudaStream_t s = (cudaStream_t)stream;
    if (layer < 0 || layer >= n->L) return fail("bad layer");
    if (which <= 1) {
        if (n->kind != 0) return fail("which=0/1 need a denoiser handle");
        Bump b(ws, ws_bytes); DenWs w; carve_den(b, n, g, w);
        if (b.off > ws_bytes) return fail("workspace too small");
        for (int r = 0; r < repeats; ++r) {
            DenEdgeArgs a;
            memset(&a, 0, sizeof(a));
            a.g = g; a.P = w.P; a.attention = which == 0 ? n->attention : 0; a.use_tanh = n->use_tanh;
            a.norm_constant = n->norm_constant; a.normf = n->normf; a.coords_range = n->coords_range;
            a.x = w.x0; a.x0 = w.x0;
            if (which == 0) {
                const DenGcl& G = n->gcl[(size_t)layer * n->n_sub];
                a.ext = n->p(G.e.l1_ext); a.wt2 = n->p(G.e.l2_wt); a.b2 = n->p(G.e.l2_b); a.vecw = n->p(G.att_w); a.att_b = G.att_b;
                a.agg = w.agg;
            } else {
                const DenEquiv& E = n->eq[layer];
                a.ext = n->p(E.c.l1_ext); a.wt2 = n->p(E.c.l2_wt); a.b2 = n->p(E.c.l2_b); a.vecw = n->p(E.last_w);
                a.x_out = w.xb;
            }
            const float* img = which == 0 ? n->p(n->gcl[(size_t)layer * n->n_sub].e.l2_tc) : n->p(n->eq[layer].c.l2_tc);
            if (n->tc_den) launch_den_edge_tc(n->HP, which, a, img, s); else launch_den_edge(n->HP, which, a, s);
            GB_LAUNCHED(1);
        }
        return check_launch("profile den_edge");
    }
    if (n->kind != 1) return fail("which=2..4 need a predictor handle");
    Bump b(ws, ws_bytes); PredWs w; carve_pred(b, n, g, true, w);
    if (b.off > ws_bytes) return fail("workspace too small");
    const PredLayer& Lr = n->pl[layer];
    for (int r = 0; r < repeats; ++r) {
        if (which == 2) {
            PredEdgeArgs a = pred_edge_args(n, Lr, g, w, layer, true);
            a.x_out = w.gx2;                 // scratch target: keep the saved coordinates of layer+1 intact
            if (n->tc_pred) launch_pred_edge_fwd_tc(n->HP, true, a, n->p(Lr.e.l2_tc), n->p(Lr.c_tc), s);
            else launch_pred_edge_fwd(n->HP, true, a, s);
        } else if (which == 3) {
            PredEdgeArgs e = pred_edge_args(n, Lr, g, w, layer, true);
            e.w2_nt = n->p(Lr.e.l2_nt); e.wc_nt = n->p(Lr.c_nt);
            e.g_agg = w.gcat + n->HP; e.ld_gagg = 2 * n->HP; e.g_xout = w.gx; e.g_Pa = w.gPa; e.g_Pb = w.gPb; e.g_x = w.gx2; e.g_attr = w.gattr;
            e.g_pre1 = w.gpre1; e.g_d = w.gd;
            if (n->tc_pred) launch_pred_edge_bwd_tc(n->HP, e, n->p(Lr.c_nt_tc), n->p(Lr.e.l2_nt_tc), s);
            else launch_pred_edge_bwd(n->HP, e, s);
        } else {
            LinArgs a = lin_base(g.n_nodes);
            a.A1 = w.h; a.lda1 = n->HP; a.K1 = n->HP; a.A2 = w.agg; a.lda2 = n->HP; a.K2 = n->HP;
            a.wt = n->p(Lr.n.l1_wt); a.wt_tc = n->p(Lr.n.l1_tc); a.bias = n->p(Lr.n.l1_b); a.out = w.s; a.ldo = n->HP; a.epi = EPI_SILU;
            run_lin(n, a, s);
        }
        GB_LAUNCHED(1);
    }
    return check_launch("profile pred kernel");
}
