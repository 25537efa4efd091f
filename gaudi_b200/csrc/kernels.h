// Internal launcher interface between api.cu (orchestration, C-ABI) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#define GB_TM_HOST 128   // rows per tile (== GB_TM of common.cuh)
#define GB_PS_ROWS_HOST 96   // node-projection rows staged per K-atom by the tcgen05 edge kernels (== tc::PS_ROWS)

namespace gb {

// Compacted, row-sorted edge structure of a batch of small graphs (built once per mask set).
//   nodes   : B*N padded nodes, node id = b*N + i
//   edges   : all (i,j) with edge_mask != 0 in the reference's dense row-major order
//             (edm/egnn/models.py:154-175), so the edges of one row node are contiguous
//   tiles   : consecutive node ranges whose edges (<=128) form one GEMM tile; a row segment never
//             straddles two tiles, so row aggregation is atomics-free and order-deterministic
//   tc_*    : per tile, the tile's edges grouped by column node (for the backward column scatter)
struct Graph {
    int n_nodes, n_edges, n_tiles, B, N;
    const int* rowptr;    // [n_nodes+1]
    const int* erow;      // [n_edges]
    const int* ecol;      // [n_edges]
    const int* tile_ptr;  // [n_tiles+1] node boundaries
    const int* tc_ptr;    // [n_tiles+1] -> range in tc_node/tc_start
    const int* tc_node;   // [n_tc] distinct column nodes per tile
    const int* tc_start;  // [n_tc+1] start in cperm (absolute edge slot), sentinel = n_edges
    const int* cperm;     // [n_edges] tile-local row index, grouped by column node
    const int* colptr;    // [n_nodes+1] CSC pointer
    const int* cedge;     // [n_edges] edge ids grouped by column node
    const float* node_mask;  // [n_nodes]
    const int4* tile_info;   // [n_tiles] (node_lo, n_nodes, e_lo, n_edges) of every tile: one load instead of a dependent chain
    int ps_rows;             // max over tiles of (row nodes + nodes of the touched graphs): node-projection rows staged per K-atom
};

enum LinEpi { EPI_BIAS = 0, EPI_SILU = 1, EPI_RES_MASK = 2, EPI_MUL_DSILU = 3, EPI_ADD_RES = 4 };

// out[:, cb*HP:(cb+1)*HP] = epi( [A1 | A2] @ Wt_cb + bias_cb )   for cb < ncb
struct LinArgs {
    const float* A1; int lda1; int K1;
    const float* A2; int lda2; int K2;      // optional K-concatenated second input (K2 = 0: none)
    const float* rowscale;                  // optional [M]: A rows are multiplied by it while loading
    const float* wt;                        // packed [ncb][(K1+K2)][HP]
    const float* wt_tc;                     // tensor-core image [ncb][atoms(K1)+atoms(K2)][hi|lo][NP][32] (swizzled) or null
    const float* bias;                      // [ncb*HP] or null
    float* out; int ldo;
    float* out2; int ldo2;                  // EPI_SILU: optional copy of the pre-activation
    const float* res; int ldr;              // EPI_RES_MASK / EPI_ADD_RES
    const float* mask;                      // EPI_RES_MASK: [M]; EPI_ADD_RES: optional scale of res rows
    const float* aux; int ldaux;            // EPI_MUL_DSILU: pre-activation
    int M; int ncb; int epi;
    int res_cb;                             // EPI_ADD_RES: column block that receives the residual (-1: all)
    int mix;                                // tcgen05 path: wt_tc is a hi + bf16-mix image (TF32 MMA + one bf16 MMA per K step) instead of hi | lo
};
void launch_lin(int HP, const LinArgs& a, cudaStream_t s);
// tcgen05 / 3xTF32 version (tc_lin_kernel.cu); H = row stride / hidden width of the node tensors (== HP)
void launch_lin_tc(int H, const LinArgs& a, const float* wimg, cudaStream_t s);
int tc_np(int H);
// C[m][n] += sum_e G[e][m] X[e][n] on tcgen05 (both operands MN-major, 3xTF32); C must be zeroed / hold the value to add to
void launch_wgrad_tc(int K, int M, int N, const float* G, int ldg, const float* X, int ldx, float* C, int ldc, int accumulate,
                     float* scratch, float* colsum, cudaStream_t s);
size_t wgrad_tc_scratch_bytes(int M, int N);
// fmt: 0 = [hi | lo] fp32 images (three TF32 MMAs: node Linear, training); 1 / 2 = [hi fp32 | hi+lo as fp16 / bf16] (edge kernels)
void launch_pack_tc(float* dst, const float* src, int ld, int k_off, int n_off, int Kv, int Nv, int NP, int atoms, int transpose, cudaStream_t s, int fmt = 0);

// tiny-K input embedding of both networks:  h0 = W [feat*mask , t] + b ;  x = z[:, :3]*mask
struct EmbedInArgs {
    const float* z; int D;                  // [n_nodes, D]
    const float* t_ptr; int t_per_mol;      // time value(s): one float or one per molecule
    const float* w; const float* b;         // Linear [H][F+1] row-major (unpacked), bias [H]
    const float* node_mask; int n_nodes; int N; int H; int HP;
    float* h; float* x;                     // [n_nodes, HP], [n_nodes, 3]
};
void launch_embed_in(const EmbedInArgs& a, cudaStream_t s);

// small-N output head: out[m, o] = (sum_k h[m,k] W[o,k] + b[o]) * mask[m]
struct EmbedOutArgs {
    const float* h; int HP; int H; const float* w; const float* b; int n_out; const float* node_mask; int n_nodes;
    float* out; int ldo;
};
void launch_embed_out(const EmbedOutArgs& a, cudaStream_t s);
// plain small-K Linear h0 = W h_in + b (EGNN.forward of the sub-module API)
void launch_embed_plain(const float* h_in, int K, const float* w, const float* b, int n_nodes, int H, int HP, float* h, cudaStream_t s);

// ---- denoiser edge kernels (GCL message / EquivariantUpdate) -----------------------------------
struct DenEdgeArgs {
    Graph g;
    const float* P;            // [n_nodes, 2*HP]  (W1a h + b1 | W1b h)
    const float* ext;          // [2][HP] rows of the first Linear acting on (radial, d0)
    const float* wt2;          // packed [HP][HP]
    const float* b2;           // [HP]
    const float* vecw;         // att_mlp weight (mode 0) / last coord Linear (mode 1), [HP]
    float att_b; int attention; int use_tanh;
    float norm_constant, normf, coords_range;
    const float* x;            // current coords [n_nodes,3]
    const float* x0;           // coords at network input (second edge attribute)
    const float* eattr;        // optional explicit [n_edges][2] edge attributes (module-level API); else null
    const float* cdiff;        // optional explicit [n_edges][3] coord_diff (module-level API)
    const float* d0_edge;      // optional explicit second edge attribute per compacted edge (EquivariantBlock.forward)
    float* agg;                // mode 0 out [n_nodes, HP]
    float* x_out;              // mode 1 out [n_nodes, 3]
};
void launch_den_edge(int HP, int mode, const DenEdgeArgs& a, cudaStream_t s);
void launch_den_edge_tc(int H, int mode, const DenEdgeArgs& a, const float* wimg, cudaStream_t s);

// ---- predictor E_GCL edge kernels ---------------------------------------------------------------
struct PredEdgeArgs {
    Graph g;
    const float* P;            // [n_nodes, 2*HP]
    const float* ext;          // [2][HP]: rows acting on (radial, edge_attr)
    const float* wt2; const float* b2;       // edge_mlp.2 packed^T, bias
    const float* att_w; float att_b; int attention;
    const float* wtc; const float* bc;       // coord_mlp.0 packed^T, bias
    const float* wc_last;                     // coord_mlp.2 weight [HP]
    int use_tanh; float coords_range;
    const float* x; const float* x0;          // current / input coordinates
    const float* a_edge;                      // optional explicit edge_attr per compacted edge (E_GCL.forward); else |x0_i-x0_j|^2
    float* agg; float* x_out;                 // forward outputs
    // activations saved for the input-gradient pass (null = inference only)
    float* sv_d1; float* sv_pre2; float* sv_d3; float* sv_tau;   // [n_tiles][HP][128] x3, [n_edges]
    // ---- backward only ----
    const float* w2_nt; const float* wc_nt;   // original-orientation [out][in] packed blocks
    const float* g_agg; int ld_gagg;   // [n_nodes, ld_gagg]  dL/d agg
    const float* g_xout;       // [n_nodes, 3]   dL/d x_{l+1}
    float* g_Pa;               // [n_nodes, HP]  (plain stores, row segments)
    float* g_Pb;               // [n_nodes, HP]  (zero-initialised, RED column scatter)
    float* g_x;                // [n_nodes, 3]   (zero-initialised, atomics)
    float* g_attr;             // [n_edges] running dL/d edge_attr (+=)
    float* g_pre1;             // tensor-core path: [n_edges, H] dL/d pre1 handed to the node-parallel reduction kernel
    float* g_d;                // tensor-core path: [n_edges, 3] dL/d (x_row - x_col)
};
void launch_pred_edge_fwd(int HP, bool save, const PredEdgeArgs& a, cudaStream_t s);
void launch_pred_edge_bwd(int HP, const PredEdgeArgs& a, cudaStream_t s);
void launch_pred_edge_fwd_tc(int H, bool save, const PredEdgeArgs& a, const float* w2img, const float* wcimg, cudaStream_t s);
void launch_pred_edge_bwd_tc(int H, const PredEdgeArgs& a, const float* wcimg_nt, const float* w2img_nt, cudaStream_t s);
// node-parallel, deterministic reductions of the backward: g_Pa (row sums), g_Pb (column sums), g_x
void launch_pred_bwd_reduce(int H, const PredEdgeArgs& a, cudaStream_t s);

size_t tile_kernel_smem_bytes(int HP);

// ---- small per-molecule / per-node kernels (step_kernels.cu) -----------------------------------
struct DenFinishArgs {   // eps = [ CoM-removed NaN-scrubbed (x_final - x_in)*mask , h_out[:, :F] ]
    const float* x_fin; const float* x_in; const float* h_out; int ld_h; const float* node_mask;
    int B, N, F; float* eps; int scrub_all;   // scrub_all: eps.nan_to_num(0.) of the guided step
    float* stats;                             // optional [4]: max|zt_x|, max|cog zt|, max|eps_x|, max|cog eps|
    const float* zt;
    const long long* stats_step;              // optional device step index: stats += 8 * (*stats_step)
};
void launch_den_finish(const DenFinishArgs& a, cudaStream_t s);

struct StepArgs {
    const float* zt; const float* eps; const float* noise;   // noise: injected [B,N,D] or null -> Philox
    const float* coef;                                        // [3]: alpha_ts, eps_coef, sigma  (device)
    const float* node_mask; int B, N, D;
    unsigned long long seed; unsigned long long draw;         // Philox stream (per draw index)
    float noise_std; int project;                             // project: remove CoM of x part (unguided path)
    float* zs;
    const unsigned long long* draw_ptr; size_t noise_stride;   // optional device draw index (overrides draw; noise += draw*stride)
};
void launch_step_pre(const StepArgs& a, cudaStream_t s);

struct GuideArgs {   // zs = nan_to_num(project(zs_pre - sigma * project(clip(grad))))
    const float* zs_pre; const float* grad; const float* coef; const float* node_mask; int B, N, D;
    float max_norm; float* zs;
};
void launch_step_guide(const GuideArgs& a, cudaStream_t s);

struct DecodeArgs {
    const float* z0; const float* eps; const float* noise; const float* coef;   // coef[3]: sigma0, alpha0, sigma_x
    const float* node_mask; int B, N, D; unsigned long long seed, draw;
    float norm_x, norm_h, bias_h;
    float* x; float* one_hot; float* cog_max;                                   // cog_max: device float (atomic max)
};
void launch_decode(const DecodeArgs& a, cudaStream_t s);
void launch_cog_fix(float* x, const float* node_mask, const float* cog_max, float thresh, int B, int N, cudaStream_t s);
void launch_noise(float* out, const float* node_mask, int B, int N, int D, float std, unsigned long long seed,
                  unsigned long long draw, cudaStream_t s);

// predictor head backward + pooled mean forward
void launch_pool_mean(const float* hout, const float* unused, int B, int N, int n_out, float* pred, cudaStream_t s);
struct HeadBwdArgs { const float* g_pred; const float* w; int n_out; int H; int HP; const float* node_mask; int B, N; float* gh; };
void launch_head_bwd(const HeadBwdArgs& a, cudaStream_t s);
struct InBwdArgs {   // dz from dL/dh0 (embedding), dL/dx0 and the accumulated edge-attribute gradient
    Graph g; const float* gh0; int HP; int H; const float* w_in; int F;   // w_in [H][F+1]
    const float* gx0; const float* g_attr; const float* x0; int D; float* gz;
};
void launch_in_bwd(const InBwdArgs& a, cudaStream_t s);

// weight packing
void launch_pack(float* dst, const float* src, int ld, int k_off, int n_off, int Kv, int Nv, int Kp, int HP,
                 int transpose, cudaStream_t s);

}  // namespace gb
