/* gaudi_b200 -- C ABI of the B200-native guided-sampling hot path of GaUDI.
 *
 * The reference (tomer196/GaUDI) has no FFI: its boundary is the Python module API
 * (SURVEY.md section 8b).  This header is the C boundary a maintainer would bind underneath that API
 * (ctypes stub in INTEGRATION.md); every entry point names the reference code it replaces.
 *
 * Conventions
 *   - plain C types, raw DEVICE pointers (fp32 / int32) unless a parameter says "host"
 *   - every function returns 0 on success, non-zero on error; gb_last_error() gives the thread-local message
 *   - "setup" functions may allocate device memory and synchronise; "hot path" functions never allocate,
 *     never synchronise, and only enqueue work on the given stream (CUDA-graph capturable)
 *   - tensors are contiguous, row-major, with the reference's shapes:
 *       z / eps / grad  [B, N, D]  D = 3 + F      node_mask [B*N]      t  one float, or B floats
 *   - stream is a cudaStream_t passed as void*
 */
#ifndef GAUDI_B200_H
#define GAUDI_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gb_net gb_net;       /* packed weights of one network (denoiser or predictor) */
typedef struct gb_graph gb_graph;   /* compacted edge structure of one (node_mask, edge_mask) pair */

int gb_abi_version(void);
const char* gb_last_error(void);
/* number of kernels launched by this library on the calling thread since the last reset (bench.py's gpu_launches) */
long long gb_launch_count(int reset);

/* ---- setup: networks --------------------------------------------------------------------------------
 * params: host array of DEVICE pointers to the fp32 parameters in state_dict order, without the
 * "buffer"/"gamma.gamma" entries.
 * denoiser (EGNN_dynamics.egnn, edm/egnn/egnn_new.py:238-321):
 *   embedding.{weight,bias}, embedding_out.{weight,bias}, then per e_block_i:
 *     per gcl_s: edge_mlp.0.{w,b} edge_mlp.2.{w,b} node_mlp.0.{w,b} node_mlp.2.{w,b} [att_mlp.0.{w,b}]
 *     gcl_equiv: coord_mlp.0.{w,b} coord_mlp.2.{w,b} coord_mlp.4.weight
 * predictor (EGNN_predictor.egnn, edm/egnn_predictor/models.py:492-560):
 *   embedding.{w,b}, embedding_out.{w,b}, then per gcl_i:
 *     edge_mlp.0.{w,b} edge_mlp.2.{w,b} node_mlp.0.{w,b} node_mlp.2.{w,b} coord_mlp.0.{w,b} coord_mlp.2.weight [att_mlp.0.{w,b}]
 */
int gb_denoiser_create(gb_net** out, int in_node_nf, int hidden_nf, int n_layers, int inv_sublayers, int attention,
                       int use_tanh, float coords_range, float norm_constant, float normalization_factor,
                       const float* const* params, int n_params, void* stream);
int gb_predictor_create(gb_net** out, int in_node_nf, int out_nf, int hidden_nf, int n_layers, int attention,
                        int use_tanh, float coords_range, const float* const* params, int n_params, void* stream);
int gb_net_destroy(gb_net* net);
int gb_net_hidden_padded(const gb_net* net);

/* ---- setup: graph topology (replaces get_adj_matrix, edm/egnn/models.py:154-175, and the per-call
 * h[row]/h[col] index tensors).  All arrays are DEVICE int32, borrowed for the lifetime of the handle.
 *   rowptr[n_nodes+1], erow/ecol[n_edges]: CSR of the edges with edge_mask != 0 in dense row-major order
 *   tile_ptr[n_tiles+1]: node boundaries of GEMM tiles (<=128 edges and <=128 nodes each, see gb_tile_pack)
 *   tc_ptr[n_tiles+1], tc_node[n_tc], tc_start[n_tc+1], cperm[n_edges]: each tile's edges grouped by column node
 *   colptr[n_nodes+1], cedge[n_edges]: CSC view (edge ids grouped by column node, ascending edge id inside a group) */
int gb_tile_pack(const int32_t* rowptr_host, int n_nodes, int32_t* tile_ptr_host_out, int* n_tiles_out);
/* Same greedy packing for a batch of graphs with nodes_per_graph padded nodes each (edges never leave their graph): a tile
 * additionally satisfies  n_nodes(tile) + nodes_per_graph * graphs_touched(tile) <= gb_stage_rows(),  the number of node-projection
 * rows the tcgen05 edge kernels stage in shared memory per K-atom.  gb_graph_create requires tiles packed this way. */
int gb_tile_pack_graphs(const int32_t* rowptr_host, int n_nodes, int nodes_per_graph, int32_t* tile_ptr_host_out, int* n_tiles_out);
int gb_stage_rows(void);
int gb_graph_create(gb_graph** out, int B, int N, int n_edges, int n_tiles, int n_tc, const int32_t* rowptr,
                    const int32_t* erow, const int32_t* ecol, const int32_t* tile_ptr, const int32_t* tc_ptr,
                    const int32_t* tc_node, const int32_t* tc_start, const int32_t* cperm, const int32_t* colptr,
                    const int32_t* cedge, const float* node_mask);
int gb_graph_destroy(gb_graph* g);

/* ---- workspace sizes (bytes of device scratch the hot-path calls need) */
size_t gb_denoiser_workspace_bytes(const gb_net* net, const gb_graph* g);
size_t gb_predictor_workspace_bytes(const gb_net* net, const gb_graph* g, int with_grad);

/* ---- hot path: networks ---------------------------------------------------------------------------------
 * gb_denoiser_forward   EGNN_dynamics._forward (edm/egnn/models.py:83-152) = phi of en_diffusion.py:352-355.
 *   scrub_all != 0 also applies the guided step's eps.nan_to_num(0.) (en_diffusion.py:881).
 *   stats (nullable, 8 floats, atomic-max accumulated): [0] max|z_x| [1] max|sum_n z_x| [2] max|eps_x|
 *   [3] max|sum_n eps_x| [4] max|z*(1-mask)| -- the quantities of assert_mean_zero_with_mask / assert_correctly_masked
 *   (edm/equivariant_diffusion/utils.py:52-65), checked by the caller once after the loop.
 * gb_predictor_forward  EGNN_predictor.forward (edm/egnn_predictor/models.py:433-457). save_for_grad keeps the
 *   activations the input-gradient pass needs inside the workspace.
 * gb_predictor_input_grad  d(sum_b <g_pred_b, pred_b>)/dz, replaces autograd.grad at en_diffusion.py:903.
 *   g_pred is [B,out] (or [out] shared by all molecules when g_pred_broadcast != 0).  Must follow a
 *   gb_predictor_forward(save_for_grad=1) on the same workspace. */
int gb_denoiser_forward(const gb_net* net, const gb_graph* g, const float* z, const float* t, int t_per_mol,
                        float* eps, int scrub_all, float* stats, void* workspace, size_t workspace_bytes, void* stream);
int gb_predictor_forward(const gb_net* net, const gb_graph* g, const float* z, const float* t, int t_per_mol,
                         float* pred, int save_for_grad, void* workspace, size_t workspace_bytes, void* stream);
int gb_predictor_input_grad(const gb_net* net, const gb_graph* g, const float* g_pred, int g_pred_broadcast,
                            float* g_z, void* workspace, size_t workspace_bytes, void* stream);

/* ---- hot path: reverse-diffusion step pieces --------------------------------------------------------------
 * coef: 3 device floats of the step  {alpha_t|s, sigma2_t|s/alpha_t|s/sigma_t, sigma_t|s*sigma_s/sigma_t}
 *       (en_diffusion.py:869-894), pre-computed by the caller with the reference's fp32 op order.
 * noise: injected [B,N,D] tensor (already masked / centred, parity mode) or NULL -> Philox(seed, draw).
 * gb_step_sample  zs = zt/coef[0] - coef[1]*eps + coef[2]*noise ; project != 0 removes the centre of gravity of the
 *                 x part (sample_p_zs_given_zt, en_diffusion.py:831-851); project == 0 is the pre-guidance z_s (:888-897)
 * gb_step_guide   clip(grad, 10) -> remove CoG -> zs = zs_pre - coef[2]*grad -> remove CoG -> nan_to_num (:905-934)
 * gb_decode       sample_p_xh_given_z0 (:533-560): coef = {sigma_0, alpha_0, exp(0.5*gamma_0)}; writes x [B,N,3],
 *                 one_hot [B,N,F] and atomically maxes |sum_n x| into cog_max (device float, nullable)
 * gb_cog_fix      re-projects x when *cog_max > thresh (:1059-1065)
 * gb_noise        sample_combined_position_feature_noise (:937-956) from Philox */
int gb_step_sample(const float* zt, const float* eps, const float* noise, const float* coef, const float* node_mask,
                   int B, int N, int D, unsigned long long seed, unsigned long long draw, int project, float* zs,
                   void* stream);
int gb_step_guide(const float* zs_pre, const float* grad, const float* coef, const float* node_mask, int B, int N,
                  int D, float max_norm, float* zs, void* stream);
int gb_decode(const float* z0, const float* eps, const float* noise, const float* coef, const float* node_mask, int B,
              int N, int D, unsigned long long seed, unsigned long long draw, float norm_x, float norm_h, float bias_h,
              float* x, float* one_hot, float* cog_max, void* stream);
int gb_cog_fix(float* x, const float* node_mask, const float* cog_max, float thresh, int B, int N, void* stream);
int gb_noise(float* out, const float* node_mask, int B, int N, int D, float std, unsigned long long seed,
             unsigned long long draw, void* stream);

/* ---- hot path: whole loop ----------------------------------------------------------------------------------
 * gb_sample_loop  EnVariationalDiffusion.sample / sample_guidance (en_diffusion.py:958-1067) for steps
 *   s = s_hi-1 ... s_lo (t = s+1), in place on z.  pred == NULL -> unguided.  Guidance target must be affine in the
 *   predictor outputs: energy_b = <target_w, pred_b> (+const); target_w [out] already multiplied by `scale`.
 *   sched: device [T][3] step coefficients; tvals: device [T+1] with tvals[k] = float(k)/T.
 *   noise: NULL (Philox; draw index of step s is T - s) or injected [T+2,B,N,D] indexed the same way.
 *   stats: nullable [T][8] per-step invariant maxima (see gb_denoiser_forward).
 *   use_graph != 0 captures one step into a CUDA graph and replays it (launch-bound small batches). */
int gb_sample_loop(const gb_net* den, const gb_net* pred, const gb_graph* g, float* z, int T, int s_hi, int s_lo,
                   const float* sched, const float* tvals, const float* target_w, const float* noise,
                   unsigned long long seed, float* stats, void* workspace, size_t workspace_bytes, int use_graph,
                   void* stream);
size_t gb_sample_loop_workspace_bytes(const gb_net* den, const gb_net* pred, const gb_graph* g);

/* ---- hot path: sub-module forwards (inference; node tensors [n_nodes, hidden_nf], hidden_nf in {64,192,196,256}) -------
 * Edge-level inputs are given per COMPACTED edge (the order of gb_graph's erow/ecol), i.e. the reference's dense
 * edge_attr / coord_diff tensors gathered at the edges with edge_mask != 0.
 * gb_den_gcl_forward     GCL.forward (edm/egnn/egnn_new.py:75-89): eattr [n_edges][2]
 * gb_den_equiv_forward   EquivariantUpdate.forward (:142-155): cdiff [n_edges][3], eattr [n_edges][2]
 * gb_den_block_forward   EquivariantBlock.forward (:214-235): d0_edge [n_edges] (the `edge_attr` argument)
 * gb_den_egnn_forward    EGNN.forward (:299-321): h_in/h_out [n_nodes, in_node_nf+1]
 * gb_pred_layer_forward  E_GCL.forward (edm/egnn_predictor/gcl.py:281-306): a_edge [n_edges]
 * gb_pred_egnn_forward   EGNN.forward (edm/egnn_predictor/models.py:543-560): h_out [n_nodes, out_nf]
 * workspace: gb_denoiser_workspace_bytes / gb_predictor_workspace_bytes(with_grad = 0). */
int gb_den_gcl_forward(const gb_net* net, const gb_graph* g, int block, int sub, const float* h_in, const float* eattr,
                       float* h_out, void* workspace, size_t workspace_bytes, void* stream);
int gb_den_equiv_forward(const gb_net* net, const gb_graph* g, int block, const float* h_in, const float* x_in,
                         const float* cdiff, const float* eattr, float* x_out, void* workspace, size_t workspace_bytes,
                         void* stream);
int gb_den_block_forward(const gb_net* net, const gb_graph* g, int block, const float* h_in, const float* x_in,
                         const float* d0_edge, float* h_out, float* x_out, void* workspace, size_t workspace_bytes,
                         void* stream);
int gb_den_egnn_forward(const gb_net* net, const gb_graph* g, const float* h_in, const float* x_in, float* h_out,
                        float* x_out, void* workspace, size_t workspace_bytes, void* stream);
int gb_pred_layer_forward(const gb_net* net, const gb_graph* g, int layer, const float* h_in, const float* x_in,
                          const float* a_edge, float* h_out, float* x_out, void* workspace, size_t workspace_bytes,
                          void* stream);
int gb_pred_egnn_forward(const gb_net* net, const gb_graph* g, const float* h_in, const float* x_in, const float* a_edge,
                         float* h_out, float* x_out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- training step (BASELINE config 5): unfused ops of the EDM denoising loss and its backward.
 * Replaces torch autograd over EnVariationalDiffusion.forward (edm/equivariant_diffusion/en_diffusion.py:777-797,
 * 644-775) -> EGNN_dynamics._forward (edm/egnn/models.py:76-152) -> EquivariantBlock (edm/egnn/egnn_new.py:42-235) as
 * driven by train_edm.compute_loss (train_edm.py:36-49).  All pointers are fp32 device memory, row-major.
 * gb_gemm        mode 0: C[M,N] = A[M,K] B[N,K]^T (+bias[N])   1: C = A[M,K] B[K,N]   2: C = A[K,M]^T B[K,N];
 *                accumulate != 0 adds into C.  (nn.Linear forward / dgrad / wgrad)
 * gb_colsum      out[k] (+)= sum_m w[m] X[m][k]   (w NULL: 1)      bias and weight-column gradients
 * gb_rowdot      out[m] = bias[0] + X[m,:] . v (bias: device scalar or NULL)                       att_mlp logit / coord_mlp last layer / dL/dr
 * gb_silu_*      SiLU and its backward;  gb_outer_dsilu: G[m][k] = s[m] v[k] SiLU'(pre[m][k])
 * gb_edge_pre    pre[e] = Pa[row_e] + Pb[col_e] + wr r_e + wd d0_e, act = SiLU(pre) if non-NULL (first edge Linear,
 *                factorised; egnn_new.py:43-47)
 * gb_rowcol_reduce  out_row[i] = scale * sum_{e: row_e = i} G[e], out_col[j] = scale * sum_{e: col_e = j} G[e]
 *                   (unsorted_segment_sum, egnn_new.py:403-419, and the backward of gb_edge_pre); either may be NULL
 * gb_gather_rows out[e] = scale * X[row_e]                          backward of the row segment sum
 * gb_gate_*      attention gate ef = m * sigmoid(logit) (egnn_new.py:49-56) and its backward (coef = dL/dlogit)
 * gb_geom_*      coord2diff (egnn_new.py:394-400) and its backward into the coordinates (g_d_scratch [n_edges,3])
 * gb_coord_*     x' = (x + sum_row u * tanh(phi) * range / normf) * mask (egnn_new.py:122-155) and its backward
 * gb_resmask     out = (a + b) * mask[row]  (b NULL: a * mask)
 * gb_den_finish_bwd  backward of the EGNN_dynamics tail (models.py:116-152): g_xfin [n,3], g_h3 [n,F+1]
 * gb_train_loss  loss[b] of compute_loss(t0_always = False), training mode, loss_type 'l2', include_charges = False,
 *                and g_net = d loss[b] / d net_out.  gamma_t [B], t_int [B] (as float). */
 /* gb_linear: out[M,N] = epi([A1 | A2] op(W) + bias) on the tcgen05 (3xTF32) node-Linear kernel of the sampler; the weight
  * image is re-packed into `wimg` (gb_linear_scratch_bytes) on every call.  transpose_w 0: op(W)[k][n] = W[n*ldw+k]
  * (nn.Linear forward), 1: W[k*ldw+n] (dgrad).  epi 0: none; 1: SiLU, pre-activation also stored to out2 if non-NULL;
  * 2: (res + y) * mask[row]; 3: y * SiLU'(aux); 4: y + res.  N <= 256; N, K1, K2, lda multiples of 4. */
size_t gb_linear_scratch_bytes(int N, int K1, int K2);
int gb_linear(int M, int N, int K1, int K2, const float* A1, int lda1, const float* A2, int lda2, const float* W, int ldw,
              int transpose_w, const float* bias, int epi, float* out, float* out2, const float* res_or_aux,
              const float* mask, void* wimg, size_t wimg_bytes, void* stream);
 /* gb_wgrad: C[M,N] (+)= G[K,M]^T X[K,N] on tcgen05 (3xTF32, both operands read row-major = MN-major, reduction split
  * across CTAs).  With `scratch` (gb_wgrad_scratch_bytes) the per-CTA slabs are summed by a second kernel in a fixed order
  * (bit-reproducible); without it they meet in C through fp32 atomics.  colsum (optional, scratch path, N < 256): also
  * returns colsum[m] = sum_k G[k][m] -- the bias gradient of the same Linear -- through a column of ones appended to X.
  * M, N <= 256 and multiples of 4; ldg, ldx multiples of 4; 16-byte aligned G, X. */
size_t gb_wgrad_scratch_bytes(int M, int N);
int gb_wgrad(int K, int M, int N, const float* G, int ldg, const float* X, int ldx, float* C, int ldc, int accumulate,
             float* colsum, void* scratch, size_t scratch_bytes, void* stream);
int gb_gemm(int mode, int M, int N, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
            const float* bias, int accumulate, void* stream);
int gb_colsum(const float* X, int ld, int M, int N, const float* w, float* out, int accumulate, void* stream);
int gb_rowdot(const float* X, int ld, int M, int N, const float* v, const float* bias, float* out, void* stream);
int gb_silu_fwd(const float* x, float* y, size_t n, void* stream);
int gb_silu_bwd(const float* x, const float* gy, float* gx, size_t n, void* stream);
int gb_outer_dsilu(const float* s_row, const float* v, const float* pre, float* G, int M, int N, void* stream);
int gb_edge_pre(const gb_graph* g, const float* Pa, const float* Pb, const float* r, const float* d0, const float* wr,
                const float* wd, int H, float* pre, float* act, void* stream);
int gb_rowcol_reduce(const gb_graph* g, const float* G, int H, float scale, float* out_row, float* out_col, void* stream);
int gb_gather_rows(const gb_graph* g, const float* X, int H, float scale, float* out, void* stream);
int gb_gate_fwd(const float* m, const float* logit, int E, int H, float* ef, float* gate, void* stream);
int gb_gate_bwd(const float* m, const float* gate, const float* wa, const float* g_ef, int E, int H, float* g_m,
                float* coef, void* stream);
int gb_geom_fwd(const gb_graph* g, const float* x, float norm_constant, float* r, float* u, void* stream);
int gb_geom_bwd(const gb_graph* g, const float* x, float norm_constant, const float* g_r, const float* g_u,
                float* g_d_scratch, float* g_x, void* stream);
int gb_coord_fwd(const gb_graph* g, const float* x, const float* u, const float* phi, float range, int use_tanh,
                 float normf, float* x_out, float* tau, void* stream);
int gb_coord_bwd(const gb_graph* g, const float* u, const float* tau, const float* g_xout, float range, int use_tanh,
                 float normf, float* g_phi, float* g_u, float* g_x, void* stream);
int gb_resmask(const float* a, const float* b, const float* mask, int M, int N, float* out, void* stream);
int gb_den_finish_bwd(const float* g_eps, const float* mask, int B, int N, int F, float* g_xfin, float* g_h3,
                      void* stream);
 /* gb_den_finish_fwd: eps [B,N,3+F] from x_fin/x_in [n,3], h3 [n,F+1].  gb_make_zt: normalize (en_diffusion.py:384-404) and
  * z_t = alpha_t xh + sigma_t eps (:661-685); gamma [T+1] schedule table, writes xh, zt [B,N,3+F] and gamma_t [B]. */
int gb_den_finish_fwd(const float* x_fin, const float* x_in, const float* h3, const float* mask, int B, int N, int F,
                      float* eps, void* stream);
int gb_make_zt(const float* x, const float* h, const float* mask, const float* eps, const float* gamma,
               const float* t_int, float norm_x, float norm_h, float bias_h, int B, int N, int F, float* xh, float* zt,
               float* gamma_t, void* stream);
/* gb_vlb_loss: -log p(x,h) estimate of forward() in eval mode, compute_loss(t0_always = True) (en_diffusion.py:644-804):
 * net_t / eps_t at (z_t, t_int in 1..T), net_0 / eps_0 / z0 at t = 0; gamma [T+1] schedule table; loss [B], error_out [B] or NULL. */
int gb_vlb_loss(const float* net_t, const float* eps_t, const float* net_0, const float* eps_0, const float* z0,
                const float* xh, const float* mask, const float* t_int, const float* gamma, int T, float norm_x,
                float norm_h, float bias_h, int B, int N, int F, float* loss, float* error_out, void* stream);
/* gb_pool_mean[_bwd]: pred [B,C] = mean over the N padded nodes of h [B*N,C] (edm/egnn_predictor/models.py:456-457). */
int gb_pool_mean(const float* h, int B, int N, int C, float* pred, void* stream);
int gb_pool_mean_bwd(const float* g_pred, int B, int N, int C, float* g_h, void* stream);
int gb_train_loss(const float* net, const float* eps, const float* zt, const float* xh, const float* mask,
                  const float* t_int, const float* gamma_t, float gamma_T, float norm_h, float bias_h, int B, int N,
                  int F, float* loss, float* g_net, void* stream);

/* ---- fused optimizer step of the training loops (replaces edm/utils.py:51-70 gradient_clipping + Queue :31-48 and the
 * torch.optim.AdamW(amsgrad=True) step of train_edm.py:22-24, 71-82) over ONE flat fp32 bucket of n parameters, no host sync:
 * global gradient norm (fixed-order double sums), max_norm = 1.5 mean + 2 std of the last `window` applied norms, clip
 * coefficient min(1, max_norm / (norm + 1e-6)), AdamW-amsgrad update in torch's single-tensor op order.
 * state: gb_adamw_state_doubles(window) device doubles, zero-initialised by the caller except the window entries it seeds
 * ([4] = length, [8..] = values; the reference seeds one entry of 3000); after a call [0] = step, [1] = ||g||, [2] = max_norm,
 * [3] = clip coefficient.  scratch: gb_adamw_scratch_doubles() device doubles.  clip = 0: plain AdamW-amsgrad. */
size_t gb_adamw_state_doubles(int window);
size_t gb_adamw_scratch_doubles(void);
int gb_adamw_amsgrad_clip(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, float* max_exp_avg_sq, size_t n,
                          float lr, float beta1, float beta2, float eps, float weight_decay, double* state, int window, int clip,
                          double* scratch, void* stream);

/* ---- geometric validity of generated ring graphs -----------------------------------------------
 * gb_check_stability replaces the per-molecule Python of check_stability (analyze/analyze.py:50-100): positions2adj
 * (utils/helpers.py:172-196), minimum-distance test, connectivity (networkx), find_triplets_quads / angel3 / angel4
 * (analyze.py:234-318) and check_angels3 / check_angels4 (:19-47), one thread per molecule.
 *   x [B,N,3] fp32, ring_type [B,N] int32, node_mask [B,N] fp32 (cat[m,m] layout when orientation_type >= 0)
 *   pair_lo/pair_hi [n_types^2]: lo*(1-tol), hi*(1+tol) of the ring-pair distance table (+inf / -inf where absent)
 *   a3_lo/a3_hi [n_types*4], a3_cnt [n_types] (-1: symbol absent): angle ranges per centre ring; a4_hi = a4["180"]*(1-tol),
 *   a4_lo = a4["0"]*(1+tol); check_a4 = 1 for cata.
 *   flags [B][8] uint8: orientation_nodes, dist_stable, connected, angels3, angels4, all-of-the-five, error bits
 *   (1 no ring nodes, 2 more than 16 rings, 4 centre symbol missing from the angle table), number of rings.
 * gb_positions2adj: dist, adj [B,N,N] fp32 exactly as positions2adj returns them (no mask argument there either). */
int gb_check_stability(const float* x, const int* ring_type, const float* node_mask, int B, int N, int n_types,
                       int orientation_type, const float* pair_lo, const float* pair_hi, float min_dist,
                       const float* a3_lo, const float* a3_hi, const int* a3_cnt, float a4_hi, float a4_lo, int check_a4,
                       unsigned char* flags, void* stream);
int gb_positions2adj(const float* x, const int* ring_type, int B, int N, int n_types, const float* pair_lo,
                     const float* pair_hi, float* dist, float* adj, void* stream);

/* ---- measurement aid (bench.py roofline leg): re-launch ONE kernel `repeats` times on the workspace left by the
 * last forward / input-gradient call.  which: 0 denoiser GCL edge, 1 denoiser EquivariantUpdate edge (denoiser
 * workspace); 2 predictor edge forward, 3 predictor edge backward, 4 node-MLP Linear (predictor grad workspace). */
int gb_profile_kernel(const gb_net* net, const gb_graph* g, int which, int layer, void* workspace,
                      size_t workspace_bytes, int repeats, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* GAUDI_B200_H */
