/*
 * hps_b200.h - C ABI of the B200-native render-and-optimise kernels for Hierarchical Point-SLAM.
 *
 * The reference (guo-han/Hierarchical-Point-SLAM) is pure Python/PyTorch + faiss-gpu: it has no FFI,
 * plugin or operator registry for this path.  The seam it does have is its Python call signatures
 * (SURVEY.md section 8b); each entry point below names the reference call site whose arithmetic it
 * replaces.  The Python host side (package hierarchical-point-slam_b200) binds these with ctypes and
 * re-exposes the reference signatures; INTEGRATION.md shows the binding a reference maintainer adds.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter comment says "host";
 *   - all tensors are contiguous row-major fp32 / int32 / uint8 as documented; sizes are explicit;
 *   - `stream` is a cudaStream_t passed as void*; all work is stream-ordered, no host sync inside
 *     unless stated;
 *   - return 0 on success, negative on error; hps_last_error() gives the message (thread-local);
 *   - no hidden allocation except inside an hps_grid_t handle (grow-only, owned by the handle).
 */
#ifndef HPS_B200_H
#define HPS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HPS_K 8          /* nn_num (configs/point_slam.yaml:189) - compiled in                    */
#define HPS_C 32         /* c_dim  (configs/point_slam.yaml:15)  - compiled in                    */
#define HPS_MLP_GEO 0    /* MLP_geometry: E=93 (sin), hidden 32, ReLU, 1 output  (decoder.py:108)  */
#define HPS_MLP_COL 1    /* MLP_color: E=40 (sin|cos), hidden 128, Softplus(100), 3 outputs (:399) */
#define HPS_MLP_NPARAM 23 /* [B, (W_i, b_i, Wc_i, bc_i) i=0..4, Wo, bo]                            */

typedef struct hps_grid hps_grid_t;

const char* hps_last_error(void);
int hps_version(void);

/* ------------------------------------------------------------------------------------------------
 * Neighbour index.  Replaces faiss.IndexIVFFlat train/add/search (neural_point.py:211-214,322-327,393)
 * with an exact search: hashed voxel grid built on device + warp-per-query top-8.
 * ---------------------------------------------------------------------------------------------- */
int hps_grid_create(hps_grid_t** out);
int hps_grid_destroy(hps_grid_t* g);
/* (Re)build over n points pos[n][3].  cell_size > 0.  May cudaMalloc (grow-only).  Stream-ordered.  */
int hps_grid_build(hps_grid_t* g, const float* pos, int64_t n, float cell_size, void* stream);
/* index.add(x) on a built index (neural_point.py:327): hashes ONLY points [n_old, n_new) of pos[n_new][3] (the first
 * n_old rows are the points the grid already holds, unchanged) into the existing structure - cells are built with room to
 * grow, new cells take ranges behind the old ones - so that an insertion costs O(new points), not O(cloud).  Keeps every
 * device address and the hash mask (kernels captured in CUDA graphs stay valid).  Synchronises `stream` once to hand back
 * *status (host int): 0 = done; otherwise the index is unusable until hps_grid_build is called again: bit 0 a new point
 * lies outside the coordinate range of the grid's origin, bit 2 / bit 3 the point arrays / the hash table are full (bit 3
 * is found before anything is touched).  A cell that outgrows its range moves to a larger one inside the point array.   */
int hps_grid_append(hps_grid_t* g, const float* pos, int64_t n_old, int64_t n_new, int* status, void* stream);
/* Host-visible facts; synchronises `stream` to read the device status word.
 * status bit0: a point fell outside the 21-bit cell range (cell_size too small for the extent).     */
int hps_grid_info(const hps_grid_t* g, int64_t* n_points, int64_t* table_size, int64_t* n_cells,
                  float* cell_size, int* status, void* stream);
/* Exact radius-limited 8-NN of q[nq][3].
 *   t_up[nq], t_dn[nq]: fp32 thresholds equivalent to the reference's float64 compares
 *     (D < r^2  <=>  D < t_up, neural_point.py:412;  D > r^2  <=>  D > t_dn, decoder.py:236).
 *   I[nq][8]   original point indices, ascending (D, index); unused slots -1
 *   D[nq][8]   fl(fl(dx*dx+dy*dy)+dz*dz) fp32 without FMA; unused slots FLT_MAX
 *   count[nq]  number of slots with D < t_up   (== reference neighbor_num)
 *   cand[nq]   optional (may be NULL): candidates examined for the query's group (for the roofline figure)
 *   group      1, or 5: consecutive queries [5g, 5g+5) - the samples of one ray - are answered by one warp that
 *              shares the voxel probes and candidate loads among them (results are independent of `group`)
 * Slots j < count are bit-identical to the reference's exhaustive search.                           */
int hps_knn_query(const hps_grid_t* g, const float* q, const float* t_up, const float* t_dn,
                  int64_t nq, int32_t* I, float* D, int32_t* count, int32_t* cand, int group, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Inverse-distance interpolation (decoder.py:195-272 and :491-559; K2 in SURVEY 2.2).
 * One neighbour list feeds up to two feature tables (geometry + colour).
 *   cloud_pos != NULL selects the tracker variant: D is recomputed from cloud_pos[I]-p (decoder.py:
 *   209-218) and the backward also produces dp.
 *   w[P][8]: normalised weights (saved for the backward); has[P]: count >= min_nn.
 * ---------------------------------------------------------------------------------------------- */
int hps_interp_fwd(const float* p, const int32_t* I, const float* D, const float* t_dn,
                   const int32_t* count, const float* cloud_pos,
                   const float* feats_a, const float* feats_b,
                   const float* fill_a, const float* fill_b, int64_t P, int min_nn,
                   float* w, float* c_a, float* c_b, uint8_t* has, void* stream);
/* dfeats_* are ACCUMULATED into (atomic adds); dp (P,3) is written (tracker variant only).
 * row_map == NULL: dfeats_* are (N,32), row = table row.  row_map (N) int32: dfeats_* are compact (M,32) buffers
 * of the live (frustum-selected, Mapper.py:512-536) rows, row_map[row] = slot or -1 for frozen rows.          */
int hps_interp_bwd(const float* p, const int32_t* I, const float* w, const uint8_t* has,
                   const float* cloud_pos, const float* feats_a, const float* feats_b,
                   const float* c_a, const float* c_b,
                   const float* dc_a, const float* dc_b, int64_t P,
                   float* dfeats_a, float* dfeats_b, float* dp, const int32_t* row_map, void* stream);

/* table[idx[r]] = rows[r] (32 floats per row) for up to two tables in ONE launch: the per-iteration write-back of the
 * optimised leaves into the full feature tables, `self.npc.update_geo_feats(geo_feats, indices=indices)` +
 * `update_col_feats` - Mapper.py:615-630 (neural_point.py:149-181).  idx_*: int64 table rows, no duplicates;
 * either triple may be empty (n == 0, pointers NULL).                                                            */
int hps_rows_store(float* table_a, const float* rows_a, const int64_t* idx_a, int64_t n_a,
                   float* table_b, const float* rows_b, const int64_t* idx_b, int64_t n_b, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Decoder MLPs (decoder.py:274-342 MLP_geometry.forward, :561-618 MLP_color.forward; K3).
 *   params: HOST array of HPS_MLP_NPARAM device pointers in the reference's own layouts
 *           (Linear.weight = (out,in) row-major; embedder._B = (3,E)).
 *   affine: colour only, may be NULL: 12 floats = mlp_exposure(exposure_feat); when given the output
 *           is sigmoid(out @ A[:9].view(3,3) + A[9:]) (decoder.py:606-612), else raw rgb (:613-614).
 *   act:    NULL (inference) or workspace of hps_mlp_act_floats(kind) floats per point, written by
 *           fwd and consumed by bwd.
 * ---------------------------------------------------------------------------------------------- */
int64_t hps_mlp_act_floats(int kind);     /* activation workspace floats per point (P rounded up to 64) */
int64_t hps_mlp_packed_floats(int kind);  /* size of the transposed-weight workspace in floats          */
/* Transpose the Linear weights into the k-major layout the forward tiles read (one tiny launch; call
 * after every optimiser step, before hps_mlp_fwd).                                                  */
int hps_mlp_pack(int kind, const float* const* params, float* packed, void* stream);
int hps_mlp_fwd(int kind, const float* const* params, const float* packed, const float* p, const float* c,
                const float* affine, int64_t P, float* out, float* act, void* stream);
/* Tensor-core edition of the forward (tcgen05.mma + TMEM, fp16 hi + fp16 lo split operands, three MMAs per
 * product, fp32 accumulation): same inputs, outputs and `act` layout as hps_mlp_fwd, so hps_mlp_bwd consumes
 * its workspace unchanged.  `packed` here is the swizzled 16-bit weight-slab image written by
 * hps_mlp_tc_pack (hps_mlp_tc_packed_bytes(kind) bytes, 16-byte aligned).                             */
int64_t hps_mlp_tc_packed_bytes(int kind);
int hps_mlp_tc_pack(int kind, const float* const* params, void* packed, void* stream);
/* `ws` (may be NULL): workspace of hps_mlp_tc_ws_bytes(kind, P) bytes for the tensor-core backward (fp16 hi/lo
 * slab images of every layer input, the pre-activations and the raw output); `act` (may be NULL) is the SIMT
 * workspace of hps_mlp_bwd.                                                                            */
int64_t hps_mlp_tc_ws_bytes(int kind, int64_t P);
/* max_ctas (hps_mlp_tc_fwd / hps_mlp_tc_bwd): 0 = one persistent CTA per SM; n > 0 caps the grid at n CTAs so that
 * two decoders launched on two streams share the 148 SMs in proportion to their per-tile cost instead of queueing
 * behind each other (the kernels hold an SM exclusively: their shared memory does not fit twice).              */
int hps_mlp_tc_fwd(int kind, const float* const* params, const void* packed, const float* p, const float* c,
                   const float* affine, int64_t P, float* out, float* act, void* ws, int max_ctas, void* stream);
/* Tensor-core backward: dgrad chain kernel + streaming wgrad kernel (MN-major operands, accumulators in TMEM)
 * + a small finalisation kernel for dWc / dbc.
 * Same outputs as hps_mlp_bwd except that dp (P,3) must be ZERO-initialised by the caller (it is accumulated).
 * `packed` = hps_mlp_tc_bwd_pack image (transposed weight slabs, hps_mlp_tc_bwd_packed_bytes(kind) bytes).    */
int64_t hps_mlp_tc_bwd_packed_bytes(int kind);
int64_t hps_mlp_tc_bwd_absmax_offset(int kind);   /* byte offset of the max|dout| word inside `packed` */
int hps_mlp_tc_bwd_pack(int kind, const float* const* params, void* packed, void* stream);
/*
 * The backward runs on gradients scaled by a power of two derived from max|dout|; that maximum is reduced into the
 * uint32 word (float bits) at packed + hps_mlp_tc_bwd_absmax_offset(kind), zeroed by hps_mlp_tc_bwd_pack (the
 * image also carries Wo Wc_4 and a scratch area for the (H x 32) sums behind dWc / dbc, see csrc/mlp_tc_bwd.cu).
 * absmax_done != 0: the caller has already reduced max|dout| into that word (hps_composite_bwd can do it while
 * it writes dout), so the extra reduction kernel is skipped.  wgrad_stream: stream for the weight-gradient kernel
 * (ordered after the dgrad kernel through an event); dc / dp / daffine are complete on `stream`, the parameter
 * gradients on `wgrad_stream`.  gscale: device scalar g - `dout` then is d loss / d raw for an upstream gradient of 1
 * (written in the forward pass by hps_composite_mapper_loss, max |dout| folded: absmax_done must be set) and every
 * output is g times what it would be for that dout; only without the exposure tail (affine == NULL).            */
int hps_mlp_tc_bwd(int kind, const float* const* params, const void* packed, void* ws, const float* p,
                   const float* affine, const float* dout, int64_t P, float* const* dparams, float* dc, float* dp,
                   float* daffine, int absmax_done, int max_ctas, void* wgrad_stream /* may be NULL = stream */,
                   const float* gscale /* may be NULL = 1 */, void* stream);
/* Single-CTA tcgen05 self-test: D[128][N] = A . B^T with fp16 operands staged K-major (major = 0, source
 * [rows][K]) or MN-major (major = 1, source [K][rows]) - pins the shared-memory descriptor encodings.        */
int hps_tc_selftest(const float* A, const float* B, float* D, int N, int K, int a_major, int b_major, void* stream);
/* Diagnostics: install (buf != NULL) or remove (NULL) a device buffer of 256 uint64 in which CTA 0 of the
 * tensor-core decoder kernels stamps %globaltimer (ns) at its phase boundaries: [0,64) forward epilogue warp,
 * [64,128) forward MMA warp, [128,192) dgrad epilogue warp, [192,256) dgrad MMA warp (scripts/tc_trace.py).
 * No counterpart in the reference. */
int hps_tc_trace(void* buf);
/* dparams: HOST array of HPS_MLP_NPARAM device pointers (entries or the array may be NULL = skip);
 * parameter gradients are ACCUMULATED (atomic adds).  dc (P,32) written.  dp (P,3) written if not
 * NULL.  daffine (12) accumulated if not NULL.                                                     */
int hps_mlp_bwd(int kind, const float* const* params, const float* p, const float* c,
                const float* affine, const float* act, const float* dout, int64_t P,
                float* const* dparams, float* dc, float* dp, float* daffine, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Alpha compositing (common.py:385-428 raw2outputs_nerf_color, occupancy=True; Renderer.py:237-239).
 *   rgb/occ are addressed as rgb[i*rgb_stride + c], occ[i*occ_stride] for sample i in [0, R*S):
 *   a packed raw (R,S,4) is (raw, 4, raw+3, 4).  rgb may be NULL (geometry stages: colour = 0).
 *   point_mask may be NULL; where it is 0 the occupancy logit is taken as -100 in the forward and
 *   the gradient is passed straight through.
 * ---------------------------------------------------------------------------------------------- */
int hps_composite_fwd(const float* rgb, int rgb_stride, const float* occ, int occ_stride,
                      const uint8_t* point_mask, const float* z, int64_t R, int S, float coef,
                      float* depth, float* var, float* color, float* weights,
                      int min_valid, uint8_t* valid_ray /* may be NULL: #samples with neighbours >= min_valid */,
                      const float* gt_depth, float* tracker_sum /* may be NULL; else += sum_r |gt - depth| /
                      sqrt(var + 1e-10), the first pass of the tracker loss (hps_tracker_loss_stats) */,
                      int zero_missing /* depth counts as 0 where gt_depth <= 0 (Renderer.py:250-251) */,
                      void* stream);
int hps_composite_bwd(const float* rgb, int rgb_stride, const float* occ, int occ_stride,
                      const uint8_t* point_mask, const float* z, int64_t R, int S, float coef,
                      const float* g_depth, const float* g_color,
                      const float* gscale /* device scalar multiplied into g_depth / g_color; may be NULL = 1 */,
                      float* d_rgb, int d_rgb_stride, float* d_occ, int d_occ_stride,
                      uint32_t* absmax_rgb, uint32_t* absmax_occ /* may be NULL: atomicMax of |d_rgb| / |d_occ| float bits */,
                      void* stream);
/* hps_composite_fwd fused with hps_mapper_loss (Mapper.py:791-856): same depth / var / color / valid_ray, plus
 * loss += sum over rays with gt_depth > 0, valid and a finite depth of |gt_depth - depth|
 *         (+ w_color * |gt_color - sigmoid(color x affine[0:9] + affine[9:12])| when gt_color is given; affine may
 *         be NULL = identity without the sigmoid's affine, see hps_mapper_loss),
 * g_depth (R), g_color (R,3) = d loss / d depth, d color, and d_affine (12) += d loss / d affine.
 * loss and d_affine are ACCUMULATED: zero them first.
 * d_occ (R*S), d_rgb (R*S,3) (may be NULL): additionally the compositing BACKWARD for an upstream gradient of 1 -
 * d loss / d occ, d loss / d rgb, what hps_composite_bwd would write from g_depth / g_color - with max |.| reduced
 * into absmax_occ / absmax_rgb (may be NULL); pass them as `dout` of hps_mlp_tc_bwd with its gscale = the upstream
 * gradient and the separate compositing-backward launch between the decoders' forward and backward is gone.
 * zero / zero_floats (may be NULL / 0; 16-byte aligned, a multiple of 4 floats): a buffer this launch clears on the
 * side - the gradient buffers of the backward that follows (no memset between the loss and the first dgrad kernel). */
int hps_composite_mapper_loss(const float* rgb, int rgb_stride, const float* occ, int occ_stride,
                              const uint8_t* point_mask, const float* z, int64_t R, int S, float coef,
                              int min_valid, const float* gt_depth, const float* gt_color /* may be NULL */,
                              const float* affine /* may be NULL */, const int32_t* frame /* may be NULL */,
                              float w_color, float* depth, float* var,
                              float* color, uint8_t* valid_ray, float* loss, float* g_depth, float* g_color,
                              float* d_affine, float* d_rgb, float* d_occ, uint32_t* absmax_rgb, uint32_t* absmax_occ,
                              float* zero, int64_t zero_floats, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Adam over the live (frustum-selected) rows of a feature table, fused with the write-back into the table
 * (Mapper.py:512-536, 592-603, 615-630, 859, 862-874: clone rows -> torch.optim.Adam -> copy rows back).
 *   param, grad, exp_avg, exp_avg_sq: (M,C) compact, C % 4 == 0;  step: 2 device ints {step count, scratch = 0}
 *   (the kernel uses t = step[0] + 1 and stores it; replayable from a CUDA graph);  table (N,C) + rows (M) int64:
 *   optional, table[rows[r]] = new param[r].  Update rule = torch.optim.Adam without amsgrad / weight decay.
 * ---------------------------------------------------------------------------------------------- */
int hps_adam_rows(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t M, int C, float lr,
                  float beta1, float beta2, float eps, int32_t* step, float* table, const int64_t* rows, void* stream);
/* The same update for up to 24 small dense fp32 tensors in one launch (tracker: pose, exposure latent, exposure MLPs -
 * Tracker.py:383-416 builds a torch.optim.Adam over exactly these).  params / grads / exp_avg / exp_avg_sq: HOST arrays
 * of n_tensors device pointers; numel, lr: HOST arrays (per-tensor size and learning rate = parameter groups);
 * step as above (one counter for the call).  zero_grad != 0: every gradient is cleared after it has been used.   */
int hps_adam_multi(int n_tensors, float* const* params, float* const* grads, float* const* exp_avg,
                   float* const* exp_avg_sq, const int64_t* numel, const float* lr, float beta1, float beta2,
                   float eps, int32_t* step, int zero_grad, void* stream);
/* hps_adam_multi with one step counter PER TENSOR: torch.optim.Adam keeps state[p]['step'] per parameter and skips
 * parameters whose .grad is None, which matters for the mapper's single optimiser over all decoders, feature tables,
 * poses and the exposure latent (Mapper.py:592-603; a stage gives gradients to only some of them, :637-660).
 * steps: HOST array of n_tensors device pointers, each to one int32 counter (distinct);  done_word: one device int32
 * of scratch that is 0 between launches.  Replayable from a CUDA graph.                                           */
int hps_adam_multi_own_steps(int n_tensors, float* const* params, float* const* grads, float* const* exp_avg,
                             float* const* exp_avg_sq, const int64_t* numel, const float* lr, int32_t* const* steps,
                             float beta1, float beta2, float eps, int32_t* done_word, int zero_grad, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Exposure MLP (decoder.py:380-396 MLP_exposure: linear1 -> Softplus(beta=100) -> linear2) on one latent
 * vector: out (n_out = 12) = W2 softplus(W1 x + b1) + b2; W1 (hidden, n_in), W2 (n_out, hidden) row-major as in
 * nn.Linear.  a1 (hidden) receives the pre-activations for the backward.  One launch each way instead of the
 * ~10 library kernels of the op-by-op version.  hps_exposure_bwd WRITES dW1, db1, dW2, db2 (and dx if not NULL).
 * Limits: n_in <= 64, hidden <= 1024, n_out <= 32.
 * ---------------------------------------------------------------------------------------------- */
int hps_exposure_fwd(const float* x, const float* W1, const float* b1, const float* W2, const float* b2, int n_in,
                     int hidden, int n_out, float* a1, float* out, void* stream);
int hps_exposure_bwd(const float* x, const float* W1, const float* W2, const float* a1, const float* dout,
                     const float* gscale /* device scalar multiplied into dout; may be NULL = 1 */, int n_in,
                     int hidden, int n_out, float* dW1, float* db1, float* dW2, float* db2, float* dx, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Depth-guided sampling (Renderer.py:163-229, N_samples = 0) + per-point radius thresholds.
 *   z[r][s] = near*d*(1-t_s) + far*d*t_s, t = linspace(0,1,S); pts = rays_o + rays_d*z (no FMA);
 *   rays with gt_depth == 0 take z_zero[r*z_zero_stride + s]: z_zero_stride 0 = one shared row of S floats
 *   (Renderer.py:207-208), S = one row per ray (the sample_near_pcl mode, Renderer.py:197-206); NULL -> 0.
 *   radius: per-RAY float64 radii (may be NULL -> radius_scalar), repeated per sample into t_up/t_dn.
 * hps_sample_rays_bwd: g_rays_o = sum_s dp, g_rays_d = sum_s dp * z with dp = dp + dp2 + dp3 (tracker pose path).
 * ---------------------------------------------------------------------------------------------- */
int hps_sample_rays(const float* rays_o, const float* rays_d, const float* gt_depth, const double* radius,
                    double radius_scalar, const float* z_zero, int64_t z_zero_stride, int64_t R, int S, float near,
                    float far, float* z, float* pts, float* t_up, float* t_dn, void* stream);
/* hps_sample_rays + hps_knn_query (group 1) in one launch: the search computes its query points itself and writes
 * z / pts / t_up / t_dn for the kernels downstream.  Same outputs, bit for bit.                                    */
int hps_sample_knn(const hps_grid_t* g, const float* rays_o, const float* rays_d, const float* gt_depth,
                   const double* radius, double radius_scalar, const float* z_zero, int64_t z_zero_stride, int64_t R,
                   int S, float near, float far, float* z, float* pts, float* t_up, float* t_dn, int32_t* I, float* D,
                   int32_t* count, int32_t* cand /* may be NULL */, void* stream);
int hps_sample_rays_bwd(const float* dp, const float* dp2 /* may be NULL */, const float* dp3 /* may be NULL */,
                        const float* z, int64_t R, int S, float* g_rays_o, float* g_rays_d, void* stream);

/* Tracker pose helpers (common.py:304-346 get_camera_from_tensor / quad2rotation, :103-119 get_rays_from_uv) as
 * single kernels with analytic backward; arithmetic is rounded like the reference's torch expressions.
 *   hps_camera_from_tensor:      cam7 = [quaternion (w,x,y,z), translation] -> RT (3,4) row-major (q is not normalised)
 *   hps_camera_from_tensor_bwd:  g_RT (3,4) -> g_cam7 (7), written
 *   hps_rays_from_uv:            pixels (pi, pj) (R each, float) -> rays_d = sum(dirs * c2w[:3,:3], -1), rays_o = c2w[:3,3];
 *                                c2w rows have `pitch` floats (4 for (3,4) and (4,4) matrices)
 *   hps_rays_from_uv_bwd:        g_c2w (3,4) += [sum_r g_rays_d (x) dirs | sum_r g_rays_o]  (caller zero-initialises)  */
int hps_camera_from_tensor(const float* cam7, float* RT, void* stream);
int hps_camera_from_tensor_bwd(const float* cam7, const float* g_RT, float* g_cam7, void* stream);
int hps_rays_from_uv(const float* pi, const float* pj, const float* c2w, int pitch, int64_t R, float fx, float fy,
                     float cx, float cy, float* rays_o, float* rays_d, void* stream);
int hps_rays_from_uv_bwd(const float* pi, const float* pj, const float* g_rays_o, const float* g_rays_d, int64_t R,
                         float fx, float fy, float cx, float cy, float* g_c2w, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Losses fused with their gradients (Mapper.py:791-856; Tracker.py:213-231).  `loss` (1 float) and
 * `d_affine` (12 floats) are ACCUMULATED into; g_depth (R) / g_color (R,3) are written.
 *   mapper: color may be NULL (geometry stages); affine may be NULL (no exposure correction); frame (R int32, may be
 *   NULL = one frame): rays of several keyframes in one batch (Mapper.py:805-822) - ray r uses affine[12*frame[r]..]
 *   and accumulates into d_affine[12*frame[r]..] (F x 12 floats each).
 *   tracker: two passes - hps_tracker_loss_stats accumulates sum_r |gt-d|/sqrt(var+1e-10) into sum_tmp
 *   (all-reduce it across ranks when rays are sharded), hps_tracker_loss applies the 10*mean outlier mask
 *   with mean = sum_tmp / count.
 * ---------------------------------------------------------------------------------------------- */
int hps_mapper_loss(const float* depth, const float* color, const uint8_t* valid, const float* gt_depth,
                    const float* gt_color, const float* affine, const int32_t* frame, float w_color, int64_t R,
                    float* loss, float* g_depth, float* g_color, float* d_affine, void* stream);
int hps_tracker_loss_stats(const float* depth, const float* var, const float* gt_depth, int64_t R,
                           float* sum_tmp, void* stream);
int hps_tracker_loss(const float* depth, const float* var, const float* color, const float* gt_depth,
                     const float* gt_color, const float* sum_tmp, double count, float w_color, int64_t R,
                     float* loss, float* g_depth, float* g_color, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Gradient exchange over NVLink peer memory (the data-parallel form of Mapper.py:783-860: the ray batch is
 * sharded over the GPUs of one box, point-feature and decoder gradients are summed over the ranks).  One kernel:
 * cross-rank barrier, reduce-scatter by slice owner (P2P vector loads, or ONE multimem.ld_reduce per vector when the
 * buffer has an NVLS multicast address), broadcast of the sums into every rank's copy (P2P stores / multimem.st),
 * barrier.  In place; every rank ends up with bit-identical sums; capturable in a CUDA graph (binary-semaphore flags).
 *   bufs[world]   HOST array of device pointers: rank p's copy of the symmetric buffer as mapped into THIS process
 *   pads[world]   HOST array of device pointers: rank p's signal pad (hps_peer_allreduce_pad_words() uint32, zeroed
 *                 once before the first call, never touched by anything else)
 *   multicast     the buffer's multicast address, or NULL (P2P loads / stores)
 *   n_floats      multiple of 4; every rank passes the same value
 * Every rank of the group must call it once per exchange, the same number of times (the flags are launch counters): it
 * waits for its peers' flags for at most two seconds - then it gives up and sets pad[ctas * 8] (this rank's pad, bit 1),
 * and the buffer's contents are undefined.                                                                           */
int hps_peer_allreduce_ctas(void);
int64_t hps_peer_allreduce_pad_words(void);
int hps_peer_allreduce(float* const* bufs, uint32_t* const* pads, float* multicast, int rank, int world,
                       int64_t n_floats, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HPS_B200_H */
