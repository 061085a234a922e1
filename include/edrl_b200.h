/*
 * edrl_b200.h -- C-ABI of the B200-native EDRL hot path.
 *
 * Two parts:
 *   Part A  multi-bandwidth Gaussian MMD           reference: code/MMD.py:3-74
 *   Part B  Essence-Point scoring + top-k select   reference: code/fusion_net.py:133-255 (class EPRL)
 *
 * The reference has no FFI layer (it is pure PyTorch; the boundary there is Python
 * name binding, code/fusion_train.py:11 and code/fusion_net.py:817-821), so these
 * entry points are what a ctypes binding inside the reference's MMD.py /
 * fusion_net.py would call -- see INTEGRATION.md.  Each entry point cites the
 * reference lines it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer (sm_100a, CUDA 12.9) unless it says "host";
 *   - every matrix is dense row-major fp32; index tensors are int32 unless stated;
 *   - `stream` is a cudaStream_t passed as void*; work is enqueued, never synchronised;
 *   - the return value is 0 on success, non-zero on error; edrl_last_error() then
 *     returns a thread-local human-readable message (the Python host raises it as
 *     ValueError / RuntimeError, mirroring the torch exceptions of the reference);
 *   - there is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef EDRL_B200_H_
#define EDRL_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define EDRL_ABI_VERSION 1

/* flags for the MMD entry points */
#define EDRL_MMD_TF32        0   /* Gram and G.Z on tcgen05 kind::tf32, operands rounded to TF32        */
#define EDRL_MMD_3XTF32      1   /* hi/lo split operands, 3 MMAs per logical MMA (fp32-level accuracy)    */
#define EDRL_MMD_TF32H       2   /* TF32 Gram; the G.Z product of the fused sweep reads its two operands as
                                    power-of-two-scaled binary16 -- the SAME 11-bit significands as their TF32
                                    roundings (the Z copy is bit-identical in value), half the bytes.  Only
                                    edrl_mmd_forward_grad differs; every other entry point treats it as TF32.  */
#define EDRL_MMD_F16S        4   /* TF32H, and the Gram of the fused sweep also reads a binary16 copy of the TF32-rounded
                                    centred operand, scaled by one power of two for the whole matrix (same 11-bit
                                    significands, exact products, fp32 accumulation): both contractions run at the
                                    kind::f16 rate with half the shared-memory bytes.  Differs from TF32 only where a
                                    value underflows binary16 (|z| < 2^-39 max|z|).  edrl_mmd_forward_grad only.      */

/* slots of the `stats` vector written by edrl_mmd_forward (8 floats) */
#define EDRL_MMD_STAT_M       0  /* signed mean discrepancy  XX + YY - XY - YX          (code/MMD.py:66-69) */
#define EDRL_MMD_STAT_SIGMA0  1  /* smallest bandwidth sigma_0                          (code/MMD.py:31-34) */
#define EDRL_MMD_STAT_D       2  /* d M / d sigma_0 (the undetached-bandwidth term)        */
#define EDRL_MMD_STAT_C       3  /* D / ((n^2 - n) * mul^(num//2))                                         */
#define EDRL_MMD_STAT_SUMR    4  /* sum_i |z_i - mean|^2                                                    */
#define EDRL_MMD_NUM_STATS    8

int         edrl_abi_version(void);
const char *edrl_last_error(void);
/* number of kernels this library has launched since load (bench.py reports it as gpu_launches) */
uint64_t    edrl_launch_count(void);
/* Name the CUDA device the calling host thread's next compute calls run on.  The library links its own
 * static CUDA runtime, so a host that selects devices through another runtime instance (PyTorch) calls this
 * before the compute entry points.  Every entry point that takes a stream binds the thread to that device
 * for the duration of the call and restores the caller's current device before it returns (the two runtimes
 * share the primary contexts: a lasting cudaSetDevice would move the caller's later allocations). */
int         edrl_set_device(int device);

/* ------------------------------------------------------------------------------------------
 * Part A -- MK_MMD                                                        code/MMD.py:46-74
 * ---------------------------------------------------------------------------------------- */

/* Bytes of scratch the MMD calls need for this shape.  The same buffer must be passed to
 * forward and backward of one loss evaluation (it holds the centred TF32 operands, the row
 * norms and the block weights that backward re-uses). */
size_t edrl_mmd_workspace_bytes(int n_s, int n_t, int d, int flags);

/* gaussian_kernel + MK_MMD forward, fused: the n x n kernel matrix is never written.
 *   X [n_s, d], Y [n_t, d]                                        code/MMD.py:16-21
 *   loss  -> 1 float   |XX + YY - XY - YX|                        code/MMD.py:60-72
 *   stats -> EDRL_MMD_NUM_STATS floats (see above), consumed by edrl_mmd_backward
 * tile_rank / tile_world shard the upper-triangular tile list across the ranks of a
 * row-block-sharded evaluation (single GPU: 0 / 1).  With tile_world > 1 the call only
 * accumulates this rank's partial sums into `partial` (2 doubles: sum a_i a_j K_ij,
 * sum a_i a_j L_ij Q_ij); all-reduce them and call edrl_mmd_finalize. */
int edrl_mmd_forward(const float *X, const float *Y, int n_s, int n_t, int d,
                     float kernel_mul, int kernel_num, int flags,
                     int tile_rank, int tile_world,
                     float *loss, float *stats, double *partial,
                     void *workspace, size_t workspace_bytes, void *stream);

/* loss / stats from all-reduced partial sums (sharded evaluation only). */
int edrl_mmd_finalize(const double *partial, int n_s, int n_t, float kernel_mul, int kernel_num,
                      float *loss, float *stats, void *workspace, size_t workspace_bytes, void *stream);

/* Backward of MK_MMD w.r.t. the rows [row_begin, row_begin + row_count) of Z = [X; Y]
 * (what autograd does to code/MMD.py:16-72, bandwidth NOT detached, clamp mask L >= 0):
 *   dZ[i, :] = grad_out * sign(M) * 4 * ( rowsum(G)_i z_i - (G Z)_i )
 * Kernel tiles are recomputed from the operands kept in `workspace`; nothing n x n is stored.
 *   grad_out : 1 float (device), the incoming gradient of the scalar loss
 *   dZ       : [row_count, d] */
int edrl_mmd_backward(int n_s, int n_t, int d, float kernel_mul, int kernel_num, int flags,
                      const float *stats, const float *grad_out,
                      int row_begin, int row_count, float *dZ,
                      void *workspace, size_t workspace_bytes, void *stream);

/* Fused training pass (every precision mode): forward block sums AND the bandwidth-independent part of the gradient in one
 * sweep over the Gram tiles of rows [row_begin, row_begin + row_count) -- a training step then visits every tile
 * once instead of 1.5 times (edrl_mmd_forward + edrl_mmd_backward).  Same math as code/MMD.py:16-72 + autograd:
 *   U[i, :] = -(G' Z)_i,  G'_ij = -a_i a_j Q_ij / sigma_0;  rowsum(G')_i goes to the workspace (apply_grad adds
 *     rowsum(G')_i z_i)
 *     (the clamp mask [L_raw >= 0] of code/MMD.py:27 is not applied to G': a pair with L_raw < 0 is a numerical
 *      duplicate, z_i = z_j up to rounding, and its term G'_ij (z_i - z_j) vanishes whatever G'_ij is)
 *                                          -> U [edrl_mmd_grad_slabs(...), rows, d], partial sums over column slabs
 *   partial sums (sum a_i a_j K_ij, sum a_i a_j L_ij Q_ij over the rows of this call) are ADDED into the
 *   workspace accumulators; finalize != 0 (the call covers all rows): loss / stats are written as by
 *   edrl_mmd_forward; finalize == 0 (row-block sharded): they are returned in `partial` for the all-reduce.
 * edrl_mmd_apply_grad then yields dZ = grad_out * sign(M) * 4 * (U + (rowsum(G')_i + c n) z_i - c sum_j z_j): the
 * uniform bandwidth term c of G in closed form (sum_j c (z_i - z_j) = c n z_i - c sum_j z_j, with the column sums of
 * the rounded centred operand).  dZ may alias U.  An optional second row range
 * (row_count2 > 0, after the first) lets a rank of a sharded evaluation cover its source rows and its target rows
 * in one launch; U / dZ hold range 1's rows followed by range 2's. */
/* Number of partial output slabs the fused pass may write for row_count + row_count2 output rows: the row panels that
 * do not fill a whole wave of SM pairs have their column sweep split into that many slabs (one partial output each).
 * U must hold slabs * (row_count + row_count2) * d floats; apply_grad sums the slabs a row was split into. */
int edrl_mmd_grad_slabs(int n_s, int n_t, int d, int flags, int row_count, int row_count2);
/* The work list behind it, for inspection and host-side tests (no device work; sms <= 0: the current device's SM count,
 * 148 without a device): plan[10] = { row panels, virtual panels (x feature passes), virtual panels swept whole, column
 * slabs of each later virtual panel, work items, persistent clusters launched, 256-column groups per sweep, padded
 * feature count, 1 if clusters are 4 CTAs (d_pad > 768: two MMA pairs share the S phase) else 0 (2 CTAs), feature
 * columns per pass (1024 / 512) }.  Item i < plan[2] sweeps virtual panel i over all groups; item i >= plan[2] sweeps
 * virtual panel plan[2] + (i - plan[2]) / plan[3], slab s = (i - plan[2]) % plan[3], groups [s G / plan[3], (s+1) G / plan[3]). */
int edrl_mmd_sweep_plan(int n_s, int n_t, int d, int flags, int row_count, int row_count2, int sms, int *plan);
int edrl_mmd_forward_grad(const float *X, const float *Y, int n_s, int n_t, int d, float kernel_mul, int kernel_num,
                          int flags, int row_begin, int row_count, int row_begin2, int row_count2, int finalize,
                          float *loss, float *stats, double *partial, float *U, void *workspace,
                          size_t workspace_bytes, void *stream);
int edrl_mmd_apply_grad(int n_s, int n_t, int d, int flags, const float *stats, const float *grad_out,
                        const float *U, int row_begin, int row_count, int row_begin2, int row_count2, float *dZ,
                        void *workspace, size_t workspace_bytes, void *stream);

/* gaussian_kernel(source, target) materialised (code/MMD.py:3-44) -- API completeness and
 * parity tests only; K is [n, n] with n = n_s + n_t.  Uses the same tcgen05 Gram tiles. */
int edrl_mmd_kernel_matrix(const float *X, const float *Y, int n_s, int n_t, int d,
                           float kernel_mul, int kernel_num, int flags, float *K,
                           void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Part B -- Essence-Point scoring and selection                code/fusion_net.py:133-255
 * ---------------------------------------------------------------------------------------- */

/* Token statistics (F.normalize(z, dim=1) hoisted with the token mean, fusion_net.py:149,224-225):
 *   z [B,T,F] -> zbar[b,f] = sum_t z[b,t,f] / (T * max(|z[b,:,f]|_2, 1e-12)),
 *   also colsum[b,f] = sum_t z and colnorm[b,f] = |z[b,:,f]|_2 (kept for backward). */
int edrl_token_stats_fwd(const float *z, int B, int T, int F,
                         float *zbar, float *colsum, float *colnorm, void *stream);
/* dz[b,t,f] from dzbar[b,f]. */
int edrl_token_stats_bwd(const float *z, const float *colsum, const float *colnorm, const float *dzbar,
                         int B, int T, int F, float *dz, void *stream);
/* eval only (fusion_net.py:163): zmean[b,t] = mean_f z[b,t,f] / max(colnorm[b,f], 1e-12). */
int edrl_token_featmean(const float *z, const float *colnorm, int B, int T, int F, float *zmean, void *stream);

/* Proxy sampling + sample-dim normalisation (fusion_net.py:143-146,150):
 *   z_p[c,s,f] = mu[c,f] + sigma[c,f] * eps[c,s,f];  z_pn = z_p / max(|z_p[c,:,f]|_2, 1e-12)
 *   pnorm[c,f] = |z_p[c,:,f]|_2 (kept for backward). */
int edrl_proxy_normalize_fwd(const float *mu, const float *sigma, const float *eps, int C, int S, int F,
                             float *z_pn, float *pnorm, void *stream);
/* dmu[c,f], dsigma[c,f] from dz_pn[c,s,f]. */
int edrl_proxy_normalize_bwd(const float *mu, const float *sigma, const float *eps, const float *pnorm,
                             const float *dz_pn, int C, int S, int F, float *dmu, float *dsigma, void *stream);

/* Scores (fusion_net.py:221-225): att[b, r] = sum_f zbar[b,f] * z_pn[r,f], r = c*S + s in [0, R). */
int edrl_score_fwd(const float *zbar, const float *z_pn, int B, int R, int F, float *att, void *stream);
/* dzbar[b,f] = sum_r datt[b,r] z_pn[r,f];  dz_pn[r,f] = sum_b datt[b,r] zbar[b,f]. */
int edrl_score_bwd(const float *datt, const float *zbar, const float *z_pn, int B, int R, int F,
                   float *dzbar, float *dz_pn, void *stream);

/* Generic row-wise top-k (torch.topk(x, k, dim=1), fusion_net.py:236-238): the k largest of each
 * row, ties lowest-index-first.  sorted != 0: descending by value like torch.topk(sorted=True);
 * sorted == 0: the same set in unspecified order (the warp kernel emits ascending index order; cheaper, and
 * enough for consumers that reduce over or gather the selection, which is all the reference does with it).  x is [R, W] with row stride ld (floats).
 * vals [R,k], idx [R,k] (int32).  k > W is an error, like torch. */
int edrl_topk_rows(const float *x, int R, int W, int ld, int k, int sorted, float *vals, int32_t *idx, void *stream);

/* Label-addressed select (fusion_net.py:227-238) without masked_select:
 *   positives of row b = att[b, y_b, :]           -> pos_val/pos_idx [B,k]
 *   negatives of row b = concat_{c != y_b} att[b,c,:] (class-major) -> neg_val/neg_idx [B,k]
 * y is int64 [B]; labels outside {0,1} are rejected by the host (proxies_dict, fusion_net.py:101). */
int edrl_select_topk_fwd(const float *att, const int64_t *y, int B, int C, int S, int k, int sorted,
                         float *pos_val, int32_t *pos_idx, float *neg_val, int32_t *neg_idx, void *stream);

/* proxy_loss = mean_b exp(-mean(pos_val_b) + mean(neg_val_b))  (fusion_net.py:240-243);
 * rowexp[b] keeps the per-row exponential for backward. */
int edrl_proxy_loss_fwd(const float *pos_val, const float *neg_val, int B, int k,
                        float *loss, float *rowexp, void *stream);
/* Backward of loss + top-k + split in one pass: datt [B,C,S] is zero except at the selected
 * positions, -g e_b/(B k) for positives and +g e_b/(B k) for negatives. grad_out: 1 float (device). */
int edrl_select_loss_bwd(const float *rowexp, const int32_t *pos_idx, const int32_t *neg_idx,
                         const int64_t *y, const float *grad_out, int B, int C, int S, int k,
                         float *datt, void *stream);

/* Fused train path of EPRL.forward after the encoder (fusion_net.py:137-150, 220-243): one call forward, one call
 * backward -- the same kernels as the entry points above, launched back to back so the host issues 2 calls instead
 * of 10 and never touches softplus / slicing ops.  proxies is the [C, 2F] parameter (mu | raw sigma; softplus with
 * torch's threshold is applied in the kernel, fusion_net.py:116-119); eps [C,S,F]; y int64 [B].
 *   saved   : edrl_essence_saved_floats() floats written by forward, read by backward
 *   scratch : edrl_essence_scratch_floats() floats, backward only
 *   dz [B,T,F] and/or dproxies [C,2F] may be NULL when not needed. */
size_t edrl_essence_saved_floats(int B, int T, int F, int C, int S, int k);
size_t edrl_essence_scratch_floats(int B, int T, int F, int C, int S, int k);
int edrl_essence_train_fwd(const float *z, const float *proxies, const float *eps, const int64_t *y, int B, int T,
                           int F, int C, int S, int k, float *loss, float *saved, void *stream);
int edrl_essence_train_bwd(const float *z, const float *proxies, const float *eps, const int64_t *y, int B, int T,
                           int F, int C, int S, int k, const float *saved, const float *grad_out, float *scratch,
                           float *dz, float *dproxies, void *stream);

/* North-star extension (no reference code, oracle = torch.topk o torch.gather):
 *   out[b, j, :] = features[b, idx[b,j], :],  features [B,T,D], idx [B,k] int32, out [B,k,D]. */
int edrl_gather_rows_fwd(const float *features, const int32_t *idx, int B, int T, int D, int k,
                         float *out, void *stream);
/* dfeatures [B,T,D] = scatter of dout [B,k,D] (indices of one row are distinct), zero elsewhere. */
int edrl_gather_rows_bwd(const float *dout, const int32_t *idx, int B, int T, int D, int k,
                         float *dfeatures, void *stream);

/* ------------------------------------------------------------------------------------------
 * Next row -- DILR Barlow-Twins cross-correlation loss      code/fusion_net.py:656-677
 *
 * z1, z2 [B, D] fp32 (D = 2048 in the reference).  Replaces `c = bn1(z1).T @ bn2(z2); c.div_(4 batch_size)` and the
 * four masked sums over the common block c[:dc, :dc] and the unique block c[dc:, dc:]; the D x D matrix is never
 * stored.  BatchNorm1d(affine=False): training != 0 normalises with the batch statistics (biased variance, eps) and
 * updates run_mean / run_var with `momentum` (unbiased variance), like nn.BatchNorm1d (code/fusion_net.py:653-654);
 * training == 0 normalises with the running statistics.  out6 = (loss_c, on_diag_c, off_diag_c, loss_u, on_diag_u,
 * off_diag_u), the reference's return tuple (:677).  The workspace keeps the normalised operands for the backward.
 * ------------------------------------------------------------------------------------------ */
size_t      edrl_dilr_workspace_bytes(int B, int D);
int         edrl_dilr_bt_loss_fwd(const float *z1, const float *z2, int B, int D, int common_dim, int batch_size,
                                  float eps, int training, float momentum, float *run_mean1, float *run_var1,
                                  float *run_mean2, float *run_var2, float *out6, void *workspace,
                                  size_t workspace_bytes, void *stream);
/* grad_out6: upstream gradients of the six outputs (device); dz1 / dz2 [B, D] (either may be NULL). */
int         edrl_dilr_bt_loss_bwd(int B, int D, int common_dim, int batch_size, int training, const float *grad_out6,
                                  float *dz1, float *dz2, void *workspace, size_t workspace_bytes, void *stream);

/* ------------------------------------------------------------------------------------------
 * Next row -- the loader's two views on the device             code/data_harvard.py:698-783
 *
 * x: `items` preprocessed samples of `per_item` values each -- fp32 in [0, 1], or uint8 (is_u8: the / 255 of :694-695
 * is applied here).  low = clip(x, 0, 1) (the reference's zero-variance draw, :722-731), high = clip(x + N(0, sigma), 0, 1)
 * (:769-783).  noise == NULL: the field comes from Philox4x32-10(seed, element index) + Box-Muller; shared_field != 0
 * gives every item the same field, as the reference's per-item np.random.seed(seed_idx) does (:698).  noise != NULL:
 * that tensor (already scaled, same shape as x) is added instead -- the parity mode.
 * ------------------------------------------------------------------------------------------ */
int         edrl_noise_views(const void *x, int is_u8, long long per_item, int items, float sigma,
                             unsigned long long seed, int shared_field, const float *noise, float *low, float *high,
                             void *stream);

/* ------------------------------------------------------------------------------------------
 * Next row -- the small losses that close MedFusion.forward   code/fusion_net.py:929-942, 390-402
 *
 * pred [B, >= C] (row stride ldp; the reference slices pred[:, :2], :930), y int64 [B]: label-smoothed cross-entropy
 * (:931-939).  mu_* / sig_* [B, Cm, F]: KL(N(mu, sigma) || N(0, 1)) as KL_between_normals / get_KL_loss compute it
 * (summed over the class axis, averaged over batch and features).  out3 = (loss1, kl_fundus, kl_oct); one launch each
 * way.  Gradient outputs may be NULL.
 * ------------------------------------------------------------------------------------------ */
int         edrl_head_losses_fwd(const float *pred, int ldp, const int64_t *y, int B, int C, float smoothing,
                                 const float *mu_f, const float *sig_f, const float *mu_o, const float *sig_o, int Cm,
                                 int F, float *out3, void *stream);
int         edrl_head_losses_bwd(const float *pred, int ldp, const int64_t *y, int B, int C, float smoothing,
                                 const float *mu_f, const float *sig_f, const float *mu_o, const float *sig_o, int Cm,
                                 int F, const float *grad3, float *dpred, float *dmu_f, float *dsig_f, float *dmu_o,
                                 float *dsig_o, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* EDRL_B200_H_ */
