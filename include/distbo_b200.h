/* distbo_b200 - C ABI of the B200-native (sm_100a) distBO joint-GP hot path.
 *
 * Drop-in boundary.  The reference (hcllaw/distBO) has no FFI: its hot path
 * is a TensorFlow-1.7 graph evaluated by sess.run.  Each entry point below replaces the TF ops
 * cited next to it (paths relative to the reference repository).  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *  - all pointers are DEVICE pointers unless a name ends in _host; matrices are row-major;
 *  - dense N x N matrices are padded to n_pad = roundup(N,128) with leading dimension n_pad, the
 *    pad block being the identity (K, L, W=L^-1) or zero (everything else);
 *  - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream and
 *    never synchronises the device (distbo_plan_create / destroy excepted);
 *  - return value: 0 ok, <0 bad argument (-1) or CUDA error (-2).  A failed Cholesky is reported
 *    asynchronously through the device int `info` (LAPACK style: 1-based index of the first
 *    non-positive pivot, 0 = success; DISTBO_INFO_TIMEOUT (< 0) = an inter-CTA wait inside the
 *    factorisation ran out, i.e. the device is shared and a producer CTA was not resident: the
 *    factor is invalid); the caller reads it back with the loss.
 *  - not re-entrant per plan; one plan per (device, n_pad).
 */
#ifndef DISTBO_B200_H
#define DISTBO_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DISTBO_INFO_TIMEOUT (-1000) /* value of `info` after a bounded inter-CTA wait expired */

#define DISTBO_ACQ_EI 0  /* bayes_opt/helpers.py:165-171 */
#define DISTBO_ACQ_UCB 1 /* bayes_opt/helpers.py:151-156 */
#define DISTBO_ACQ_LCB 2 /* bayes_opt/helpers.py:158-163 */

/* Library / build identification: returns the compiled arch as an int (1000 for sm_100a). */
int distbo_version(void);
/* Number of CUDA kernels this library has launched since load (or since the last reset != 0).
 * Host-side counter, used by bench.py for its "gpu_launches" claim. */
int64_t distbo_launch_count(int reset);

/* Debug aid: device buffer (2 * n_pad/128 * 4 uint64; slot 0 of every 4 initialised to ~0, the rest to 0)
 * that distbo_potrf_f64's kernels fill with globaltimer stamps; NULL switches it off. */
int distbo_debug_timeline(void* buf);

/* ---- K1: dataset embedding  (train/distbo_net.py:7-48, train/base.py:9-43) ----------------------
 * Fused tanh-MLP feature map phi_x (and phi_y), per-sample outer product phi_x (x) phi_y, and
 * per-dataset mean over contiguous row segments seg_off[t]..seg_off[t+1].
 *   X [S,dx], Y [S,dy] row-major; nets: W1 [d_in,h], b1 [h], W2 [h,p] (no bias on the last layer).
 *   x net with three layers (nn_net's weights_3 branch, train/distbo_net.py:11-13): pm > 0, W2x [hx,pm], b2x [pm],
 *   W3x [pm,p]: phi_x = tanh(tanh(x W1x + b1x) W2x + b2x) W3x.  pm == 0 (b2x, W3x NULL): the two-layer map.
 *   The 'nn' operator (distbo_net.py:18-20) is this x net applied to the caller's [x | y] rows with y_mode 0 or 4.
 *   y_mode 0: embed_op 'none'  -> out cols = p                      (phi_x only)
 *          1: 'joint' regression -> p*q  (index i*q+j), y-net (W1y,b1y,W2y) applied to Y
 *          2: classification     -> p*dy + dy  (outer product with one-hot Y, then class ratios)
 *          3: 'concat' regression -> p + q*p: mean(phi_x), then the conditional embedding operator
 *             C = Psi^T Phi (Phi^T Phi + lambda I)^-1 flattened [q][p], lambda = exp(*log_reg)
 *             (train/distbo_net.py:49-89); raw_sums [T, p + q*p + p*p] doubles receives the per-dataset sums
 *             the backward pass needs.  log_reg / raw_sums may be NULL for y_mode 0..2 and 4.
 *          4: phi_x only, then the class ratios mean(Y) -> p + dy   ('nn' on classification data, :43-48)
 *   E [T, lde]: columns [0,P') written, the caller owns any extra columns (size-ratio column).
 *   S = seg_off[T] (total rows, known to the host).  workspace: distbo_embed_workspace(...) bytes.
 * _f32 / _f64 select the arithmetic type of samples, weights and output. */
int distbo_embed_fwd_f64(const double* X, const double* Y, const int32_t* seg_off, int T, int64_t S, int dx,
                         int dy, const double* W1x, const double* b1x, const double* W2x, int hx, int p, int pm,
                         const double* b2x, const double* W3x, const double* W1y, const double* b1y,
                         const double* W2y, int hy, int q, int y_mode, const double* log_reg, double* raw_sums,
                         double* E, int lde, void* workspace, int64_t workspace_bytes, void* stream);
int distbo_embed_fwd_f32(const float* X, const float* Y, const int32_t* seg_off, int T, int64_t S, int dx, int dy,
                         const float* W1x, const float* b1x, const float* W2x, int hx, int p, int pm,
                         const float* b2x, const float* W3x, const float* W1y, const float* b1y, const float* W2y,
                         int hy, int q, int y_mode, const float* log_reg, double* raw_sums, float* E, int lde,
                         void* workspace, int64_t workspace_bytes, void* stream);
/* 1 when distbo_embed_fwd_f32 runs this shape on the tensor-core kernel (two-layer x net, x-net-only poolings:
 * y_mode 0 and 2 with dy <= 4, shared-memory ring fits): rows fetched by bulk async copies, layer 1 as 3xTF32
 * mma.sync, the linear last layer applied after the per-dataset mean.  0: the generic kernel.
 * (DISTBO_EMBED_TC=0 forces 0.) */
int distbo_embed_f32_on_tensor_cores(int dx, int dy, int hx, int p, int pm, int y_mode);
/* Backward of the above (replaces TF autodiff through distbo_net, train/train.py:31-40):
 * dE [T,lde] -> gradients of the net weights (same shapes as the weights; overwritten; gb2x / gW3x only with
 * pm > 0); y_mode 3 also needs log_reg, the raw_sums its forward pass stored, and returns g_log_reg (device scalar). */
int distbo_embed_bwd_f64(const double* X, const double* Y, const int32_t* seg_off, int T, int64_t S, int dx,
                         int dy, const double* W1x, const double* b1x, const double* W2x, int hx, int p, int pm,
                         const double* b2x, const double* W3x, const double* W1y, const double* b1y,
                         const double* W2y, int hy, int q, int y_mode, const double* log_reg,
                         const double* raw_sums, const double* dE, int lde, double* gW1x, double* gb1x,
                         double* gW2x, double* gb2x, double* gW3x, double* gW1y, double* gb1y, double* gW2y,
                         double* g_log_reg, void* workspace, int64_t workspace_bytes, void* stream);
/* Bytes of device workspace the calls above need for S = seg_off[T] total rows (chunk table + per-chunk
 * partial sums; elem_bytes 4 or 8; backward != 0 for distbo_embed_bwd_f64). */
int64_t distbo_embed_workspace(int64_t S, int T, int dx, int dy, int hx, int p, int pm, int hy, int q, int y_mode,
                               int elem_bytes, int backward);

/* ---- host helper of the mini-batch sampler (train/gp_utils.py:38-84) ---------------------------------
 * CPython's random.shuffle(x) on an int64 array, bit for bit: mt_state [624] and *mt_index are the Mersenne-Twister
 * state of a random.Random (getstate()[1][:624], [624]), advanced in place.  The reference shuffles class-member
 * lists with the global `random`; this keeps its stream without Python-speed shuffles on the training thread.
 * Host memory only; no device work. */
int distbo_py_shuffle_i64(uint32_t* mt_state, int32_t* mt_index, int64_t* x, int64_t n);

/* ---- K2: Matern-3/2 product Gram  (train/kernels.py:19-28,58-88; log_gaus_likelihood.py:86-122) --
 * Task-level Gram KE[t,t'] = matern32(E[t]/exp(log_bw_embed), E[t']/exp(log_bw_embed)), [T,T]. */
int distbo_task_gram_f64(const double* E, int lde, int T, int P, const double* log_bw_embed,
                         double* KE, void* stream);
/* theta_s[i,k] = theta[i,k] / exp(log_bw[k]) and its row norms; pad rows (i >= N) are zero. */
int distbo_scale_theta_f64(const double* theta, int N, int d, const double* log_bw, double* theta_s,
                           double* sqnorm, int n_rows_out, void* stream);
/* K[i,j] = exp(log_scale) * matern32(theta_i, theta_j) * KE[task_id[i], task_id[j]]
 *          + (exp(2 log_sigma) + jitter) * [i == j]           for i,j < N (lower 128-tiles only),
 * identity on the pad block.  KE may be NULL (kernel 'params': factor 1). */
int distbo_gram_f64(const double* theta_s, const double* sqnorm, int N, int n_pad, int d,
                    const int32_t* task_id, const double* KE, int T, const double* log_scale,
                    const double* log_sigma, double jitter, double* K, void* stream);

/* ---- K3: GP fit  (tf.cholesky log_gaus_likelihood.py:123,211; gp_utils.py:118-147) ---------------
 * A plan caches the tile lists / streams / events for one padded size. */
int distbo_plan_create(int n_pad, void** plan_out);
int distbo_plan_destroy(void* plan);
/* A (n_pad x n_pad, lower tiles valid) is overwritten by its lower Cholesky factor L (upper part of
 * the diagonal blocks zeroed); the 128x128 diagonal blocks of W receive inv(L_kk). */
int distbo_potrf_f64(void* plan, double* A, double* W, int32_t* info, void* stream);
/* W <- L^-1 (lower), given L and the diagonal sub-block inverses left in W by distbo_potrf_f64. T: n_pad^2 temp. */
int distbo_trtri_f64(void* plan, const double* L, double* W, double* T, void* stream);
/* Kinv (lower 128-tiles) <- W^T W. */
int distbo_lauum_f64(void* plan, const double* W, double* Kinv, void* stream);
/* V = W (y - mean), alpha = W^T V, nll = 0.5|V|^2 + 0.5 N log(2 pi) + sum log L_ii.
 * y [N]; V, alpha [n_pad]; mean, nll device scalars; workspace: distbo_nll_workspace(n_pad) doubles. */
int64_t distbo_nll_workspace(int n_pad);
int distbo_nll_f64(const double* L, const double* W, int N, int n_pad, const double* y,
                   const double* mean, double* V, double* alpha, double* nll, double* workspace,
                   void* stream);
/* Closed-form gradient contraction with G = 0.5 (Kinv - alpha alpha^T), replacing TF autodiff
 * through the Cholesky (train/train.py:33-40):
 *   g_log_bw [d], g_scalars = {d/dlog_scale, d/dmean, d/dlog_sigma}, H [T,T] = dNLL/dKE (full).
 * workspace: n_pad * (T + d + 3) + T * T doubles. */
int distbo_gram_grad_f64(const double* Kinv, const double* alpha, const double* theta_s,
                         const double* sqnorm, int N, int n_pad, int d, const int32_t* task_id,
                         const int32_t* task_off, const double* KE, int T, const double* log_scale,
                         const double* log_sigma, double* g_log_bw, double* g_scalars, double* H,
                         double* workspace, void* stream);
/* Backward of distbo_task_gram_f64: H [T,T] -> g_log_bw_embed [P], dE [T,lde] (cols < P).
 * workspace: T * P doubles. */
int distbo_task_gram_bwd_f64(const double* E, int lde, int T, int P, const double* log_bw_embed,
                             const double* H, double* g_log_bw_embed, double* dE, double* workspace,
                             void* stream);
/* Multi-task kernel (train/kernels.py:32-55, log_gaus_likelihood.py:68-75,114-116): the task covariance
 * KE = D^-1/2 (L L^T) D^-1/2 with L = tril(fill_triangular(exp(log_tri))), D = diag(L L^T); log_tri has
 * T(T+1)/2 entries in tf.contrib.distributions.fill_triangular order.  It takes the place of
 * distbo_task_gram_f64 for kernel 'multi'.  workspace: 3*T*T doubles (both calls). */
int distbo_task_tri_f64(const double* log_tri, int T, double* KE, double* workspace, void* stream);
/* Backward: H = dNLL/dKE [T,T] (from distbo_gram_grad_f64) -> g_log_tri [T(T+1)/2]. */
int distbo_task_tri_bwd_f64(const double* log_tri, int T, const double* H, double* g_log_tri,
                            double* workspace, void* stream);
/* ---- BLR surrogate on deep features (algorithm='BLR')  (train/log_gaus_likelihood_feature.py:13-225,
 *      train/feature_map.py:12-46) ------------------------------------------------------------------
 * Feature row F_i = [emb[task_id[i]] (Pe) | theta_i (d)];  Phi = tanh(F W1 + b1) W2 (n_layers 2, D = h2) or
 * tanh(tanh(F W1 + b1) W2 + b2) W3 (n_layers 3); W1 [Pe+d,h1], W2 [h1,h2], W3 [h2,D] row-major, widths <= 64,
 * d <= 16.  A = I + (alpha/sigma^2) Phi^T Phi = L L^T, e = L^-1 Phi^T y,
 * loss = (1/sigma^2)(|y|^2 - (alpha/sigma^2)|e|^2) + N log sigma^2 + 2 sum log L_ii      (:107-143).
 * Outputs: Linv [64*64] (row-major, ld 64, zero padded), evec [64], loss, info (0, or k+1 when pivot k is not
 * positive).  want_grad != 0 also fills the gradients of every weight, of log_sigma / log_alpha, and dE [T,ldE]
 * (d loss / d emb; may be NULL) - the closed form of what TF autodiff returns (train/train.py:31-40).
 * Pe == 0: no embedding part (kernel 'params'); E, task_id, task_off are ignored.
 * The random Fourier map (num_of_rff > 0, feature_map.py:19-26) is not part of this entry point.
 * workspace: distbo_blr_workspace(N, T, d, h1, h2, D) doubles. */
int64_t distbo_blr_workspace(int64_t N, int T, int d, int h1, int h2, int D);
int distbo_blr_fit_f64(const double* theta, const double* y, int N, int d, const double* E, int ldE, int Pe,
                       const int32_t* task_id, const int32_t* task_off, int T, const double* W1,
                       const double* b1, const double* W2, const double* b2, const double* W3, int n_layers,
                       int h1, int h2, int D, const double* log_sigma, const double* log_alpha, double* Linv,
                       double* evec, double* loss, int32_t* info, double* gW1, double* gb1, double* gW2,
                       double* gb2, double* gW3, double* g_log_sigma, double* g_log_alpha, double* dE,
                       double* workspace, int want_grad, void* stream);
/* Posterior + acquisition + arg-max over candidates (build_predict_feature, :145-190; helpers.py:51-68,151-171):
 * mean = (alpha/sigma^2) e^T L^-1 phi, var = alpha |L^-1 phi|^2, std = sqrt(max(var, 1e-32)).  cand [M,d] raw
 * (cand_shift / cand_scale = frozen StandardScaler, or both NULL), emb [Pe] = embedding of the candidates' task.
 * block_best_*: distbo_blr_score_blocks() entries.  Optional outputs may be NULL. */
int distbo_blr_score_blocks(void);
int distbo_blr_score_f64(const double* cand, const double* cand_shift, const double* cand_scale, int64_t M,
                         int d, const double* emb, int Pe, const double* W1, const double* b1,
                         const double* W2, const double* b2, const double* W3, int n_layers, int h1, int h2,
                         int D, const double* Linv, const double* evec, const double* log_sigma,
                         const double* log_alpha, int acq_kind, double y_max, double xi, double kappa,
                         double* out_mean, double* out_std, double* out_acq, double* block_best_val,
                         int64_t* block_best_idx, double* best_val, int64_t* best_idx, int64_t index_offset,
                         void* stream);
/* TF-1.7 Adam update on a flat parameter vector (train/gp_regressor.py:141; epsilon-hat form):
 * lr_t = lr sqrt(1 - beta2^step) / (1 - beta1^step); p -= lr_t m / (sqrt(v) + eps).  step is 1-based.
 * clip_norm <= 0 disables tf.clip_by_global_norm (train/train.py:33-35).
 * lr_t_dev (optional): device scalar holding lr_t; when non-NULL it is used instead of (lr, step) so
 * that the launch can be replayed from a CUDA graph with a value the host refreshes per step.
 * ctrl (optional, DISTBO_CTRL_SLOTS doubles on the device): the training loop's control block - the
 * early-stopping rule of train/train.py:53-65 evaluated ON THE DEVICE so that the host never has to read the
 * loss between epochs.  With ctrl: the launch is a no-op once ctrl[DISTBO_CTRL_STOP] != 0; if *info != 0 (failed
 * factorisation) STOP = 2, INFO = *info and the update is NOT applied (TF raises inside sess.run); otherwise
 * after the update the epoch's loss *nll goes to loss_hist[EPOCHS % hist_cap] and LAST_LOSS, EPOCHS += 1, and for
 * EPOCHS-before >= FIRST_EPOCH: (CUR_MIN - loss) > THRESHOLD ? (COUNTDOWN = PATIENCE, CUR_MIN = loss)
 * : COUNTDOWN -= 1; COUNTDOWN <= 0 -> STOP = 1.  The caller initialises CUR_MIN = COUNTDOWN = +inf, STOP = EPOCHS = 0,
 * FIRST_EPOCH, THRESHOLD (1e-4) and PATIENCE (100) before the first epoch of a fit. */
#define DISTBO_CTRL_CUR_MIN 0
#define DISTBO_CTRL_COUNTDOWN 1
#define DISTBO_CTRL_STOP 2
#define DISTBO_CTRL_EPOCHS 3
#define DISTBO_CTRL_LAST_LOSS 4
#define DISTBO_CTRL_FIRST_EPOCH 5
#define DISTBO_CTRL_INFO 6
#define DISTBO_CTRL_THRESHOLD 7
#define DISTBO_CTRL_PATIENCE 8
#define DISTBO_CTRL_SLOTS 16
int distbo_adam_f64(double* params, const double* grads, double* m, double* v, int64_t n, double lr,
                    double beta1, double beta2, double eps, int64_t step, double clip_norm,
                    const double* lr_t_dev, double* ctrl, const double* nll, const int32_t* info,
                    double* loss_hist, int64_t hist_cap, void* stream);

/* ---- K4: posterior + acquisition + argmax  (log_gaus_likelihood.py:128-218; gp_regressor.py:197;
 *      bayes_opt/helpers.py:51-68,151-171) ---------------------------------------------------------
 * For candidates theta_c [M,d] (already scaled by the frozen StandardScaler; NOT yet divided by the
 * bandwidths): k_* = scale * matern32(theta, theta_c) * kvec, A = W k_*, mean = A^T V + m,
 * var = scale - |A|^2, std = sqrt(max(var,1e-32)), acq in {EI, UCB, LCB}.
 * kvec [n_pad] = KE[task(i), target] (1 for kernel 'params'; 0 on pad rows).
 * Outputs (any may be NULL): mean [M], std [M], acq [M]; best_val[1], best_idx[1] (int64; first
 * index of the maximum, as numpy argmax).  Kst: workspace of distbo_score_workspace(n_pad) doubles;
 * block_best_val / block_best_idx: distbo_score_chunk() / 128 entries each. */
int distbo_score_chunk(void);
/* doubles of workspace distbo_score_f64 needs behind `Kst` (K_* chunk + per-slice partial sums) */
int64_t distbo_score_workspace(int n_pad);
/* Measurement aid: while enabled, distbo_score_f64 brackets every launch of its contraction kernel
 * with CUDA events on the launching stream; _read synchronises them and returns the summed device
 * time (ms) and the number of launches since the last enable/read. */
int distbo_score_timing(int enable);
int distbo_score_timing_read(double* total_ms_host, int64_t* launches_host);
int distbo_score_f64(const double* W, const double* V, int N, int n_pad, const double* theta_s,
                     const double* sqnorm, int d, const double* log_bw, const double* kvec,
                     const double* log_scale, const double* mean_param, const double* theta_c,
                     const double* cand_shift, const double* cand_scale, int64_t M, int acq_kind, double y_max, double xi, double kappa, double* Kst,
                     double* out_mean, double* out_std, double* out_acq, double* block_best_val,
                     int64_t* block_best_idx, double* best_val, int64_t* best_idx,
                     int64_t index_offset, void* stream);

/* ---- K4 on the 5th-generation tensor cores (tcgen05, int8 digits, exact integer accumulation) ------------
 * Same operator, same results to fp64 rounding level, for the same reference code as distbo_score_f64
 * (log_gaus_likelihood.py:199-217: tf.matrix_triangular_solve + reduce_sum(square)).  fp64 has no tcgen05
 * kind; W = L^-1 and K_* are cut into 7 signed base-256 digits of a 55-bit fixed-point number each and the
 * 28 digit-pair products that matter are exact int32 tcgen05.mma kind::i8 accumulations in tensor memory,
 * folded into fp64 once per (candidate, row of W).
 * distbo_score_i8_prepare: W [n_pad,n_pad] (lower triangle read) -> W8 [7][n_pad][n_pad] int8 digit planes,
 *   rowscale [n_pad] = 2^exponent(max_k |W[i,k]|); once per fit.
 * distbo_score_i8_f64: as distbo_score_f64 with (W8, rowscale) in place of W and `ws` a workspace of
 *   distbo_score_i8_workspace(n_pad) doubles (digit planes of one K_* chunk + per-slice partial sums). */
int64_t distbo_score_i8_workspace(int n_pad);
int distbo_score_i8_prepare(const double* W, int n_pad, int8_t* W8, double* rowscale, void* stream);
int distbo_score_i8_f64(const int8_t* W8, const double* rowscale, const double* V, int N, int n_pad,
                        const double* theta_s, const double* sqnorm, int d, const double* log_bw,
                        const double* kvec, const double* log_scale, const double* mean_param,
                        const double* theta_c, const double* cand_shift, const double* cand_scale, int64_t M,
                        int acq_kind, double y_max, double xi, double kappa, double* ws, double* out_mean,
                        double* out_std, double* out_acq, double* block_best_val, int64_t* block_best_idx,
                        double* best_val, int64_t* best_idx, int64_t index_offset, void* stream);

/* ---- LAUUM on the tcgen05 tensor cores: Kinv = W^T W (lower 128-tiles), the same digit-plane contraction with
 * T = W^T split per column of W; replaces distbo_lauum_f64 from n_pad = 2048 (autodiff through tf.cholesky /
 * tf.matrix_triangular_solve in the reference, log_gaus_likelihood.py:119-125).
 * distbo_lauum_i8_plan fills `sched` (distbo_lauum_i8_sched_words(n_pad) int32, device) with the static work
 * lists of the persistent CTAs - once per n_pad, synchronous, outside any capture.
 * distbo_lauum_i8_f64: ws = distbo_lauum_i8_workspace(n_pad) doubles; capture-safe. */
int64_t distbo_lauum_i8_sched_words(int n_pad);
int distbo_lauum_i8_plan(int n_pad, int32_t* sched);
int64_t distbo_lauum_i8_workspace(int n_pad);
int distbo_lauum_i8_f64(const double* W, int n_pad, const int32_t* sched, double* ws, double* Kinv, void* stream);

/* ---- Triangular inverse W = L^-1 by levels.  distbo_trtri_levels: number of levels of the halving tree over the
 * 128-row diagonal blocks (level 0 = the diagonal blocks themselves).  distbo_trtri_range_f64 runs levels
 * [level_begin, level_end) on the FP64 DMMA tile GEMM (distbo_trtri_f64 = all of them).
 * The levels whose blocks have at least `min_rows` rows can run on the tcgen05 digit-plane kernel instead:
 * distbo_trtri_i8_first_level = the first such level; distbo_trtri_i8_plan fills `sched`
 * (distbo_trtri_i8_sched_words int32, device; once per n_pad, synchronous); distbo_trtri_i8_f64 runs those levels
 * (ws: distbo_trtri_i8_workspace(n_pad) doubles; T: the n_pad x n_pad scratch distbo_trtri_f64 uses; capture-safe).
 * Replaces tf.matrix_triangular_solve of log_gaus_likelihood.py:183-184,211-214. */
int distbo_trtri_levels(void* plan);
int distbo_trtri_range_f64(void* plan, const double* L, double* W, double* T, int level_begin, int level_end,
                           void* stream);
int distbo_trtri_i8_first_level(int n_pad, int min_rows);
int64_t distbo_trtri_i8_sched_words(int n_pad, int min_rows);
int distbo_trtri_i8_plan(int n_pad, int min_rows, int32_t* sched);
int64_t distbo_trtri_i8_workspace(int n_pad);
int distbo_trtri_i8_f64(const double* L, double* W, double* T, int n_pad, int min_rows, const int32_t* sched,
                        double* ws, void* stream);

/* Mini-batch gather: dst[r, :] = src[idx[r], :], rows of row_bytes bytes (multiple of 8), idx int64 on the device.
 * The device part of loop_batches (train/gp_utils.py:28-104); capture-safe. */
int distbo_gather_rows(const void* src, int64_t row_bytes, const int64_t* idx, int64_t n_rows, void* dst, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DISTBO_B200_H */
