/*
 * ctr_b200.h -- C ABI of libctr_b200.so, the B200 (sm_100a) CTR feature-interaction engine.
 *
 * This is the drop-in boundary for the hot path of lambdaji/tf_repos'
 * deep_ctr/Model_pipeline/*.py `model_fn`s.  The reference has no FFI of its own (it is
 * TensorFlow-1.4 Python); each entry point below replaces the cluster of TF ops cited next to
 * it (file:line relative to the reference checkout).  INTEGRATION.md shows the ctypes binding a
 * maintainer of the reference would add.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller unless the name ends in `_host`;
 *    tensors are contiguous row-major fp32 / int32 (ids may be int32 or int64, see id_bits);
 *  - no allocation inside the library: scratch is caller-provided, sized by *_workspace_bytes();
 *  - every call enqueues work on `stream` and returns immediately (no host sync, graph-capturable);
 *  - return value: 0 = CTR_OK, <0 = ctr_status; text via ctr_last_error() (thread-local);
 *  - no C++ exception crosses the ABI; no global mutable state except the launch counter.
 */
#ifndef CTR_B200_H_
#define CTR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* ctr_stream_t; /* a cudaStream_t / CUstream */

enum ctr_status {
  CTR_OK = 0,
  CTR_ERR_INVALID_ARG = -1,
  CTR_ERR_UNSUPPORTED = -2,
  CTR_ERR_CUDA = -3,
  CTR_ERR_WORKSPACE = -4
};

/* interaction modes of the embedding kernels */
enum ctr_fm_mode {
  CTR_FM_DEEPFM = 0, /* y_v[B] = 0.5*sum_k((sum_f e)^2 - sum_f e^2)        DeepFM.py:129-135 */
  CTR_FM_NFM = 1,    /* bi[B,K] = 0.5*((sum_f e)^2 - sum_f e^2)            NFM.py:122-128    */
  CTR_FM_PLAIN = 2   /* gather+scale only                                  DCN.py:135-138, PNN.py:134-136, AFM.py:128-130 */
};

enum ctr_optimizer {
  CTR_OPT_ADAM = 0,     /* tf.train.AdamOptimizer      DeepFM.py:205 */
  CTR_OPT_ADAGRAD = 1,  /* tf.train.AdagradOptimizer   DeepFM.py:207 */
  CTR_OPT_MOMENTUM = 2, /* tf.train.MomentumOptimizer  DeepFM.py:209 */
  CTR_OPT_FTRL = 3      /* tf.train.FtrlOptimizer      DeepFM.py:211 */
};

/* ---- library ------------------------------------------------------------------------------- */
int ctr_abi_version(void);
const char* ctr_last_error(void);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t ctr_launch_count(void);
/* number of SMs the library sized its persistent grids for (148 on B200); <0 on error */
int ctr_device_sm_count(void);

/* ---- K1: gather + scale + FM first/second order + emit x ------------------------------------
 * Replaces tf.nn.embedding_lookup(FM_W/FM_V) + multiply + reduce_sum/square chain,
 * DeepFM.py:125-135,151 (NFM.py:118-128; plain gather DCN.py:135-138).
 *   ids   [B,F]  int32 (id_bits=32) or int64 (id_bits=64); vals [B,F] f32
 *   V     [N,K]  f32;   W [N] f32 or NULL (no first-order term)
 *   x     [B,F*K] = V[ids]*vals            (NULL to skip; NFM does not need it)
 *   y_w   [B]     = sum_f W[ids]*vals      (NULL iff W NULL)
 *   y2    mode DEEPFM: [B]; mode NFM: [B,K]; mode PLAIN: ignored (may be NULL)
 *   S     [B,K]   = sum_f e  (saved for the backward; NULL in PLAIN mode)
 *   oob   optional int32[2] device word: {count, first bad id}; TF raises InvalidArgument for ids
 *         outside [0,N) on CPU -- here such an occurrence contributes 0 and is counted.
 */
int ctr_fm_embed_fwd(const void* ids, int id_bits, const float* vals, const float* V, const float* W,
                     int64_t N, int B, int F, int K, int mode, float* x, float* y_w, float* y2,
                     float* S, int32_t* oob, ctr_stream_t stream);

/* ---- K2: backward of K1 w.r.t. the gathered rows ---------------------------------------------
 * Replaces the autodiff of DeepFM.py:125-135,151 that optimizer.minimize (DeepFM.py:213) builds:
 * per-occurrence IndexedSlices values for FM_V and FM_W.
 *   x   [B,F*K] the scaled embeddings saved by the forward;  S [B,K] saved by the forward
 *   dX  [B,F*K] upstream grad of x (NULL = 0);  dy2: DEEPFM [B] upstream of y_v, NFM [B,K] of bi
 *   dyw [B] upstream of y_w (NULL iff g_w NULL)
 *   g_rows [B*F,K] = (dy2*(S-e) + dX) * val ;  g_w [B*F] = dyw*val
 */
int ctr_fm_embed_bwd(const float* vals, const float* x, const float* S, const float* dX,
                     const float* dy2, const float* dyw, int B, int F, int K, int mode,
                     float* g_rows, float* g_w, ctr_stream_t stream);

/* ---- K3: de-duplication of IndexedSlices ------------------------------------------------------
 * Replaces optimizer._deduplicate_indexed_slices = tf.unique + unsorted_segment_sum that
 * optimizer.minimize applies to the embedding gradients (DeepFM.py:213, [TF-sem]).
 * Stable LSD radix sort of (id, position), then run-length encoding.  Integer outputs are
 * bit-exact w.r.t. numpy.unique(return_inverse=True) + a stable argsort:
 *   perm        [n]  positions 0..n-1 sorted by (id, position)
 *   uniq        [n]  ascending distinct ids (first *n_uniq valid)
 *   inverse     [n]  inverse[p] = index into uniq of ids[p]
 *   seg_offsets [n+1] run starts in the sorted order; seg_offsets[*n_uniq] = n
 *   n_uniq      int32[1] device scalar
 *   long_list   int32[n+1] scratch-out: [0] = number of runs longer than CTR_LONG_SEG, then their
 *               uniq indices (consumed by ctr_segment_sum_rows)
 */
#define CTR_LONG_SEG 128
size_t ctr_unique_segment_workspace_bytes(int64_t n, int64_t N);
int ctr_unique_segment(const int32_t* ids, int64_t n, int64_t N, int32_t* perm, int32_t* uniq,
                       int32_t* inverse, int32_t* seg_offsets, int32_t* n_uniq, int32_t* long_list,
                       void* ws, size_t ws_bytes, ctr_stream_t stream);

/* g_uniq[u,:] = sum over the run u of g_rows[perm[i],:]  (fixed-shape trees: deterministic).
 * g_w / gw_uniq are the optional scalar column of the first-order table (NULL to skip). */
int ctr_segment_sum_rows(const float* g_rows, const float* g_w, const int32_t* perm,
                         const int32_t* seg_offsets, const int32_t* n_uniq,
                         const int32_t* long_list, int64_t n, int K, float* g_uniq, float* gw_uniq,
                         void* ws /* optional scratch (e.g. the ctr_unique_segment workspace, free by now): runs longer than
                         CTR_LONG_SEG are cut into 1024-occurrence chunks summed by separate CTAs and added in chunk order
                         (as many runs as the scratch has partial rows for); NULL = one CTA per long run */,
                         size_t ws_bytes, ctr_stream_t stream);

/* ---- K4: optimizer apply on table rows ---------------------------------------------------------
 * TF-1.x arithmetic, [TF-sem] (see oracle/tf_semantics.py for the restatement and its sources).
 * `hyper` is a DEVICE float[8]: {lr_t, beta1, beta2, eps, l2_reg, aux0, aux1, aux2} so that a
 * captured CUDA graph can be replayed while the step counter advances.
 *   Adam:     lr_t = lr*sqrt(1-b2^t)/(1-b1^t) precomputed (ctr_adam_tick);
 *   Adagrad:  lr_t = lr;            slot0 = accumulator
 *   Momentum: lr_t = lr; aux0 = momentum; slot0 = accumulator
 *   Ftrl:     lr_t = lr; aux0 = lr_power, aux1 = l1, aux2 = l2(ftrl);  slot0 = accum, slot1 = linear
 *
 * sparse apply: rows uniq[0..*n_uniq) get g = g_uniq + l2_reg*var and the optimizer's *sparse*
 * update.  If stage != NULL the new (var, slot0, slot1) rows are written to stage[3][n][K]
 * instead of in place (exact mode: the dense sweep runs next, then ctr_opt_patch_rows).
 */
int ctr_opt_sparse_rows(int opt, float* var, float* slot0, float* slot1, const int32_t* uniq,
                        const int32_t* n_uniq, const float* g_uniq, int64_t n_max, int K,
                        const float* hyper, float* stage, ctr_stream_t stream);
/* dense sweep: every one of the n_elem elements takes the step with g = l2_reg*var (what TF does
 * to rows no gather touched, because l2_loss densifies the gradient and sparse Adam decays every
 * row).  Also accumulates sum(var_old^2) into sumsq_partials[grid] (for tf.nn.l2_loss in the
 * loss, DeepFM.py:189-190) when non-NULL; *n_partials returns the number written. */
int ctr_opt_dense_sweep(int opt, float* var, float* slot0, float* slot1, int64_t n_elem,
                        const float* hyper, float* sumsq_partials, int* n_partials_host,
                        ctr_stream_t stream);
int ctr_opt_patch_rows(float* var, float* slot0, float* slot1, const int32_t* uniq,
                       const int32_t* n_uniq, const float* stage, int64_t n_max, int K, int n_slots,
                       ctr_stream_t stream);
/* dense variables with an explicit gradient (MLP weights, biases, fm_bias, cross_w/b):
 * TF's fused ApplyAdam/ApplyAdagrad/ApplyMomentum/ApplyFtrl kernels; g += l2_reg*var first when
 * l2_reg (hyper[4]) != 0. */
int ctr_opt_dense_grad(int opt, float* var, float* slot0, float* slot1, const float* grad,
                       int64_t n_elem, const float* hyper, ctr_stream_t stream);
/* Called once at the start of a step: lr_t = lr*sqrt(1-b2p)/(1-b1p) from the CURRENT beta powers is
 * written to hyper[8*r] for r < n_hyper (consecutive hyper records: tables, dense variables ...),
 * then the powers advance the way AdamOptimizer._finish does (fp32 running products) and the step
 * counter increments.  state = device float[4] {beta1_power, beta2_power, lr, global_step}. */
int ctr_adam_tick(float* state, float* hyper, int n_hyper, ctr_stream_t stream);
/* deterministic sum of n floats -> out[0] (two-level fixed tree) ; scale applied at the end */
int ctr_reduce_sum(const float* in, int64_t n, float scale, float* out, float* ws, size_t ws_bytes,
                   ctr_stream_t stream);
/* 0.5*sum(t^2) (tf.nn.l2_loss) of a whole tensor, deterministic */
size_t ctr_l2_loss_workspace_bytes(int64_t n);
/* out[0] = scale * 0.5 * sum(t^2) */
int ctr_l2_loss(const float* t, int64_t n, float scale, float* out, void* ws, size_t ws_bytes, ctr_stream_t stream);

/* ---- K4': exact-deferred ("epoch") table update ---------------------------------------------------
 * Same results, bit for bit, as ctr_opt_sparse_rows(stage) + ctr_opt_dense_sweep + ctr_opt_patch_rows
 * every step, at 1/P of the HBM traffic: the update TF gives a row that nothing gathered
 * (g = l2_reg*var, DeepFM.py:189-190 + non-lazy sparse Adam, [TF-sem]) is an element-wise recurrence,
 * so it is replayed lazily -- when a batch gathers the row (ctr_epoch_rows) or once per epoch of P
 * steps for all rows in one pass over HBM (ctr_epoch_sweep).
 *   last      uint8[N] per row: steps of the current epoch already applied to the stored state
 *   lr_table  device float[ctr_epoch_max_steps()]: lr_t of every step of the current epoch
 *   ss        device double[ctr_epoch_max_steps()]: sum(var^2) seen by the row kernels, per step
 * Step j of an epoch:  ctr_epoch_tick(j) ; ctr_unique_segment(ids) ;
 *   ctr_epoch_rows(apply=0, j)  -> gathered rows hold the state at the start of step j
 *   forward / backward / ctr_segment_sum_rows ;  ctr_epoch_rows(apply=1, j) ;
 *   after the last step (or to flush mid-epoch): ctr_epoch_sweep(upto, reset) ; ctr_epoch_reg_loss.
 */
int ctr_epoch_max_steps(void);
int ctr_epoch_tick(float* state, float* hyper, int n_hyper, float* lr_table, int j, int is_adam,
                   ctr_stream_t stream);
int ctr_epoch_rows(int opt, int apply, float* var, float* slot0, float* slot1, uint8_t* last,
                   const int32_t* uniq, const int32_t* n_uniq, const float* g_uniq, int64_t n_max, int K,
                   const float* hyper, const float* lr_table, int j, double* ss, ctr_stream_t stream);
/* ctr_epoch_rows for the [N,K] table AND a scalar table [N] gathered with the same ids (fm_v + fm_w, DeepFM.py:115-116)
 * in one launch; K in {4, 8, 16, 32, 64, 128, 256}.  Same arithmetic as two ctr_epoch_rows calls. */
int ctr_epoch_rows2(int opt, int apply, float* var, float* slot0, float* slot1, uint8_t* last, float* w_var, float* w_slot0,
                    float* w_slot1, uint8_t* w_last, const int32_t* uniq, const int32_t* n_uniq, const float* g_uniq,
                    const float* gw_uniq, int64_t n_max, int K, const float* hyper, const float* lr_table, int j, double* ss,
                    double* ss_w, ctr_stream_t stream);
/* All rows -> state after `upto` steps of this epoch.  Rows whose `last` byte equals `from` (nothing gathered
 * them since the previous sweep; from = 0 after an epoch-end sweep) replay steps from..upto-1; the others replay
 * last..upto-1.  reset != 0: epoch end, every `last` byte returns to 0; reset == 0: mid-epoch flush, `last` = upto.
 * ss_partials: device double[ctr_epoch_max_steps()][*n_partials_host] (*n_partials_host is always
 * 6 * ctr_device_sm_count(), so it can be sized before the first call), ZERO-INITIALISED ONCE by the caller: a call
 * rewrites, for every step < upto, one entry per CTA it launches (fewer than *n_partials_host) and ctr_epoch_reg_loss
 * sums whole rows, so the entries no CTA owns must hold 0.
 * list / list_cap / list_count / ss_rows (optional, Adam): scratch for the packed-pipe sweep (csrc/epoch_adam.cu):
 * device int32[list_cap] with list_cap >= min(n_rows, ids gathered since `from`), a device int32 counter, and the
 * row kernels' per-step sum(var^2) accumulator (the `ss` of ctr_epoch_rows).  NULL selects the scalar kernels.
 * *list_count returns the number of gathered rows found; if it exceeds list_cap the precondition was violated and
 * the rows beyond list_cap were NOT caught up (the engine sizes the list so that this cannot happen). */
int ctr_epoch_sweep(int opt, float* var, float* slot0, float* slot1, uint8_t* last, int64_t n_rows, int K,
                    const float* hyper, const float* lr_table, int from, int upto, int reset, double* ss_partials,
                    int* n_partials_host, int32_t* list, int64_t list_cap, int32_t* list_count, double* ss_rows,
                    ctr_stream_t stream);
/* reg[s] (+)= scale*(ss_rows[s] + sum_b ss_partials[s][b]) for s < upto; clears ss_rows[s] */
int ctr_epoch_reg_loss(double* ss_rows, const double* ss_partials, int n_partials, int upto, float scale,
                       float* reg, int accumulate, ctr_stream_t stream);
/* Diagnostics: the epoch sweeps evaluate sqrt/div through hand-scheduled IEEE fast paths with one range
 * check per four elements (optim_steps.cuh).  This compares them, bit for bit, with the compiler's
 * sqrt.rn / div.rn on n pseudo-random in-range operands (seeded; uniform mantissas, hard mantissa
 * patterns mixed in).  mismatches: device int64[2] = {sqrt mismatches, div mismatches} (overwritten). */
int ctr_selftest_divsqrt(uint64_t seed, int64_t n, int64_t* mismatches, ctr_stream_t stream);
/* The packed (FMUL2/FADD2/FFMA2) untouched-row Adam loops of the epoch sweep (csrc/adam_packed.cuh) against the
 * scalar step, bit for bit, on n random 8-element states of a regime (0: normal range; 1: tiny / denormal / zero
 * first moments; 2: also denormal / zero second moments), `steps` (1..4) steps each.
 * out3: device int64[3] = {elements whose (var, m, v) bits differ, trajectories the validity check rejected, total}. */
int ctr_selftest_adam_packed(int regime, uint64_t seed, int64_t n, int steps, float lr, float l2, int64_t* out3,
                             ctr_stream_t stream);

/* ---- loss head -----------------------------------------------------------------------------------
 * y = ((bias + y_a) + y_b) + y_c (NULL terms skipped; DeepFM.py:172-175), pred = sigmoid(y)
 * (:176), loss_ce = sum(max(y,0) - y*t + log1p(exp(-|y|)))/B_total (:188), dy = (pred - t)/B_total,
 * dbias = sum(dy).  B_total = B on one GPU; 1 = summed loss (canned estimators); the global batch under data parallelism (the per-rank
 * loss_ce / dbias / gradients then SUM to the global mean).  labels NULL => inference (y, pred only).
 */
int ctr_logit_loss(const float* bias, const float* y_a, const float* y_b, const float* y_c,
                   const float* labels, int B, int B_total, float* y, float* pred, float* loss_ce,
                   float* dy, float* dbias, ctr_stream_t stream);

/* ---- dense layers: fp32 SIMT GEMMs with fused epilogues -------------------------------------------
 * tf.contrib.layers.fully_connected (DeepFM.py:156,165) + tf.nn.dropout (:162) and their autodiff.
 *   fwd: out[M,Nd] = dropout(act(in[M,Kd] @ Wt[Kd,Nd] + b)),  act 0 = identity, 1 = relu;
 *        drop_mask = binary keep mask [M,Nd] (NULL = no dropout): out = x / keep_prob * mask.
 *   bwd: dOut is overwritten with dZ = (dOut*mask/keep) * (out > 0);  db = colsum(dZ);
 *        dW = in^T @ dZ (split over M, deterministic);  dIn = dZ @ Wt^T (NULL to skip).
 * fc1: the N = 1 output layer over the concatenation [in_a | in_b] (in_b NULL/Kb = 0 for DeepFM's
 *      deep_out; DCN's out_layer takes [x_L, x_deep], DCN.py:178-181).
 * Accumulation order is fixed => bit-reproducible. */
int ctr_fc_fwd(const float* in, const float* Wt, const float* b, const float* drop_mask, float keep_prob,
               int M, int Kd, int Nd, int act, float* out, ctr_stream_t stream);
/* same, plus a per-row-group bias: out = act(in@Wt + b + group_bias[row / group_P]) -- DIN's attention
 * layer, where the ad-embedding part of [e, e-a, a] @ W is one row per sample (DIN.py:161-164) */
int ctr_fc_fwd_grouped(const float* in, const float* Wt, const float* b, const float* group_bias, int group_P,
                       const float* drop_mask, float keep_prob, int M, int Kd, int Nd, int act, float* out,
                       ctr_stream_t stream);
size_t ctr_fc_bwd_workspace_bytes(int M, int Kd, int Nd);
/* accumulate_din != 0: dIn += dZ @ Wt^T (instead of =) */
int ctr_fc_bwd(const float* in, const float* Wt, const float* out, const float* drop_mask, float keep_prob,
               float* dOut, int M, int Kd, int Nd, int act, float* dIn, int accumulate_din, float* dW, float* db,
               void* ws, size_t ws_bytes, ctr_stream_t stream);
int ctr_fc1_fwd(const float* in_a, int Ka, const float* in_b, int Kb, const float* w, const float* b, int M,
                float* y, ctr_stream_t stream);
size_t ctr_fc1_bwd_workspace_bytes(int M, int Ka, int Kb);
int ctr_fc1_bwd(const float* in_a, int Ka, const float* in_b, int Kb, const float* w, const float* dy, int M,
                float* d_a, float* d_b, float* dw, float* db, void* ws, size_t ws_bytes, ctr_stream_t stream);
/* binary keep mask: mask[i] = (hash(seed, *step_dev, i) < keep_prob) -- tf.nn.dropout's
 * floor(keep_prob + uniform); step_dev = device float holding the global step (NULL = 0) so that a
 * captured graph draws a fresh mask every replay.  Not TF's Philox stream. */
int ctr_dropout_mask(float* mask, int64_t n, float keep_prob, uint64_t seed, const float* step_dev,
                     ctr_stream_t stream);

/* ---- batch normalisation after the relu of a hidden layer ------------------------------------------
 * Replaces batch_norm_layer (DeepFM.py:159-160,231-235; DCN.py:171,241-247; PNN.py:181; NFM.py:143; DIN.py:206):
 * tf.contrib.layers.batch_norm(decay, center=True, scale=True, updates_collections=None), epsilon 0.001 [TF-sem],
 * followed by the layer's dropout (drop_mask NULL = none).  x, out: [n, H] row-major.
 * train != 0: batch moments (biased variance) -> save_mean / save_var [H]; moving_mean / moving_var are updated in
 * place (moving -= (moving - batch)*(1 - decay)).  train == 0: moving statistics, no dropout, nothing written but out.
 * ctr_bn_bwd: gradients through the batch moments; d_x [n,H], d_gamma / d_beta [H] (overwritten). */
size_t ctr_bn_workspace_bytes(int H);   /* scratch for the chunked column reductions (TRAIN forward and backward) */
int ctr_bn_fwd(const float* x, int n, int H, const float* gamma, const float* beta, float* moving_mean,
               float* moving_var, int train, float decay, float eps, const float* drop_mask, float keep_prob, float* out,
               float* save_mean, float* save_var, void* ws, size_t ws_bytes, ctr_stream_t stream);
int ctr_bn_bwd(const float* d_out, const float* x, int n, int H, const float* save_mean, const float* save_var,
               const float* gamma, float eps, const float* drop_mask, float keep_prob, float* d_x, float* d_gamma,
               float* d_beta, void* ws, size_t ws_bytes, ctr_stream_t stream);

/* ---- K5: DCN cross network (DCN.py:140-145) ------------------------------------------------------
 * x_{l+1} = x0 * (x_l . w_l) + x_l + b_l,  l = 0..L-1;  w,b: [L,D];  x0: [B,D], D = F*K (D%4==0, <=2048)
 * fwd saves the L scalars s[b,l] = x_l . w_l;  bwd recomputes x_l from x0 and s.
 * bwd: dx0 = dx_in (NULL = 0) + dL/dx0 through the cross network;  dw, db: [L,D] (deterministic). */
int ctr_cross_fwd(const float* x0, const float* w, const float* b, int B, int D, int L, float* xL, float* s,
                  ctr_stream_t stream);
size_t ctr_cross_bwd_workspace_bytes(int B, int D, int L);
int ctr_cross_bwd(const float* x0, const float* w, const float* b, const float* s, const float* dxL,
                  const float* dx_in, int B, int D, int L, float* dx0, float* dw, float* db, void* ws,
                  size_t ws_bytes, ctr_stream_t stream);

/* ---- K6/K9: DIN embedding + field-wise pooling layers (DIN.py:143-183) ---------------------------
 * gather_scale_rows: out[(i/G)*ld_group + (i%G)*K + k] = V[ids[i]][k] * (wgt ? wgt[i] : 1)
 *     (tf.nn.embedding_lookup of feat_ids / a_catids / padded behaviour ids, DIN.py:143-147,155-156;
 *      G, ld_group let the rows land directly inside the concatenated MLP input, DIN.py:199)
 * bag_sum: tf.nn.embedding_lookup_sparse(combiner="sum") over CSR bags (a_intids, DIN.py:148; the
 *     non-attention pooling branch :180-183) and its gradient g_rows[i] = d_out[bag(i)] * w_i
 * din_pool: att = sigmoid(z); u[b] = sum_p (ids[b,p] > 0) * att[b,p] * E[b,p,:]   (DIN.py:169-172)
 *     bwd: dE = mask*att*du (written, not accumulated); dz = mask*att*(1-att)*(E . du)
 * group_sum: dU[b] = sum_p dZ[b*P+p]   (gradient of ctr_fc_fwd_grouped's group bias)
 * scale_rows: out[i,:] = (x[(i/G)*ld_group + (i%G)*K : +K] + add[i,:]) * w[i]  (add, w optional)
 * axpby: out = alpha*a + beta*b */
int ctr_gather_scale_rows(const int32_t* ids, const float* wgt, const float* V, int64_t N, int64_t n, int K,
                          int G, int64_t ld_group, float* out, int32_t* oob, ctr_stream_t stream);
int ctr_bag_sum_fwd(const int32_t* ids, const float* wgt, const int32_t* offsets, const float* V, int64_t N,
                    int B, int K, int64_t ld, float* out, ctr_stream_t stream);
int ctr_bag_sum_bwd(const float* d_out, int64_t ld, const float* wgt, const int32_t* offsets, int B, int K,
                    float* g_rows, ctr_stream_t stream);
int ctr_scale_rows(const float* x, const float* add, const float* w, int64_t n, int K, int G, int64_t ld_group,
                   float* out, ctr_stream_t stream);
int ctr_din_pool_fwd(const float* E, const float* z, const int32_t* ids, int B, int P, int K, float* att, float* u,
                     int64_t ld_u, ctr_stream_t stream);
int ctr_din_pool_bwd(const float* E, const float* att, const int32_t* ids, const float* du, int64_t ld_u, int B,
                     int P, int K, float* dE, float* dz, ctr_stream_t stream);
int ctr_group_sum(const float* dZ, int B, int P, int N, float* dU, ctr_stream_t stream);
/* Attention unit backward through the N=1 output and the hidden layer's relu/dropout in ONE pass over the [B*P, H] hidden
 * activations Hh (autodiff of DIN.py:164-169): dZ[r][c] = dz[r]*w2[c] (*mask/keep) where Hh > 0; dU[b][c] = sum_p dZ
 * (the group-bias gradient; colsum(dU) is the layer's bias gradient); gw2_part[b][c] = sum_p dz*Hh (colsum = the output
 * layer's weight gradient).  Then ctr_fc_bwd(act = 2) takes dZ as is. */
int ctr_din_att_dz(const float* Hh, const float* drop_mask, float keep_prob, const float* dz, const float* w2, int B, int P,
                   int H, float* dZ, float* dU, float* gw2_part, ctr_stream_t stream);
/* out[k] = sum_r part[r*ld + k], k < ncols: deterministic column sums of many rows (fixed order) */
int ctr_colsum_rows(const float* part, int rows, int ld, int ncols, float* out, ctr_stream_t stream);
int ctr_axpby(const float* a, float alpha, const float* b, float beta, int64_t n, float* out, ctr_stream_t stream);

/* ---- K7/K8: all-pairs interactions (pair order i < j row-major; P = F(F-1)/2) -----------------------
 * pnn_product: z[b] = [x[b] (F*K) | inner (P)]            (outer = 0; PNN.py:148-153)
 *              z[b] = [x[b] (F*K) | outer (P*K*K)]        (outer = 1; PNN.py:164-167, "NOT ready yet")
 *   bwd: dX = dz[:, :F*K] + product-rule terms of the tail.
 * afm_pairs:  pw[b,p,:] = e_i * e_j (AFM.py:132-138);  bwd: dX[b,f,:] = sum_o dpw[b,pair(f,o),:] * e_o
 * afm_pool:   att = softmax_p(logit) (AFM.py:151), w = dropout(att) (:152-153), y_emb = sum_p w_p pw_p (:156);
 *   bwd: dpw = w * dy_emb (written), dlogit = softmax backward.
 * dropout_apply: out = x / keep * mask  (NFM's dropout on the bi-interaction vector, NFM.py:136-137;
 *   AFM's on y_emb, AFM.py:157-158) */
int ctr_pnn_product_fwd(const float* x, int B, int F, int K, int outer, float* z, ctr_stream_t stream);
int ctr_pnn_product_bwd(const float* x, const float* dz, int B, int F, int K, int outer, float* dX,
                        ctr_stream_t stream);
int ctr_afm_pairs_fwd(const float* x, int B, int F, int K, float* pw, ctr_stream_t stream);
int ctr_afm_pairs_bwd(const float* x, const float* dpw, int B, int F, int K, float* dX, ctr_stream_t stream);
int ctr_afm_pool_fwd(const float* pw, const float* logit, const float* mask, float keep, int B, int P, int K,
                     float* att, float* y_emb, ctr_stream_t stream);
int ctr_afm_pool_bwd(const float* pw, const float* att, const float* mask, float keep, const float* dy_emb, int B,
                     int P, int K, float* dpw, float* dlogit, ctr_stream_t stream);
int ctr_dropout_apply(const float* x, const float* mask, float keep, int64_t n, float* out, ctr_stream_t stream);

/* ---- row-sharded table routing (not in the reference) ---------------------------------
 * owner(id) = id % G, local row = id / G.  bucket_ids: the *n_uniq sorted unique ids of a batch are
 * grouped by owner: counts[G]; order[pos] = index into uniq; pos_of[u] = pos; local_ids[pos] = id / G
 * (bucket-major: the send buffer of the id all-to-all).  remap_ids: out[i] = pos_of[inverse[i]] turns
 * the batch's ids into indices of the received row cache.  gather_scalar: out[i] = W[ids[i]]. */
int ctr_a2a_bucket_ids(const int32_t* uniq, const int32_t* n_uniq, int64_t n_max, int G, int32_t* counts,
                       int32_t* cursor, int32_t* order, int32_t* pos_of, int32_t* local_ids, ctr_stream_t stream);
int ctr_remap_ids(const int32_t* inverse, const int32_t* pos_of, int64_t n, int32_t* out, ctr_stream_t stream);
/* Composite routing keys (what tf_repos_b200/sharded.py uses): key(id) = (id % G) * ceil(N/G) + id / G, so that ONE
 * ctr_unique_segment of the keys yields the unique ids in bucket order (owner-major, ascending id inside an owner),
 * `inverse` = the occurrence's position in the received row cache, and perm / seg_offsets = the gradient segments in
 * cache order.  ids outside [0, N) are counted in oob (may be NULL) and routed to row 0.
 * shard_split: counts[o] = unique ids owned by rank o, local_ids[u] = local row of the u-th unique key. */
int ctr_shard_keys(const int32_t* ids, int64_t n, int64_t N, int G, int32_t* keys, int32_t* oob, ctr_stream_t stream);
int ctr_shard_split(const int32_t* uniq_keys, const int32_t* n_uniq, int64_t n_max, int64_t N, int G, int32_t* counts,
                    int32_t* local_ids, ctr_stream_t stream);
int ctr_gather_scalar(const int32_t* ids, const float* W, int64_t N, int64_t n, float* out, ctr_stream_t stream);

/* ---- wide_n_deep feature columns (wide_n_deep.py:92-107) ------------------------------------
 * categorical_column_with_identity(num_buckets=NB, default_value=0) + embedding_column(K) per categorical
 * column, numeric columns appended in the name-sorted order (num_perm), and the linear_model over the same
 * columns.  The Fc per-column tables are stacked: column f owns rows [f*NB, (f+1)*NB) of emb [Fc*NB, K] and
 * wide_cat [Fc*NB]; flat_ids[b,f] = f*NB + (0 <= id < NB ? id : 0).
 *   x   [B, Fc*K + Fd] = [emb rows of columns 0..Fc-1 | dense[b, num_perm[0..Fd-1]]]      (emb != NULL)
 *   lin [B] = sum_f wide_cat[flat_ids[b,f]] + sum_j dense[b,j]*wide_num[j] + wide_bias     (wide_* != NULL)
 * bwd: g_rows[b*Fc+f,:] = dX[b, f*K:(f+1)*K], g_cat[b*Fc+f] = dy[b], g_num[j] = sum_b dy[b]*dense[b,j],
 *      g_bias = sum_b dy[b]  (fixed reduction trees: deterministic).  NULL outputs are skipped. */
int ctr_wd_input_fwd(const int32_t* ids, const float* dense, const float* emb, const float* wide_cat,
                     const float* wide_num, const float* wide_bias, const int32_t* num_perm, int B, int Fc, int Fd,
                     int NB, int K, int32_t* flat_ids, float* x, float* lin, ctr_stream_t stream);
int ctr_wd_input_bwd(const float* dX, const float* dy, const float* dense, int B, int Fc, int Fd, int K, float* g_rows,
                     float* g_cat, float* g_num, float* g_bias, ctr_stream_t stream);

/* ---- libsvm input (HOST buffers) -------------------------------------------------------------------
 * decode_libsvm of input_fn (DeepFM.py:65-81): "<label> <id>:<val> ..." lines -> ids int32 [rows,F],
 * vals f32 [rows,F], labels f32 [rows].  Parses complete lines of buf_host[0,len) up to max_rows;
 * returns rows parsed (or <0), *consumed_host = bytes consumed.  All pointers are HOST pointers. */
int64_t ctr_parse_libsvm(const char* buf_host, size_t len, int F, int64_t max_rows, int final_chunk,
                         int32_t* ids_host, float* vals_host, float* labels_host, size_t* consumed_host);
int ctr_libsvm_count_fields(const char* buf_host, size_t len);

/* ---- libsvm input (DEVICE buffers) -----------------------------------------------------
 * The same decode_libsvm (DeepFM.py:65-81) for text that is already in device memory: the first max_rows
 * complete lines of text[0,len) (plus an unterminated last line when final_chunk != 0) are tokenised by one
 * thread each.  Every value this path emits has exactly the bits ctr_parse_libsvm (strtof/strtol) gives;
 * whatever it cannot guarantee is only COUNTED and the caller re-parses the chunk with the host entry point:
 *   info (device int64[5]) = { rows, bytes consumed, blank lines, malformed lines (bad token / pair count != F),
 *                             lines holding a number for the host (inf/nan/hex, > 15 digits, |exp10| > 22,
 *                             fp32 subnormal/overflow, or within one double-ulp of an fp32 rounding boundary) }
 * ids/vals/labels rows are valid iff info[2] == info[3] == info[4] == 0.  len < 2^32. */
size_t ctr_parse_libsvm_device_workspace_bytes(size_t len, int64_t max_rows);
int ctr_parse_libsvm_device(const char* text, size_t len, int F, int64_t max_rows, int final_chunk, int32_t* ids,
                            float* vals, float* labels, int64_t* info, void* ws, size_t ws_bytes, ctr_stream_t stream);

/* ---- table initialisation (glorot_normal_initializer, DeepFM.py:115-116; truncated at 2 sigma) --- */
int ctr_init_trunc_normal(float* t, int64_t n, float stddev, uint64_t seed, ctr_stream_t stream);
int ctr_fill(float* t, int64_t n, float value, ctr_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CTR_B200_H_ */
