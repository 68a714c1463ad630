/*
 * nplda.h -- C ABI of libnplda.so: B200 (sm_100a) kernels for the pairwise
 * trial-scoring hot path of iiscleap/NeuralPlda.
 *
 * The reference has no FFI: the path sits behind a Python torch.nn.Module API
 * (/root/reference/utils/models.py).  Each entry point below names the
 * reference interface it replaces; neuralplda_b200/models.py binds them with
 * ctypes and keeps the reference's class / method signatures.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer on the current device unless its name
 *    ends in _host; fp32 tensors are contiguous row-major; the caller owns all
 *    buffers (PyTorch allocations); the library keeps no global mutable state
 *    and allocates nothing that outlives a call except explicit workspaces
 *    the caller passes in (one exception: NPLDA_IMPL_TC_F8 keeps an 8 KB ring of
 *    range-guard flags per device, allocated on first use).
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it
 *    and nothing synchronises unless stated.
 *  - return value: 0 = OK; > 0 = a cudaError_t; < 0 = NPLDA_ERR_*.
 *    nplda_error_string() renders either.  Nothing throws or exits.
 *  - there is no CPU fallback anywhere in this library.
 */
#ifndef NPLDA_H_
#define NPLDA_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPLDA_OK 0
#define NPLDA_ERR_BAD_ARG (-1)        /* null pointer, negative size, misaligned buffer     */
#define NPLDA_ERR_UNSUPPORTED_DIM (-2) /* layer widths beyond what the kernels are built for */
#define NPLDA_ERR_WORKSPACE (-3)      /* workspace too small                                 */
#define NPLDA_ERR_NO_DEVICE (-4)      /* no sm_100 device / kernel image not loadable        */
#define NPLDA_ERR_IO (-5)             /* file cannot be opened / read / written               */
#define NPLDA_ERR_FORMAT (-6)         /* trial file rows with differing numbers of fields     */

/* kernel selection for the score kernels */
#define NPLDA_IMPL_AUTO 0   /* the tcgen05 kernel when the shape allows (NPLDA_IMPL_TC_BF16 arithmetic for the score-only
                               entries -- the fastest, ~1e-5 of the fp32 reference; NPLDA_IMPL_TC arithmetic for the training
                               forwards, whose activations feed the gradients), else SIMT */
#define NPLDA_IMPL_SIMT 1   /* fp32 FFMA2 register-tiled kernel (any supported shape) */
#define NPLDA_IMPL_TC 2     /* tcgen05 kernel, both layers as split fp16 ("fp16x3": hi*hi + lo*hi + hi*lo, weights scaled into
                               range at pack time; inputs outside fp16's range are detected on the device and the call is
                               recomputed by the bf16x3 kernel on the same stream); error if shape unsupported */
#define NPLDA_IMPL_TC_BF16 4 /* tcgen05 kernel, both layers as split bf16 (any fp32 range, ~8x the rounding error of fp16x3) */
#define NPLDA_IMPL_TC_PAIR 5 /* NPLDA_IMPL_TC_BF16 arithmetic by CTA pairs (tcgen05.mma.cta_group::2, csrc/score_tcp.cu): each CTA of a
                               cluster converts its own 128 rows and holds half of the weight rows.  What NPLDA_IMPL_AUTO takes
                               for materialised NeuralPlda pairs from 9 472 pairs on (one 64-pair tile per SM); same scores */
#define NPLDA_IMPL_TC_PAIR_F8 6 /* the CTA-pair kernel with NPLDA_IMPL_TC_F8's layer-1 arithmetic (fp16 + two e4m3 products: four MMAs
                               per K = 32 instead of six); inputs outside its range are recomputed by NPLDA_IMPL_TC_PAIR on the stream */
#define NPLDA_IMPL_TC_F8 3  /* tcgen05 kernel, layer 1 as fp16*fp16 + two e4m3*e4m3 correction products on the same
                               accumulator (same MAC count, 2/3 of the MMA instructions).  Inputs outside the range the
                               e4m3 terms cover (typical |x| in [2^-3, 2^8)) are detected on the device and the call is
                               recomputed by the NPLDA_IMPL_TC kernel on the same stream, without a host round trip.
                               Opt-in: measured no faster than NPLDA_IMPL_TC on B200. */

/* loss ids, matching the reference's case-sensitive `lossfn` strings (models.py:395-399) */
#define NPLDA_LOSS_SOFTCDET 0      /* 'SoftCdet'      */
#define NPLDA_LOSS_CROSSENTROPY 1  /* 'crossentropy'  */

#define NPLDA_MAX_BETAS 8

int nplda_version(void);
const char *nplda_error_string(int code);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
int64_t nplda_launch_count(void);

/* ---------------------------------------------------------------------------
 * Packed weights.  The score kernels consume the affine layers transposed
 * (k-major), zero-padded and, for the tensor-core path, split into bf16 hi/lo
 * images in the tcgen05 shared-memory layout.  Packing is three small kernels
 * (~6 us); it must be re-run whenever the parameters may have changed -- the
 * Python layer re-packs on every call, because neither `.data.copy_()` (the
 * reference's own idiom, models.py:449-457, :420) nor fused optimisers bump
 * tensor._version.  nplda_pack_bytes() gives the workspace size for given dims;
 * ZERO the workspace once when it is allocated.
 * flags: NPLDA_PACK_MIXED also builds the layer-1 images of NPLDA_IMPL_TC_F8 / NPLDA_IMPL_TC_PAIR_F8
 * (two more kernels; without it that impl falls back to NPLDA_IMPL_TC on the
 * device); NPLDA_PACK_EPOCH_ODD selects which of the two fingerprint slots this
 * call fills -- alternate it between successive packs of one workspace.  The
 * fingerprint (a 64-bit content hash of the packed parameters) is what
 * nplda_table_prepare(NPLDA_PREPARE_IF_CHANGED) validates a cached row table
 * against, on the device.
 * ------------------------------------------------------------------------- */
#define NPLDA_PACK_MIXED 1
#define NPLDA_PACK_EPOCH_ODD 2
#define NPLDA_PACK_PAIR 4      /* also build the CTA-pair weight images nplda_score_fwd_split reads (two more kernels) */
int64_t nplda_pack_bytes(int d_in, int d1, int d2);

/* NeuralPlda parameters (models.py:349-363): W1 [d1,d_in], b1 [d1], W2 [d2,d1],
 * b2 [d2], P_sqrt [d2], Q [d2]. */
int nplda_pack_weights(const float *W1, const float *b1, const float *W2, const float *b2,
                       const float *p_sqrt, const float *q, int d_in, int d1, int d2,
                       void *pack, int64_t pack_bytes, int flags, void *stream);

/* DPlda parameters (models.py:464-476): W1 [d1,d_in], b1 [d1],
 * logistic_regres.weight [2*d1*d1+d1], logistic_regres.bias [1]. */
int dplda_pack_weights(const float *W1, const float *b1, const float *w_lr, const float *c_lr,
                       int d_in, int d1, void *pack, int64_t pack_bytes, int flags, void *stream);

/* ---------------------------------------------------------------------------
 * K1: fused score forward.
 * Replaces NeuralPlda.forward (models.py:378-382 = extract_plda_embeddings
 * 366-370 twice + forward_from_plda_embeddings 372-376):
 *   a = W1 x + b1 ; u = a / max(||a||, 1e-12) ; y = W2 u + b2
 *   S = sum_k Q_k y1_k^2 + Q_k y2_k^2 + 2 P_sqrt_k^2 y1_k y2_k
 * x1, x2: [n, d_in] fp32, 16-byte aligned rows when d_in % 4 == 0.
 * scores: [n] fp32.  n == 0 is a no-op.
 * ------------------------------------------------------------------------- */
int nplda_score_fwd(const float *x1, const float *x2, int64_t n, int d_in, int d1, int d2,
                    const void *pack, float *scores, int impl, void *stream);

/* Replaces DPlda.forward (models.py:491-495; closed form of 483-489):
 *   u = normalize(W1 x + b1)
 *   S = u1^T(Wb+Wb^T)u2 + u1^T Ww u1 + u2^T Ww u2 + ws.(u1+u2) + c          */
int dplda_score_fwd(const float *x1, const float *x2, int64_t n, int d_in, int d1,
                    const void *pack, float *scores, int impl, void *stream);
/* The entry the module uses: for 512-170-like shapes NPLDA_IMPL_AUTO evaluates layer 1 and both 170 x 170 forms on the
 * tensor cores in ONE pass over x (the DPL form of the tcgen05 score kernel: Y_P = a Pm^T, then Y_R = a R^T into the same
 * accumulator, a re-read from shared memory); for other shapes it is dplda_score_fwd.  The kernel needs no workspace:
 * dplda_fwd_workspace_bytes returns 0 and `workspace` may be NULL (both kept so that callers written against the
 * two-pass version keep working). */
int64_t dplda_fwd_workspace_bytes(int64_t n, int d_in, int d1);
int dplda_score_fwd_ws(const float *x1, const float *x2, int64_t n, int d_in, int d1, const void *pack,
                       float *scores, int impl, void *workspace, int64_t workspace_bytes, void *stream);

/* The two half-steps of forward, which the reference exposes as public methods:
 * embeddings (models.py:366-370: y = W2 normalize(W1 x + b1) + b2, [n, d2]; DPlda 478-481:
 * u = normalize(W1 x + b1), [n, d1]) and the score from materialised embeddings (372-376 / 483-489).
 * Not on the trial-scoring hot path (forward never materialises embeddings); fp32 SIMT kernels. */
int nplda_embed_fwd(const float *x, int64_t n, int d_in, int d1, int d2, const void *pack, float *emb,
                    int is_dplda, void *stream);
int nplda_score_from_embeddings(const float *y1, const float *y2, int64_t n, int d2, const float *p_sqrt,
                                const float *q, float *scores, void *stream);
int dplda_score_from_embeddings(const float *u1, const float *u2, int64_t n, int d_in, int d1,
                                const void *pack, float *scores, void *stream);

/* Same scores with the pair gather fused in: row i scores table[idx1[i]] vs
 * table[idx2[i]].  Replaces load_xvec_trials_from_numbatch
 * (sv_trials_loaders.py:418-426) + forward.  table: [n_rows, d_in] fp32;
 * idx1, idx2: [n] int64 in [0, n_rows).  Out-of-range indices are an error
 * reported through *bad_index_flag (device int32, set non-zero), never a fault. */
int nplda_score_fwd_indexed(const float *table, int64_t n_rows, const int64_t *idx1,
                            const int64_t *idx2, int64_t n, int d_in, int d1, int d2,
                            const void *pack, float *scores, int32_t *bad_index_flag, int impl,
                            void *stream);
int dplda_score_fwd_indexed(const float *table, int64_t n_rows, const int64_t *idx1,
                            const int64_t *idx2, int64_t n, int d_in, int d1, const void *pack,
                            float *scores, int32_t *bad_index_flag, int impl, void *stream);

/* ---------------------------------------------------------------------------
 * K1x: the same scores from a PRE-SPLIT table, on the tensor cores (csrc/score_tcx.cu).
 * The reference's callers hold a static x-vector table (the pickled dict,
 * xvector_NeuralPlda_pytorch.py:117) and index batches (sv_trials_loaders.py:418-426).
 * nplda_table_split converts the fp32 table ONCE into the tensor cores' operand format
 * (fp16 hi/lo of x 2^kx, kx from the table's absolute maximum: x 2^kx = hi + lo + O(2^-22 |x|);
 * nplda_split_bytes(n_rows, d_in) = n_rows * d_in * 4 + 128 bytes, 128-byte aligned, d_in % 32 == 0, rows 16-byte
 * aligned); nplda_score_fwd_split then scores trials (table[idx1[k]], table[idx2[k]]): TMA row gather
 * (tile::gather4) straight into the shared-memory A operand, tcgen05.mma.cta_group::2 over CTA pairs, no
 * conversion work per call.  Replaces load_xvec_trials_from_numbatch (sv_trials_loaders.py:418-426) +
 * NeuralPlda.forward (models.py:378-382) for index batches.  Needs a pack built with NPLDA_PACK_PAIR and
 * d_in % 32 == 0, d_in >= 64, d1, d2 <= 176 (NPLDA_ERR_UNSUPPORTED_DIM otherwise: use nplda_score_fwd_indexed).
 * Out-of-range indices set *bad_index_flag (device int32) and never fault.
 * ------------------------------------------------------------------------- */
int64_t nplda_split_bytes(int64_t n_rows, int d_in);
int nplda_table_split(const float *table, int64_t n_rows, int d_in, void *split, void *stream);
int nplda_score_fwd_split(const void *split, int64_t n_rows, const int64_t *idx1, const int64_t *idx2, int64_t n,
                          int d_in, int d1, int d2, const void *pack, float *scores, int32_t *bad_index_flag,
                          void *stream);

/* Batch gather alone (the device half of load_xvec_trials_from_numbatch / _from_idbatch,
 * sv_trials_loaders.py:418-437, for callers that want the materialised [n, d] pair the
 * reference's training loop passes to forward): x1[t] = table[idx1[t]], x2[t] = table[idx2[t]].
 * Rows outside [0, n_rows) set *bad_index_flag (device or pinned host memory) and are zero-filled. */
int nplda_gather_pairs(const float *table, int64_t n_rows, int d, const int64_t *idx1, const int64_t *idx2,
                       int64_t n, float *x1, float *x2, int32_t *bad_index_flag, void *stream);

/* ---------------------------------------------------------------------------
 * Trial-list scoring with every utterance transformed ONCE.
 * Replaces the per-batch gather + double forward of the scoring loop
 * (scorefile_generator.py:29-36 / 46-53 with sv_trials_loaders.py:418-437 and
 * models.py:378-382 / 491-495) for a trial list over a device-resident x-vector
 * table: both models' scores split as S(i,j) = r[i] + r[j] + A[i] . B[j] with
 * per-utterance rows (NeuralPlda: A = y, B = 2 P_sqrt^2 y, r = sum Q y^2;
 * DPlda: A = u, B = (Wb + Wb^T) u, r = u^T Ww u + ws.u + c/2).
 * nplda_table_prepare: table [n_rows, d_in] fp32 -> rowtab (nplda_rowtab_bytes(n_rows)
 * bytes, caller-owned); re-run when the table or the parameters change.  flags:
 * NPLDA_PACK_EPOCH_ODD as given to the pack call that filled `pack`;
 * NPLDA_PREPARE_IF_CHANGED makes the call a device-side no-op (three empty
 * launches) when rowtab was built by an earlier call from a pack with the same
 * parameter fingerprint -- pass it when the same table is scored again, so that
 * "pack + prepare" can precede every scoring call.
 * nplda_score_pairs: scores[t] = S(idx1[t], idx2[t]); indices outside [0, n_rows)
 * set *bad_index_flag and score 0, never a fault.  Layer widths up to 175.
 * ------------------------------------------------------------------------- */
int64_t nplda_rowtab_bytes(int64_t n_rows);
#define NPLDA_PREPARE_IF_CHANGED 1
int nplda_table_prepare(const float *table, int64_t n_rows, int d_in, int d1, int d2, const void *pack,
                        int is_dplda, float *rowtab, int flags, void *stream);
int nplda_score_pairs(const float *rowtab, int64_t n_rows, const int64_t *idx1, const int64_t *idx2,
                      int64_t n, float *scores, int32_t *bad_index_flag, void *stream);

/* Dense trial lists (key files written enrol-major, every model against the same segments): when the list covers a
 * sizeable fraction of (its enrol rows) x (its test rows), one grid product over those rows (nplda_score_grid) and a
 * 4-byte gather per trial replace the per-trial row gather of nplda_score_pairs (scorefile_generator.py:29-36 / 46-53
 * gathers and transforms both sides of every trial).
 *   nplda_trial_rows: flags [2][n_rows] int32 (scratch), pos [2][n_rows] int32 (rank of a row among the rows the list
 *     uses on that side, -1 if unused), list [2][n_rows] int64 (the used rows, ascending), counts [2] int32; side 0 =
 *     idx1 (enrol), side 1 = idx2 (test).  The caller reads `counts` to size the grid [counts[0]][counts[1]].
 *   nplda_trial_grid_gather: scores[t] = grid[pos[0][idx1[t]] * ld + pos[1][idx2[t]]]  (0 for rows outside the table,
 *     which nplda_trial_rows reported through bad_index_flag). */
int nplda_trial_rows(const int64_t *idx1, const int64_t *idx2, int64_t n, int64_t n_rows, int32_t *flags, int32_t *pos,
                     int64_t *list, int32_t *counts, int32_t *bad_index_flag, void *stream);
int nplda_trial_grid_gather(const float *grid, int64_t ld, const int32_t *pos, int64_t n_rows, const int64_t *idx1,
                            const int64_t *idx2, int64_t n, float *scores, void *stream);
/* nplda_score_grid: the full enrol x test grid of a trial list (BASELINE.json configs[2]/[3];
 * the id x cohort matrix utils/adaptive_score_normalization.py:32 reads) as one
 * [n_enrol,176] x [176,n_test] fp32 product over the row table:
 * scores[i * ld_scores + j] = S(enrol_rows[i], test_rows[j]), ld_scores >= n_test
 * (enrol-major trial order).  4 bytes of HBM traffic per trial instead of 20 + two
 * row gathers.  Rows outside [0, n_rows) set *bad_index_flag; their row / column of
 * the grid scores 0. */
int nplda_score_grid(const float *rowtab, int64_t n_rows, const int64_t *enrol_rows, int64_t n_enrol,
                     const int64_t *test_rows, int64_t n_test, float *scores, int64_t ld_scores,
                     int32_t *bad_index_flag, void *stream);
/* The same with the kernel chosen: NPLDA_IMPL_AUTO / NPLDA_IMPL_TC = the tcgen05 kernel (csrc/grid_tc.cu: rows gathered by
 * TMA from the fp16 hi/lo grid operands nplda_table_prepare leaves behind the rows, r[i] + r[j] folded into the
 * contraction), NPLDA_IMPL_SIMT = the fp32 FFMA2 kernel. */
int nplda_score_grid_impl(const float *rowtab, int64_t n_rows, const int64_t *enrol_rows, int64_t n_enrol,
                          const int64_t *test_rows, int64_t n_test, float *scores, int64_t ld_scores,
                          int32_t *bad_index_flag, int impl, void *stream);

/* ---------------------------------------------------------------------------
 * K2: loss / detection-cost accumulators.
 * One pass over (scores, labels) producing the raw fp64 sums every loss of the
 * reference is built from.  acc layout (4*K + 4 doubles), ADDED INTO (the
 * caller zeroes it; per-GPU shards are all-reduced as raw sums before
 * nplda_loss_finalize):
 *   for k < K:  acc[4k+0] = sum_i t_i     * sigmoid(alpha (th_k - s_i))   soft miss
 *               acc[4k+1] = sum_i (1-t_i) * sigmoid(alpha (s_i - th_k))   soft false alarm
 *               acc[4k+2] = sum_i t_i     * [s_i < th_k]                  hard miss   (cdet)
 *               acc[4k+3] = sum_i (1-t_i) * [s_i > th_k]                  hard false alarm
 *   acc[4K+0] = sum t, acc[4K+1] = sum (1-t),
 *   acc[4K+2] = sum_i BCE(sigmoid(s_i - th_xent), t_i)  (log clamped at -100)
 *   acc[4K+3] = n
 * thresholds: [K] fp32 device; th_xent: [1] fp32 device or NULL (= 0, DPlda).
 * Replaces the reductions inside softcdet (models.py:384-388), crossentropy
 * (390-393 / 503-506) and cdet (401-404).
 * ------------------------------------------------------------------------- */
int nplda_loss_accum(const float *scores, const float *labels, int64_t n, const float *thresholds,
                     int K, float alpha, const float *th_xent, double *acc, void *stream);

/* out[0] = softcdet, out[1] = crossentropy, out[2] = cdet (hard), all fp32,
 * computed on the device from (all-reduced) acc; betas_host: [K] doubles. */
int nplda_loss_finalize(const double *acc, const double *betas_host, int K, float *out,
                        void *stream);

/* dL/ds_i for loss_id, scaled by *grad_out (device fp32 scalar; NULL = 1), plus
 * the threshold gradients: dth[k] for SoftCdet (k < K), dth[K] for threshold_Xent.
 * dth is ADDED INTO (caller zeroes; all-reduce for multi-GPU).  acc must hold
 * the global label sums (after the all-reduce). */
int nplda_loss_bwd(const float *scores, const float *labels, int64_t n, const float *thresholds,
                   const double *betas_host, int K, float alpha, const float *th_xent,
                   const double *acc, int loss_id, const float *grad_out, float *dscores,
                   double *dth, void *stream);

/* ---------------------------------------------------------------------------
 * K3: score backward.  Recomputes the activations from x (nothing is saved by
 * the forward), and ADDS parameter gradients into fp32 buffers shaped like the
 * parameters (caller zeroes; all-reduce for multi-GPU).  Any gradient pointer
 * may be NULL to skip it.  workspace: nplda_bwd_workspace_bytes(n, ...) bytes.
 * Replaces autograd through models.py:366-376 / 478-489.
 * ------------------------------------------------------------------------- */
int64_t nplda_bwd_workspace_bytes(int64_t n, int d_in, int d1, int d2);
int nplda_score_bwd(const float *x1, const float *x2, int64_t n, int d_in, int d1, int d2,
                    const float *W1, const float *b1, const float *W2, const float *b2,
                    const float *p_sqrt, const float *q, const float *dscores, float *dW1,
                    float *db1, float *dW2, float *db2, float *dp_sqrt, float *dq, float *dx1,
                    float *dx2, void *workspace, int64_t workspace_bytes, void *stream);
int dplda_score_bwd(const float *x1, const float *x2, int64_t n, int d_in, int d1,
                    const float *W1, const float *b1, const float *w_lr, const float *dscores,
                    float *dW1, float *db1, float *dw_lr, float *dc_lr, float *dx1, float *dx2,
                    void *workspace, int64_t workspace_bytes, void *stream);

/* Training with saved activations.  nplda_score_fwd_train / dplda_score_fwd_train compute the scores like
 * nplda_score_fwd / dplda_score_fwd_ws and also leave, in the caller-owned buffer `act` of
 * nplda_act_floats(n, is_dplda) floats, the rows the backward needs (NeuralPlda: a = W1 x + b1 and y;
 * DPlda: a, (Ww + Ww^T) u and (Wb + Wb^T) u; each [2 n][176] fp32, side 1 of pair p n rows after side 0).
 * Passing that buffer to nplda_score_bwd_act / dplda_score_bwd_act (otherwise identical to the functions above;
 * act == NULL makes them the same) saves the backward its passes of the tensor-core forward kernel.  `act` is
 * only read.  NPLDA_ERR_UNSUPPORTED_DIM for shapes the tcgen05 kernel does not take: use the plain entries. */
int64_t nplda_act_floats(int64_t n, int kind);   /* kind: 0 NeuralPlda, 1 DPlda (trainable LDA), 2 DPlda rows u only */
int nplda_score_fwd_train(const float *x1, const float *x2, int64_t n, int d_in, int d1, int d2, const void *pack,
                          float *scores, float *act, void *stream);
int dplda_score_fwd_train(const float *x1, const float *x2, int64_t n, int d_in, int d1, const void *pack,
                          float *scores, float *act, void *stream);
/* DPlda with the LDA frozen -- what the reference's driver trains (xvector_DPlda_pytorch.py:140-147 sets
 * requires_grad = False on centering_and_LDA): the forward keeps only the normalised rows u = a / |a|
 * (urows: nplda_act_floats(n, 2) floats, [2 n][176], side 1 of pair p n rows after side 0) and the gradient of
 * logistic_regres (autograd through models.py:483-489) is two contractions over the batch on the tensor cores,
 *   dWw = (Ps + Pd) / 2, dWb = (Ps - Pd) / 2,  Ps = sum_p g_p s_p s_p^T, Pd = sum_p g_p d_p d_p^T,  s = u1 + u2, d = u1 - u2,
 *   dws = sum_p g_p s_p, dc = sum_p g_p   (g = dscores), ADDED into dw_lr [2 d1^2 + d1] and dc_lr [1] (either may be NULL). */
int dplda_score_fwd_train_u(const float *x1, const float *x2, int64_t n, int d_in, int d1, const void *pack,
                            float *scores, float *urows, void *stream);
int64_t dplda_lr_bwd_workspace_bytes(int64_t n, int d1);     /* caller-owned, 256-byte aligned */
int dplda_lr_bwd(const float *urows, int64_t n, int d1, const float *dscores, float *dw_lr, float *dc_lr,
                 void *workspace, int64_t workspace_bytes, void *stream);
int nplda_score_bwd_act(const float *x1, const float *x2, int64_t n, int d_in, int d1, int d2,
                        const float *W1, const float *b1, const float *W2, const float *b2,
                        const float *p_sqrt, const float *q, const float *dscores, float *dW1,
                        float *db1, float *dW2, float *db2, float *dp_sqrt, float *dq, float *dx1,
                        float *dx2, const float *act, void *workspace, int64_t workspace_bytes, void *stream);
int dplda_score_bwd_act(const float *x1, const float *x2, int64_t n, int d_in, int d1,
                        const float *W1, const float *b1, const float *w_lr, const float *dscores,
                        float *dW1, float *db1, float *dw_lr, float *dc_lr, float *dx1, float *dx2,
                        const float *act, void *workspace, int64_t workspace_bytes, void *stream);

/* ---------------------------------------------------------------------------
 * minC threshold sweep.  Replaces the O(N_t * N) Python loop of
 * NeuralPlda.minc (models.py:406-421) given the two score populations already
 * sorted ascending: for every target score s_j, pmiss_j = (count(tgt < s_j) - 1)
 * or 1 when the count is 0, pfa_j = (count(non >= s_j) - 1) or 1 when 0, both in
 * fp32 divided by n_t / n_n as fp32, cdet = pmiss + beta_k * pfa, first argmin.
 * out_min: [K] fp32, out_arg: [K] int64 (index into tgt_sorted).
 * ------------------------------------------------------------------------- */
int nplda_minc_sweep(const float *tgt_sorted, int64_t n_t, const float *non_sorted, int64_t n_n,
                     float sum_t, float sum_n, const double *betas_host, int K, float *out_min,
                     int64_t *out_arg, void *stream);

/* The two sorted score populations the sweep takes (models.py:407-408: torch.sort(output[target > 0.5]) and
 * torch.sort(output[target < 0.5])), hand-written:
 *   nplda_split_by_label: scores with label > 0.5 -> tgt, < 0.5 -> non (capacity n each, order unspecified),
 *     counts[0] / counts[1] = how many (device uint64 pair, zeroed by the call);
 *   nplda_sort_f32: in-place ascending sort of n floats (bitonic network, NaN last like torch.sort). */
int nplda_split_by_label(const float *scores, const float *labels, int64_t n, float *tgt, float *non,
                         unsigned long long *counts, void *stream);
int nplda_sort_f32(float *keys, int64_t n, void *stream);

/* ---------------------------------------------------------------------------
 * Host-buffer entry (what a non-PyTorch caller binds, and what bench.py's e2e
 * leg times): x1_host/x2_host are host buffers (pinned for full speed), scores
 * come back in scores_host.  Copies are chunked and overlapped with the score
 * kernel on internal streams; the call returns after the last D2H completes.
 * pack is a device pointer from nplda_pack_weights.  dev_scratch: device buffer
 * of nplda_host_scratch_bytes(chunk_pairs, d_in) bytes.
 * ------------------------------------------------------------------------- */
int64_t nplda_host_scratch_bytes(int64_t chunk_pairs, int d_in);
int nplda_score_fwd_host(const float *x1_host, const float *x2_host, int64_t n, int d_in, int d1,
                         int d2, const void *pack, float *scores_host, int64_t chunk_pairs,
                         void *dev_scratch, int64_t dev_scratch_bytes, int is_dplda, int impl);

/* ---------------------------------------------------------------------------
 * Cohort score normalisation: the arithmetic of the reference's
 * utils/adaptive_score_normalization.py on the device, in float64 like the script.
 * nplda_cohort_stats: scores [m, c] fp32 row-major (one row per enrol / test id, its
 * scores against the c cohort utterances, e.g. an id x cohort grid from
 * nplda_score_pairs) -> stats [m, 4] doubles: mean, std (population, np.std) over
 * the row (:33-34), and mean, std over the top_n LOWEST scores of the row -- the
 * first top_n of the ascending sort, as :32/:35-36 compute them (all c if c < top_n).
 * nplda_score_norm: for trial i with rows e = enrol_row[i], t = test_row[i] of stats
 * (:62-66): out[0*n+i] = znorm, out[1*n+i] = tnorm, out[2*n+i] = snorm,
 * out[3*n+i] = asnorm1.  Rows outside [0, m) set *bad_index_flag.
 * ------------------------------------------------------------------------- */
int nplda_cohort_stats(const float *scores, int64_t m, int64_t c, int64_t top_n, double *stats, void *stream);
int nplda_score_norm(const float *raw, const int64_t *enrol_row, const int64_t *test_row, int64_t n,
                     const double *stats, int64_t m, double *out, int32_t *bad_index_flag, void *stream);

/* ---------------------------------------------------------------------------
 * Host-side text I/O either side of the path; no device work.
 * Trial-list reader: replaces np.genfromtxt(file, dtype='str') + the per-row
 * dict lookups of sv_trials_loaders.py:376-383 / 399-406 and
 * scorefile_generator.py:26 / 45.  Fields are split on runs of blanks, '#' starts
 * a comment, empty lines are skipped, all rows must have the same field count
 * (else NPLDA_ERR_FORMAT, genfromtxt's ValueError).
 * nplda_trials_map_ids: column `col` of rows [first_row, rows) looked up in a table
 * of n_ids ids (one buffer, every id terminated by '\n'; values[r] or r when
 * values == NULL; later duplicates win).  mode 0: field as is; 1: os.path.splitext
 * root (sv_trials_loaders.py:403); 2: basename + splitext (:432).  Unknown -> -1.
 * nplda_trials_col_float: Python float() of a column (labels), ok[i] = 0 where it
 * does not parse (the reference drops such rows, :379-383).
 * nplda_scores_write: rows [first_row, rows): first ncols_keep fields + the score as
 * numpy prints a float32 (str(np.float32(s)): shortest round-trip digits), tab
 * separated, '\n' terminated, after an optional header line -- the files of
 * scorefile_generator.py:37-38 (sre: all columns, header + "\tLLR") and :54-55
 * (voices: 2 columns, no header), byte for byte.
 * ------------------------------------------------------------------------- */
typedef struct nplda_trials nplda_trials;
int nplda_trials_open(const char *path, nplda_trials **out);
void nplda_trials_close(nplda_trials *t);
int64_t nplda_trials_rows(const nplda_trials *t);
int nplda_trials_cols(const nplda_trials *t);
int64_t nplda_trials_field(const nplda_trials *t, int64_t row, int col, const char **ptr);   /* length; not NUL-terminated */
int nplda_trials_map_ids(const nplda_trials *t, int col, int mode, const char *ids, int64_t ids_bytes,
                         int64_t n_ids, const int64_t *values, int64_t first_row, int64_t *idx_out);
int nplda_trials_col_float(const nplda_trials *t, int col, int64_t first_row, float *out, uint8_t *ok);
int nplda_scores_write(const char *path, const nplda_trials *t, int64_t first_row, int ncols_keep,
                       const float *scores, const char *header_line);
int nplda_format_f32(float v, char *out24);   /* str(np.float32(v)) into out24 (not NUL-terminated); returns the length */

/* Test hook: forces the pieces of the score backward for A/B comparisons -- weight-gradient contraction (gemm),
 * activations from the tensor-core forward (emit), dL/du pass (du): 0 = automatic (the default), 1 = the all-fp32
 * piece, 2 = the tensor-core piece.  Process-wide; production code never calls it. */
void nplda_debug_backward_paths(int gemm, int emit, int du);

#ifdef __cplusplus
}
#endif
#endif /* NPLDA_H_ */
