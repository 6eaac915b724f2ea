/*
 * dae_sm100.h -- C ABI of libdae_sm100.so: the sm_100a kernels behind
 * DenoisingAutoencoder.fit / transform (DAE-with-triplet-loss training hot path).
 *
 * The reference (louislung/DAE_RNN_News_Recommendation) has NO FFI: its arithmetic is a
 * TensorFlow-1.12 graph run by `tf.Session.run` (autoencoder/autoencoder.py:233,241).  Each entry
 * point below replaces the TF ops cited next to it; INTEGRATION.md shows the ctypes stub a
 * maintainer of the reference would add.  All citations are relative to the reference root.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; sizes are element counts
 *   - `stream` is a cudaStream_t passed as void*; every call is asynchronous on that stream
 *   - return value: 0 = ok, <0 = DAE_ERR_*; dae_last_error() gives the message (thread local)
 *   - the library owns no persistent device memory; the caller (Python/torch) owns every buffer
 *   - CSR: indptr int64[N+1], indices int32[nnz] (sorted inside a row), values float32[nnz]
 *   - parameters: ONE flat fp32 buffer theta = [ W (F x H row-major) | bh (H) | bv (F) ]
 *     gradients / optimizer slots use the same flat layout
 *   - activations: 0 = identity ('none'), 1 = sigmoid, 2 = tanh      (autoencoder.py:380-387,402-409)
 *   - losses: 0 = cross_entropy, 1 = mean_squared, 2 = cosine_proximity (triplet_loss_utils.py:268-273)
 *   - strategies: 0 = none, 1 = batch_all, 2 = batch_hard            (autoencoder.py:70)
 *   - optimizers: 0 = gradient_descent, 1 = ada_grad, 2 = momentum, 3 = adam (autoencoder.py:451-472)
 */
#ifndef DAE_SM100_H
#define DAE_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DAE_OK 0
#define DAE_ERR_BAD_ARG (-1)
#define DAE_ERR_CUDA (-2)
#define DAE_ERR_UNSUPPORTED (-3)

#define DAE_ACT_NONE 0
#define DAE_ACT_SIGMOID 1
#define DAE_ACT_TANH 2

#define DAE_LOSS_CE 0
#define DAE_LOSS_MSE 1
#define DAE_LOSS_COSINE 2

#define DAE_TRIPLET_NONE 0
#define DAE_TRIPLET_BATCH_ALL 1
#define DAE_TRIPLET_BATCH_HARD 2

#define DAE_OPT_SGD 0
#define DAE_OPT_ADAGRAD 1
#define DAE_OPT_MOMENTUM 2
#define DAE_OPT_ADAM 3

/* per-step scalar slots written by the kernels (float64 each), see dae_step_finalize */
#define DAE_STAT_COST 0
#define DAE_STAT_AE_LOSS 1
#define DAE_STAT_TRIPLET_LOSS 2
#define DAE_STAT_FRACTION 3
#define DAE_STAT_NUM 4
#define DAE_STAT_SUM_W 5
#define DAE_STAT_N_VALID 6
#define DAE_STAT_SUM_LW 7
#define DAE_STAT_TRIPLET_SUM 8
#define DAE_STAT_N_ACTIVE 9
#define DAE_STAT_SLOTS 16

int dae_version(void);
/* copies the calling thread's last error message into buf (NUL terminated); returns its length */
int dae_last_error(char* buf, size_t len);

/* ---- batching -------------------------------------------------------------------------------
 * Replaces utils.gen_batches' per-batch fancy indexing + get_sparse_ind_val_shape
 * (autoencoder/utils.py:53-66,162-180) and the label-only parts of batch_all
 * (triplet_loss_utils.py:47-76,110-111,129): takes rows perm[offset : offset+B] of the epoch's
 * permutation, orders them by label (the loss is invariant to the order of rows in a batch),
 * and emits for each batch row its dataset row id, label, class segment [seg_lo, seg_hi) and -
 * for batch_all - the closed-form data weight w_i and N_valid.  strategy none: order kept, w = 1.
 * One CTA; B <= 4096.  stats: float64[DAE_STAT_SLOTS], zeroed here, SUM_W / N_VALID filled.
 * ctl (optional, device int64[4]): per-step cursors kept in device memory so that a captured CUDA graph of the
 * step can be replayed without host-side argument changes -- ctl[0] is added to `offset`, ctl[1] is the row of the
 * stats log dae_step_finalize writes, ctl[2] the 1-based optimizer step; dae_step_advance moves all three.
 */
int dae_step_advance(int64_t* ctl, int64_t row_stride, void* stream);
int dae_batch_prepare(const int32_t* perm, int64_t offset, const int64_t* ctl, int32_t B, const float* labels_all,
                      int32_t strategy, int32_t* rows_out, float* labels_out, int32_t* seg_lo,
                      int32_t* seg_hi, float* weight_out, double* stats, void* stream);
/* The one-CTA sort is 20 us of pure latency, so a graph-replayed step prepares the NEXT batch (cursor ctl[0] + stride) on a side
 * branch into staging buffers (*_s) while the current step computes -- nothing is written when that batch would run past
 * n_perm -- and the next step starts with dae_batch_commit, a copy of the staged batch into the live buffers. */
int dae_batch_prepare_next(const int32_t* perm, int64_t n_perm, int64_t stride, const int64_t* ctl, int32_t B,
                           const float* labels_all, int32_t strategy, int32_t* rows_s, float* labels_s,
                           int32_t* seg_lo_s, int32_t* seg_hi_s, float* weight_s, double* stats_s, void* stream);
int dae_batch_commit(int32_t B, const int32_t* rows_s, const float* labels_s, const int32_t* seg_lo_s,
                     const int32_t* seg_hi_s, const float* weight_s, const double* stats_s, int32_t* rows,
                     float* labels_b, int32_t* seg_lo, int32_t* seg_hi, float* weight, double* stats, void* stream);
/* explicit (org, pos, neg) triplets (autoencoder/utils.py:73-91, autoencoder_triplet.py:106-147): rows_out[3B] = the batch's
 * rows in the three blocks of the stacked [org; pos; neg] matrix (n_each rows per block); stats zeroed, SUM_W = B. */
int dae_batch_prepare_explicit(const int32_t* perm, int64_t offset, const int64_t* ctl, int32_t B, int64_t n_each,
                               int32_t* rows_out, double* stats, void* stream);

/* ---- K1: CSR x dense encode ---------------------------------------------------------------------
 * E[r,:] = f( in_scale * X[rows[r],:] . W + bh ) - f(bh)      (autoencoder.py:377,389; transform :494-497;
 * in_scale folds utils.decay_noise, utils.py:147-159).  rows == NULL means rows[r] = r.
 * E is written fp32 with leading dimension ldE.  Entries whose value is exactly 0 (masked) are skipped.
 * col_count (optional, int32[F]): zeroed, then receives the number of kept entries per feature column of the batch --
 * the bucket sizes dae_encode_csr_bwd_gather needs.
 * e_hi / e_lo (optional, bf16 [n_rows x ld_split]): columns [0, H) of the bf16 hi/lo operand copy of E for the tensor-core
 * GEMMs (the caller keeps the padding columns zero and the all-ones column set).
 */
int dae_encode_csr_fwd(const int64_t* indptr, const int32_t* indices, const float* values,
                       const int32_t* rows, int32_t n_rows, int32_t F, int32_t H, float in_scale,
                       const float* W, const float* bh, int32_t enc_act, float* E, int64_t ldE,
                       int32_t* col_count, void* e_hi, void* e_lo, int64_t ld_split, void* stream);

/* ---- K5: encode backward -------------------------------------------------------------------------
 * dA = dE * f'(A);  dbh = sum_i dA_i - f'(bh) * sum_i dE_i;  dW[c,:] += v * dA[r,:] for every stored
 * (r,c,v) of the corrupted batch (autodiff of autoencoder.py:389).  dE is overwritten with dA.
 * dW is accumulated with fp32 atomics on top of whatever the decode backward wrote.
 * dE_add (optional, same layout as dE): a second contribution to dL/dE, added before f' is applied -- the triplet term
 * alpha (G + G^T) E, computed on the mining branch of the step, meets the decode term here instead of in a GEMM of its own.
 * dbh_zeroed != 0: the caller has already zeroed dbh (it zeroes the whole flat gradient buffer), skip the memset node.
 */
int dae_encode_csr_bwd(const int64_t* indptr, const int32_t* indices, const float* values,
                       const int32_t* rows, int32_t n_rows, int32_t F, int32_t H, float in_scale,
                       const float* E, const float* bh, int32_t enc_act, float* dE, const float* dE_add,
                       int64_t ldE, float* dW, float* dbh, int32_t dbh_zeroed, void* stream);

/* Same result with ~6x fewer atomics on dW: the batch's kept entries are bucketed by feature column (col_count from the
 * forward call; col_start int32[F+1], col_cursor int32[F], ent_col/ent_row int32[cap], ent_val f32[cap] are
 * caller-provided scratch, cap >= kept entries of the batch); CTAs then walk fixed-size chunks of the bucketed entry
 * list, accumulate v * dA[r,:] in registers per column run and issue one vector red.global.add per (chunk, column) run.
 * Supports H <= 1024 (H % 4 == 0) / 512 (H % 2 == 0) / 256; larger H: use dae_encode_csr_bwd.
 */
/* exclusive scan of the per-column counts into col_start[F+1] / col_cursor[F]; dae_encode_csr_bwd_gather runs it itself unless
 * it is called with col_count == NULL (then the scan must already have been issued, e.g. on a parallel stream) */
int dae_col_scan(const int32_t* col_count, int32_t F, int32_t* col_start, int32_t* col_cursor, void* stream);
int dae_encode_csr_bwd_gather(const int64_t* indptr, const int32_t* indices, const float* values,
                              const int32_t* rows, int32_t n_rows, int32_t F, int32_t H, float in_scale,
                              const float* E, const float* bh, int32_t enc_act, float* dE, const float* dE_add,
                              int64_t ldE, float* dW, float* dbh, int32_t dbh_zeroed, const int32_t* col_count,
                              int32_t* col_start, int32_t* col_cursor, int32_t* ent_col, int32_t* ent_row,
                              float* ent_val, void* stream);

/* transform-sized K1 (many rows, one launch): persistent CTAs stage the K most frequent rows of W (hot_cols[K], 2 kB each at
 * H = 500) in shared memory with 1-D bulk-TMA copies and serve the entries of those columns from there; the cold tail gathers from
 * L2 as in dae_encode_csr_fwd.  hot_slot[F] = index of the column in hot_cols, or -1.  K * H * 4 <= 200 KB, H % 4 == 0.
 * groups = 4 or 8 row groups of 128 threads per CTA; the number of CTAs per SM follows from the size of the staged set.
 * Same results as dae_encode_csr_fwd up to fp32 summation order inside a row (identical: entries are added in CSR order). */
int dae_encode_csr_fwd_hot(const int64_t* indptr, const int32_t* indices, const float* values, int32_t n_rows,
                           int32_t F, int32_t H, float in_scale, const float* W, const float* bh, int32_t enc_act,
                           float* E, int64_t ldE, const int32_t* hot_cols, const int32_t* hot_slot, int32_t K,
                           int32_t groups, void* stream);

/* ---- fp32 reference GEMM (CUDA cores) ----------------------------------------------------------
 * C[m,n] = alpha * sum_k A[m*sam + k*sak] * B[n*sbn + k*sbk] + beta * C[m,n]; generic strides.
 * The v1 / validation path for the dense contractions (autoencoder.py:411 and its autodiff,
 * triplet_loss_utils.py:93,219); the production path is dae_gemm_bf16x3 / dae_decode_fused_bf16x3.
 */
int dae_sgemm(int32_t M, int32_t N, int32_t K, float alpha, const float* A, int64_t sam, int64_t sak,
              const float* B, int64_t sbn, int64_t sbk, float beta, float* C, int64_t ldc, void* stream);

/* ---- tcgen05 path for the dense contractions (K2/K3, Gram matrix) -----------------------------------
 * fp32-accurate GEMM on the 5th-gen tensor cores: each fp32 operand is carried as bf16 hi + bf16 lo
 * (dae_split_bf16) and D = A_hi.B_hi + A_lo.B_hi + A_hi.B_lo is accumulated in fp32 in TMEM.
 * TMA-fed (128B swizzle), persistent, warp specialised; operands are read K-major or MN-major straight
 * from row-major arrays (majorness flags), so dZ^T / E^T / W^T are never materialised.
 *
 * dae_split_bf16: hi/lo [rows x ld_dst] <- src [rows x cols] * scale, zero padded; column `ones_col`
 *   (if >= 0) is set to 1.0 -- the [E | 1] trick that makes dW = dZ^T.[E | 1] also deliver dbv.
 * dae_sym_split_bf16: hi/lo <- alpha * (G + G^T)   (dE_tri = alpha (G + G^T) E).
 * dae_gemm_bf16x3: C[m,n] (+)= alpha * sum_k A(m,k) B(n,k).
 *   a_mn_major = 0: A stored [M x lda] (K contiguous); 1: A stored [K x lda] (M contiguous).  Same for B/N.
 *   columns n < n_store go to C; column special_col (if special_out != NULL) goes to special_out[m].
 *   k_splits > 1 (uniform split-K), k_splits < 0 (stream-K: the tile x k-block units are shared evenly by the SMs, chosen
 *   automatically when the 128x256 tiling does not fill whole waves) or accumulate != 0: fp32 atomics into C (C is zeroed
 *   first unless accumulate).
 * dae_decode_fused_bf16x3: Z = E.W^T with the decode-loss epilogue fused (D = g(Z+bv), CE/MSE row loss
 *   against the clean CSR rows, dZ written directly as bf16 hi/lo [B x ld_dz]); row_loss_part is the
 *   [B] row-loss vector (zeroed here, accumulated with fp32 atomics, one add per half tile).  Replaces autoencoder.py:411 +
 *   triplet_loss_utils.py:262-275 and their autodiff without materialising Z, D or dense X.
  *   tile_ptr is int32 scratch [Brows x (2*ceil(F/256) + 1)] (per-row CSR offsets of every 128-column part, filled here).
 */
int dae_split_bf16(const float* src, int32_t rows, int32_t cols, int64_t ld_src, void* hi, void* lo,
                   int64_t ld_dst, int32_t ones_col, float scale, void* stream);
int dae_sym_split_bf16(const float* G, int32_t B, int64_t ldg, float alpha, void* hi, void* lo, int64_t ld,
                       void* stream);
int dae_gemm_bf16x3(int32_t M, int32_t N, int32_t K, float alpha, const void* a_hi, const void* a_lo,
                    int64_t lda, int32_t a_mn_major, const void* b_hi, const void* b_lo, int64_t ldb,
                    int32_t b_mn_major, float* C, int64_t ldc, int32_t n_store, int32_t special_col,
                    float* special_out, int32_t k_splits, int32_t accumulate, void* stream);
/* C[m,n] (+)= alpha * sum_k (G[m,k] + G[k,m]) * B[k,n] for a square G [M x M] (bf16 hi/lo, row-major) and B stored [M x ldb] row-major:
 * dE2 = alpha (G + G^T) E of the triplet backward in ONE launch -- the k loop runs over G's columns and then over G's rows (the same
 * array through an M-contiguous tensor map), so G + G^T is never formed.  Stream-K with fp32 atomics; C is zeroed first unless accumulate. */
int dae_gemm_sym_bf16x3(int32_t M, int32_t N, float alpha, const void* g_hi, const void* g_lo, int64_t ldg,
                        const void* b_hi, const void* b_lo, int64_t ldb, float* C, int64_t ldc, int32_t accumulate,
                        void* stream);
/* Tile engine selection of the two entry points above (test hook).  pair_mode: -1 = CTA pairs (cta_group::2: two SMs share one
 * 256-row UMMA tile, each staging half of the B operand) for the large store GEMMs only -- the default --, 0 = never, 1 = whenever
 * possible (fused decode included).  lean: 0 (default) = deepest operand rings; 1 = 2-stage rings (128 KB per CTA), which leave
 * ~70 KB of every SM's shared memory to kernels running concurrently with the GEMM.
 * tile_ptr of dae_decode_prepare is laid out for the configuration current at the time of the call. */
int dae_gemm_config(int32_t pair_mode, int32_t lean);
/* The part of the fused decode that needs only the batch's row ids: zero row_loss_part and fill tile_ptr.  dae_decode_fused_bf16x3
 * runs it in line unless it is called with prepared != 0 (a graph-replayed step issues it on a parallel branch, next to K1). */
int dae_decode_prepare(int32_t Brows, int32_t F, const int64_t* indptr, const int32_t* indices, const int32_t* rows,
                       float* row_loss_part, int32_t* tile_ptr, void* stream);
int dae_decode_fused_bf16x3(int32_t Brows, int32_t F, int32_t K, const void* e_hi, const void* e_lo,
                            int64_t lde, const void* w_hi, const void* w_lo, int64_t ldw,
                            const int64_t* indptr, const int32_t* indices, const float* values,
                            const int32_t* rows, const float* bv, int32_t dec_act, int32_t loss_func,
                            const float* weight, const double* stats, void* dz_hi, void* dz_lo,
                            int64_t ld_dz, float* row_loss_part, int32_t* tile_ptr, int32_t prepared, void* stream);
/* out[i] = sum_p parts[p * n + i] (deterministic reduction of the per-tile row-loss partials) */
int dae_reduce_parts(const float* parts, int32_t n_parts, int32_t n, float* out, void* stream);

/* ---- decode loss + dZ (elementwise part of K2) ------------------------------------------------------
 * In place on Z (B x F, leading dim ldz), where Z = E.W^T (no bias yet):
 *   D = g(Z + bv); row loss l_i per triplet_loss_utils.py:268-273 against the CLEAN batch rows (CSR,
 *   densified on the fly); dZ = (w_i / (sum_w + 1e-16)) * dl_i/dD * g'(Z)   (autodiff of :269-275, :411)
 * Z is overwritten with dZ; row_loss[B] receives l_i.  weight == NULL means w = 1 (strategy none);
 * sum_w is read from stats[DAE_STAT_SUM_W].
 */
int dae_decode_loss_bwd(const int64_t* indptr, const int32_t* indices, const float* values,
                        const int32_t* rows, int32_t n_rows, int32_t F, const float* bv, int32_t dec_act,
                        int32_t loss_func, const float* weight, const double* stats, float* Z, int64_t ldz,
                        float* row_loss, void* stream);

/* column sums: out[f] = sum_r M[r*ld + f]  (dbv = sum_i dZ_i) */
int dae_colsum(const float* M, int32_t n_rows, int32_t n_cols, int64_t ld, float* out, void* stream);

/* ---- K4: triplet mining -------------------------------------------------------------------------
 * batch_all (triplet_loss_utils.py:79-131, pos_triplets_only=False as called at autoencoder.py:430):
 *   rows must be label sorted (dae_batch_prepare). S = E.E^T is an input (B x B, ld lds).
 *   Writes G (B x B): dL_tri/dS, accumulates loss sum / positive count into stats.
 *   pos_only != 0 (pos_triplets_only=True, :118-120; never used by the model): the loss sum covers positive triplets
 *   only and G receives raw COUNTS of positive triplets (G[i,j] = -#k, G[i,k] = +#j) from which the caller derives the weights.
 *   g_hi / g_lo (optional, bf16 [B x ld_split]): G also leaves as the hi / lo operand pair dae_gemm_sym_bf16x3 reads.
 * batch_hard (triplet_loss_utils.py:202-259): also writes the data weight (w) and sum_w.
 */
int dae_triplet_batch_all(const float* S, int64_t lds, int32_t B, const int32_t* seg_lo, const int32_t* seg_hi,
                          float* G, int64_t ldg, double* stats, int32_t pos_only, void* g_hi, void* g_lo,
                          int64_t ld_split, void* stream);
int dae_triplet_batch_hard(const float* S, int64_t lds, int32_t B, const float* labels, float* G, int64_t ldg,
                           float* weight, double* stats, void* stream);
/* explicit triplets (autoencoder_triplet.py:308-311): loss = mean softplus(e.en - e.ep); ACCUMULATES alpha * dloss
 * into dE/dEp/dEn (on top of the reconstruction gradient) and the loss sum into stats[DAE_STAT_TRIPLET_SUM]. */
int dae_triplet_explicit(const float* E, const float* Ep, const float* En, int32_t B, int32_t H, int64_t ld,
                         float alpha, float* dE, float* dEp, float* dEn, double* stats, void* stream);

/* ---- step epilogue ---------------------------------------------------------------------------------
 * Reduces row_loss -- or, if parts != NULL, the [n_parts x B] per-tile partials of dae_decode_fused_bf16x3 -- (x weight)
 * deterministically and fills COST / AE_LOSS / TRIPLET_LOSS / FRACTION / NUM
 * of `stats` (autoencoder.py:438,441; triplet_loss_utils.py:127,131,257,259,275); then copies the
 * DAE_STAT_SLOTS doubles to stats_log (one row of the per-epoch log) if non-NULL.
 */
int dae_step_finalize(const float* row_loss, const float* parts, int32_t n_parts, const float* weight, int32_t B,
                      int32_t strategy, float alpha, double* stats, double* stats_log, const int64_t* ctl, void* stream);

/* ---- K6: optimizer ------------------------------------------------------------------------------------
 * theta <- update(theta, grad * grad_scale) over the flat buffer (autoencoder.py:451-472; TF-1.12 rules:
 * SGD; Adagrad accum(0)=0.1, no eps; Momentum accum=mu*accum+g, theta-=lr*accum; Adam b1 .9 b2 .999 eps 1e-8
 * with lr_t = lr*sqrt(1-b2^t)/(1-b1^t), t = step (1-based)).
 * If w_hi/w_lo are non-NULL the updated W (first F*H entries) is also written as the bf16 hi/lo pair [F x ld_split]
 * consumed by the tensor-core GEMMs (fuses dae_split_bf16 of W into the update).
 */
int dae_optimizer_step(float* theta, const float* grad, float* slot1, float* slot2, int64_t n, int32_t opt,
                       float lr, float momentum, float grad_scale, int32_t step, const int64_t* ctl, void* w_hi,
                       void* w_lo, int32_t F, int32_t H, int64_t ld_split, void* stream);

/* ---- corruption -----------------------------------------------------------------------------------------
 * values_out[p] = keep[p] ? values[p] : 0 where keep is a host-generated byte mask (bit-parity mode with
 * np.random.rand(nnz) >= v, utils.py:111) -- or, if keep == NULL, a Philox4x32-10 draw keyed by
 * (seed, epoch, p):  u >= corr_frac  (device mode; same distribution, different stream).
 */
int dae_mask_values(const float* values, const uint8_t* keep, int64_t nnz, float corr_frac, uint64_t seed,
                    uint64_t epoch, float* values_out, void* stream);

/* ---- "next" row: pairwise similarity of embeddings + nearest-article lookup -----------------
 * Replaces helpers.pairwise_similarity (helpers.py:11-50: sklearn cosine_similarity / linear_kernel, optional normalize,
 * zeroed diagonal) and the nanargmax lookup of main_autoencoder.py:352-353.  sim = normalize(E).normalize(E)^T runs on
 * dae_gemm_bf16x3; these are the two kernels around it.
 * dae_rownorm_split_bf16: rows scaled by 1/||x||_2 (norm_kind 2; all-zero rows untouched, like sklearn), 1/||x||_1 (1),
 *   1/max|x| (3) or 1 (0), written as bf16 hi/lo [rows x ld_dst] (zero padded) and/or as fp32 x_out.
 * dae_row_argmax: per row arg-max / max of S skipping column row+diag_offset (optionally zeroing it in place).
 */
int dae_rownorm_split_bf16(const float* X, int32_t rows, int32_t cols, int64_t ld, int32_t norm_kind, void* hi, void* lo,
                           int64_t ld_dst, float* x_out, int64_t ld_out, void* stream);
int dae_row_argmax(float* S, int32_t rows, int32_t cols, int64_t ld, int64_t diag_offset, int32_t zero_diag,
                   int32_t* idx_out, float* val_out, void* stream);

/* ---- "next" row: related-vs-unrelated AUROC of a pairwise similarity matrix ---------------------
 * Replaces the numeric part of helpers.visualize_pairwise_similarity (helpers.py:88-100).
 * dae_pair_partition: for every pair i > j of the strict lower triangle with labels[i] >= 0 and labels[j] >= 0
 *   (-1 = missing, helpers.py:91) append S[i, j] to `related` if labels[i] == labels[j] (helpers.py:92-95) else to
 *   `unrelated` (helpers.py:96-97).  cursors[0..1] are device counters the caller zeroes; on return they hold the
 *   group sizes.  Order inside a group is unspecified.  Capacity: R = sum_c n_c(n_c-1)/2, U = M(M-1)/2 - R floats.
 * dae_auroc_count: *twice_u += sum_q 2*#{t < q} + #{t == q} (query_is_positive = 1: queries are the related scores,
 *   sorted_targets the ascending unrelated scores) or sum_q 2*#{t > q} + #{t == q} (0: roles swapped).  Then
 *   AUROC = twice_u / (2 R U) -- the area sklearn's roc_curve + auc (helpers.py:99-100) return, ties included.
 */
int dae_pair_partition(const float* S, int64_t lds, int32_t n, const int32_t* labels, float* related, float* unrelated,
                       uint64_t* cursors, void* stream);
int dae_auroc_count(const float* queries, int64_t n_queries, const float* sorted_targets, int64_t n_targets,
                    int32_t query_is_positive, uint64_t* twice_u, void* stream);

/* ---- data-parallel exchange step: in-switch all-reduce of the flat gradient buffer -------------------
 * The reference is single-process; row-sharded training adds ONE sum over ranks of [dW | dbh | dbv] between the
 * gradient kernels and dae_optimizer_step.  Default transport: ncclAllReduce.  dae_allreduce_multimem is the
 * in-graph alternative: `multicast_grad` is the NVSwitch multicast address bound to every rank's gradient buffer
 * (symmetric memory, identical offset on every rank, n floats, 16-byte aligned); `peer_flags` is a DEVICE array of
 * `world` pointers to each rank's flag words (2 * n_blocks * world zero-initialised uint32, peer-mapped); `epochs` is
 * this rank's own uint32[n_blocks] (zero-initialised, ordinary device memory): the per-CTA exchange counter the
 * flags carry, advanced by the kernel itself so that it stays in step with CUDA-graph replays.  Every rank calls it
 * with the same n and n_blocks on its step stream, the same number of times; on return (stream order) every rank's
 * buffer holds the sum.  Rank r reduces float4 packets [r*ceil(n4/P), ...) with multimem.ld_reduce and writes them
 * back with multimem.st; CTA-level flag barriers (one remote store per peer to arrive, local polling to wait) open and
 * close the exchange and trap after 5 s instead of hanging a replica.
 */
int dae_allreduce_multimem(float* multicast_grad, void* const* peer_flags, uint32_t* epochs, int32_t rank,
                           int32_t world, int64_t n, int32_t n_blocks, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DAE_SM100_H */
