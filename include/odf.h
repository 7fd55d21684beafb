/* libodf — C ABI of the B200-native FALKON hot path (online-detection on-line learners).
 *
 * The reference (hsp-iit/online-detection) has no FFI of its own: its seam is the Python duck
 * type between first-party code and the third-party `falkon` package
 * (github.com/FalkonML/falkon @ 0d96c685, INSTALLATION_GUIDE.md:67-77).  Every entry point below
 * names the reference interface it stands in for.  All pointers are DEVICE pointers to
 * caller-owned, row-major fp32 buffers (16-byte aligned); `stream` is a cudaStream_t passed as
 * void*; every call is asynchronous on that stream unless stated; return 0 = OK, negative =
 * error (text via odf_last_error()).  No torch types cross this boundary; no internal threads;
 * one process per GPU.  There is no CPU fallback: without a CUDA device every compute entry
 * point fails with ODF_ERR_CUDA.
 */
#ifndef ODF_H_
#define ODF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ODF_OK 0
#define ODF_ERR_ARG (-1)
#define ODF_ERR_CUDA (-2)
#define ODF_ERR_WORKSPACE (-3)
#define ODF_ERR_LINALG (-4)

/* which triangular solve odf_precond_solve applies (FalkonPreconditioner.invT / invTt / invA /
 * invAt) */
#define ODF_SOLVE_T 0   /* B <- T^-1  B */
#define ODF_SOLVE_TT 1  /* B <- T^-T  B */
#define ODF_SOLVE_A 2   /* B <- A^-1  B */
#define ODF_SOLVE_AT 3  /* B <- A^-T  B */

/* operand kinds of the fused tile: how the pre-pass stores the hi/lo split of a point set */
#define ODF_KIND_TF32 0 /* tf32 in fp32 words, kind::tf32 MMAs ("3xTF32")                         */
#define ODF_KIND_F16 1  /* fp16 after a power-of-two scaling, kind::f16 MMAs (same 2x11 bits, 2x rate) */

const char* odf_last_error(void);
int odf_version(void);
/* kind used by the convenience entry points (odf_gauss_mmv/dmmv/kmm); default ODF_KIND_F16 */
int odf_set_default_kind(int kind);
int odf_default_kind(void);

/* ---- layout helpers ------------------------------------------------------------------------ */
/* row pitch, in ELEMENTS, of a prepared operand array: features padded to the k-block width
 * (32 tf32 / 64 fp16 elements = 128 bytes) plus the seed block and the ones block             */
int64_t odf_operand_pitch(int64_t d, int kind);
int64_t odf_operand_bytes(int64_t n, int64_t d, int kind); /* bytes of ONE of the hi / lo arrays */
int64_t odf_pad_rows(int64_t n);  /* length of norm vectors / pitch of V^T: round_up(n, 128)    */
int odf_tpad(int64_t T);          /* padded RHS count: 16 if T<=16, 32 if T<=32, else -1        */
/* Number of column splits the tile launcher wants for this shape (size of the partial slab). */
int odf_tile_splits(int64_t n_rows, int64_t n_cols, int64_t d, int kind);

/* ---- operand preparation ------------------------------------------------------------------- */
/* Fused z-score + hi/lo split + squared norms (two passes over X).
 * Replaces OnlineRegionClassifier.zScores (src/modules/region-classifier/
 * OnlineRegionClassifier.py:224-227) and the norm pre-pass of falkon GaussianKernel.
 *   x' = (X[r,:] - mean) * scale                      (mean may be NULL -> 0)
 *   sqnorm[r] = |x'|^2   -> odf_pad_rows(n) floats, zero padded
 *   opscale[0] = s       -> power of two (1 for ODF_KIND_TF32); opscale is a 2-float device buffer
 *   hi = rn11(s x'), lo = rn11(s x' - hi)  -> [n x odf_operand_pitch(d, kind)] elements each
 *   hi additionally carries the seed block (-s^2 |x'|^2 / 2 as hi, lo) and the ones block.   */
int odf_prepare_points(const float* X, int64_t n, int64_t d, int64_t ldx, const float* mean,
                       float scale, int kind, void* hi, void* lo, float* sqnorm, float* opscale,
                       void* stream);
/* EXPERIMENTAL: the fused tile as a split GEMM (3-pass fp16 / tf32 products, fp32 accumulation in TMEM) for the blocked
 * preconditioner build (odf/precond_blocked.py; replaces cuBLAS sgemm inside falkon FalkonPreconditioner.init).
 * odf_prepare_points_linear = odf_prepare_points without z-score and with a ZERO accumulator seed; odf_gemm_nt_split:
 * C[m x n] = alpha A B^T + beta C for A [m x k], B [n x k] in that prepared form (k <= ~2048 per call: the tensor
 * core adds with truncation, long contractions are sliced on the host and accumulated with beta = 1).            */
int odf_prepare_points_linear(const float* X, int64_t n, int64_t d, int64_t ldx, int kind, void* hi, void* lo,
                              float* sqnorm, float* opscale, void* stream);
int odf_gemm_nt_split(int kind, const void* a_hi, const void* a_lo, const float* a_sqnorm,
                      const float* a_opscale, int64_t m, const void* b_hi, const void* b_lo,
                      const float* b_sqnorm, const float* b_opscale, int64_t n, int64_t k, float alpha,
                      float beta, float* C, int64_t ldc, void* stream);
/* In-place z-score only (same reference lines). */
int odf_zscore(float* X, int64_t n, int64_t d, int64_t ldx, const float* mean, float scale,
               void* stream);
/* V [m x T] (pitch ldv) -> V^T split: vt_hi, vt_lo [T_pad x ldvt] = split(scale * V^T), zero
 * padded (ldvt = odf_pad_rows(m)). */
int odf_split_rhs(const float* V, int64_t m, int64_t T, int64_t ldv, float scale, float* vt_hi,
                  float* vt_lo, int64_t ldvt, int T_pad, void* stream);

/* ---- fused Gaussian-kernel tile (tcgen05 / TMEM / TMA) -------------------------------------- */
/* partial[s][r][0..T_pad) = sum over the columns of split s of K(row r, col q) * V[q, :]
 * with K = exp(-|r-q|^2 / (2 sigma^2)).  Stands in for falkon GaussianKernel.mmv
 * (call sites: FALKONWrapper_with_centers_selection_incore.py:75-82, roi_box_predictors.py:158,
 * roi_mask_predictors.py:90, rpn.py:225) and for each half of GaussianKernel.dmmv.
 * n_splits must come from odf_tile_splits(n_rows, n_cols, d, kind).                                   */
int odf_gauss_mmv_prepared(int kind, const void* r_hi, const void* r_lo, const float* r_sqnorm,
                           const float* r_opscale, int64_t n_rows, const void* q_hi,
                           const void* q_lo, const float* q_sqnorm, const float* q_opscale,
                           int64_t n_cols, int64_t d, const float* vt_hi, const float* vt_lo,
                           int64_t ldvt, int T_pad, int n_splits, float sigma, float* partial,
                           void* stream);
/* Same as odf_gauss_mmv_prepared, and additionally spills the K tiles (fp32, exactly the K_hi + K_lo
 * the contraction used) into the transient row panel P [n_rows x ldp], ldp >= odf_pad_rows(n_cols),
 * for odf_panel_tmm.  The panel covers one launch (a chunk of rows), never the whole K_nm.      */
int odf_gauss_mmv_prepared_spill(int kind, const void* r_hi, const void* r_lo, const float* r_sqnorm,
                                 const float* r_opscale, int64_t n_rows, const void* q_hi,
                                 const void* q_lo, const float* q_sqnorm, const float* q_opscale,
                                 int64_t n_cols, int64_t d, const float* vt_hi, const float* vt_lo,
                                 int64_t ldvt, int T_pad, int n_splits, float sigma, float* partial,
                                 float* panel, int64_t ldp, void* stream);
/* out_partial[s][c][0..T_pad) = sum over the rows of split s of P[r][c] * W[r][0..T_pad):
 * the K_blk^T w half of GaussianKernel.dmmv on a spilled panel (fp32 FMA, HBM-streaming).
 * W is [n_rows x T_pad] (zero padded columns); n_splits = odf_panel_splits(n_rows, M).           */
int odf_panel_splits(int64_t n_rows, int64_t M);
int odf_panel_tmm(const float* P, int64_t ldp, const float* W, int64_t n_rows, int64_t M, int T_pad,
                  int n_splits, float* out_partial, void* stream);
/* Tensor-core variant of the spill + panel pair (the default sweep): the tile writes its K tiles as two planes,
 * hi = rn16(K) (fp16) and lo = rni((K - hi) 2^19) + 128 (ONE BYTE: the residual in fixed point, 2^-20 absolute on
 * K <= 1; 3 B per kernel value), both tile-blocked
 * [plane][column tile][row block][16 groups of 8 centres][128 rows][8] (odf_panel16_bytes(n_rows, n_cols) bytes,
 * 128-byte aligned; the panel kernels widen the lo bytes to fp16 in shared memory); odf_finish_w16 reduces the partial slabs of the first contraction into W = K v (+ addend)
 * [n_rows x T_pad] and splits it into W16 [round_up(n_rows,128) x 64] fp16 (hi | lo, per-column power-of-two
 * scales derived from max|W[:, t]|, left in absmax[32]); odf_panel16_tmm streams the planes once from HBM and contracts
 * out_partial[s][c][0..T_pad) = sum_r K[r][c] W[r][.] with tcgen05 kind::f16 MMAs on MN-major operands
 * (three split products, fp32 accumulation chains of 512 rows).  n_splits = odf_panel16_splits(n_rows, M).
 * Replaces the K_blk^T w half of falkon GaussianKernel.dmmv (...incore.py:68).                               */
size_t odf_panel16_bytes(int64_t n_rows, int64_t n_cols);
/* CTA-pair variant of the fused tile (cta_group::2, M = 256 per MMA: each CTA of a 2-CTA cluster loads half of every
 * column tile): same operator and outputs as odf_gauss_mmv_prepared / _spill16 (panel16 may be NULL), for launches
 * with odf_tile_pair_eligible(n_rows) != 0.  Its K.V contraction runs on fp16 hi/lo pairs, so the right-hand sides
 * come from odf_split_rhs16: vt_hi16 / vt_lo16 [T_pad x ldvt] fp16 (ldvt = odf_pad_rows(m)), hi = rn16(s_t v),
 * lo = rn16((s_t v - hi) 2^11), s_t the per-column power-of-two scale fixed by absmax[32] (written there).      */
int odf_tile_pair_eligible(int64_t n_rows);
int odf_split_rhs16(const float* V, int64_t m, int64_t T, int64_t ldv, float scale, void* absmax,
                    void* vt_hi16, void* vt_lo16, int64_t ldvt, int T_pad, void* stream);
int odf_gauss_mmv_pair(int kind, const void* r_hi, const void* r_lo, const float* r_sqnorm,
                       const float* r_opscale, int64_t n_rows, const void* q_hi, const void* q_lo,
                       const float* q_sqnorm, const float* q_opscale, int64_t n_cols, int64_t d,
                       const void* vt_hi16, const void* vt_lo16, int64_t ldvt, const void* v_absmax,
                       int T_pad, int n_splits, float sigma, float* partial, void* panel16,
                       void* stream);
int odf_gauss_mmv_prepared_spill16(int kind, const void* r_hi, const void* r_lo, const float* r_sqnorm,
                                   const float* r_opscale, int64_t n_rows, const void* q_hi,
                                   const void* q_lo, const float* q_sqnorm, const float* q_opscale,
                                   int64_t n_cols, int64_t d, const float* vt_hi, const float* vt_lo,
                                   int64_t ldvt, int T_pad, int n_splits, float sigma, float* partial,
                                   void* panel16, void* stream);
int odf_finish_w16(const float* partial, int n_splits, int64_t n_rows, int T_pad, int64_t T,
                   const float* addend, int64_t ld_add, float* w_f32, void* absmax, void* w16,
                   void* stream);
int odf_panel16_splits(int64_t n_rows, int64_t M);
int odf_panel16_tmm(const void* panel16, int64_t n_rows, int64_t M, const void* w16, const void* absmax,
                    int T_pad, int n_splits, float* out_partial, void* stream);
/* The other contraction on the SAME panel, for sweeps over panels that stay resident in HBM (sweep mode
 * "resident"): out_partial[s][r][0..T_pad) = sum_{c in column range s} K[r][c] V[c][.], i.e. the K_blk v half of
 * falkon GaussianKernel.dmmv / mmv without evaluating a kernel value.  v16 [round_up(M,128) x 64] fp16 is the
 * odf_finish_w16 split of V (hi | lo, per-column scales in absmax[32]).  A [8 rows x 8 centres] block of the panel
 * is a K-major core matrix for this product (rows = the MMA's M dimension), so the planes stream through plain
 * 16 KB bulk copies.  n_splits = odf_panel16_mmv_splits(n_rows, M).                                          */
int odf_panel16_mmv_splits(int64_t n_rows, int64_t M);
int odf_panel16_mmv(const void* panel16, int64_t n_rows, int64_t M, const void* v16, const void* absmax,
                    int T_pad, int n_splits, float* out_partial, void* stream);
/* EXPERIMENTAL precision tier of the two panel contractions: only the hi plane is streamed (2 B per kernel value,
 * K to 11 bits; W / V keep their hi | lo split).  Same arguments and outputs as odf_panel16_tmm / odf_panel16_mmv.
 * A CPU emulation of the whole fit (tools/precision_study.py, profiles/r1_precision_study_cpu.log) puts the effect
 * on the decision scores at 3e-7 .. 3e-5 relative on learnable data and at 5e-4 on an over-fitted random-label
 * problem, against the 1e-3 parity bar: an opt-in tier, never the default; not yet validated on the GPU.        */
int odf_panel16_tmm_hi(const void* panel16, int64_t n_rows, int64_t M, const void* w16, const void* absmax,
                       int T_pad, int n_splits, float* out_partial, void* stream);
int odf_panel16_mmv_hi(const void* panel16, int64_t n_rows, int64_t M, const void* v16, const void* absmax,
                       int T_pad, int n_splits, float* out_partial, void* stream);
/* out[r, t] = scale * sum_s partial[s][r][t] + addend[r, t]   (addend may be NULL) */
int odf_finish_rows(const float* partial, int n_splits, int64_t n_rows, int T_pad, int64_t T,
                    float scale, const float* addend, int64_t ld_add, float* out, int64_t ldo,
                    void* stream);
/* same reduction, written as the split transposed right-hand side of a following pass */
int odf_finish_split(const float* partial, int n_splits, int64_t n_rows, int T_pad, int64_t T,
                     float scale, const float* addend, int64_t ld_add, float* wt_hi, float* wt_lo,
                     int64_t ldwt, void* stream);
/* K[i, j] = exp(-|c_i - c_j|^2 / (2 sigma^2)), M x M, pitch ldk.  falkon Kernel.__call__ (K_MM
 * for FalkonPreconditioner.init, reached through InCoreFalkon.fit, ...incore.py:68).            */
int odf_gauss_kmm_prepared(int kind, const void* c_hi, const void* c_lo, const float* c_sqnorm,
                           const float* c_opscale, int64_t M, int64_t d, float sigma, float* K,
                           int64_t ldk, void* stream);

/* ---- convenience entry points on plain fp32 operands (prepare internally in `ws`) ----------- */
#define ODF_OP_MMV 0
#define ODF_OP_DMMV 1
#define ODF_OP_KMM 2
#define ODF_OP_PRECOND 3
size_t odf_workspace_bytes(int op, int64_t n, int64_t M, int64_t d, int64_t T);
/* potrf scratch of odf_precond_init / odf_potrf_upper for an M x M matrix as cuSOLVER reports it on the current
 * device (cusolverDnSpotrf_bufferSize; 0 if no device can be queried).  odf_workspace_bytes(ODF_OP_PRECOND, ..)
 * returns max(this, a closed-form lower bound). */
size_t odf_precond_workspace_bytes(int64_t M);
/* out[n x T] = K(X, C) V          — kernel.mmv(X1, X2, v, out) */
int odf_gauss_mmv(const float* X, int64_t n, int64_t ldx, const float* C, int64_t M, int64_t ldc,
                  int64_t d, const float* V, int64_t T, int64_t ldv, float sigma, float* out,
                  int64_t ldo, void* ws, size_t ws_bytes, void* stream);
/* out[M x T] = K(X,C)^T (K(X,C) V + W); V or W may be NULL — kernel.dmmv(X1, X2, v, w) */
int odf_gauss_dmmv(const float* X, int64_t n, int64_t ldx, const float* C, int64_t M, int64_t ldc,
                   int64_t d, const float* V, int64_t ldv, const float* W, int64_t ldw, int64_t T,
                   float sigma, float* out, int64_t ldo, void* ws, size_t ws_bytes, void* stream);
int odf_gauss_kmm(const float* C, int64_t M, int64_t ldc, int64_t d, float sigma, float* K,
                  int64_t ldk, void* ws, size_t ws_bytes, void* stream);

/* ---- preconditioner: FalkonPreconditioner.init / invT / invTt / invA / invAt ---------------- */
/* In: Tm = K_MM (M x M, pitch M).  Out: Tm = T (upper, K_MM + eps*M*I = T^T T, strict lower
 * zeroed), Am = A (upper, T T^T / M + lam*I = A^T A).  cuSOLVER potrf + cuBLAS syrk.
 * SYNCHRONOUS (reads the factorisation status); ODF_ERR_LINALG if a pivot fails.               */
int odf_precond_init(float* Tm, float* Am, int64_t M, float lam, float eps, void* ws,
                     size_t ws_bytes, void* stream);
/* The same factors (and, when Tinv / Ainv are not NULL, their explicit inverses: all four upper triangular with a zero
 * strict lower triangle, pitch M) with every O(M^3) flop on the tensor cores: blocked Cholesky on row-major lower
 * factors -- cuSOLVER potrf only on the diagonal blocks (2048 wide, ODF_PRECOND_NB), cuBLAS trsm for the row panels; the
 * trailing updates, T T^T and the divide-and-conquer triangular inverses are 3-pass split-fp16 tcgen05 GEMMs
 * (odf_gemm_nt_split's tile) with contraction chains of at most 1024; T^-1 is computed on an internal stream beside
 * T T^T and the second factorisation.  In: K = K_MM (DESTROYED: it ends up holding the lower factor of T).
 * SYNCHRONOUS (reads the pivots' status); ODF_ERR_LINALG if a diagonal block is not positive definite.
 * ws >= odf_precond_build_workspace_bytes(M).                                                                      */
size_t odf_precond_build_workspace_bytes(int64_t M);
int odf_precond_build(float* K, float* Tm, float* Am, float* Tinv, float* Ainv, int64_t M, float lam, float eps,
                      void* ws, size_t ws_bytes, void* stream);
/* B [M x T] (pitch ldb) <- op(Tri)^-1 B with Tri upper triangular (pitch M), op by `which`.   */
int odf_precond_solve(const float* Tri, int64_t M, float* B, int64_t T, int64_t ldb, int which,
                      void* stream);

/* Explicit inverse of an upper-triangular factor (Inv = Tri^-1, upper, pitch M; done once per fit) and its application
 *   Bout = Inv B (transposed = 0)   or   Bout = Inv^T B (transposed = 1),   B, Bout: M x T, pitch ldb,
 * by an own kernel that reads only the triangle of Inv and accumulates in fp64 (csrc/odf_tri.cu; ODF_PRECOND_APPLY=cublas
 * selects the full-square cuBLAS sgemm of round 1).  This replaces the 4 latency-bound triangular solves per CG iteration;
 * it is algebraically the same preconditioner (any invertible T~, A~ used consistently leaves the fixed point of the
 * preconditioned system unchanged up to the regulariser lam T~^T T~).                                             */
int odf_precond_invert(const float* Tri, float* Inv, int64_t M, void* stream);
int odf_precond_apply(const float* Inv, int64_t M, const float* Bin, float* Bout, int64_t T,
                      int64_t ldb, int transposed, void* stream);
/* Row-major fp32 GEMM on sub-blocks, C (m x n) = alpha op(A) op(B) + beta C (cuBLAS sgemm, true fp32): used by the
 * row-sharded fit for its column block of T T^T (only the non-zero ranges of the triangular factor are multiplied). */
int odf_gemm(int trans_a, int trans_b, int64_t m, int64_t n, int64_t k, float alpha, const float* A,
             int64_t lda, const float* B, int64_t ldb, float beta, float* C, int64_t ldc, void* stream);
/* Rows [r0, r1) of the same product for an UPPER-triangular Inv (reads only its non-zero part): Bout_rows is
 * (r1 - r0) x T with pitch ldo.  Used by the row-sharded fit: one row block per rank, then an all-gather.      */
int odf_precond_apply_rows(const float* Inv, int64_t M, int64_t r0, int64_t r1, const float* Bin,
                           float* Bout_rows, int64_t T, int64_t ldb, int64_t ldo, int transposed,
                           void* stream);

/* ---- conjugate-gradient vector kernels on M x T blocks (per-column scalars) ------------------ */
/* FalkonConjugateGradient / ConjugateGradient.solve.  `state` is a device
 * array of 4*T+4 floats owned by the caller: rs_old[T], a[T], b[T], rs_new[T], then flags:
 * state[4T] = 1.0 once converged (every later update becomes a no-op, which freezes the
 * iterate exactly where the reference's `break` leaves it).                                    */
int odf_cg_init(const float* R, int64_t M, int64_t T, int64_t ld, float* state, void* ws,
                size_t ws_bytes, void* stream);           /* rs_old = colsum(R^2), flag = 0      */
int odf_cg_alpha(const float* P, const float* AP, int64_t M, int64_t T, int64_t ld, float eps,
                 float* state, void* ws, size_t ws_bytes, void* stream); /* a = rs_old/(P.AP+eps)*/
/* Y[:, t] += sign * a[t] * X[:, t]  with a = state[T..2T) */
int odf_cg_axpy_a(float* Y, const float* X, int64_t M, int64_t T, int64_t ld, float sign,
                  const float* state, void* stream);
/* R = Bm - H   (full-gradient restart), skipped when converged */
int odf_cg_residual(float* R, const float* Bm, const float* H, int64_t M, int64_t T, int64_t ld,
                    const float* state, void* stream);
/* rs_new = colsum(R^2); converged |= sqrt(max rs_new) < tol; b = rs_new/(rs_old+eps);
 * rs_old = rs_new */
int odf_cg_beta(const float* R, int64_t M, int64_t T, int64_t ld, float eps, float tol,
                float* state, void* ws, size_t ws_bytes, void* stream);
/* P[:, t] = R[:, t] + b[t] * P[:, t] */
int odf_cg_xpby_b(float* P, const float* R, int64_t M, int64_t T, int64_t ld, const float* state,
                  void* stream);
/* out = alpha * A + beta * Bm (elementwise on M x T; A/Bm may alias out; Bm may be NULL) */
int odf_axpby(float* out, float alpha, const float* A, float beta, const float* Bm, int64_t M,
              int64_t T, int64_t ld, void* stream);
size_t odf_cg_workspace_bytes(int64_t M, int64_t T);

/* Pieces of odf_precond_init for the row-sharded multi-GPU fit, where the O(M^3) steps are split over the
 * ranks (column blocks of T T^T and of the two explicit inverses, all-gathered by the host side):
 * in-place Cholesky A = U^T U of a row-major symmetric matrix (upper factor left in A, strict lower triangle
 * zeroed; synchronises the stream to read cuSOLVER's info), diagonal shift, triangle clear.
 * ws >= odf_workspace_bytes(ODF_OP_PRECOND, 0, M, 0, 1).                                              */
int odf_potrf_upper(float* A, int64_t M, void* ws, size_t ws_bytes, void* stream);
int odf_add_diag(float* A, int64_t M, float value, void* stream);
int odf_zero_strict_lower(float* A, int64_t M, void* stream);

/* ---- RLS box-refinement regressors, all classes / anchors in one call ------------------------------------------- */
/* RegionRefinerTrainer.train / solve (src/modules/region-refiner/region_refiner_trainer/train_region_refiner.py:25-119), fp64 as
 * the reference: for every class c over ITS rows, Xi = [X 1],  w_k = (Xi^T Xi + lam I)^-1 Xi^T y'_k (k < 4),
 * losses = (Xi w_k - y'_k)^2 / 2.  X [n x d] fp32 (pitch ldx); Yw [n x 4] fp64 = the centred, whitened targets (per ORIGINAL
 * row); perm [n] = row indices sorted by class; seg_host [n_classes + 1] (HOST) = class boundaries in perm; row_class [n] =
 * class of perm[r].  Out: W [n_classes][4][d + 1] fp32 (bias last), losses [n][4] fp32 in perm order.  The normal matrices
 * of all classes come from ONE launch on the fp64 tensor cores (upper 64 x 64 tiles of [X 1 y']^T [X 1 y']), the
 * factorisations from cuSOLVER Dpotrf / Dpotrs on internal side streams.  SYNCHRONOUS; ODF_ERR_LINALG on a failed pivot;
 * classes without rows are skipped.  ws >= odf_rls_workspace_bytes(n, d, n_classes).                               */
size_t odf_rls_workspace_bytes(int64_t n, int64_t d, int64_t n_classes);
int odf_rls_train(const float* X, int64_t n, int64_t d, int64_t ldx, const double* Yw, const int64_t* perm,
                  const int64_t* seg_host, const int* row_class, int64_t n_classes, double lam, float* W, float* losses,
                  void* ws, size_t ws_bytes, void* stream);
/* RegionPredictor.predict (src/modules/region-refiner/region_predictor/predict_regions.py:16-80; the *_parallel heads
 * roi_box_predictors.py:97-124, rpn.py:158-187) for one image, fused: y = feat Wp + bias (Wp [d x 4 C], class-major columns),
 * per class y T_inv + mu, box decode with the predictor's eps-width convention, clamp to the image; out [n][C + 1][4] with
 * the un-refined box in slot 0.  mean != NULL fuses zScores ((feat - mean) * zscale) into the feature load.        */
int odf_rls_apply(const float* feat, int64_t n, int64_t d, int64_t ldf, const float* Wp, const float* bias,
                  const float* Tinv, const float* mu, const float* ex_boxes, int64_t C, float img_w, float img_h, float eps,
                  const float* mean, float zscale, float* out, void* stream);

/* ---- index / integer side: minibootstrap selection, box decode, detection post-processing ---- */
/* Stable stream compaction: idx_out[0..*count_out) = ascending indices i in [0, n) with
 * scores[i * stride] > thresh (strict != 0) or >= thresh (strict == 0) -- bit-identical to
 * `torch.where(scores > thr)[0]`, the hard- / easy-negative selection of the minibootstrap
 * (src/modules/region-classifier/OnlineRegionClassifier_incore.py:117-118, 133-134).  idx_out holds n
 * entries; count_out is a DEVICE int; ws >= odf_select_workspace_bytes(n).                        */
size_t odf_select_workspace_bytes(int64_t n);
int odf_select_indices(const float* scores, int64_t n, int64_t stride, float thresh, int strict, int64_t* idx_out,
                       int* count_out, void* ws, size_t ws_bytes, void* stream);
/* dst[k, 0:d] = src[idx[k], 0:d] for k < *count (device count, at most max_rows): `X[idx]` appended to a
 * pre-allocated cache instead of torch.cat (same file, :119, :135).                                 */
int odf_gather_rows(const float* src, int64_t ld_src, const int64_t* idx, const int* count, int64_t max_rows, int64_t d,
                    float* dst, int64_t ld_dst, void* stream);
/* py_od_utils.decode_boxes_detector (src/py_od_utils.py:247-274): ex_boxes [R x 4], deltas / out [R x 4*Tc],
 * legacy +1 widths, x2 = ctr + w/2 - 1, clamped to [0, img-1].                                      */
int odf_decode_boxes(const float* ex_boxes, const float* deltas, int64_t R, int64_t Tc, float img_w, float img_h, float* out,
                     void* stream);
/* OnlineDetectionPostProcessor.filter_results (src/modules/accuracy-evaluator/OnlineDetectionPostProcessor.py:35-79):
 * boxes [R x 4*Tc], scores [R x Tc] (column 0 = background); per class 1..Tc-1: score > score_thresh,
 * greedy NMS in descending score order, suppress when IoU(+1 convention) > nms_thresh (maskrcnn-benchmark
 * boxlist_nms / _C.nms semantics, kept boxes in ascending RoI order); classes concatenated; if more than
 * dets_per_img survive keep those with score >= the dets_per_img-th largest (kthvalue rule, ties kept).
 * Outputs hold R*(Tc-1) entries; out_rois = source RoI index; out_count is a DEVICE int.           */
size_t odf_postprocess_workspace_bytes(int64_t R, int64_t Tc);
int odf_detect_postprocess(const float* boxes, const float* scores, int64_t R, int64_t Tc, float score_thresh,
                           float nms_thresh, int dets_per_img, float* out_boxes, float* out_scores, int64_t* out_labels,
                           int64_t* out_rois, int* out_count, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* ODF_H_ */
