/* libsdpcutsel -- C ABI of the B200 (sm_100a) cut-selection hot path.
 *
 * Drop-in boundary for the per-round selection path of rb2309/SDPCutSel-via-NN.  The reference has one
 * FFI crossing on this path -- ctypes -> neural_nets/NNs.so: `double neural_net_{2..5}D(const double*)`
 * (cut_select_qp.py:284-303, call sites 579-582) -- and otherwise a Python method surface
 * (CutSolver._get_sdp_vertex_cover :377, ._sel_eigcut_by_ordering_on_measure :543,
 * ._gen_eigcuts_selected :705, ._get_eigendecomp :788, .__preprocess_triangle_ineq :799,
 * .__separate_and_add_triangle :823).  Each entry point below names the reference lines it replaces.
 *
 * Conventions: plain pointers and sizes only; all pointers are HOST pointers owned by the caller unless the
 * name says `_dev`; arrays are C-contiguous; every function returns 0 (SDPCS_OK) or a negative error code
 * and never throws; sdpcs_last_error() gives the message.  One context drives one GPU; calls on a context
 * are serialised by the caller.  There is no CPU fallback: without a CUDA device sdpcs_create fails.
 */
#ifndef SDPCUTSEL_H
#define SDPCUTSEL_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SDPCS_OK 0
#define SDPCS_ERR_INVALID (-1) /* bad argument */
#define SDPCS_ERR_CUDA (-2)    /* CUDA runtime error (message in sdpcs_last_error) */
#define SDPCS_ERR_STATE (-3)   /* weights / instance / cover / scores not set */
#define SDPCS_ERR_NOMEM (-4)   /* shard does not fit in device memory */

#define SDPCS_MAX_N 250 /* north_star: n <= 250 */
#define SDPCS_MAX_RHO 5

/* selection strategies, numbering of cut_select_qp.py:80-81 */
#define SDPCS_STRAT_FEAS 1
#define SDPCS_STRAT_OPT 2
#define SDPCS_STRAT_EXACT 3 /* optimality by the exact rho-dimensional SDP (Mosek in the reference) */
#define SDPCS_STRAT_COMB 4

typedef struct sdpcs_ctx sdpcs_ctx;

/* Algorithmic constants, mirrors the class attributes at cut_select_qp.py:22-41. */
typedef struct sdpcs_params {
    double thres_min_opt;    /* _THRES_MIN_OPT    = 0      */
    double thres_neg_eigval; /* _THRES_NEG_EIGVAL = -1e-15 */
    double big_m;            /* _BIG_M            = 1000   */
    double thres_tri_viol;   /* _THRES_TRI_VIOL   = 1e-7   */
    int32_t thres_tri_dense; /* _THRES_TRI_DENSE  = 2      */
    int32_t jacobi_sweeps;   /* scoring: 0 = Householder tridiagonalisation + Laguerre (default), > 0 = cyclic
                              * Jacobi with that many sweeps; cut generation always uses Jacobi (needs vectors) */
    int32_t nn_engine;       /* NN_rhoD evaluation: SDPCS_NN_TCGEN05 (default) or SDPCS_NN_DMMA */
    int32_t nn_fused_prep;   /* tcgen05 engine: 0 = layer-0 digit images staged through HBM by a separate kernel (default), which
                              * also computes lam_min when a call asks for both scores (one unranking + gather per candidate);
                              * 1 = images built in shared memory by producer warps of the MLP kernel (no image in HBM);
                              * 2 = like 0 but lam_min always by its own kernel (the arrangement 0 replaced, kept for A/B) */
    /* Near-tie guard (SURVEY 7 hard part 1).  The reference's order among candidates whose scores agree to rounding
     * noise is decided by LAPACK dsyevd / libm exp round-off (cut_select_qp.py:647-654, 599-601); device scores differ
     * from those by <= ~1e-14 (lam) / ~1e-10 (obj).  With a guard > 0 a selection also returns every candidate ranked
     * after the k-th whose score is within the guard of the k-th score (sdpcs_last_band), classifies with thresholds
     * relaxed by the guard in modes 1 and 3 (no candidate the reference could count as violated / positive is dropped on
     * the device), and counts the candidates whose classification lies inside the guard.  The caller re-scores near-tie
     * runs with the reference's own arithmetic (the Python mirror does: numpy eigvalsh / exact NN) and re-sorts them.
     * 0 disables the relaxation; the band then holds the exact ties of the k-th score only. */
    double guard_lam;        /* absolute guard on lam_min scores, default 1e-12 */
    double guard_obj;        /* absolute guard on optimality measures, default 1e-9 (mirror: 4e-12 * rho * max|Q_arr|) */
    int64_t band_cap;        /* room for near ties after the k winners, default 65536 entries */
    double sdp_mu_final;     /* exact SDP measure (strat 3): final barrier parameter, value within d * mu of the optimum; 1e-12 */
} sdpcs_params;

/* NN engines.  Both evaluate neural_net_{2..5}D (cut_select_qp.py:579-582) to FP64 accuracy:
 *   TCGEN05: FP64-accurate int8-sliced contraction on the 5th-generation tensor cores (tcgen05.mma kind::i8,
 *            accumulators in TMEM), FP64 only for bias / tansig; NN inputs must lie in (-2, 2) after mapminmax
 *            (always true for LP points in the McCormick box) -- otherwise the call transparently re-scores
 *            with the DMMA engine and counts it in sdpcs_timings.nn_fallbacks; sdpcs_set_weights bounds the
 *            pre-activations of every layer from the weights and serves a net whose scaled pre-activations could
 *            reach 960 (tansig has no clamp) with the DMMA engine altogether;
 *   DMMA:    FP64 tensor-core contraction (mma.sync.m8n8k4.f64). */
#define SDPCS_NN_TCGEN05 0
#define SDPCS_NN_DMMA 1
/*   SCREEN:  the same tcgen05 pipeline with 4-digit (32-bit fixed point) operands: 10 digit pairs instead of 28, two
 *            TMEM accumulator stages; NN outputs accurate to ~1e-7 (north_star asks 1e-5).  Meant as the first tier of a
 *            screen-and-refine selection: every candidate is scored by it, the contenders for the k places (everything
 *            within the guard of the k-th score, sdpcs_last_band) are re-evaluated by the FP64-accurate engine, so the
 *            selection stays the reference's (distributed.ShardedSelector(screen=True), bench.py `screened`). */
#define SDPCS_NN_SCREEN 2

/* Device timings (ms, CUDA events on the context's stream) of the last sdpcs_score / sdpcs_topk calls. */
typedef struct sdpcs_timings {
    double score_ms;      /* fused unrank+gather+eig(+NN) kernels                      */
    double select_ms;     /* key build + radix select + collect + sort                 */
    double h2d_ms;        /* upload of vars_values                                      */
    int64_t score_launches;
    int64_t select_launches;
    int64_t nn_fallbacks; /* sdpcs_score calls that re-scored with the DMMA engine (input range), cumulative */
    double nn_ms;         /* the NN part of score_ms (K1+K2+K4 launches), 0 if the last score had no NN part */
} sdpcs_timings;

int sdpcs_default_params(sdpcs_params *p);
int sdpcs_create(sdpcs_ctx **out, int device);
int sdpcs_destroy(sdpcs_ctx *ctx);
const char *sdpcs_last_error(const sdpcs_ctx *ctx); /* ctx may be NULL: last creation error */
int sdpcs_set_stream(sdpcs_ctx *ctx, void *cuda_stream); /* cudaStream_t; NULL = context-owned stream */
int sdpcs_set_params(sdpcs_ctx *ctx, const sdpcs_params *p);
int sdpcs_get_timings(const sdpcs_ctx *ctx, sdpcs_timings *t);

/* NN_rhoD weights, rho in 2..5; blob layout in sdpcutsel-via-nn_b200/nn_weights.py.
 * Replaces ctypes.cdll.LoadLibrary('neural_nets/NNs.so') + getattr(lib, 'neural_net_%dD')
 * (cut_select_qp.py:297-303). */
int sdpcs_set_weights(sdpcs_ctx *ctx, int rho, const double *blob, int64_t len);

/* Instance arrays: Q_arr = upper triangle of Q, row-major, n(n+1)/2 doubles (cut_select_qp.py:318-321,
 * cut_select_qcqp.py:259).  Invalidates cover and scores. */
int sdpcs_set_instance(sdpcs_ctx *ctx, int n, const double *Q_arr);

/* Candidate set = all rho-subsets with lexicographic rank in [rank_begin, rank_end) (rank_end < 0: C(n,rho)).
 * Candidate index (agg_idx) == lex rank.  Nothing is stored: subsets are unranked on the fly.
 * Replaces the nested loops at cut_select_qp.py:451-455 and the aggregation at 524-540. */
int sdpcs_set_cover_all(sdpcs_ctx *ctx, int rho, int64_t rank_begin, int64_t rank_end);

/* Candidate set = explicit list (pattern-E vertex cover, cut_select_qp.py:401-522): idx is N x rho int16,
 * rows padded with -1 for cliques smaller than rho (sizes 2..rho may be mixed); candidate i has
 * agg_idx = agg_offset + i. */
int sdpcs_set_cover_list(sdpcs_ctx *ctx, int rho, const int16_t *idx, int64_t N, int64_t agg_offset);

/* Candidate set = pattern-E vertex cover P^E_rho built ON THE DEVICE from the sparsity pattern: every rho-clique of
 * the off-diagonal graph of adj (n x n, non-zero = edge, either triangle) plus every smaller clique (size >= 2)
 * contained in no clique one size larger, in the order of the reference's nested loops = lexicographic tuple order.
 * Replaces cut_select_qp.py:401-449 (rho = 3), 457-483 (rho = 4), 485-522 (rho = 5).  Candidate i has
 * agg_idx = agg_offset + i; *out_N (may be NULL) receives the number of candidates. */
int sdpcs_set_cover_pattern(sdpcs_ctx *ctx, int rho, const uint8_t *adj, int64_t agg_offset, int64_t *out_N);

/* The current list cover (sdpcs_set_cover_list / sdpcs_set_cover_pattern) as N x rho int16 rows padded with -1, in
 * candidate order: the index tuples the reference keeps in agg_list[i][0] (cut_select_qp.py:525-540). */
int sdpcs_get_cover_rows(sdpcs_ctx *ctx, int16_t *out_idx, int64_t cap_rows);

/* Multi-GPU sharding of any cover: keep the candidates [begin, end) of the cover as currently set (local
 * candidate indices); their agg_idx does not change.  For the all-subsets cover this moves the rank range, for a list
 * cover (built or shipped in full on every rank) it narrows a view, nothing is copied.  A list cover can be restricted
 * once; set it again to choose another shard. */
int sdpcs_cover_restrict(sdpcs_ctx *ctx, int64_t begin, int64_t end);

/* Cover algebra of the QCQP caller on the device (CutSolverQCQP.__get_vertex_cover,
 * cut_select_qcqp.py:319-333: `[el for el in agg_list_cons if el in agg_list]` and its complement, O(N^2) list scans in
 * the reference).  Keeps the candidates of ctx's list cover that occur in other's list cover (keep_members = 1:
 * P(E_m) intersected with P(E_0), in P(E_m) order) or that do not (keep_members = 0: the difference); both contexts hold
 * covers of the same instance on the same device.  *out_N (may be NULL) = candidates left; their agg_idx is their new
 * position (+ agg_offset). */
int sdpcs_cover_filter(sdpcs_ctx *ctx, const sdpcs_ctx *other, int keep_members, int64_t *out_N);

int sdpcs_num_candidates(const sdpcs_ctx *ctx, int64_t *N);

/* Score every candidate of the cover at the LP point vars_values = [X upper-tri row-major | x]
 * (cut_select_qp.py:547).  want bit 0: lam_min of [1 x^T; x X]_rho (cut_select_qp.py:643-647, 788-797);
 * bit 1: optimality measure max_elem*(NN(x_rho, Q~_rho) - <Q~_rho, X_rho>) (cut_select_qp.py:573-582);
 * bit 2 (instead of bit 1): the EXACT measure max_elem*(v - <Q~_rho, X_rho>), v = min <Q~_rho, X> over
 * [[X, x_rho], [x_rho^T, 1]] PSD, diag(X) <= x_rho -- what the reference obtains from Mosek per sub-problem
 * (strat 3, cut_select_qp.py:555-567, 584-598) -- by a batched barrier solver, value within rho * sdp_mu_final.
 * Scores stay resident on the device for sdpcs_topk / sdpcs_scores.  vars_values == NULL re-uses the LP point
 * already resident on the device from the previous call (device-resident timing). */
int sdpcs_score(sdpcs_ctx *ctx, const double *vars_values, int want);

/* Copy resident scores of local candidates [i0, i1) to the host (either pointer may be NULL). Parity tests. */
int sdpcs_scores(sdpcs_ctx *ctx, int64_t i0, int64_t i1, double *out_lam, double *out_obj);

/* Counters over the resident scores: out[0] = N, out[1] = #violated (lam < thres_neg_eigval),
 * out[2] = #strong (obj > thres_min_opt and violated).  (cut_select_qp.py:604-615, 629) */
int sdpcs_counts(sdpcs_ctx *ctx, int64_t *out3);

/* Largest obj among the candidates with obj > thres_min_opt that are NOT violated, as seen by the last sdpcs_topk pass
 * (-inf if none).  With it a sharded caller can tell when the combined rule (cut_select_qp.py:603-625) reduces to the
 * strong list itself: n_strong >= k and max - big_m < pivot_obj + big_m (every one of the k strong elements up to the
 * pivot is re-scored obj + big_m, nothing else can reach them), so that the second selection pass can be skipped. */
int sdpcs_max_pos_nonviolated(sdpcs_ctx *ctx, double *out);

/* Top-k of the resident scores.  mode:
 *   1  feasibility: violated only, key -lam desc, ties agg_idx asc          (cut_select_qp.py:639-654)
 *   2  optimality:  all, key obj desc, ties agg_idx asc                      (cut_select_qp.py:599-601)
 *   3  strong set:  obj > thres_min_opt and violated, key obj desc           (cut_select_qp.py:607-612)
 *   4  combined final: key = re-scored measure of cut_select_qp.py:603-625 given the pivot
 *      (pivot_obj, pivot_idx = the k-th strong candidate, all_walked = |S| < k), ties by (obj desc, agg_idx asc)
 * Outputs (length k, first *out_n valid, already in final order): agg_idx, key score, lam, obj. */
int sdpcs_topk(sdpcs_ctx *ctx, int mode, int64_t k, double pivot_obj, int64_t pivot_idx, int all_walked,
               int64_t *out_idx, double *out_score, double *out_lam, double *out_obj, int64_t *out_n);

/* Guard band of the last sdpcs_topk / sdpcs_select pass (see sdpcs_params.guard_*): the candidates ranked right after the
 * winners whose primary score (-lam in mode 1, obj in modes 2 and 3, the re-scored measure in mode 4; for sdpcs_select
 * strat 4 on the strong-prefix path: obj) is within the guard of the k-th score, in selection order; first min(cap, n)
 * are written, *out_n = how many.  out_info[4] (may be NULL) = {n_band, band_open, n_unc_lam, n_unc_obj}:
 * n_band near ties after the winners; band_open = 1 if more near ties exist than could be collected (a tie class larger
 * than band_cap); n_unc_lam / n_unc_obj = candidates of the whole cover with |lam - thres_neg_eigval| <= guard_lam /
 * |obj - thres_min_opt| <= guard_obj, i.e. whose violated / positive classification round-off could flip.
 * A selection is free of near-tie ambiguity iff n_band = 0, no two consecutive winners are closer than the guard and,
 * where the strategy classifies, n_unc_* = 0. */
int sdpcs_last_band(sdpcs_ctx *ctx, int64_t cap, int64_t *out_idx, double *out_score, double *out_lam,
                    double *out_obj, int64_t *out_n, int64_t *out_info);

/* Merge m entries gathered from several shards into the global top-k with the same comparator
 * (score desc, obj2 desc, agg_idx asc); perm receives the positions of the winners, in order. */
int sdpcs_merge_topk(sdpcs_ctx *ctx, int64_t m, const double *score, const double *obj2, const int64_t *idx,
                     int64_t k, int64_t *out_perm, int64_t *out_n);

/* Multi-GPU exchange without a host hop.  sdpcs_topk_pack_dev = sdpcs_topk whose result stays on the device,
 * packed into the caller's DEVICE buffer d_block of (2 + k + band_rows) x 4 doubles: two header rows
 * [len, N_local, #violated, #strong], [max positive non-violated obj, n_unc_lam, n_unc_obj, band open] and the local
 * winners followed by up to band_rows near ties as rows [agg_idx, score, lam, obj].  The ranks all-gather their blocks
 * (ncclAllGather / torch.distributed on the context's stream: 16 (k + band_rows + 2) B per rank), then
 * sdpcs_merge_packed_dev merges the `world` gathered blocks on the device with the selection comparator
 * (score desc, obj desc if use_obj2, agg_idx asc) and downloads only the global winners (*out_n <= k) followed by
 * their guard band (*out_band rows within delta of the k-th score); out_* have capacity out_cap.
 * out_hdr[8] = {rows merged, sum N, sum #violated, sum #strong, max of the extras, sum n_unc_lam, sum n_unc_obj,
 * band open}. */
int sdpcs_topk_pack_dev(sdpcs_ctx *ctx, int mode, int64_t k, double pivot_obj, int64_t pivot_idx, int all_walked,
                        int64_t band_rows, void *d_block);
int sdpcs_merge_packed_dev(sdpcs_ctx *ctx, const void *d_gathered, int world, int64_t rows_cap, int64_t k, int use_obj2,
                           double delta, int64_t out_cap, int64_t *out_idx, double *out_score, double *out_lam,
                           double *out_obj, int64_t *out_n, int64_t *out_band, double *out_hdr);

/* One-call selection on a single GPU with HOST buffers (upload + score + select + download):
 * the whole of _sel_eigcut_by_ordering_on_measure (cut_select_qp.py:543-654) for strat 1, 2, 3, 4, returning the
 * prefix of length <= k of the ranked list.  out_counts = {N, #violated walked, #strong}; out_new_strat as
 * cut_select_qp.py:629 (strat 4 only, else = strat).  vars_values == NULL: use the resident LP point. */
int sdpcs_select(sdpcs_ctx *ctx, int strat, const double *vars_values, int64_t k,
                 int64_t *out_idx, double *out_score, double *out_lam, double *out_obj,
                 int64_t *out_n, int64_t *out_counts, int *out_new_strat);

/* Lexicographic unranking (host utility, no GPU): ranks[m] -> out_idx[m x rho]. */
int sdpcs_unrank(int n, int rho, const int64_t *ranks, int64_t m, int32_t *out_idx);
int sdpcs_binom(int n, int k, int64_t *out);

/* Eigenvector cuts for m selected subsets (cut_select_qp.py:737-751): sets is m x rho int16 (-1 padded).
 * width = rho + rho(rho+1)/2.  out_ind (m x width, -1 padded): LP columns [x vars | X vars];
 * out_val: coefficients; out_rhs = -v0^2; out_lam: lam_min (eigh); out_violated: lam < thres_neg_eigval;
 * out_gap (may be NULL): second smallest eigenvalue minus lam_min -- when it is ~0 the eigenvector (and with it the
 * cut) is not unique and the reference's row is whatever LAPACK returned; the Python mirror recomputes those rows
 * with numpy eigh.  vars_values holds n(n+1)/2 + n doubles. */
int sdpcs_gen_cuts(sdpcs_ctx *ctx, int rho, const int16_t *sets, int64_t m, const double *vars_values,
                   int64_t *out_ind, double *out_val, double *out_rhs, double *out_lam, uint8_t *out_violated,
                   double *out_gap);

/* Row emission in one shot.  The violated cuts of sdpcs_gen_cuts as CSR rows -- row starts
 * out_rowptr[rows+1], LP columns out_ind, coefficients out_val (capacity m x width each), right-hand sides out_rhs[m],
 * sense >= for every row: the arrays CPXaddrows takes, instead of one cplex.SparsePair per cut
 * (cut_select_qp.py:747-754).  out_src[r] (may be NULL) = index into `sets` of row r; out_lam / out_gap (may be NULL) =
 * lam_min and eigenvalue gap of row r; *out_nrows = #rows.  Rows are emitted for lam_min < thres_neg_eigval + guard_lam
 * (with the default guard a subset within 1e-12 of the threshold is emitted and left to the caller). */
int sdpcs_gen_cuts_csr(sdpcs_ctx *ctx, int rho, const int16_t *sets, int64_t m, const double *vars_values,
                       int64_t *out_rowptr, int64_t *out_ind, double *out_val, double *out_rhs, int64_t *out_src,
                       double *out_lam, double *out_gap, int64_t *out_nrows);

/* Triangle-inequality rows (cut_select_qp.py:846-860) for m (triple lex rank, type) pairs as returned by
 * sdpcs_triangles, as CSR: 4 entries per row for types 0..2 (rhs 0), 6 for type 3 (rhs -1), sense >=.
 * Host utility (no GPU): out_rowptr[m+1], out_ind / out_val capacity 6 m, out_rhs[m]. */
int sdpcs_triangle_rows_csr(int n, const int64_t *triple_rank, const int8_t *type, int64_t m, int64_t *out_rowptr,
                            int64_t *out_ind, double *out_val, double *out_rhs);

/* Dense eigenvalue cuts, strat 0 (CutSolver.__gen_dense_eigcuts, cut_select_qp.py:757-786): one
 * eigen-decomposition of the full [1 x^T; x X] (order n + 1 <= 256) on the device; every eigenvalue among the n
 * smallest below thres_neg_eigval gives one dense row of width n + n(n+1)/2 over the LP columns
 * [nb_lifted + i, i < n | 0 .. nb_lifted) (x variables first, then X upper-triangular row-major):
 * coefficients out_val (row-major, capacity max_cuts rows), right-hand side out_rhs = -v0^2, sense >=.
 * out_eigvals (may be NULL): the n + 1 eigenvalues ascending.  *out_ncuts = number of rows; if it exceeds max_cuts
 * the call fails with SDPCS_ERR_INVALID after setting *out_ncuts (max_cuts = n always suffices). */
int sdpcs_dense_eigcuts(sdpcs_ctx *ctx, const double *vars_values, int64_t max_cuts, double *out_eigvals,
                        int64_t *out_ncuts, double *out_val, double *out_rhs);

/* Eigen-decomposition of one [1 x^T; x X] matrix of order d+1 (cut_select_qp.py:788-797): eigenvalues
 * ascending, eigenvectors as columns of V (row-major (d+1)x(d+1)); out_vecs may be NULL. */
int sdpcs_eigendecomp(sdpcs_ctx *ctx, int d, const double *curr_pt, const double *X_slice,
                      double *out_vals, double *out_vecs);

/* Triangle inequalities (cut_select_qp.py:799-863).  adj: n x n uint8 sparsity pattern (NULL = dense).
 * sdpcs_triangles scores all triples with >= thres_tri_dense edges, keeps violations >= thres_tri_viol,
 * orders by (density, violation) desc with ties in (triple lex rank, type) order, and returns the first
 * min(kmax, #violated): lex rank of the triple, type 0..3, violation, density; *out_n_violated = #violated. */
int sdpcs_set_tri_pattern(sdpcs_ctx *ctx, const uint8_t *adj);
int sdpcs_triangles(sdpcs_ctx *ctx, const double *vars_values, int64_t kmax, int64_t *out_triple_rank,
                    int8_t *out_type, double *out_viol, int8_t *out_density, int64_t *out_n,
                    int64_t *out_n_violated, int64_t *out_n_triples);

/* Batched NN_rhoD forward pass on the GPU for m input rows of length rho(rho+3)/2 -- the replacement of
 * `neural_net_%dD(input_arr)` (cut_select_qp.py:579-582). */
int sdpcs_nn_eval(sdpcs_ctx *ctx, int rho, const double *inputs, int64_t m, double *out);

/* Test hook for the TCGEN05 engine: the scaled pre-activations z = -2 log2(e) (W a + b) of tansig layer
 * `layer` (0-based) for m input rows, out_z is m x 64 (neurons beyond the layer width are zero padded). */
int sdpcs_nn_debug_layer(sdpcs_ctx *ctx, int rho, const double *inputs, int64_t m, int layer, double *out_z);

/* Batched exact SDP values v(x, C) = min <C, X> s.t. [[X, x], [x^T, 1]] PSD, diag(X) <= x for m sub-problems of size d
 * in 2..5 (the Mosek model of cut_select_qp.py:555-567 and of the training-data sampler utilities.py:40-47).  in: m rows
 * [x (d) | C upper triangle row-major (d(d+1)/2)] with <C, X> = sum_{i<=j} C_ij X_ij; out[m]; out_iters (may be NULL):
 * Newton steps taken. */
int sdpcs_sdp_solve(sdpcs_ctx *ctx, int d, const double *in, int64_t m, double *out, int32_t *out_iters);

/* FP64 roofline denominators measured on this device: DFMA and DMMA.8x8x4 peak TFLOP/s. */
int sdpcs_fp64_peak(sdpcs_ctx *ctx, double *dfma_tflops, double *dmma_tflops);

#ifdef __cplusplus
}
#endif
#endif
