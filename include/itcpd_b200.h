/* libitcpd_b200 -- B200-native (sm_100a) CP-ALS engine behind ITensorCPD.jl's algorithm hooks.
 *
 * Plain C ABI: opaque handle, raw pointers, 64-bit sizes, int status returns (0 = ok), no
 * exceptions, no torch / CUDA types in any signature.  Host buffers are caller-owned and only
 * touched during the call; every device buffer is owned by the handle.  One handle = one driver
 * thread (like the reference's C helper it is not re-entrant).  All calls are synchronous on
 * return unless stated otherwise.
 *
 * Layout conventions are the reference's: the dense target T is
 * column-major in the user's index order (first index fastest), factor A_n is I_n x R
 * column-major, Gram matrices R x R, lambda length R.  Sample / pivot matrices are int64,
 * 1-based (as Julia stores them), nsamp x (N-1) column-major.
 *
 * Each entry point cites the reference interface (file:line under the ITensorCPD.jl tree) that
 * it replaces.  INTEGRATION.md shows the `ccall` binding for each.
 */
#ifndef ITCPD_B200_H
#define ITCPD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct itcpd_ctx itcpd_ctx;

/* status codes */
enum {
    ITCPD_OK = 0,
    ITCPD_ERR_CUDA = 1,        /* a CUDA runtime / driver call failed (see itcpd_last_error) */
    ITCPD_ERR_ARG = 2,         /* bad argument / call order */
    ITCPD_ERR_NO_DEVICE = 3,   /* no sm_100 device: the library has NO CPU fallback */
    ITCPD_ERR_NAN = 4,         /* NaN fit (reference: throw("Error NAN"), fit_check.jl:40-42) */
    ITCPD_ERR_COMM = 5,        /* NCCL / peer-memory failure */
    ITCPD_ERR_UNSUPPORTED = 6
};

/* solve paths reported by itcpd_solve (ldiv_solve.jl:13-29) */
enum { ITCPD_SOLVE_CHOLESKY = 0, ITCPD_SOLVE_QRCP = 1 };

/* MTTKRP algorithms (all give the same M_n; they differ in how T is streamed) */
enum {
    ITCPD_MTTKRP_TREE = 0,   /* default: two-pass dimension tree, TMA + FP64 DMMA GEMM            */
    ITCPD_MTTKRP_DIRECT = 1  /* one fused pass per mode, plain FMA kernel (debug / cross-check)   */
};

int         itcpd_version(void);
const char *itcpd_last_error(void);

/* ---- context ------------------------------------------------------------------------------ */
/* Replaces nothing in the reference (which has no device state); mirrors the build/load
 * convention of deps/build.jl:1-28 + src/ITensorCPD.jl:25-36. */
int itcpd_create(itcpd_ctx **out, int device);
int itcpd_destroy(itcpd_ctx *ctx);
int itcpd_device_info(itcpd_ctx *ctx, int *sm_count, int *cc_major, int *cc_minor, int64_t *hbm_bytes);
int itcpd_synchronize(itcpd_ctx *ctx);
/* number of kernels this handle has launched so far (bench.py's gpu_launches) */
int64_t itcpd_launch_count(itcpd_ctx *ctx);
/* runtime options (name, values; * = default).  Every path below has run on B200 hardware:
 *   "mttkrp_alg"     0* dimension tree (KRPFreeNormal / KRPNormal results), 1 direct one-pass-per-mode MTTKRP
 *   "split_a","split_b"  force the dimension-tree split points (0* = traffic cost model)
 *   "tile_warps"     4 | 8*  warps per GEMM CTA;  "swizzle" 1* | 0 (debug);  "tma3d" 1* | 0;  "stream_k" 0 | 1* | 2
 *   "overlap_factor" 1* Gram-Hadamard + Cholesky on a side stream under the GEMM;  "use_graph" 1* CUDA-graph replay of sweeps
 *   "graph_single"   1* | 0: repeated itcpd_sweep(1) calls -- the per-iteration loop of the reference API -- capture the sweep graph
 *                    on the second call and replay it afterwards (env ITCPD_GRAPH_SINGLE)
 *   "chol_alg"       0 block kernel (any R), 1 team kernel (R <= 128, bitwise equal to 0), 2 right-looking kernels (R <= 128),
 *                    3* right-looking where the factorisation is exposed, where R > 64 and under short GEMM passes, team kernel elsewhere
 *   "chol_short_gflop" 50*: GEMM passes below this many GFLOP count as short for "chol_alg" = 3 (0: never)
 *   "seqrcs_use_omega" 0* | 1: itcpd_seqrcs lists its candidate columns in increasing order, the order SEQRCS(...; use_omega = true)
 *                    produces (SEQRCS.jl:109-113; the sketch itself is bitwise the same either way, see itcpd_sketch_unfolding_csc);
 *                    0 = the matrix-free variant's order (:159), the reference's default
 *   "sketch_unfold"  1* | 0: the SE-QRCS set-up sketches every mode but the first from an explicit unfolding (a second copy of the tensor
 *                    for the duration of the call, made by a tiled transpose; skipped when HBM has no room) instead of gathering
 *                    its strided columns in place; bitwise the same sketch, 10x faster.  2 = also in itcpd_sketch_unfolding[_csc]
 *                    (how the tests compare the two paths)
 *   "staged_upload"  1* | 0: a pageable host tensor is uploaded through pinned staging buffers filled by host threads
 *   "peer_graph"     1* | 0: sharded sweeps without an NCCL call inside (device-side exchange epochs, small all-reduces over the
 *                    peer-mapped buffer), so that they are captured like single-GPU sweeps; set before itcpd_peer_export
 *   "early_pass_b"   0* | 1: pass B of the dimension tree on its own stream as soon as the modes it contracts are updated (bitwise the
 *                    default schedule; measured: no gain) (env ITCPD_EARLY_B)
 *   "gemm_i8"        0* | 1 | 2: OPT-IN contraction on the INT8 tensor cores (tcgen05.mma kind::i8, TMEM accumulators) on 6 / 7 base-256
 *                    digits per operand -- 48/56-bit fixed point per row, NOT the FP64 arithmetic of the default path; 1 converts T on the
 *                    fly, 2 keeps T's digit planes pre-packed in HBM (6 B / element / unfolding) (csrc/gemm_i8.cu)
 *   "i8_spare_sms"   0* .. 63: SMs the persistent INT8 GEMM leaves to the side-stream factorisation
 *   "time_gemm"      1: CUDA events around every GEMM launch (itcpd_gemm_timing); disables the graph
 *   "time_phases"    1: CUDA events after every phase of a mode update (itcpd_phase_timing); disables the graph
 * environment at itcpd_create: ITCPD_CHOL=0|1|2|3, ITCPD_NO_GRAPH=1, ITCPD_NO_SWIZZLE=1, ITCPD_GEMM_I8=1|2, ITCPD_EARLY_B=1,
 * ITCPD_GRAPH_SINGLE=0|1, ITCPD_I8_SPARE_SMS=n */
int itcpd_set_option(itcpd_ctx *ctx, const char *name, int64_t value);

/* ---- target tensor (ALS.target, als_optimizer.jl:5-10; decompose.jl:5-7 wraps without copy) - */
int itcpd_set_tensor(itcpd_ctx *ctx, int order, const int64_t *dims, const double *host_colmajor);
/* Shape only: tells the handle the extents of the target without allocating or touching any tensor storage (the
 * handle then serves reconstruct / CPD-only calls; every call that reads T fails with ITCPD_ERR_ARG).  Replaces the
 * `ITensor(inds(target))` shape carrier of the reference (optimizers/.../randomized/qr_lev_score_sampled.jl:77). */
int itcpd_set_shape(itcpd_ctx *ctx, int order, const int64_t *dims);
/* i.i.d. N(0,1) entries from a counter-based generator (Philox4x32-10 + Box-Muller); element with
 * global column-major linear index e gets the value of counter (seed, e + elem_offset), so slabs of
 * one big tensor can be generated independently on several GPUs. */
int itcpd_generate_tensor(itcpd_ctx *ctx, int order, const int64_t *dims, uint64_t seed, int64_t elem_offset);
/* synthetic low-rank + noise target generated on the device: T = sum_r prod_n A_n[i_n, r] + noise * N(0,1) with
 * A_n = column-normalised randn(I_n, rank) from the counter-based stream `seed` (accuracy studies at sizes where
 * a host tensor is impractical).  Leaves the handle's CPD state at the planted factors. */
int itcpd_generate_lowrank_tensor(itcpd_ctx *ctx, int order, const int64_t *dims, int rank, uint64_t seed, double noise);
int itcpd_get_tensor(itcpd_ctx *ctx, double *host_colmajor);
int itcpd_tensor_norm(itcpd_ctx *ctx, double *fro_norm);            /* norm(T), README.md:96 */

/* ---- CPD state (CPD{TargetT}, cpd.jl:7-46) ------------------------------------------------- */
int itcpd_set_rank(itcpd_ctx *ctx, int rank);
int itcpd_set_factor(itcpd_ctx *ctx, int mode, const double *host);  /* cp[n]   cpd.jl:28 */
int itcpd_get_factor(itcpd_ctx *ctx, int mode, double *host);
int itcpd_set_lambda(itcpd_ctx *ctx, const double *host);            /* cp[]    cpd.jl:29 */
int itcpd_get_lambda(itcpd_ctx *ctx, double *host);
int itcpd_get_gram(itcpd_ctx *ctx, int mode, double *host);          /* :part_grammian[n] */
/* random_factors (cpd.jl:48-60): randn(I_n,R) per mode from one counter-based stream, column
 * normalised; lambda = norms of the last factor.  (Julia's MersenneTwister stream is not
 * reproducible outside Julia; the Julia extension passes its own factors via itcpd_set_factor.) */
int itcpd_random_cpd(itcpd_ctx *ctx, uint64_t seed);

/* ---- the five per-mode hooks of optimize.jl:19-30 ------------------------------------------ */
/* compute_als(::MttkrpAlgorithm) (optimizers/.../standard/tensor.jl:3-14): all Grams G_n = A_n'A_n */
int itcpd_compute_grams(itcpd_ctx *ctx);
/* compute_krp(::MttkrpAlgorithm) (MttkrpAlgorithm.jl:18-31): Gamma = hadamard_{m != mode} G_m.
 * host_out (R x R) may be NULL. */
int itcpd_gram_hadamard(itcpd_ctx *ctx, int mode, double *host_out);
/* matricize_tensor(::KRPFreeNormal / ::KRPNormal) (algorithms/.../standard/tensor.jl:12-44 ->
 * had_contract.jl:72-124): M_mode = T_(mode) * KRP(other factors).  host_out (I_mode x R) may be NULL. */
int itcpd_mttkrp(itcpd_ctx *ctx, int mode, double *host_out);
/* solve_ls_problem (MttkrpAlgorithm.jl:34-41) -> ldiv_solve! (ldiv_solve.jl:13-29): pivoted
 * Cholesky (dpstrf semantics, absolute tol, upper) + permuted triangular solves; on rank
 * deficiency the column-pivoted-QR min-norm solve (dgelsy semantics, rcond = R*eps).
 * Operates on the device-resident Gamma and M_mode; path_out/rank_out may be NULL. */
int itcpd_solve(itcpd_ctx *ctx, int mode, double chol_tol, int *path_out, int *rank_out);
/* (path, rank) of the most recent R x R solve on this handle, whichever entry point ran it (itcpd_solve, the sampled / projected
 * updates, the last mode of a sweep): ITCPD_SOLVE_CHOLESKY, or ITCPD_SOLVE_QRCP when the pivoted Cholesky met a pivot <= tol and
 * the pivoted-QR min-norm fallback of ldiv_solve.jl:19-21 was taken. */
int itcpd_last_solve_status(itcpd_ctx *ctx, int mode_slot, int *path_out, int *rank_out);
/* row_norm (math_tools/row_norm.jl:4-24): lambda_r = ||X[:,r]||, A_mode = X ./ lambda */
int itcpd_normalize(itcpd_ctx *ctx, int mode);
/* post_solve (tensor.jl:46-49): G_mode = A_mode' A_mode */
int itcpd_post_solve(itcpd_ctx *ctx, int mode);
/* check_converge(::FitCheck) scalars (fit_check.jl:28-29, converge_checks.jl:5-11):
 * inner = sum(M_N .* A_N .* lambda), model_norm2 = lambda' (hadamard_n G_n) lambda.  Uses the
 * saved last-mode MTTKRP: no extra tensor pass. */
int itcpd_fit_terms(itcpd_ctx *ctx, double *inner, double *model_norm2);

/* CPDiffCheck / CPAngleCheck (converge_checks/cp_diff_check.jl:20-71, cp_angle_check.jl:20-73): the checks compare the
 * CPD of consecutive sweeps through factor matrices only.  snapshot = `check.PrevCP = CPD(factors, lambda)`;
 * diff_terms = `(PrevCP.lambda * cp_cp_contract(PrevCP, currCP)[1] * lambda)[]` and `norm_factors(grams, lambda)`. */
int itcpd_cpd_snapshot(itcpd_ctx *ctx);
int itcpd_cpd_diff_terms(itcpd_ctx *ctx, double *inner_prev_curr, double *norm2_curr);

/* ---- whole sweeps (optimize.jl:15-32 body), device resident --------------------------------- */
/* Runs `nsweeps` ALS sweeps (modes 1..N in order, Gram refresh after each mode, fit scalars after
 * each sweep).  inner[] / model_norm2[] (length nsweeps) may be NULL.  The host keeps the
 * FitCheck / NoCheck state machine (fit_check.jl:30-65) and decides when to stop. */
int itcpd_sweep(itcpd_ctx *ctx, int nsweeps, double chol_tol, double *inner, double *model_norm2);
/* same, but only enqueues the work on the handle's stream; results land in the handle's pinned
 * staging area and are fetched by itcpd_sweep_results after itcpd_synchronize. */
int itcpd_sweep_async(itcpd_ctx *ctx, int nsweeps, double chol_tol);
int itcpd_sweep_results(itcpd_ctx *ctx, int nsweeps, double *inner, double *model_norm2, int *qrcp_fallbacks);
/* als_optimize (als_optimizer.jl:15-25) end to end from HOST buffers: upload T and the N initial
 * factors, run nsweeps sweeps, download factors + lambda + per-sweep fit scalars. */
int itcpd_als_from_host(itcpd_ctx *ctx, int order, const int64_t *dims, const double *host_T, int rank,
                        const double *const *host_factors_in, int nsweeps, double chol_tol,
                        double *const *host_factors_out, double *host_lambda_out,
                        double *inner, double *model_norm2);

/* ---- reconstruct (algebra/reconstruct.jl:2-9) and residual ------------------------------------ */
int itcpd_reconstruct(itcpd_ctx *ctx, double *host_colmajor);
/* ||T - [[lambda; A_1..A_N]]||_F without materialising the reconstruction */
int itcpd_residual_norm(itcpd_ctx *ctx, double *fro_norm);

/* ---- sampled / randomized path ---------------------------------------------------------------- */
/* compute_leverage_score_probabilitiy (math_tools/probability.jl:3-10): p_i = ||Q[i,:]||^2/min(I,R) */
int itcpd_leverage_scores(itcpd_ctx *ctx, int mode, double *host_out);
/* sample_factor_matrices (probability.jl:23-33): nsamp i.i.d. weighted draws with replacement for
 * every mode != skip_mode from the current device-side leverage scores. out: int64 1-based,
 * nsamp x (N-1) column-major. */
int itcpd_sample_factor_matrices(itcpd_ctx *ctx, int skip_mode, int64_t nsamp, uint64_t seed, int64_t *host_out);
/* pivot_hadamard (algebra/had_contract.jl:277-295): K[s,r] = prod_{m != mode} A_m[piv[s,m], r] */
int itcpd_pivot_hadamard(itcpd_ctx *ctx, int mode, int64_t nsamp, const int64_t *host_pivots, double *host_out);
/* fused_flatten_sample (algebra/pivot_mapping.jl:59-85): out[:, s] = mode fibre of T at piv[s,:] */
int itcpd_gather_fibers(itcpd_ctx *ctx, int mode, int64_t nsamp, const int64_t *host_pivots, double *host_out);
/* column_to_multi_coords / multi_coords_to_column (pivot_mapping.jl:17-47), host-side integer maps */
int itcpd_column_to_multi_coords(int64_t ncols, const int64_t *cols, int ndims, const int64_t *dims, int64_t *out);
int itcpd_multi_coords_to_column(int64_t ncols, const int64_t *coords, int ndims, const int64_t *dims, int64_t *out);
/* sparse-sign embeddings: same signature and libc-rand() stream as the reference's C helper
 * (algebra/sparse_sign.c:25-70, algebra/sparsestack.c:24-79; bound at SEQRCS.jl:41-60). */
void itcpd_sparse_sign(int l, int n, int s, double *vals, int *rows, int *colstarts);
void itcpd_sparsestack(int l, int n, int s, double *vals, int *rows, int *colstarts);
/* 1 if the two generators above read libc's rand() stream without its per-call lock (glibc: the state array is borrowed through
 * initstate()/setstate(), advanced in place and handed back, after a self-test on private state; 2.3x faster, the same numbers and
 * the same continuation of the stream), 0 if they call rand() (any other libc, or ITCPD_PLAIN_RAND set in the environment).
 * While a generator call is in progress libc sits on a scratch state: a rand() call made by ANOTHER thread during that window draws
 * from the scratch state instead of interleaving with the generator (with plain rand() the two would interleave; neither is
 * reproducible, and the reference's generators have the same single-stream assumption). */
int itcpd_sparse_sign_fast_stream(void);
/* sketched_matricization (pivot_mapping.jl:111-140): A_sk = T_(mode) * Omega' (I_mode x l) from the
 * (rows 0-based, vals) arrays the generators above fill; s non-zeros per column. */
int itcpd_sketch_unfolding(itcpd_ctx *ctx, int mode, int l, int s, const int *rows0, const double *vals, double *host_out);
/* The sparse-matrix variant of the same sketch (src/algebra/pivot_mapping.jl:90-104; what SEQRCS(...; use_omega = true) calls,
 * src/algebra/SEQRCS.jl:89-134): Omega (l x ncols) as Julia's SparseMatrixCSC stores it -- colptr (ncols + 1), rowval, both 1-based
 * Int64, and nzval.  Same kernel, same summation order as the matrix-free variant. */
int itcpd_sketch_unfolding_csc(itcpd_ctx *ctx, int mode, int l, int64_t ncols, const int64_t *colptr, const int64_t *rowval,
                               const double *nzval, double *host_out);
/* one sampled ALS mode update (ProjectionAlgorithm.jl:57-68): given 1-based pivots for `mode`, gathers T_s and K on
 * the device and solves  normal != 0: (K'K) \ (T_s K)'  (pivoted Cholesky, QRCP fallback)
 *                        normal == 0: qr(K, ColumnNorm()) \ T_s'  (pivoted-QR min-norm least squares, nsamp >= R);
 * then normalises and refreshes the Gram and the leverage scores of `mode` (post_solve of LevScoreSampled).
 * On a slab-sharded handle (itcpd_comm_init) the pivots are GLOBAL coordinates (the last mode spans all slabs) and must
 * be identical on every rank: each rank contributes the samples whose last-mode coordinate lies in its slab and the
 * sampled normal equations are all-reduced (csrc/sampled_sharded.cu); itcpd_leverage_scores then returns the scores of
 * the LOCAL rows of the sharded factor, itcpd_sample_factor_matrices draws from the all-gathered scores. */
int itcpd_sampled_update(itcpd_ctx *ctx, int mode, int64_t nsamp, const int64_t *host_pivots, double chol_tol, int normal);

/* Device-resident sweeps of the leverage-score sampled solver: for every mode a weighted draw of nsamp[mode] samples, the sampled
 * least-squares update, row_norm and the leverage refresh -- optimize.jl:17-28 with the ProjectionAlgorithm hooks of LevScoreSampled
 * (algorithms/als_algorithms/randomized/krp_lev_score_sampled.jl:9-58, ProjectionAlgorithm.jl:57-68) and no host round trip.
 * The k-th draw of the call is seeded with draw_counter + k, exactly the seeds a host loop of itcpd_sample_factor_matrices(seed) +
 * itcpd_sampled_update passes (the two drivers are bitwise equal); the sweep body is replayed from a CUDA graph.  Asynchronous:
 * returns once the sweeps are enqueued.  Single-GPU handles only. */
int itcpd_sampled_sweep_async(itcpd_ctx *ctx, int nsweeps, const int64_t *nsamp, uint64_t draw_counter, double chol_tol, int normal);

/* qr(T_(mode), ColumnNorm()) of the pivot-projected setup (optimizers/.../randomized/qr_lev_score_sampled.jl:22-23,126-127):
 * column-pivoted Householder QR of the mode unfolding on the device.  piv_out: the full pivot order, n = P / I_mode
 * int64 entries, 1-based; rdiag_out: diag(R), min(I_mode, n) doubles. */
int itcpd_qrcp_unfolding(itcpd_ctx *ctx, int mode, int64_t *piv_out, double *rdiag_out);
/* the same factorisation for a caller-supplied column-major m x n host matrix; `steps` eliminations (<= min(m,n)) */
int itcpd_qrcp_matrix(itcpd_ctx *ctx, int64_t m, int64_t n, const double *host_A, int64_t steps, int64_t *piv_out, double *rdiag_out);
/* SEQRCS, matrix-free variant with compute_r = false (algebra/SEQRCS.jl:139-182): sparse-sign sketch of the unfolding
 * (libc rand() stream of the reference's generators), QRCP of the sketch, first t sketch pivots -> candidate columns,
 * QRCP of the gathered candidate columns, p = [candidates[p_subset]; remaining columns].  piv_out: n int64 1-based;
 * rdiag_out: diag(R) of the candidate QR (nrdiag_out values, caller provides I_mode doubles). */
int itcpd_seqrcs(itcpd_ctx *ctx, int mode, int l, int s, int t, int injective, int64_t *piv_out, double *rdiag_out,
                 int64_t *nrdiag_out, int64_t *ncand_out);
/* The same for several modes in one call -- the loop of optimizers/.../randomized/qr_lev_score_sampled.jl:80-176 over its random
 * modes.  Embeddings are generated in the order of `modes`, so the libc rand() stream is consumed exactly as by per-mode calls in
 * that order; the host half of mode i+1 (the reference's generator, the sort of the sketch by row) runs on a helper thread while
 * the device factorises mode i.  seeds: NULL, or one value per mode -- >= 0: srand(seed) right before that mode's generator
 * (tests / reproducible runs), < 0: the stream continues.  piv_out[i]: P / I_modes[i] int64; rdiag_out: NULL or one pointer per mode
 * (NULL entries allowed); nrdiag_out / ncand_out: NULL or nmodes values. */
int itcpd_seqrcs_modes(itcpd_ctx *ctx, int nmodes, const int *modes, const int *l, const int *s, const int *t, int injective,
                       const int64_t *seeds, int64_t *const *piv_out, double *const *rdiag_out, int64_t *nrdiag_out, int64_t *ncand_out);
/* KRP-structured SE-QRCS (algebra/SEQRCS.jl:184-241, compute_r = false): the same procedure applied to the Khatri-Rao
 * product of the handle's CURRENT factors of every mode != `mode` (R x n, never formed): the sketch is
 * omega_hadamard (algebra/had_contract.jl:300-329), candidates are gathered with pivot_hadamard.  Outputs as itcpd_seqrcs. */
int itcpd_seqrcs_krp(itcpd_ctx *ctx, int mode, int l, int s, int t, int injective, int64_t *piv_out, double *rdiag_out,
                     int64_t *nrdiag_out, int64_t *ncand_out);
/* pivot-projected solvers: cache the projector of `mode` and its sampled target T_s = fused_flatten_sample(T, mode, piv)
 * on the device (qr_lev_score_sampled.jl:64-65, 162-163); then one mode update per call
 * (ProjectionAlgorithm.jl:57-68, `normal` as in itcpd_sampled_update; post_solve is a no-op for these solvers). */
int itcpd_set_projector(itcpd_ctx *ctx, int mode, int64_t nsamp, const int64_t *host_pivots);
int itcpd_projected_update(itcpd_ctx *ctx, int mode, double chol_tol, int normal);
/* after the setup only the samples are touched: release the dense tensor (ALS(ITensor(inds(target)), ...), :77,175) */
int itcpd_drop_tensor(itcpd_ctx *ctx);

/* ---- multi-GPU (slab sharding along the last mode; one process per GPU) ------------------------ */
/* The handle's tensor is the local slab T[..., slab]; factor N holds only the slab's rows.
 * After itcpd_comm_init every MTTKRP of a non-sharded mode is all-reduced, and the norms / Gram /
 * fit scalars of the sharded mode are all-reduced. */
int itcpd_comm_unique_id(void *out128);                       /* rank 0: ncclGetUniqueId */
int itcpd_comm_init(itcpd_ctx *ctx, int nranks, int rank, const void *id128);
int itcpd_comm_destroy(itcpd_ctx *ctx);
int itcpd_allgather_factor(itcpd_ctx *ctx, int mode, int64_t rows_total, double *host_out);
/* Fused all-reduce + solve over NVLink peer memory (optional, after itcpd_comm_init; same-node ranks, CUDA IPC).
 * export: allocates this rank's exchange buffer and returns its 64-byte IPC handle; the launcher all-gathers the
 * handles; import: maps every peer's buffer.  From then on the M_n all-reduce of the non-sharded modes is not an NCCL
 * call: the partial MTTKRP is published in the exchange buffer and every rank's row-solve kernel sums the peers'
 * partials (fixed rank order, so all ranks get identical bits) while loading its right-hand sides. */
int itcpd_peer_export(itcpd_ctx *ctx, void *handle64_out);
int itcpd_peer_import(itcpd_ctx *ctx, int nranks, int rank, const void *handles);
int itcpd_peer_disable(itcpd_ctx *ctx);

/* ---- measurement helpers ------------------------------------------------------------------- */
/* average device time (ms, CUDA events on the handle's stream) of the dominant GEMM kernel over
 * the launches since the last reset, and how many launches that was */
int itcpd_gemm_timing(itcpd_ctx *ctx, int reset, double *avg_ms, int64_t *launches);
/* With option "time_phases" = 1 (disables the graph) the sweep driver records a CUDA event after every phase of a mode
 * update; this returns the accumulated milliseconds per phase id since the last reset:
 * 0 between modes, 1 MTTKRP (pack + GEMM + second level [+ NCCL all-reduce]), 2 peer signal, 3 solve (join with the
 * side-stream factorisation, peer wait, row solves), 4 normalise, 5 Gram refresh, 6 fit scalars.  Measurement only. */
int itcpd_phase_timing(itcpd_ctx *ctx, int reset, int nphases, double *ms_by_phase, int64_t *marks);
/* FP64 tensor-pipe peak probe: register-resident DMMA.8x8x4 issue loop on every SM; returns
 * achieved TFLOP/s (2*8*8*4 flops per instruction) -- a roofline denominator for this box. */
int itcpd_probe_dmma_peak(itcpd_ctx *ctx, double *tflops);
int itcpd_probe_dfma_peak(itcpd_ctx *ctx, double *tflops);
/* CUDA events on the handle's stream (slots 0..15): bench.py times exactly K sweeps on the device */
int itcpd_event_record(itcpd_ctx *ctx, int slot);
int itcpd_event_elapsed_ms(itcpd_ctx *ctx, int slot_start, int slot_stop, double *ms);
/* pinned (page-locked) host memory for the end-to-end path: host buffers the caller owns */
int itcpd_host_alloc(int64_t bytes, void **out);
int itcpd_host_free(void *p);
/* write `bytes` of a device scratch buffer (L2 flush between timed iterations of small problems) */
int itcpd_flush_l2(itcpd_ctx *ctx, int64_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* ITCPD_B200_H */
