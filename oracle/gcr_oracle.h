/*
 * oracle/gcr_oracle.h -- CPU restatement of the reference's hypothesise-and-verify hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it, and only as the checker / baseline.
 *
 * Every function restates, in plain FP64 C++ with no Eigen/OpenCV, one function of danini/graph-cut-ransac
 * (paths relative to /root/reference/src/pygcransac/include/), cited next to each declaration.  Arithmetic is
 * written in the reference's operation order and compiled with -ffp-contract=off so that integer results are
 * bit-exact and FP64 results differ from the reference only where the reference calls an Eigen decomposition
 * (Eigen is absent from the reference tree and from this image; those solvers follow Eigen 3.4 semantics and are
 * cross-checked against numpy/scipy -- see DESIGN.md "Oracle pinning").
 *
 * Pinning status: the reference ships no operator-level golden vectors.  The functions marked
 * [pinned:_ref] are checked bit-for-bit (or to 1e-15) against the reference's own headers compiled verbatim into
 * oracle/_ref/libgcr_ref.so (oracle/build_ref.sh); fixtures generated from that library are committed under
 * tests/golden/.  Functions marked [pinned:source] restate Eigen-dependent code: Eigen itself cannot be had here, but the
 * reference's own solver / estimator SOURCE compiles verbatim over a stand-in for the part of Eigen's API it uses whose
 * decompositions are oracle/linalg.h (oracle/shim_solvers, oracle/build_ref_solvers.sh -> oracle/_ref/libgcr_ref_solvers.so),
 * and tests/test_reference_solver_sources.py holds restatement and source side by side: coefficient layouts, polynomial
 * expansions, elimination templates, normalisation, orientation tests, the order and sign / scale of the returned models
 * are pinned to the reference; Eigen's own floating-point behaviour stays restated (oracle/linalg.h, pinned against numpy).
 */
#ifndef GCR_ORACLE_H
#define GCR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* problem ids (shared with include/gcr/gcr.h) */
enum { GCRO_H = 0, GCRO_F = 1, GCRO_E = 2, GCRO_PNP = 3 };
/* scoring ids */
enum { GCRO_MSAC = 0, GCRO_MAGSAC = 1 };

/* ---- residuals -------------------------------------------------------------------------------------------- */
/* estimators/homography_estimator.h:181-199  [pinned:_ref via scoring]                                         */
double gcro_sqres_h(const double *pt4, const double *H9);
/* estimators/fundamental_estimator.h:195-222 == estimators/essential_estimator.h:186-213                       */
double gcro_sqres_sampson(const double *pt4, const double *F9);
/* estimators/fundamental_estimator.h:225-253 (symmetric epipolar)                                              */
double gcro_sqres_symepi(const double *pt4, const double *F9);
/* estimators/perspective_n_point_estimator.h:148-184 (pt = u v X Y Z, model 3x4 row-major)                     */
double gcro_sqres_pnp(const double *pt5, const double *P12);
/* dispatch: problem -> squared residual; cols = 4 (H,F,E) or 5 (PNP)                                           */
double gcro_sqres(int problem, const double *pt, const double *model);
int gcro_model_len(int problem); /* 9 or 12 */
int gcro_point_cols(int problem); /* 4 or 5 */

/* ---- scoring ---------------------------------------------------------------------------------------------- */
/* Gamma(1.5, k/1000): value of gamma_values.cpp:2012 stored_gamma_values[k], k in [0, 38682]  [pinned:_ref]    */
double gcro_gamma_lut(int k);
/* scoring_function.h:137-233 MSACScoringFunction::getScore with squared_truncated_threshold = (1.5*thr)^2 as
 * GCRANSAC.h:253-255 sets it.  Writes score (0 if no inliers), inlier count, and (if inl_idx) ascending indices. */
void gcro_score_msac(int problem, const double *pts, int n, const double *model, double thr,
                     double *score, uint64_t *n_inl, uint32_t *inl_idx);
/* scoring_function.h:246-348 MAGSACScoringFunction::getScore (sigma_max = thr, inlier iff residual <= thr)      */
void gcro_score_magsac(int problem, const double *pts, int n, const double *model, double thr,
                       double *score, uint64_t *n_inl, uint32_t *inl_idx);
/* K models at once (models: k x model_len) */
void gcro_score_batch(int problem, int scoring, const double *pts, int n, const double *models, int k,
                      double thr, double *scores, uint64_t *n_inl);
/* per-point squared residuals of one model */
void gcro_residuals(int problem, const double *pts, int n, const double *model, double *out);

/* ---- samplers (a2, a3) ------------------------------------------------------------------------------------ */
/* Philox4x32-10 block: ctr[4], key[2] -> out[4] (the counter-based generator the product's K1 kernel uses)     */
void gcro_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* uniform_sampler.h:81-99 + uniform_random_generator.h:90-103 semantics (m distinct uniform indices in
 * [0, n-1] by draw-and-reject, kept in draw order) on the counter-based stream (seed, sample_index, stream).   */
void gcro_sample_uniform(uint64_t seed, uint64_t sample_index, uint32_t stream, uint32_t n, int m, uint32_t *out);
/* prosac_sampler.h:129-174 growth function T'_n (n entries)  [pinned:_ref]                                     */
void gcro_prosac_growth(uint32_t n, int m, uint64_t convergence_iters, uint64_t *growth);
/* prosac_sampler.h:65-88,207-236: subset size n_k the k-th (1-based) PROSAC draw uses  [pinned:_ref]           */
uint32_t gcro_prosac_subset_size(const uint64_t *growth, uint32_t n, int m, uint64_t k);
/* k-th PROSAC sample (1-based k) on the counter-based stream                                                   */
void gcro_sample_prosac(uint64_t seed, uint64_t k, uint32_t n, int m, const uint64_t *growth,
                        uint64_t convergence_iters, uint32_t *out);

/* importance_sampler.h:67-82: the cumulative table of the reference's std::discrete_distribution<int> (libstdc++
 * bits/random.tcc: normalise by the sum, partial sums, last entry 1); n entries
 * [pinned: std::discrete_distribution over the same word stream, oracle/_ref ref_sample_importance]            */
void gcro_importance_cdf(const double *probabilities, uint32_t n, double *cdf);
/* importance_sampler.h:126-137: m independent draws (with replacement) of that distribution; draw j of sample
 * `sample_index` takes two words of stream 3: p = generate_canonical<double, 53> = (w0 + w1 * 2^32) / 2^64,
 * index = lower_bound(cdf, p)                                                                                  */
void gcro_sample_importance(uint64_t seed, uint64_t sample_index, uint32_t n, int m, const double *cdf, uint32_t *out);

/* ---- homography (a4, a5, a7, a21) ------------------------------------------------------------------------- */
/* homography_estimator.h:346-381 */
int gcro_h_valid_sample(const double *pts, const uint32_t *sample);
/* solver_homography_four_point.h:109-190 + math_utils.h:45-87; returns #models (0/1). sample==NULL -> rows 0..3 */
int gcro_solve_h4(const double *pts, int cols, const uint32_t *sample, double *H9);
/* homography_estimator.h:326-342 */
int gcro_h_valid_model(const double *H9);
/* homography_estimator.h:213-310: Hartley normalisation of the selected points; T1,T2 3x3 row-major */
void gcro_normalize_points(const double *pts, int cols, const uint32_t *sample, int n, double *norm_pts,
                           double *T1, double *T2);
/* homography_estimator.h:136-173 + solver_homography_four_point.h:192-264; weights may be NULL.
 * Weight indexing follows the reference literally: weights[i] with i the POSITION in the sample. */
int gcro_refit_h(const double *pts, const uint32_t *sample, int n, const double *weights, double *H9);

/* ---- fundamental (a8, a9, a11, a22) ----------------------------------------------------------------------- */
/* solver_fundamental_matrix_seven_point.h:91-291; returns #models (0..3), F row-major 9 each  [pinned:source]       */
int gcro_solve_f7(const double *pts, const uint32_t *sample, double *F27);
/* fundamental_estimator.h:737-800 oriented epipolar constraint on the sample (1 valid / 0 not) */
int gcro_f_orientation_valid(const double *F9, const double *pts, const uint32_t *sample, int m);
/* fundamental_estimator.h:161-184 estimateModel = 7pt + orientation filter */
int gcro_estimate_f7(const double *pts, const uint32_t *sample, double *F27);
/* fundamental_estimator.h:574-618 + :637-734 + solver_fundamental_matrix_eight_point.h:93-182  [pinned:source]      */
int gcro_refit_f8(const double *pts, const uint32_t *sample, int n, const double *weights, double *F9);
/* LM bundle adjustment of F (relative_pose/bundle.cpp:253-339, jacobian_impl.h:472-606): refinement of F9 on n points
 * in the working frame; the estimator's default non-minimal fit (fundamental_estimator.h:574-618 with the BA solver);
 * the post-run polish of py/gcransac_python.cpp:1851-1858                                                        */
int gcro_refine_f_lm(const double *np, int n, const double *weights, double *F9, int max_iterations);
int gcro_refit_f_ba(const double *pts, const uint32_t *sample, int n, const double *weights, double *F9);
int gcro_polish_f_ba(const double *pts, const uint32_t *inliers, int n, const double *weights, double *F9);
/* DEGENSAC (fundamental_estimator.h:341-572): H-degenerate sample test -> 1 and the plane homography             */
int gcro_degensac_sample_homography(const double *F9, const double *pts, const uint32_t *sample, double sq_h_threshold,
                                    double *H9);
/* solver_fundamental_matrix_plane_and_parallax.h:109-160 (2 points + H), and with the orientation test (:161-184)  */
int gcro_solve_f_plane_parallax(const double *pts, const uint32_t *sample, const double *H9, double *F9);
int gcro_estimate_f_plane_parallax(const double *pts, const uint32_t *sample, const double *H9, double *F9);
/* fundamental_estimator.h:287-338 symmetric-epipolar ratio test part of isValidModel (no DEGENSAC)              */
int gcro_f_valid_model_ratio(const double *F9, const double *pts, const uint32_t *inl, int n_inl, int m,
                             double truncated_thr, double min_ratio);

/* ---- essential (a12, a13, a14) ---------------------------------------------------------------------------- */
/* solver_essential_matrix_five_point_nister.h:263-615; returns #models (0..10)  [pinned:source] */
int gcro_solve_e5_nister(const double *pts, const uint32_t *sample, double *E90);
/* solver_essential_matrix_five_point_stewenius.h:114-231 (n >= 5, optional weights[sample[i]])  [pinned:source]      */
int gcro_solve_e5_stewenius(const double *pts, const uint32_t *sample, int n, const double *weights, double *E90);
/* essential_estimator.h:306-336 estimateModelNonminimal = Stewenius + Frobenius normalisation + sign (E33 >= 0);
 * weights (optional) are read by POINT index (stewenius.h:143)  [pinned:source]                                      */
int gcro_refit_e(const double *pts, const uint32_t *sample, int n, const double *weights, double *E90);
/* essential_estimator.h:152-175 estimateModel = nister + orientation filter (:338-401) */
int gcro_estimate_e5(const double *pts, const uint32_t *sample, double *E90);
/* relative-pose bundle adjustment (gcr_oracle_eba.cpp): solver_essential_matrix_bundle_adjustment.h:113-186 */
int gcro_refine_e_lm(const double *pts, const uint32_t *sample, int n, const double *weights, const double *E9, double *out36);
int gcro_polish_e_ba(const double *pts, int n_points, const uint32_t *inliers, int n, const double *weights, double *out360);
/* utils.h:65-110 normalizeCorrespondences: K^-1 on both images (K row-major 3x3) */
void gcro_normalize_correspondences(const double *pts, int n, const double *K1, const double *K2, double *out);

/* test hooks: the dense kernels of oracle/linalg.h that stand in for Eigen (pinned against numpy) */
void gcro_test_colpiv_q(const double *A, int n, double *Q);
int gcro_test_eigenvalues(const double *A, int n, double *wr, double *wi);
void gcro_test_eigenvector(const double *A, int n, double lambda, double *v);

/* ---- absolute pose (a15) ---------------------------------------------------------------------------------- */
/* solver_p3p.h:177-385 Lambda-Twist; returns #models (0..4), 3x4 row-major each  [pinned:source] */
int gcro_solve_p3p(const double *pts5, const uint32_t *sample, double *P48);
/* non-minimal absolute pose (gcr_oracle_pba.cpp): normalised DLT + the LM of refine_pnp; weights by position or NULL */
int gcro_refit_pnp(const double *pts5, const uint32_t *sample, int n, const double *weights, double *P12);
int gcro_refine_pnp_lm(const double *pts5, const uint32_t *sample, int n, const double *weights, double *P12);

/* ---- neighbourhood (a24) ---------------------------------------------------------------------------------- */
/* grid_neighborhood_graph.h:229-292: keys[i] = cell key of point i; returns neighbor_number; *n_cells = filled
 * cells.  sorted_idx (n) = point indices grouped by ascending key, ascending index inside a cell;
 * cell_offsets (n+1 capacity) = start of each filled cell in sorted_idx, terminated by n.  [pinned:_ref]       */
uint64_t gcro_grid_build(const double *pts, int n, int dims, const double *cell_sizes, uint64_t cells_per_axis,
                         uint64_t *keys, uint32_t *sorted_idx, uint32_t *cell_offsets, uint32_t *n_cells);

/* ---- fixed-K round (a1 inner loop, a19) -------------------------------------------------------------------- */
/* GCRANSAC.h:297-471 without adaptive stop / LO; the CPU "port" arm bench.py times when oracle/_ref is absent */
void gcro_round(int problem, int scoring, const double *pts, int n, uint64_t seed, uint64_t first, int count,
                double thr, double *best_score, uint64_t *best_index, uint64_t *models_scored, double *best_model);

/* ---- termination (a1) ------------------------------------------------------------------------------------- */
/* GCRANSAC.h:158-172 */
uint64_t gcro_iteration_number(uint64_t inliers, uint64_t n, int m, double log_probability);

#ifdef __cplusplus
}
#endif
#endif
