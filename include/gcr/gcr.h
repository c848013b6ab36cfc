/*
 * include/gcr/gcr.h -- C ABI of the B200-native GC-RANSAC hot path (libgcr_b200.so).
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++/torch types, one int status per call
 * (0 = ok; gcr_last_error() explains anything else).  Each entry point names the reference function(s) of
 * danini/graph-cut-ransac it replaces; paths are relative to src/pygcransac/include/ of the reference.
 * INTEGRATION.md shows the adapter a reference maintainer adds on top (a ScoringFunction<> subclass and the
 * find*_ drivers), and graph-cut-ransac_b200/host/ holds our C++ mirror of that template surface.
 *
 * Conventions
 *  - "d_" pointers are device pointers on the context's GPU; "h_" pointers are host pointers.
 *  - All work is enqueued on the context's stream; results in device buffers are valid after gcr_sync() (or any
 *    stream-ordered consumer).  Functions with host outputs synchronise internally.
 *  - Points are row-major FP64, `cols` = 4 (x1 y1 x2 y2: H, F, E) or 5 (u v X Y Z: PnP), `n_pairs` independent
 *    problems of `n` points each stored back to back (n_pairs = 1 for a single problem).
 *  - Models are row-major FP64: 9 doubles (3x3) for H/F/E, 12 (3x4) for PnP.
 *  - There is no CPU fallback anywhere behind this ABI: without a CUDA device gcr_ctx_create() fails.
 */
#ifndef GCR_GCR_H_INCLUDED
#define GCR_GCR_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gcr_ctx gcr_ctx;
typedef int gcr_status;

enum { GCR_OK = 0, GCR_ERR_CUDA = 1, GCR_ERR_ARG = 2, GCR_ERR_STATE = 3, GCR_ERR_NOMEM = 4 };
enum { GCR_H = 0, GCR_F = 1, GCR_E = 2, GCR_PNP = 3 };          /* problem */
enum { GCR_MSAC = 0, GCR_MAGSAC = 1 };                          /* scoring function */
enum { GCR_UNIFORM = 0, GCR_PROSAC = 1, GCR_IMPORTANCE = 2 };   /* sampler */
enum { GCR_WEIGHT_NONE = 0, GCR_WEIGHT_POSITIONAL = 1, GCR_WEIGHT_BY_POINT = 2 }; /* refit weight indexing */
/* gcr_score flags */
enum { GCR_SCORE_EXACT_COUNTS = 1 /* FP64 re-check of points within the FP32 error band of the threshold */ };

/* ---- context ---------------------------------------------------------------------------------------------- */
gcr_status gcr_ctx_create(int device, gcr_ctx **out);
gcr_status gcr_ctx_destroy(gcr_ctx *ctx);
/* use an existing cudaStream_t (e.g. torch's current stream); NULL restores the context's own stream */
gcr_status gcr_ctx_set_stream(gcr_ctx *ctx, void *cuda_stream);
gcr_status gcr_sync(gcr_ctx *ctx);
const char *gcr_last_error(const gcr_ctx *ctx);
const char *gcr_version(void);
/* number of SMs / SM clock (kHz) of the context's device -- used by bench.py for the FP32 roofline */
gcr_status gcr_device_info(gcr_ctx *ctx, int *sm_count, int *sm_clock_khz, size_t *l2_bytes);
/* how many kernels this context has launched so far (bench.py's gpu_launches claim) */
uint64_t gcr_launch_count(const gcr_ctx *ctx);

/* ---- tuning / instrumentation (no reference counterpart) --------------------------------------------------- */
/* width of the FP32 error band (in units of FP32 epsilon, normalised coordinates) inside which
 * GCR_SCORE_EXACT_COUNTS re-evaluates a point in FP64; default 16 */
gcr_status gcr_set_band_factor(gcr_ctx *ctx, double factor);
/* force a scoring-kernel variant (hypotheses per thread): -1 auto, 0 R=4, 1 R=6, 2 R=8 (PnP: 0 R=2, 1 R=4) */
gcr_status gcr_set_score_variant(gcr_ctx *ctx, int variant);
/* generic tuning knobs (experiments / benches; every knob has a sane default) */
enum {
	GCR_TUNE_K3_CTAS_PER_SM = 0, /* cap on the resident CTAs per SM of the scoring kernel (0 = as many as fit) */
	GCR_TUNE_ROUND_PREFETCH = 1, /* gcr_round: solve the NEXT round's samples on a side stream while this round is scored (default 1) */
	GCR_TUNE_SMEM_CARVEOUT = 2   /* preferred shared-memory carve-out (percent, -1 = driver's choice) of the solver / scoring kernels:
	                                kernels with the SAME carve-out can share an SM (set before the first launch) */
};
gcr_status gcr_set_tuning(gcr_ctx *ctx, int knob, int value);
/* per-CTA timeline of the following K3 launches: d_trace (device, >= 4 * 148 * 8 uint64, or NULL to stop) receives per
 * CTA its SM id, first and last %globaltimer tick (ns) and the number of segments it ran (tools/k3_trace.py) */
gcr_status gcr_set_score_trace(gcr_ctx *ctx, uint64_t *d_trace);
/* CUDA-event timing of every K3 launch made through gcr_round / gcr_score (on the context's stream) */
gcr_status gcr_profile_enable(gcr_ctx *ctx, int on);
/* total K3 device time (ms), number of K3 launches and hypothesis x point evaluations they covered; resets */
gcr_status gcr_profile_read(gcr_ctx *ctx, double *score_ms, uint64_t *score_launches, uint64_t *score_evals);
/* device counters since the last reset: out[0] = valid models produced by gcr_solve_minimal, out[1] = FP64
 * re-check evaluations dropped because the work list was full (must stay 0), out[2] = FP64 re-checks queued,
 * out[3] = candidates gcr_round's selection re-scored in FP64, out[4] = the largest |FP32 score - FP64 score| / FP64 score
 * seen among them, as the bit pattern of a float (the premise of the selection margin: must stay below margin / 2),
 * out[5..7] reserved */
gcr_status gcr_counters_read(gcr_ctx *ctx, uint64_t out[8], int reset);
/* the context's device-block cache (gcr_malloc / gcr_free): blocks held, blocks handed out, bytes held */
gcr_status gcr_cache_stats(gcr_ctx *ctx, uint64_t *blocks, uint64_t *blocks_in_use, uint64_t *bytes);

/* Measures the FP32-FMA-pipe roofline of the context's device (the bound of the scoring kernels): runs a
 * kernel of dependent packed-FFMA chains back to back for at least min_ms.  mode 0: pure FFMA2 (the pipe's peak);
 * mode 1: FFMA2 with K3's side work (2 MUFU.RCP + compare / count / add per 19 FFMA2: the ceiling of K3's instruction
 * mix).  tflops_out[0] = sustained over the interval, tflops_out[1] = best single launch (2 flop per FMA lane). */
gcr_status gcr_ubench_fp32(gcr_ctx *ctx, int mode, double min_ms, double tflops_out[2], double *ms_out);

/* device memory helpers so that C / ctypes callers need no other CUDA binding */
gcr_status gcr_malloc(gcr_ctx *ctx, size_t bytes, void **d_out);
gcr_status gcr_free(gcr_ctx *ctx, void *d_ptr);
gcr_status gcr_memcpy_h2d(gcr_ctx *ctx, void *d_dst, const void *h_src, size_t bytes);
gcr_status gcr_memcpy_d2h(gcr_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
/* queues a device -> host copy on the context's stream; the data is in h_dst after the next gcr_sync or gcr_memcpy_d2h
 * (several read-backs of one step then cost one synchronise) */
gcr_status gcr_memcpy_d2h_async(gcr_ctx *ctx, void *h_dst, const void *d_src, size_t bytes);
gcr_status gcr_memset(gcr_ctx *ctx, void *d_dst, int value, size_t bytes);

/* ---- points ----------------------------------------------------------------------------------------------- */
/* Replaces the cv::Mat view the drivers build (py/gcransac_python.cpp:4274).  Copies host points to the device,
 * keeps the FP64 copy and derives the centred / power-of-two-scaled FP32 copies the scoring kernels stream. */
gcr_status gcr_points_upload(gcr_ctx *ctx, const double *h_points, int n, int cols, int n_pairs);
/* The same for a float32 host array -- what the reference's own tests hand to the Python entry points
 * (tests/test_homography.py:52-60; pybind's py::array_t<double> widens it on the host, py/bindings.cpp:859-893): half the
 * bytes cross PCIe and the exact widening to FP64 happens on the device. */
gcr_status gcr_points_upload_f32(gcr_ctx *ctx, const float *h_points, int n, int cols, int n_pairs);
/* Same, for points already resident in HBM (no copy; the caller keeps d_points alive). */
gcr_status gcr_points_set_device(gcr_ctx *ctx, const double *d_points, int n, int cols, int n_pairs);
/* per-pair normalisation the FP32 copy uses: out[8*pair + 0..4] = c1x c1y c2x c2y s (PnP: cu cv su cX cY cZ sX) */
gcr_status gcr_points_get_norm(gcr_ctx *ctx, double *h_out /* 8 * n_pairs */);

/* ---- K1 sampling ------------------------------------------------------------------------------------------ */
/* UniformSampler::sample (samplers/uniform_sampler.h:81-99) + generateUniqueRandomSet
 * (uniform_random_generator.h:90-103), ProsacSampler::sample (samplers/prosac_sampler.h:207-236), on a
 * counter-based Philox4x32-10 stream: sample i of pair p is a pure function of (seed, first_index + i, p).
 * d_out: [n_pairs][count][m] uint32.  For PROSAC first_index is the 1-based k of the first draw. */
gcr_status gcr_sample(gcr_ctx *ctx, int sampler, int m, uint64_t first_index, int count, uint64_t seed,
                      uint32_t *d_out);
/* ImportanceSampler (samplers/importance_sampler.h:67-82, the NG-RANSAC sampler, `sampler=3` of the Python entry
 * points): the inlier probabilities of the n uploaded points become the cumulative table of the reference's
 * std::discrete_distribution<int> (normalise by the sum, partial sums, last entry 1).  GCR_IMPORTANCE samples are then
 * m INDEPENDENT draws from that distribution (importance_sampler.h:134-135: with replacement, a sample may repeat a
 * point), draw j of sample i = two 32-bit words of stream 3 of (seed, first_index + i, pair) -> the 53-bit uniform of
 * std::generate_canonical -> lower bound in the table.  All pairs of the context share the table.
 * Call after gcr_points_upload; n must equal the number of points. */
gcr_status gcr_set_point_probabilities(gcr_ctx *ctx, const double *h_probabilities, int n);
/* Uniform draws from explicit pools (GCRANSAC.h:842-846, the LO inner sampler): pool p has pool_sizes[p] entries
 * at d_pool + pool_offsets[p]; out[i] = pool[draw_i].  stream distinguishes independent uses of one seed. */
gcr_status gcr_sample_from_pool(gcr_ctx *ctx, const uint32_t *d_pool, uint32_t pool_size, int m,
                                uint64_t first_index, int count, uint64_t seed, uint32_t stream, uint32_t *d_out);

/* ---- K2 minimal solvers (+ fused sample / model validity tests) --------------------------------------------- */
/* H: RobustHomographyEstimator::isValidSample (estimators/homography_estimator.h:346-381) +
 *    HomographyFourPointSolver::estimateMinimalModel (estimators/solver_homography_four_point.h:109-190).
 * F: FundamentalMatrixSevenPointSolver::estimateModel (estimators/solver_fundamental_matrix_seven_point.h:91-291)
 *    + oriented epipolar constraint (estimators/fundamental_estimator.h:161-184, 737-800).
 * E: EssentialMatrixFivePointNisterSolver::estimateModel (estimators/solver_essential_matrix_five_point_nister.h
 *    :263-615) + orientation (estimators/essential_estimator.h:152-175).
 * PnP: P3PSolver::estimateModel (estimators/solver_p3p.h:177-385).
 * d_samples [n_pairs][count][m]; d_models [n_pairs][count][max_models][9|12] FP64; d_valid same shape, 1 = model
 * present and passed the tests.  max_models = 1 (H), 3 (F), 10 (E), 4 (PnP).  The contents of a model slot whose flag
 * is 0 are unspecified (the kernels do not spend stores on slots nobody reads: 89 % of the 7-point slots). */
gcr_status gcr_solve_minimal(gcr_ctx *ctx, int problem, const uint32_t *d_samples, int count, double *d_models,
                             uint8_t *d_valid);
/* FundamentalMatrixPlaneParallaxSolver::estimateModel (estimators/solver_fundamental_matrix_plane_and_parallax.h
 * :109-160) + the oriented epipolar constraint (estimators/fundamental_estimator.h:161-184): the minimal solver of
 * DEGENSAC's inner GC-RANSAC (fundamental_estimator.h:538-543).  d_samples [n_pairs][count][2]; h_homography: the
 * plane homography, 9 doubles row-major, HOST memory; d_models [n_pairs][count][9]; d_valid [n_pairs][count]. */
gcr_status gcr_solve_plane_parallax(gcr_ctx *ctx, const uint32_t *d_samples, int count, const double *h_homography,
                                    double *d_models, uint8_t *d_valid);
int gcr_max_models(int problem);
int gcr_sample_size(int problem);
int gcr_model_len(int problem);

/* ---- K3 scoring ------------------------------------------------------------------------------------------- */
/* MSACScoringFunction::getScore (scoring_function.h:137-233) / MAGSACScoringFunction::getScore (:246-348) with
 * the estimator's squaredResidual (homography_estimator.h:181-199, fundamental_estimator.h:195-222,
 * essential_estimator.h:186-213, perspective_n_point_estimator.h:148-184) for k models per pair at once.
 * threshold is the user threshold (MSAC truncates at (1.5*threshold)^2 as GCRANSAC.h:253-255 does).
 * d_valid may be NULL; models with valid == 0 get score -1 and 0 inliers.
 * d_scores [n_pairs][k] float, d_inliers [n_pairs][k] uint32 (may be NULL). */
gcr_status gcr_score(gcr_ctx *ctx, int problem, int scoring, const double *d_models, const uint8_t *d_valid,
                     int k, double threshold, int flags, float *d_scores, uint32_t *d_inliers);
/* GCR_SCORE_EXACT_COUNTS queues its FP64 re-checks in a bounded work list (4 Mi entries); what did not fit in the LAST
 * gcr_score call of the context is reported here (a queued read-back: valid after the next gcr_sync).  Non-zero = that
 * call's counts are not exact -- score fewer models per call. */
gcr_status gcr_score_dropped_async(gcr_ctx *ctx, uint32_t *h_dropped);
/* The same arithmetic in FP64, reference operation order, for a few models of pair `pair`: exact score, exact
 * inlier count and (optionally) the ascending inlier index list -- the side output of getScore
 * (scoring_function.h:166-167).  d_inlier_idx: [k][n] uint32 or NULL; d_residuals: [k][n] FP64 or NULL. */
gcr_status gcr_score_exact(gcr_ctx *ctx, int problem, int scoring, int pair, const double *d_models, int k,
                           double threshold, double *d_scores, uint32_t *d_counts, uint32_t *d_inlier_idx,
                           double *d_residuals);
/* the same for a batch of pairs: one model per pair (d_models [n_pairs][9|12]), every pair in one launch; outputs
 * [n_pairs], d_inlier_idx [n_pairs][n] or NULL.  With gcr_refit_pairs: the "argmax + one final least-squares refit" tail
 * of the batched configuration (SURVEY.md 8d cfg 4; GCRANSAC.h:588-617 per pair) */
gcr_status gcr_score_exact_pairs(gcr_ctx *ctx, int problem, int scoring, const double *d_models, double threshold,
                                 double *d_scores, uint32_t *d_counts, uint32_t *d_inlier_idx);
/* host-pointer convenience wrappers (H2D of models, D2H of results inside) */
gcr_status gcr_score_host(gcr_ctx *ctx, int problem, int scoring, const double *h_models, const uint8_t *h_valid,
                          int k, double threshold, int flags, float *h_scores, uint32_t *h_inliers);
gcr_status gcr_score_exact_host(gcr_ctx *ctx, int problem, int scoring, int pair, const double *h_models, int k,
                                double threshold, double *h_scores, uint32_t *h_counts, uint32_t *h_inlier_idx,
                                double *h_residuals);

/* ScoringFunction<Estimator>::getScore (scoring_function.h:88-96) for ONE model, host in / out: the call the drop-in
 * adapter gcr::CudaScoringFunction (graph-cut-ransac_b200/host/gcr_scoring_function.h) makes from the reference's own
 * GCRANSAC<> loop.  squared_truncated_threshold = the value initialize() received (used as is by MSAC, :163);
 * threshold = getScore's threshold_ (MAGSAC's sigma_max, :257).  h_inlier_idx: [n] uint32 or NULL. */
gcr_status gcr_getscore_host(gcr_ctx *ctx, int problem, int scoring, int pair, const double *h_model,
                             double squared_truncated_threshold, double threshold, double *h_score, uint32_t *h_count,
                             uint32_t *h_inlier_idx);

/* ---- K4 selection ----------------------------------------------------------------------------------------- */
/* Running-argmax of GCRANSAC.h:440-471 with Score::operator< (scoring_function.h:63-66): best score, ties to the
 * LOWEST index.  d_keys [n_pairs] uint64 = (orderedFloatBits(score) << 32) | (0xFFFFFFFF - (index + index_base)):
 * max over keys == best model, which is also what the hypothesis-sharded mode all-reduces (ncclMax). */
gcr_status gcr_argmax(gcr_ctx *ctx, const float *d_scores, int k, uint32_t index_base, uint64_t *d_keys);
/* decode helpers (pure functions, usable on the host) */
float gcr_key_score(uint64_t key);
uint32_t gcr_key_index(uint64_t key);

/* ---- K6 non-minimal refits -------------------------------------------------------------------------------- */
/* H: RobustHomographyEstimator::estimateModelNonminimal (homography_estimator.h:136-173) +
 *    HomographyFourPointSolver::estimateNonMinimalModel (solver_homography_four_point.h:192-264).
 * F: FundamentalMatrixEstimator::estimateModelNonminimal (fundamental_estimator.h:574-618) +
 *    FundamentalMatrixEightPointSolver (solver_fundamental_matrix_eight_point.h:93-182).
 * E: EssentialMatrixEstimator::estimateModelNonminimal (essential_estimator.h:306-336) +
 *    EssentialMatrixFivePointSteweniusSolver::estimateModel (solver_essential_matrix_five_point_stewenius.h:114-231):
 *    up to 10 models per set; weights are read BY POINT there (:143).
 * `count` index sets of pair `pair`; set j has set_sizes[j] indices at d_index_sets + set_offsets[j].
 * d_weights: per-point weights or NULL; weight_mode: GCR_WEIGHT_POSITIONAL reproduces the reference's positional
 * read, GCR_WEIGHT_BY_POINT reads weights[index].
 * d_models [count][R][9], d_ok [count][R] with R = gcr_max_refit_models(problem): 1 (H, F), 10 (E).
 * One CTA per set: 128 threads for a call with one set (all inliers of a run), one warp per set for a call with many (the
 * samples of a local-optimisation round, <= 7 m points each; larger sets are fitted correctly, just by 32 threads). */
gcr_status gcr_refit(gcr_ctx *ctx, int problem, int pair, const uint32_t *d_index_sets, const uint32_t *d_set_offsets,
                     const uint32_t *d_set_sizes, int count, const double *d_weights, int weight_mode,
                     double *d_models, uint8_t *d_ok);
/* The reference's DEFAULT non-minimal F fit (types.h:78-80): FundamentalMatrixBundleAdjustmentSolver
 * (estimators/solver_fundamental_matrix_bundle_adjustment.h:113-177) -- the 8-point model refined by the LM bundle
 * adjustment of relative_pose/bundle.cpp:253-339 (F = U diag(1, sigma, 0) V^T, Sampson residual, truncated loss,
 * <= 25 iterations, relative_pose/jacobian_impl.h:472-606).  frame = GCR_BA_HARTLEY: as called through
 * FundamentalMatrixEstimator::estimateModelNonminimal (fundamental_estimator.h:574-618: both steps on Hartley-normalised
 * points, result de-normalised); GCR_BA_PIXEL: the solver called on pixel coordinates, the post-run polish of
 * findFundamentalMatrix_ (py/gcransac_python.cpp:1851-1858; result not rescaled).  Sets with fewer than 8 points: ok = 0. */
#define GCR_BA_HARTLEY 1
#define GCR_BA_PIXEL 2
gcr_status gcr_refit_f_ba(gcr_ctx *ctx, int pair, const uint32_t *d_index_sets, const uint32_t *d_set_offsets,
                          const uint32_t *d_set_sizes, int count, const double *d_weights, int weight_mode, int frame,
                          double *d_models, uint8_t *d_ok);
/* EssentialMatrixBundleAdjustmentSolver::estimateModel (estimators/solver_essential_matrix_bundle_adjustment.h:113-186), the
 * post-run polish of findEssentialMatrix_ (py/gcransac_python.cpp:2766-2821), on ONE index set of pair `pair`: the
 * Stewenius non-minimal fit (up to 10 models) -> motion_from_essential + cheirality on the set's first point
 * (relative_pose/essential.cpp:44-117: up to 4 poses each) -> the 5-dof LM of refine_relpose (relative_pose/bundle.cpp
 * :176-250, Sampson error, TruncatedLoss(1.0), <= 25 iterations) -> E = [t]_x R.  d_weights: one weight per set ENTRY (by
 * position: what the reference's driver fills and its LM accumulator reads; its Stewenius initialiser indexes the same
 * vector by point, past its end -- here position everywhere) or NULL.  d_models [40][9], d_ok [40] in the order the
 * reference emits them (initial model major, then pose); sets of fewer than 6 points: every ok = 0. */
gcr_status gcr_refit_e_ba(gcr_ctx *ctx, int pair, const uint32_t *d_index_set, uint32_t set_size, const double *d_weights,
                          double *d_models, uint8_t *d_ok);
int gcr_max_e_ba_models(void); /* 40 */
/* one non-minimal fit per pair, all pairs in one launch: set p = the first d_sizes[p] entries of d_index_sets[p][n],
 * fitted to pair p's points (H: homography_estimator.h:136-173; F: 8-point, fundamental_estimator.h:574-618) */
gcr_status gcr_refit_pairs(gcr_ctx *ctx, int problem, const uint32_t *d_index_sets, const uint32_t *d_sizes,
                           double *d_models, uint8_t *d_ok);
int gcr_max_refit_models(int problem);

/* ---- K7 grid neighbourhood -------------------------------------------------------------------------------- */
/* GridNeighborhoodGraph<4>::initialize (neighborhood/grid_neighborhood_graph.h:229-292), pair `pair`.
 * d_keys [n] uint64 cell key per point (GridCell index, :70-82); d_sorted_idx [n] point indices grouped by
 * ascending key, ascending index inside a cell; d_cell_offsets [n+1] starts of the filled cells in d_sorted_idx
 * (terminated by n); h_n_cells = filledCellNumber(); h_neighbor_number = sum n(n-1)/2 (:283-284). */
gcr_status gcr_grid_build(gcr_ctx *ctx, int pair, const double h_cell_sizes[4], uint64_t cells_per_axis,
                          uint64_t *d_keys, uint32_t *d_sorted_idx, uint32_t *d_cell_offsets, uint32_t *h_n_cells,
                          uint64_t *h_neighbor_number);
/* FlannNeighborhoodGraph::initialize (neighborhood/flann_neighborhood_graph.h:99-137) -- the neighbourhood the Python
 * entry points use for neighborhood=1: every point's neighbours within `radius` (L2 over the four FP32 coordinates,
 * strict <, the point itself excluded).  The reference's answer comes from OpenCV's approximate FLANN search; this is
 * the exact graph.  Two passes around a host prefix sum: d_counts [n] neighbours per point; then, with d_offsets [n]
 * the exclusive prefix of the counts, d_neighbors receives every point's neighbours in ascending index order. */
gcr_status gcr_radius_graph_count(gcr_ctx *ctx, int pair, double radius, uint32_t *d_counts);
gcr_status gcr_radius_graph_fill(gcr_ctx *ctx, int pair, double radius, const uint32_t *d_offsets, uint32_t *d_neighbors);

/* The pairwise structure of GCRANSAC::labeling (GCRANSAC.h:963-1008) on the device: every unordered pair of points that
 * share a grid cell (gcr_grid_edges, from gcr_grid_build's outputs) or are radius neighbours (gcr_csr_edges, from the
 * CSR of gcr_radius_graph_*: d_offsets [n + 1]), once, as (i, j > i) in the order the reference visits them -- i
 * ascending, j ascending -- without its N x N `used_edges` matrix.  d_edge_offsets [n + 1] receives the first edge of
 * every point (its last entry = the edge count, which equals gcr_grid_build's neighbor_number / half the CSR length);
 * d_edges [capacity][2] uint32, capacity >= that count (capacity 0: offsets only).  No host synchronisation. */
gcr_status gcr_grid_edges(gcr_ctx *ctx, const uint32_t *d_sorted_idx, const uint32_t *d_cell_offsets, uint32_t n_cells,
                          uint32_t *d_edge_offsets, uint32_t *d_edges, uint64_t capacity);
gcr_status gcr_csr_edges(gcr_ctx *ctx, const uint32_t *d_offsets, const uint32_t *d_neighbors, uint32_t *d_edge_offsets,
                         uint32_t *d_edges, uint64_t capacity);
/* The energy terms of labeling() for one model of pair `pair` (d_model: device, 9 | 12 doubles): d_unary [n][2] = the
 * (source, sink) capacities of add_term1 (GCRANSAC.h:933-957, truncation at (1.5 threshold)^2), d_pairwise [n_edges] =
 * e00 * lambda of add_term2 for every edge of d_edges (:996-1003; the other three terms are lambda, lambda, 0), or NULL;
 * d_dist [n] = clamp(r^2 / T, 0, 1) per point, or NULL (a caller that feeds a host max-flow can form e00 = 0.5 (dist_i +
 * dist_j) while it adds the edge instead of reading 8 bytes per edge back).  The min-cut itself stays with the
 * reference's Boykov-Kolmogorov solver on the host. */
gcr_status gcr_labeling_terms(gcr_ctx *ctx, int problem, int pair, const double *d_model, double threshold, double lambda,
                              const uint32_t *d_edges, uint64_t n_edges, double *d_unary, double *d_pairwise, double *d_dist);

/* ---- many pairs, one launch: the operations of the batched multi-pair driver (host/gcr_batch_engine.cpp) ------------ */
/* The lock-step driver keeps P image pairs resident on one context and executes the pending operation of every pair as
 * one launch; these are the per-pair entry points above with a pair index PER ITEM instead of one per call.  Results are
 * bit-identical to the single-pair calls (same kernels, same CTA shapes). */
/* gcr_sample / gcr_solve_minimal / gcr_score work on every pair of the context: restrict the next sampling / solving
 * calls to the pairs with d_active[pair] != 0 (their model slots' valid flags are 0 for the others, so gcr_score spends
 * nothing on them).  NULL = all pairs.  The array is read at launch time on the context's stream. */
gcr_status gcr_set_active_pairs(gcr_ctx *ctx, const uint8_t *d_active);
/* gcr_sample on a context whose pairs advance independently (the seats of the batch engine): pair p draws sample i at
 * index first_index + i + d_first_offsets[p] of the stream of seed + d_seed_offsets[p] (+ the usual p * 0x9E3779B97F4A7C15;
 * 64-bit wrapping sums).  Either array may be NULL (= zeros); read at launch time on the context's stream. */
gcr_status gcr_set_pair_streams(gcr_ctx *ctx, const uint64_t *d_first_offsets, const uint64_t *d_seed_offsets);
/* Replaces the points of `count` pairs of an uploaded multi-pair context in place: pair h_pairs[r] <- row r of h_points
 * ([count][n][cols] FP64, host), its normalisation frame and FP32 copy re-derived exactly as gcr_points_upload derives
 * them (n <= 32768).  Queued on the context's stream; h_points / h_pairs may be reused when the call returns. */
gcr_status gcr_points_replace_pairs(gcr_ctx *ctx, const int32_t *h_pairs, int count, const double *h_points);
/* gcr_score_exact for k models of different pairs: model m against pair d_pair_ids[m] */
gcr_status gcr_score_exact_multi(gcr_ctx *ctx, int problem, int scoring, const int32_t *d_pair_ids, const double *d_models, int k,
                                 double threshold, double *d_scores, uint32_t *d_counts, uint32_t *d_inlier_idx, double *d_residuals);
/* gcr_labeling_terms for k models of different pairs; d_out [k][3 n]: per model the [n][2] capacities, then [n] distances */
gcr_status gcr_labeling_terms_multi(gcr_ctx *ctx, int problem, const int32_t *d_pair_ids, const double *d_models, int k,
                                    double threshold, double lambda, double *d_out);
/* gcr_sample_from_pool for several requests at once: request r draws `trials` samples of sample_size positions from its
 * pool (d_pools + pool_offset) on its own stream of draws and writes them to d_out + out_offset + t * sample_size */
typedef struct gcr_pool_request
{
	uint32_t pool_offset, pool_size, sample_size, out_offset;
	uint64_t first_index, seed;
	uint32_t stream, reserved;
} gcr_pool_request;
gcr_status gcr_sample_from_pools(gcr_ctx *ctx, const uint32_t *d_pools, const gcr_pool_request *d_requests, int n_requests,
                                 int trials, uint32_t *d_out);
/* gcr_refit / gcr_refit_f_ba for `count` index sets of different pairs: set s belongs to pair d_pair_of_set[s], reads its
 * weights (if any) from d_weights + d_weight_offsets[s] and writes its models to slot d_out_slots[s] (NULL: slot s).
 * wide != 0: every set is a whole inlier set fitted with the CTA shape of a one-set gcr_refit call; 0: one warp per set,
 * the shape of a many-set call (the block sums associate by CTA shape: a batched fit equals the unbatched one bit for bit) */
gcr_status gcr_refit_multi(gcr_ctx *ctx, int problem, int ba_frame, const int32_t *d_pair_of_set, const uint32_t *d_index_sets,
                           const uint32_t *d_set_offsets, const uint32_t *d_set_sizes, int count, const double *d_weights,
                           const uint32_t *d_weight_offsets, int weight_mode, const uint32_t *d_out_slots, int wide,
                           double *d_models, uint8_t *d_ok);
/* pinned host memory + direct asynchronous copies for the engine's large transfers (no bounce buffer), and a row gather */
gcr_status gcr_host_alloc_pinned(gcr_ctx *ctx, size_t bytes, void **h_out);
gcr_status gcr_host_free_pinned(gcr_ctx *ctx, void *h_ptr);
gcr_status gcr_memcpy_pinned_async(gcr_ctx *ctx, void *dst, const void *src, size_t bytes, int to_device);
gcr_status gcr_gather_rows(gcr_ctx *ctx, const void *d_src, uint32_t row_bytes, const uint64_t *d_rows, uint32_t count, void *d_dst);

/* ---- fused fixed-K round (the benchmarked step) ------------------------------------------------------------ */
/* One pass of the hot path over every pair on the context: K1 sample `count` minimal samples -> K2 solve -> K3
 * score (FP32, valid models only) -> K4: the near-best candidates of the FP32 ranking are re-scored in FP64 and the
 * winner is picked by (FP64 score, lowest index), i.e. GCRANSAC.h:440's strict '<' in the reference's precision.
 * Outputs per pair: packed key (see gcr_argmax; score = the winner's FP64 score rounded to float, index = sample *
 * max_models + model slot, offset by index_base; 0 if no model scored > 0) and the winning model (FP64).
 * evals_out = model slots x points of this round (valid models are counted by gcr_counters_read). */
gcr_status gcr_round(gcr_ctx *ctx, int problem, int scoring, int sampler, uint64_t first_index, int count,
                     uint64_t seed, double threshold, int flags, uint32_t index_base, uint64_t *d_keys,
                     double *d_best_models, uint64_t *evals_out);

/* FP64 score and exact inlier count of each pair's winner of the last gcr_round (host outputs, [n_pairs]). */
gcr_status gcr_round_results(gcr_ctx *ctx, double *h_best_scores, uint32_t *h_best_counts);

/* The winner of a round as ONE 128-byte device record per pair: what the chunks of a large problem accumulate and what
 * the hypothesis-sharded mode exchanges between GPUs.  Unlike the packed key of gcr_round it carries the score in FP64
 * and a 64-bit index, so GCRANSAC.h:440's strict '<' (scoring_function.h:63-66) is decided in the reference's precision
 * across chunks and ranks too, and 2^20 samples x 3 models cannot wrap. */
typedef struct gcr_winner
{
	double score;      /* FP64 score (reference arithmetic); 0 when valid == 0 */
	uint64_t index;    /* global model-slot index: index_base + sample * max_models + model slot */
	uint32_t count;    /* exact inlier count */
	uint32_t valid;    /* 1 = some model scored > 0 */
	double model[12];  /* row-major, 9 (H / F / E) or 12 (PnP) entries, the rest 0 */
	uint64_t reserved; /* 0 */
} gcr_winner;
/* gcr_round writing winner records: d_winners [n_pairs].  accumulate != 0: the record is only replaced by a strictly
 * better winner -- (FP64 score, then lower index) -- so a problem walked in chunks keeps its running best on the device. */
gcr_status gcr_round_winners(gcr_ctx *ctx, int problem, int scoring, int sampler, uint64_t first_index, int count,
                             uint64_t seed, double threshold, int flags, uint64_t index_base, int accumulate,
                             gcr_winner *d_winners, uint64_t *evals_out);

/* ---- hypothesis-sharded mode: one process (or thread) per GPU, NCCL over NVLink ------------------------------ */
/* 128-byte NCCL unique id (ncclGetUniqueId): rank 0 creates it, the caller hands it to every rank (MPI, torch.distributed,
 * a socket: host plumbing).  NCCL is loaded with dlopen("libnccl.so.2") -- inside a torch process that is torch's copy. */
gcr_status gcr_comm_unique_id(void *h_out_128);
/* joins the communicator of `world` contexts (blocking collective call: every rank must make it) */
gcr_status gcr_comm_init(gcr_ctx *ctx, const void *h_unique_id_128, int rank, int world);
gcr_status gcr_comm_destroy(gcr_ctx *ctx);
/* One fixed-K pass with the samples [0, total_samples) of the counter-based stream sharded over the ranks of the
 * context's communicator (rank g owns a contiguous, balanced range; sample i is a pure function of
 * (seed, first_index + i), so no rank needs another's samples).  Each rank walks its range in chunks of chunk_samples
 * (bounds the workspace), accumulating its winner record on the device; the records are then all-gathered (ncclAllGather,
 * 128 bytes per pair per rank) and every rank reduces them by the same rule -- larger FP64 score, then lower index -- so
 * all ranks end with the identical record, model included, with no host synchronisation and no second collective.
 * Without a communicator (or world = 1) it is the chunked single-GPU pass.  d_winners [n_pairs] device. */
gcr_status gcr_round_sharded(gcr_ctx *ctx, int problem, int scoring, int sampler, uint64_t first_index,
                             uint64_t total_samples, uint64_t seed, double threshold, int flags, int chunk_samples,
                             gcr_winner *d_winners, uint64_t *evals_out);
/* the reduction rule on its own: d_all [world][n_pairs] records -> d_out [n_pairs] */
gcr_status gcr_winners_reduce(gcr_ctx *ctx, const gcr_winner *d_all, int world, gcr_winner *d_out);

#ifdef __cplusplus
}
#endif
#endif
