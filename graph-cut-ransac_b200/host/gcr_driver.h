// gcr_driver.h -- batched host driver: GCRANSAC<>::run (GCRANSAC.h:202-627) with the data-parallel parts on the B200.
//
// The reference's loop is sequential and adaptive (best-so-far, iteration bound, local optimisation all change as it
// goes).  This driver returns what that loop returns from the same sample stream, but runs the loop body in batches
// (SURVEY.md 7.3): B minimal samples are drawn, solved and scored on the GPU at once (gcr_sample / gcr_solve_minimal /
// gcr_score); the host then replays the reference's sequential bookkeeping over the batch -- unsuccessful-sample
// accounting (:297-339), strict '<' best-so-far with isValidModel (:437-471), LO trigger, getIterationNumber
// (:158-172) and the loop condition (:282-283).  Candidates that may beat the best so far are confirmed in FP64
// (gcr_score_exact: exact score, exact ascending inlier list), so every decision is taken on the reference's numbers.
// Local optimisation (:780-910): residuals for the unary terms come from the GPU, the graph is built exactly as
// GCRANSAC::labeling does (:913-1021) from the GPU-built grid cells (gcr_grid_build) and cut by the REFERENCE's own
// BK max-flow (energy.h / graph.h, compiled from the reference tree -- the min-cut stays in the reference's host code);
// the inner RANSAC (<= lo_number x {pool sample, non-minimal refit, full scoring}) is one GPU batch
// (gcr_sample_from_pool / gcr_refit / gcr_score).  IRLS (:630-758) and the final fit use gcr_refit / gcr_score_exact.
//
// C ABI (libgcr_driver.so): see gcr_find_model below.  Python: graph-cut-ransac_b200/pygcransac/__init__.py.
#ifndef GCR_DRIVER_H_INCLUDED
#define GCR_DRIVER_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#include "gcr/gcr.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct gcr_run_settings
{
	/* utils::Settings (settings.h:43-86) as the drivers set it (py/gcransac_python.cpp:4498-4504) */
	double threshold;
	double confidence;
	double spatial_coherence_weight;
	uint64_t max_iteration_number;
	uint64_t min_iteration_number;
	uint64_t max_local_optimization_number; /* lo_number */
	uint64_t min_iteration_number_before_lo; /* 20 */
	uint64_t max_graph_cut_number;           /* 10 */
	uint64_t max_least_squares_iterations;   /* 10 */
	uint64_t max_unsuccessful_model_generations; /* 100 */
	int do_local_optimization;
	int do_final_iterated_least_squares;
	/* grid neighbourhood (GridNeighborhoodGraph<4>) */
	double cell_sizes[4];
	uint64_t cells_per_axis;
	/* driver */
	uint64_t seed;  /* counter-based sample streams: main = stream 0, local optimisation = stream 1 */
	int batch;      /* minimal samples per GPU batch (default 1024) */
	/* FundamentalMatrixEstimator(minimum_inlier_ratio_in_validity_check = 0.5, use_degensac = true,
	 * homography_threshold = 2.0) (fundamental_estimator.h:91-93); F only */
	int use_degensac;
	/* F: bit 0 = FundamentalMatrixBundleAdjustmentSolver as the non-minimal solver (the reference's default estimator,
	 * types.h:78-80; gcr_refit_f_ba in the Hartley frame), bit 1 = the post-run polish of findFundamentalMatrix_
	 * (py/gcransac_python.cpp:1832-1887; pixel frame) */
	int f_bundle_adjustment;
	/* main sampler: 0 = uniform (samplers/uniform_sampler.h), 1 = PROSAC (samplers/prosac_sampler.h: the correspondences
	 * are ordered by quality), 2 = Progressive NAPSAC (image_sizes below; host recurrence like 4), 3 = the importance sampler of NG-RANSAC (samplers/importance_sampler.h; needs
	 * point_probabilities below), 4 = the AR-Sampler (samplers/adaptive_reordering_sampler.h; point_probabilities and
	 * sampler_variance below; its heap recurrence runs on the host, gcr_host_samplers.h); the local-optimisation sampler is uniform either way (py/gcransac_python.cpp:4316-4330) */
	int sampler;
	/* neighbourhood of the graph-cut: 0 = GridNeighborhoodGraph<4> (cell_sizes / cells_per_axis above), 1 = the radius
	 * neighbourhood of FlannNeighborhoodGraph (neighborhood/flann_neighborhood_graph.h:99-137; exact instead of FLANN's
	 * approximate search), points within neighborhood_radius in the 4-D correspondence space */
	int neighborhood;
	double neighborhood_radius;
	/* sampler 3: the inlier probability of each of the n correspondences (host memory, read during the call) */
	const double *point_probabilities;
	/* sampler 4: estimator_variance_ of AdaptiveReorderingSampler (`sampler_variance` of the Python calls, default 0.1) */
	double sampler_variance;
	/* sampler 2 (Progressive NAPSAC, samplers/progressive_napsac_sampler.h): w1, h1, w2, h2 -- its grid layers have
	 * 16 / 8 / 4 / 2 cells per axis of these sizes (py/gcransac_python.cpp:4335-4343) */
	double image_sizes[4];
	/* the scoring function GCRANSAC<> is instantiated with: GCR_MSAC (MSACScoringFunction, the default template argument
	 * and what every Python entry point uses) or GCR_MAGSAC (MAGSACScoringFunction, scoring_function.h:237-349) */
	int scoring;
	/* E: the post-run relative-pose bundle adjustment of findEssentialMatrix_ (py/gcransac_python.cpp:2766-2821;
	 * gcr_refit_e_ba) -- 0 = off (GCRANSAC<>::run alone), 1 = on (what the Python entry point does) */
	int e_bundle_adjustment;
	/* PnP: find6DPose_'s post-run step (py/gcransac_python.cpp:1227-1238): the non-minimal fit on the inliers replaces the
	 * model -- 0 = off, 1 = on (what the Python entry point does) */
	int pnp_bundle_adjustment;
} gcr_run_settings;

typedef struct gcr_run_stats
{
	uint64_t iteration_number, graph_cut_number, local_optimization_number, main_samples, lo_samples;
	uint64_t gpu_batches, exact_rescores, models_scored;
	double score;
	double graph_cut_seconds; /* host min-cut, timed separately (north star) */
	uint64_t degensac_runs, degensac_updates; /* inner plane-and-parallax runs / models they replaced */
	/* host wall clock per phase (s): [0] upload + grid build, [1] GPU batches of the main loop (sample, solve, score,
	 * read-back), [2] the rest of the main loop (FP64 re-scores, isValidModel / DEGENSAC), [3] local optimisation
	 * (labeling, min-cut, inner batch), [4] iterated least squares, [5] final fit + post-run polish */
	double phase_seconds[6];
} gcr_run_stats;

/* defaults of utils::Settings + the Python drivers */
void gcr_run_settings_default(gcr_run_settings *s);

/* GCRANSAC<>::run for problem GCR_H, GCR_F, GCR_E (points already K-normalised) on `n` correspondences (row-major FP64
 * n x 4, host memory), or GCR_PNP (find6DPose_, py/gcransac_python.cpp:1026-1272: rows u v X Y Z, n x 5, image points
 * K-normalised; P3P minimal solver, non-minimal fits = normalised DLT + the LM of refine_pnp, radius neighbourhood).
 * model_out: 9 doubles (PnP: 12, the 3 x 4 pose); inlier_mask: n bytes (1 = inlier).  Returns the inlier count (0 = no model), < 0 on error
 * (gcr_last_error(ctx)). */
int gcr_find_model(gcr_ctx *ctx, int problem, const double *h_points, int n, const gcr_run_settings *settings,
                   double *model_out, uint8_t *inlier_mask, gcr_run_stats *stats);

/* findEssentialMatrix_ (py/gcransac_python.cpp:2519-2855): the neighbourhood grid is built on the pixel coordinates, the
 * run uses K^-1-normalised points and settings->threshold / mean focal length; with settings->e_bundle_adjustment its
 * final relative-pose bundle adjustment on the inliers (:2766-2821).  K1, K2: row-major 3 x 3 intrinsics.  model_out: E. */
int gcr_find_essential(gcr_ctx *ctx, const double *h_points, int n, const double *K1, const double *K2,
                       const gcr_run_settings *settings, double *model_out, uint8_t *inlier_mask, gcr_run_stats *stats);

/* samples [first, first + count) of the AR-Sampler sequence the driver uploads for sampler 4 (host only) */
int gcr_host_sample_adaptive_reordering(const double *probabilities, int n, int m, double variance, uint64_t seed,
                                        uint64_t first, int count, long long repeat_at, uint32_t *out);

/* samples [first, first + count) of the Progressive NAPSAC sequence the driver uploads for sampler 2 (host only) */
int gcr_host_sample_progressive_napsac(const double *points, int n, int m, const double *image_sizes, uint64_t seed,
                                       uint64_t first, int count, uint32_t *out);

#ifdef __cplusplus
}
#endif
#endif
