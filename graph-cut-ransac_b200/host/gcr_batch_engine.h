// gcr_batch_engine.h -- the lock-step multi-pair driver (findHomographyBatch & co.).
//
// P image pairs sit on the GPU together ("seats" of a multi-pair context).  Each pair in flight runs the SAME sequential
// driver as gcr_find_model (gcr_driver.cpp: the replay of GCRANSAC::run, GCRANSAC.h:202-627) as a FIBER (a user-level
// context on a stack of its own) over a BatchedBackend whose operations are requests: the fiber posts "main batch [first,
// first + B)", "FP64 getScore of these models", "labeling terms of this model", "LO round: pool -> samples -> refits ->
// scores", "refit of this inlier set" and switches out.  A dispatcher thread per engine takes whatever is pending, groups
// it by kind and queues every group as ONE launch sequence for all the pairs that asked (the per-item pair index entry
// points of include/gcr/gcr.h: gcr_set_active_pairs / gcr_set_pair_streams, gcr_score_exact_multi,
// gcr_labeling_terms_multi, gcr_sample_from_pools, gcr_refit_multi, gcr_points_replace_pairs), synchronises ONCE per
// super-step and makes the fibers runnable again.  About one OS thread per host core runs whichever fibers are runnable:
// the host-only work between two requests -- the reference's Boykov-Kolmogorov min-cut above all -- is fanned out over the
// cores without more runnable threads than cores and without a kernel wake-up per request.
//
// Results per pair equal gcr_find_model's with the pair's seed (tests/test_gpu_batch.py).  Measured: ahead
// of the one-pair-per-thread path where pairs are small and driver calls were the cost (configs[0]: 5.9 k against 4.6 k
// pairs/s), level on F, behind where the host cores (the min-cut, ~3 ms of one core per H pair) or one pair's own GPU
// kernels (E) are the bound.
#ifndef GCR_BATCH_ENGINE_H_INCLUDED
#define GCR_BATCH_ENGINE_H_INCLUDED

#include "gcr_driver.h"

#ifdef __cplusplus
extern "C" {
#endif

/* GCRANSAC<>::run for n_pairs independent problems of n correspondences each (h_points [n_pairs][n][4], host) on GPU
 * `device`; problem GCR_H, GCR_F or GCR_E (points already K-normalised; see gcr_find_essential_batch).  One settings
 * structure for all pairs; pair p draws from the streams of seeds[p], or of settings->seed + p * 0x9E3779B97F4A7C15 when
 * seeds is NULL, and returns what gcr_find_model returns for that pair with that seed.  seats: pairs resident on the GPU
 * at a time, over all engines (0 = $GCR_BATCH_SEATS or 512); engines: contexts + dispatcher threads (0 = $GCR_BATCH_ENGINES
 * or 2).  models_out [n_pairs][9], inlier_masks [n_pairs][n], inlier_counts [n_pairs] (0 = no model), stats [n_pairs] or
 * NULL.  Returns 0, or < 0 on error (gcr_driver_last_error() / stderr).  Not available in a batch: sampler 3 (one
 * probability table per context), GCR_PNP, more than 32768 correspondences per pair. */
int gcr_find_model_batch(int device, int problem, const double *h_points, int n, int n_pairs, const gcr_run_settings *settings,
                         const uint64_t *seeds, int seats, int engines, double *models_out, uint8_t *inlier_masks, int *inlier_counts,
                         gcr_run_stats *stats);

/* findEssentialMatrix_ for a batch: pixel correspondences; K1 / K2 row-major 3 x 3, one pair of matrices for all image
 * pairs (per_pair_K = 0) or one per image pair ([n_pairs][9], per_pair_K = 1) */
int gcr_find_essential_batch(int device, const double *h_points, int n, int n_pairs, const double *K1, const double *K2, int per_pair_K,
                             const gcr_run_settings *settings, const uint64_t *seeds, int seats, int engines, double *models_out,
                             uint8_t *inlier_masks, int *inlier_counts, gcr_run_stats *stats);

/* frees the engines (contexts, device buffers, pinned staging) the batch calls keep between calls */
void gcr_batch_engines_release(void);

/* the engine's host-side concurrency core (fibers, run queue, request / completion protocol) exercised without a GPU: n_fibers
 * fibers post `rounds` requests each, a dispatcher thread completes them, n_threads OS threads run the fibers.  Returns
 * n_fibers x rounds, or < 0 if a request was lost or a fiber resumed twice. */
long long gcr_fiber_selftest(int n_fibers, int n_threads, int rounds);

/* kernels launched so far on the engines' contexts (gcr_launch_count summed over them) */
uint64_t gcr_batch_launch_count(void);

/* the calling thread's last driver error (gcr_find_model / gcr_find_essential / the batch calls) */
const char *gcr_driver_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
