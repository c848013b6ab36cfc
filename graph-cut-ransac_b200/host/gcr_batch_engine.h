// gcr_batch_engine.h -- the lock-step multi-pair driver (findHomographyBatch & co.).
//
// P image pairs are uploaded to ONE context (n_pairs = P).  A host thread per pair in flight runs the SAME sequential
// driver as gcr_find_model (gcr_driver.cpp: the replay of GCRANSAC::run, GCRANSAC.h:202-627) over a BatchedBackend whose
// operations are requests: the thread posts "main batch [first, first + B)", "FP64 getScore of these models", "labeling
// terms of this model", "LO round: pool -> samples -> refits -> scores", "refit of this inlier set" and sleeps.  The
// dispatcher takes whatever is pending, groups it by kind and executes every group as ONE launch sequence for all the
// pairs that asked (the per-item pair index entry points of include/gcr/gcr.h: gcr_set_active_pairs,
// gcr_score_exact_multi, gcr_labeling_terms_multi, gcr_sample_from_pools, gcr_refit_multi) with one synchronise, hands
// the results back and wakes the threads.  The host-only work between two requests -- the reference's Boykov-Kolmogorov
// min-cut above all -- runs in the pair threads, i.e. fanned out over the host cores.
//
// The driver calls per pair drop from ~150 (each a trip through the CUDA context's lock, which is what bounds the
// one-pair-per-thread pipelines at ~650 k calls/s) to ~30 per super-step whatever the number of pairs in it.
// Results per pair equal gcr_find_model's with seed + pair * 0x9E3779B97F4A7C15 (the per-pair stream offset of the
// sampling kernels): tests/test_gpu_batch.py.
#ifndef GCR_BATCH_ENGINE_H_INCLUDED
#define GCR_BATCH_ENGINE_H_INCLUDED

#include "gcr_driver.h"

#ifdef __cplusplus
extern "C" {
#endif

/* GCRANSAC<>::run for n_pairs independent problems of n correspondences each (h_points [n_pairs][n][4], host), problem
 * GCR_H, GCR_F or GCR_E (K-normalised points; see gcr_find_essential_batch).  One settings structure for all pairs; pair p
 * draws from the streams of settings->seed + p * 0x9E3779B97F4A7C15.  host_threads: pairs in flight (0 = 4 x the host
 * cores, at most n_pairs).  models_out [n_pairs][9], inlier_masks [n_pairs][n], inlier_counts [n_pairs] (0 = no model),
 * stats [n_pairs] or NULL.  Returns 0, or < 0 on error (gcr_last_error(ctx) / stderr).
 * Not supported in a batch: sampler 3 (one probability table per context), GCR_PNP. */
int gcr_find_model_batch(gcr_ctx *ctx, int problem, const double *h_points, int n, int n_pairs, const gcr_run_settings *settings,
                         int host_threads, double *models_out, uint8_t *inlier_masks, int *inlier_counts, gcr_run_stats *stats);

/* findEssentialMatrix_ for a batch: pixel correspondences, one K1 / K2 (row-major 3 x 3) per pair ([n_pairs][9]) */
int gcr_find_essential_batch(gcr_ctx *ctx, const double *h_points, int n, int n_pairs, const double *K1, const double *K2,
                             const gcr_run_settings *settings, int host_threads, double *models_out, uint8_t *inlier_masks,
                             int *inlier_counts, gcr_run_stats *stats);

#ifdef __cplusplus
}
#endif
#endif
