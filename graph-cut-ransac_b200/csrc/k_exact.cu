// k_exact.cu -- K5: FP64 scoring of a few models in the reference's operation order, with the ascending inlier list
// (the side output of getScore, scoring_function.h:166-167 / :338-344) and optional per-point residuals (what
// GCRANSAC::labeling, GCRANSAC.h:945-961, and the Tukey weights, :657-668, consume).
//
// One CTA per model; threads stride over the points in 1024-point blocks; the inlier list is written with a
// ballot/prefix compaction so indices come out ascending exactly like the reference's emplace_back order.
// HBM-bound: 32 B (40 B PnP) read per point per model; used O(log K) times per run (new-best events, LO, IRLS).
#include "gcr_internal.cuh"

#include <math.h>
#include <string.h>

#define EX_THREADS 1024

struct ExactArgs
{
	const double *pts;    // [n][cols] of the selected pair
	const double *models; // [k][ML]
	const double *lut64;  // MAGSAC: Gamma(1.5, j/1000), j = 0..5000
	double *scores;
	uint32_t *counts;
	uint32_t *inlier_idx; // [k][n] or null
	double *residuals;    // [k][n] or null
	int n, k;
	double thr;
	double T_override; // > 0: squared truncated threshold handed over as is (ScoringFunction::initialize), else from thr
	size_t pair_stride; // 0: every model against the selected pair; n * cols: model m against pair m (one model per pair)
	const int32_t *pair_ids; // with pair_stride: model m against pair pair_ids[m] (gcr_score_exact_multi), or null
};

template <int PROBLEM, int SCORING>
__global__ void __launch_bounds__(EX_THREADS) k_score_exact(const ExactArgs a)
{
	constexpr int COLS = PROBLEM == GCR_PNP ? 5 : 4;
	constexpr int ML = PROBLEM == GCR_PNP ? 12 : 9;
	__shared__ uint32_t warp_counts[32];
	__shared__ double warp_sums[32];
	__shared__ uint32_t base_shared;

	const int model = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	double M[ML];
	for (int i = 0; i < ML; ++i)
		M[i] = a.models[(size_t)model * ML + i];

	// thresholds exactly as the reference forms them
	const double truncated_threshold = 3.0 / 2.0 * a.thr;                                 // GCRANSAC.h:253
	const double squared_truncated_threshold =
	    a.T_override > 0.0 ? a.T_override : truncated_threshold * truncated_threshold; // GCRANSAC.h:254
	// MAGSAC constants, scoring_function.h:257-296
	const double gamma_k = 0.0036572608340910764;
	const double gamma_difference = 0.88622692545275801365 - gamma_k; // tgamma(1.5)
	const double C_times_two_ad_dof = 0.25 * 2.8284271247461903; // C * pow(2, 1.5)
	const double squared_sigma_max_2 = a.thr * a.thr * 2.0;
	const double one_over_sigma = C_times_two_ad_dof / a.thr;
	const double weight_zero = one_over_sigma * gamma_difference;

	if (tid == 0)
		base_shared = 0;
	__syncthreads();

	double local_sum = 0.0;
	for (int start = 0; start < a.n; start += EX_THREADS)
	{
		const int i = start + tid;
		bool inlier = false;
		if (i < a.n)
		{
			double pt[COLS];
			for (int c = 0; c < COLS; ++c)
				pt[c] = a.pts[(size_t)(a.pair_ids ? a.pair_ids[model] : model) * a.pair_stride + (size_t)i * COLS + c];
			const double r2 = gcr_sqres64<PROBLEM>(pt, M);
			if (a.residuals)
				a.residuals[(size_t)model * a.n + i] = r2;
			if (SCORING == GCR_MSAC)
			{
				if (r2 < squared_truncated_threshold) // scoring_function.h:163
				{
					inlier = true;
					local_sum -= r2; // :177
				}
			}
			else
			{
				const double residual = sqrt(r2); // :300-302
				if (!(residual > a.thr))          // :304
				{
					double weight;
					if (residual < 2.220446049250313e-16)
						weight = weight_zero;
					else
					{
						const double squared_residual = residual * residual;
						unsigned long long x = (unsigned long long)round(1e4 * squared_residual / squared_sigma_max_2);
						if (36000 < x)
							x = 36000;
						if (x > 5000)
							x = 5000; // unreachable for residual <= thr; guards the 5001-entry table
						weight = one_over_sigma * (a.lut64[x] - gamma_k);
					}
					local_sum += weight / weight_zero;
					inlier = true; // residual <= threshold_ (:338)
				}
			}
		}
		// ordered compaction
		const uint32_t ballot = __ballot_sync(0xffffffffu, inlier);
		if (lane == 0)
			warp_counts[warp] = __popc(ballot);
		__syncthreads();
		uint32_t before = 0, total = 0;
		for (int w = 0; w < EX_THREADS / 32; ++w)
		{
			const uint32_t c = warp_counts[w];
			if (w < warp) before += c;
			total += c;
		}
		const uint32_t base = base_shared;
		if (inlier && a.inlier_idx)
			a.inlier_idx[(size_t)model * a.n + base + before + __popc(ballot & ((1u << lane) - 1))] = (uint32_t)i;
		__syncthreads();
		if (tid == 0)
			base_shared = base + total;
	}
	// deterministic tree sum of the score terms
	for (int o = 16; o > 0; o >>= 1)
		local_sum += __shfl_xor_sync(0xffffffffu, local_sum, o);
	if (lane == 0)
		warp_sums[warp] = local_sum;
	__syncthreads();
	if (tid == 0)
	{
		double value = 0.0;
		for (int w = 0; w < EX_THREADS / 32; ++w)
			value += warp_sums[w];
		const uint32_t inliers = base_shared;
		double score;
		if (SCORING == GCR_MSAC)
			score = inliers == 0 ? 0.0
			                     : (value + inliers * squared_truncated_threshold) / squared_truncated_threshold; // :221-229
		else
			score = value;
		a.scores[model] = score;
		a.counts[model] = inliers;
	}
}

__global__ void k_fill_lut64(double *lut)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k > 5000)
		return;
	const double x = k / 1000.0, sx = sqrt(x);
	lut[k] = 0.88622692545275801365 * erfc(sx) + sx * exp(-x);
}

gcr_status gcr_ensure_magsac_lut64(gcr_ctx *ctx)
{
	if (ctx->d_magsac_lut64)
		return GCR_OK;
	GCR_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_magsac_lut64, 5008 * sizeof(double)));
	k_fill_lut64<<<(5001 + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_magsac_lut64);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}

static gcr_status score_exact_impl(gcr_ctx *ctx, int problem, int scoring, int pair, const double *d_models, int k,
                                   double threshold, double T_override, double *d_scores, uint32_t *d_counts,
                                   uint32_t *d_inlier_idx, double *d_residuals, const int32_t *d_pair_ids = nullptr);

// k models that belong to DIFFERENT pairs of the context (model m against pair d_pair_ids[m]), one launch: the batch
// engine's getScore
extern "C" gcr_status gcr_score_exact_multi(gcr_ctx *ctx, int problem, int scoring, const int32_t *d_pair_ids, const double *d_models,
                                            int k, double threshold, double *d_scores, uint32_t *d_counts, uint32_t *d_inlier_idx,
                                            double *d_residuals)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!d_pair_ids)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_score_exact_multi: d_pair_ids is required");
	return score_exact_impl(ctx, problem, scoring, -2, d_models, k, threshold, 0.0, d_scores, d_counts, d_inlier_idx, d_residuals,
	                        d_pair_ids);
}

extern "C" gcr_status gcr_score_exact(gcr_ctx *ctx, int problem, int scoring, int pair, const double *d_models, int k,
                                      double threshold, double *d_scores, uint32_t *d_counts, uint32_t *d_inlier_idx,
                                      double *d_residuals)
{
	return score_exact_impl(ctx, problem, scoring, pair, d_models, k, threshold, 0.0, d_scores, d_counts, d_inlier_idx,
	                        d_residuals);
}

extern "C" gcr_status gcr_score_exact_pairs(gcr_ctx *ctx, int problem, int scoring, const double *d_models, double threshold,
                                            double *d_scores, uint32_t *d_counts, uint32_t *d_inlier_idx)
{
	if (!ctx)
		return GCR_ERR_ARG;
	return score_exact_impl(ctx, problem, scoring, -1, d_models, ctx->n_pairs, threshold, 0.0, d_scores, d_counts, d_inlier_idx, nullptr);
}

static gcr_status score_exact_impl(gcr_ctx *ctx, int problem, int scoring, int pair, const double *d_models, int k,
                                   double threshold, double T_override, double *d_scores, uint32_t *d_counts,
                                   uint32_t *d_inlier_idx, double *d_residuals, const int32_t *d_pair_ids)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!ctx->d_pts64 || ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_score_exact: no points uploaded");
	if (problem < GCR_H || problem > GCR_PNP || gcr_cols_of(problem) != ctx->cols)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_score_exact: problem does not match the uploaded points");
	// pair == -1 (internal, gcr_score_exact_pairs): k == n_pairs models, model m scored against pair m
	// pair == -2 (internal, gcr_score_exact_multi): model m scored against pair d_pair_ids[m]
	if (pair < -2 || pair >= ctx->n_pairs || (pair == -1 && k != ctx->n_pairs) || k <= 0 || !d_models || !d_scores || !d_counts)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_score_exact: bad arguments");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	if (scoring == GCR_MAGSAC)
	{
		gcr_status lst = gcr_ensure_magsac_lut64(ctx);
		if (lst)
			return lst;
	}
	ExactArgs a;
	a.pts = pair < 0 ? ctx->d_pts64 : ctx->d_pts64 + (size_t)pair * ctx->n * ctx->cols;
	a.models = d_models;
	a.lut64 = ctx->d_magsac_lut64;
	a.scores = d_scores;
	a.counts = d_counts;
	a.inlier_idx = d_inlier_idx;
	a.residuals = d_residuals;
	a.n = ctx->n;
	a.k = k;
	a.thr = threshold;
	a.T_override = T_override;
	a.pair_stride = 0;
	a.pair_ids = d_pair_ids;
	if (pair < 0) // gcr_score_exact_pairs: model m belongs to pair m (gcr_score_exact_multi: to pair d_pair_ids[m])
	{
		a.pts = ctx->d_pts64;
		a.pair_stride = (size_t)ctx->n * ctx->cols;
	}
#define GCR_EX(P)                                                                                        \
	do                                                                                                   \
	{                                                                                                    \
		if (scoring == GCR_MSAC)                                                                         \
			k_score_exact<P, GCR_MSAC><<<k, EX_THREADS, 0, ctx->stream>>>(a);                             \
		else                                                                                             \
			k_score_exact<P, GCR_MAGSAC><<<k, EX_THREADS, 0, ctx->stream>>>(a);                           \
	} while (0)
	if (problem == GCR_H)
		GCR_EX(GCR_H);
	else if (problem == GCR_PNP)
		GCR_EX(GCR_PNP);
	else
		GCR_EX(GCR_F);
#undef GCR_EX
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}

extern "C" gcr_status gcr_score_exact_host(gcr_ctx *ctx, int problem, int scoring, int pair, const double *h_models,
                                           int k, double threshold, double *h_scores, uint32_t *h_counts,
                                           uint32_t *h_inlier_idx, double *h_residuals)
{
	if (!ctx || !h_models || !h_scores || !h_counts || k <= 0)
		return GCR_ERR_ARG;
	const int ml = problem == GCR_PNP ? 12 : 9;
	const size_t n = ctx->n;
	void *d_m = nullptr, *d_s = nullptr, *d_c = nullptr, *d_i = nullptr, *d_r = nullptr;
	gcr_status st = gcr_malloc(ctx, (size_t)k * ml * sizeof(double), &d_m);
	if (!st) st = gcr_malloc(ctx, k * sizeof(double), &d_s);
	if (!st) st = gcr_malloc(ctx, k * sizeof(uint32_t), &d_c);
	if (!st && h_inlier_idx) st = gcr_malloc(ctx, (size_t)k * n * sizeof(uint32_t), &d_i);
	if (!st && h_residuals) st = gcr_malloc(ctx, (size_t)k * n * sizeof(double), &d_r);
	if (!st) st = gcr_memcpy_h2d(ctx, d_m, h_models, (size_t)k * ml * sizeof(double));
	if (!st)
		st = gcr_score_exact(ctx, problem, scoring, pair, (const double *)d_m, k, threshold, (double *)d_s,
		                     (uint32_t *)d_c, (uint32_t *)d_i, (double *)d_r);
	// the four read-backs are queued and delivered by ONE synchronise; the temporaries go back to the context's block
	// cache (gcr_free, on every path: a raw cudaFree would leave their cache entries dangling)
	if (!st) st = gcr_memcpy_d2h_async(ctx, h_scores, d_s, k * sizeof(double));
	if (!st) st = gcr_memcpy_d2h_async(ctx, h_counts, d_c, k * sizeof(uint32_t));
	if (!st && h_inlier_idx) st = gcr_memcpy_d2h_async(ctx, h_inlier_idx, d_i, (size_t)k * n * sizeof(uint32_t));
	if (!st && h_residuals) st = gcr_memcpy_d2h_async(ctx, h_residuals, d_r, (size_t)k * n * sizeof(double));
	const gcr_status sync_st = gcr_sync(ctx); // also orders the reuse of the blocks after the kernels that wrote them
	if (!st) st = sync_st;
	gcr_free(ctx, d_m);
	gcr_free(ctx, d_s);
	gcr_free(ctx, d_c);
	gcr_free(ctx, d_i);
	gcr_free(ctx, d_r);
	return st;
}

// ScoringFunction::getScore for ONE model with host in / out -- what gcr::CudaScoringFunction (host/gcr_scoring_function.h)
// calls.  squared_truncated_threshold is the value MSACScoringFunction::initialize received (GCRANSAC.h:253-255) and is
// used as is; threshold is getScore's threshold_ argument (MAGSAC's sigma_max).  Device buffers live in the context.
extern "C" gcr_status gcr_getscore_host(gcr_ctx *ctx, int problem, int scoring, int pair, const double *h_model,
                                        double squared_truncated_threshold, double threshold, double *h_score,
                                        uint32_t *h_count, uint32_t *h_inlier_idx)
{
	if (!ctx || !h_model || !h_score || !h_count)
		return GCR_ERR_ARG;
	if (!ctx->d_pts64 || ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_getscore_host: no points uploaded");
	if (scoring == GCR_MSAC && !(squared_truncated_threshold > 0.0))
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_getscore_host: MSAC needs the squared truncated threshold of initialize()");
	if (scoring == GCR_MAGSAC && !(threshold > 0.0))
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_getscore_host: MAGSAC needs threshold > 0");
	const int ml = problem == GCR_PNP ? 12 : 9;
	const size_t n = ctx->n;
	// scratch: model | score | count | inlier list
	const size_t o_model = 0, o_score = 128, o_count = 136, o_idx = 256;
	gcr_status st = gcr_ensure_scratch(ctx, o_idx + n * sizeof(uint32_t));
	if (st)
		return st;
	st = gcr_ensure_pinned(ctx, o_idx + n * sizeof(uint32_t));
	if (st)
		return st;
	char *d = (char *)ctx->d_scratch, *h = (char *)ctx->h_pinned;
	memcpy(h + o_model, h_model, ml * sizeof(double));
	GCR_CHECK_CUDA(ctx, cudaMemcpyAsync(d + o_model, h + o_model, ml * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
	st = score_exact_impl(ctx, problem, scoring, pair, (const double *)(d + o_model), 1, threshold,
	                      scoring == GCR_MSAC ? squared_truncated_threshold : 0.0, (double *)(d + o_score),
	                      (uint32_t *)(d + o_count), h_inlier_idx ? (uint32_t *)(d + o_idx) : nullptr, nullptr);
	if (st)
		return st;
	// score, count and the whole index buffer in ONE copy / synchronise (the list is at most n entries: 4 n bytes)
	const size_t back = h_inlier_idx ? o_idx - o_score + n * sizeof(uint32_t) : 16;
	GCR_CHECK_CUDA(ctx, cudaMemcpyAsync(h + o_score, d + o_score, back, cudaMemcpyDeviceToHost, ctx->stream));
	GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	*h_score = *(double *)(h + o_score);
	*h_count = *(uint32_t *)(h + o_count);
	if (h_inlier_idx && *h_count)
		memcpy(h_inlier_idx, h + o_idx, (size_t)*h_count * sizeof(uint32_t));
	return GCR_OK;
}
