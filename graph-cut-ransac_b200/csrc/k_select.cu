// k_select.cu -- K4: best-model selection, and gcr_round(), the fused fixed-K pass (K1 -> K2 -> K3 -> K4).
//
// The reference keeps a running best with a strict '<' on Score::value (GCRANSAC.h:440, scoring_function.h:63-66), so
// among equal scores the FIRST model wins.  Here every score becomes a 64-bit key
//     key = (orderedBits(score) << 32) | (0xFFFFFFFF - global_index)
// and the best model is max(key): larger score first, then the smaller index.  max is associative and commutative,
// so a block reduction + one atomicMax per block is deterministic, and the same key is what the hypothesis-sharded
// multi-GPU mode all-reduces with ncclMax.
#include "gcr_internal.cuh"

__global__ void __launch_bounds__(256) k_argmax(const float *__restrict__ scores, int k, uint32_t index_base,
                                                unsigned long long *__restrict__ keys)
{
	const int pair = blockIdx.y;
	const float *s = scores + (size_t)pair * k;
	unsigned long long best = 0ull;
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < k; i += gridDim.x * blockDim.x)
	{
		const unsigned long long key =
		    ((unsigned long long)gcr_ordered_u32(s[i]) << 32) | (unsigned long long)(0xFFFFFFFFu - ((uint32_t)i + index_base));
		if (key > best)
			best = key;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
	{
		const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
		if (other > best)
			best = other;
	}
	__shared__ unsigned long long warp_best[8];
	if ((threadIdx.x & 31) == 0)
		warp_best[threadIdx.x >> 5] = best;
	__syncthreads();
	if (threadIdx.x == 0)
	{
		for (int w = 1; w < 8; ++w)
			if (warp_best[w] > best)
				best = warp_best[w];
		atomicMax(&keys[pair], best);
	}
}

__global__ void k_gather_best(const unsigned long long *__restrict__ keys, uint32_t index_base,
                              const double *__restrict__ models, int k, int ml, double *__restrict__ best_models,
                              int n_pairs)
{
	const int pair = blockIdx.x * blockDim.x + threadIdx.x;
	if (pair >= n_pairs)
		return;
	const uint32_t global_index = 0xFFFFFFFFu - (uint32_t)(keys[pair] & 0xFFFFFFFFull);
	const uint32_t local = global_index - index_base;
	for (int j = 0; j < ml; ++j)
		best_models[(size_t)pair * ml + j] = local < (uint32_t)k ? models[((size_t)pair * k + local) * ml + j] : 0.0;
}

extern "C" gcr_status gcr_argmax(gcr_ctx *ctx, const float *d_scores, int k, uint32_t index_base, uint64_t *d_keys)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!d_scores || !d_keys || k <= 0 || ctx->n_pairs <= 0)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_argmax: bad arguments");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	GCR_CHECK_CUDA(ctx, cudaMemsetAsync(d_keys, 0, sizeof(uint64_t) * ctx->n_pairs, ctx->stream));
	int gx = (k + 255) / 256;
	const int cap = ctx->n_pairs >= 2 * ctx->sm_count ? 1 : (4 * ctx->sm_count + ctx->n_pairs - 1) / ctx->n_pairs;
	if (gx > cap)
		gx = cap;
	k_argmax<<<dim3(gx, ctx->n_pairs), 256, 0, ctx->stream>>>(d_scores, k, index_base, (unsigned long long *)d_keys);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}

extern "C" float gcr_key_score(uint64_t key) { return gcr_from_ordered_u32((uint32_t)(key >> 32)); }
extern "C" uint32_t gcr_key_index(uint64_t key) { return 0xFFFFFFFFu - (uint32_t)(key & 0xFFFFFFFFull); }

extern "C" gcr_status gcr_set_band_factor(gcr_ctx *ctx, double f)
{
	if (!ctx || !(f >= 0.0))
		return GCR_ERR_ARG;
	ctx->band_factor = f;
	return GCR_OK;
}
extern "C" gcr_status gcr_set_score_trace(gcr_ctx *ctx, uint64_t *d_trace)
{
	if (!ctx)
		return GCR_ERR_ARG;
	ctx->d_score_trace = (unsigned long long *)d_trace;
	return GCR_OK;
}
extern "C" gcr_status gcr_set_score_variant(gcr_ctx *ctx, int v)
{
	if (!ctx)
		return GCR_ERR_ARG;
	ctx->score_variant = v;
	return GCR_OK;
}

// ------------------------------------------------------------------------------- exact winner selection ---
// The FP32 scores of K3 rank the hypotheses, but FP32 cannot separate two models whose scores differ by less than its
// own error (~2e-5 relative for Sampson: the model coefficients themselves are rounded to FP32).  gcr_round therefore
// keeps, per pair, the best FP32 key of each of SEL_BUCKETS index classes (slot % SEL_BUCKETS; max is order-independent,
// so this is deterministic; done inside K3's finalize kernel, k_score.cu), re-scores in FP64 -- reference operation order per point -- every bucket winner within
// SEL_MARGIN of the best, and picks the winner by (FP64 score, lowest index): the strict '<' first-wins rule of
// GCRANSAC.h:440 evaluated in the reference's precision.
#define SEL_BUCKETS GCR_SEL_BUCKETS
#define SEL_MARGIN 2e-4f
#define SEL_THREADS 256
#define SEL_MAX_THREADS 1024 /* few pairs: one big CTA per pair re-scores a candidate over all points in ~1 us */

struct SelectArgs
{
	const double *pts; // [n_pairs][n][cols]
	const double *models; // [n_pairs][k][ML]
	const double *lut64;
	const unsigned long long *buckets; // [n_pairs][SEL_BUCKETS]
	unsigned long long *keys;          // [n_pairs]
	double *best_models;               // [n_pairs][ML] or null
	double *best_scores;               // [n_pairs] or null (FP64 score of the winner)
	uint32_t *best_counts;             // [n_pairs] or null (exact inlier count of the winner)
	int n, k;
	uint32_t index_base;
	double thr;
};

template <int PROBLEM, int SCORING>
__global__ void __launch_bounds__(SEL_MAX_THREADS) k_select_exact(const SelectArgs a)
{
	constexpr int COLS = PROBLEM == GCR_PNP ? 5 : 4;
	constexpr int ML = PROBLEM == GCR_PNP ? 12 : 9;
	__shared__ unsigned long long s_keys[SEL_BUCKETS];
	__shared__ int s_cand[SEL_BUCKETS];
	__shared__ int s_ncand;
	__shared__ double s_model[ML];
	__shared__ double warp_sum[SEL_MAX_THREADS / 32];
	__shared__ uint32_t warp_cnt[SEL_MAX_THREADS / 32];
	__shared__ double s_best_score;
	__shared__ uint32_t s_best_slot, s_best_count;
	const int pair = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid < SEL_BUCKETS)
		s_keys[tid] = a.buckets[(size_t)pair * SEL_BUCKETS + tid];
	__syncthreads();
	if (tid == 0)
	{
		unsigned long long best = 0ull;
		for (int b = 0; b < SEL_BUCKETS; ++b)
			if (s_keys[b] > best) best = s_keys[b];
		int nc = 0;
		if (best)
		{
			const float top = gcr_from_ordered_u32((uint32_t)(best >> 32));
			const float floor_score = top - SEL_MARGIN * top;
			for (int b = 0; b < SEL_BUCKETS; ++b)
				if (s_keys[b] && gcr_from_ordered_u32((uint32_t)(s_keys[b] >> 32)) >= floor_score)
					s_cand[nc++] = b;
		}
		s_ncand = nc;
		s_best_score = -1.0;
		s_best_slot = 0xFFFFFFFFu;
		s_best_count = 0;
	}
	__syncthreads();
	const double *P = a.pts + (size_t)pair * a.n * COLS;
	// thresholds / MAGSAC constants exactly as the reference forms them (GCRANSAC.h:253-254, scoring_function.h:257-296)
	const double truncated_threshold = 3.0 / 2.0 * a.thr;
	const double T = truncated_threshold * truncated_threshold;
	const double gamma_k = 0.0036572608340910764;
	const double gamma_difference = 0.88622692545275801365 - gamma_k;
	const double one_over_sigma = 0.25 * 2.8284271247461903 / a.thr;
	const double weight_zero = one_over_sigma * gamma_difference;
	const double squared_sigma_max_2 = a.thr * a.thr * 2.0;
	for (int c = 0; c < s_ncand; ++c)
	{
		const uint32_t gidx = 0xFFFFFFFFu - (uint32_t)(s_keys[s_cand[c]] & 0xFFFFFFFFull);
		const uint32_t slot = gidx - a.index_base;
		if (tid < ML)
			s_model[tid] = a.models[((size_t)pair * a.k + slot) * ML + tid];
		__syncthreads();
		double M[ML];
		for (int j = 0; j < ML; ++j) M[j] = s_model[j];
		double sum = 0.0;
		uint32_t cnt = 0;
		for (int i = tid; i < a.n; i += (int)blockDim.x)
		{
			double pt[COLS];
			for (int j = 0; j < COLS; ++j) pt[j] = P[(size_t)i * COLS + j];
			const double r2 = gcr_sqres64<PROBLEM>(pt, M);
			if (SCORING == GCR_MSAC)
			{
				if (r2 < T) { sum -= r2; ++cnt; }
			}
			else
			{
				const double residual = sqrt(r2);
				if (!(residual > a.thr))
				{
					double weight;
					if (residual < 2.220446049250313e-16)
						weight = weight_zero;
					else
					{
						const double squared_residual = residual * residual;
						unsigned long long x = (unsigned long long)round(1e4 * squared_residual / squared_sigma_max_2);
						if (x > 5000) x = 5000;
						weight = one_over_sigma * (a.lut64[x] - gamma_k);
					}
					sum += weight / weight_zero;
					++cnt;
				}
			}
		}
		for (int o = 16; o > 0; o >>= 1)
		{
			sum += __shfl_xor_sync(0xffffffffu, sum, o);
			cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
		}
		if (lane == 0) { warp_sum[warp] = sum; warp_cnt[warp] = cnt; }
		__syncthreads();
		if (tid == 0)
		{
			double value = 0.0;
			uint32_t inl = 0;
			for (int w = 0; w < (int)blockDim.x / 32; ++w) { value += warp_sum[w]; inl += warp_cnt[w]; }
			const double score = SCORING == GCR_MSAC ? (inl == 0 ? 0.0 : (value + inl * T) / T) : value;
			if (score > s_best_score || (score == s_best_score && gidx < s_best_slot))
			{
				s_best_score = score;
				s_best_slot = gidx;
				s_best_count = inl;
			}
		}
		__syncthreads();
	}
	if (s_ncand == 0)
	{
		if (tid == 0)
		{
			a.keys[pair] = 0ull;
			if (a.best_scores) a.best_scores[pair] = 0.0;
			if (a.best_counts) a.best_counts[pair] = 0;
		}
		if (a.best_models && tid < ML)
			a.best_models[(size_t)pair * ML + tid] = 0.0;
		return;
	}
	if (tid == 0)
	{
		a.keys[pair] = ((unsigned long long)gcr_ordered_u32((float)s_best_score) << 32) | (unsigned long long)(0xFFFFFFFFu - s_best_slot);
		if (a.best_scores) a.best_scores[pair] = s_best_score;
		if (a.best_counts) a.best_counts[pair] = s_best_count;
	}
	if (a.best_models && tid < ML)
		a.best_models[(size_t)pair * ML + tid] = a.models[((size_t)pair * a.k + (s_best_slot - a.index_base)) * ML + tid];
}

// ------------------------------------------------------------------------------------------------ gcr_round ---
extern "C" gcr_status gcr_round(gcr_ctx *ctx, int problem, int scoring, int sampler, uint64_t first_index, int count,
                                uint64_t seed, double threshold, int flags, uint32_t index_base, uint64_t *d_keys,
                                double *d_best_models, uint64_t *evals_out)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!ctx->d_pts64 || ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_round: no points uploaded");
	if (count <= 0 || !d_keys)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_round: bad arguments");
	const int m = gcr_sample_size(problem), mm = gcr_max_models(problem), ml = gcr_model_len(problem);
	if (m == 0)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_round: unknown problem %d", problem);
	const size_t P = (size_t)ctx->n_pairs, K = (size_t)count * mm;
	// carve the round buffer: samples | models | valid | scores
	size_t off = 0;
	auto take = [&](size_t bytes) {
		const size_t o = off;
		off += (bytes + 255) & ~(size_t)255;
		return o;
	};
	const size_t o_samples = take(P * count * m * sizeof(uint32_t));
	const size_t o_models = take(P * K * ml * sizeof(double));
	const size_t o_valid = take(P * K);
	const size_t o_scores = take(P * K * sizeof(float));
	const size_t o_buckets = take(P * SEL_BUCKETS * sizeof(unsigned long long));
	const size_t o_best_scores = take(P * sizeof(double));
	const size_t o_best_counts = take(P * sizeof(uint32_t));
	if (off > ctx->round_capacity)
	{
		GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_round);
		ctx->d_round = nullptr;
		ctx->round_capacity = 0;
		GCR_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_round, off));
		ctx->round_capacity = off;
	}
	char *base = (char *)ctx->d_round;
	uint32_t *d_samples = (uint32_t *)(base + o_samples);
	double *d_models = (double *)(base + o_models);
	uint8_t *d_valid = (uint8_t *)(base + o_valid);
	float *d_scores = (float *)(base + o_scores);

	(void)d_samples; // K1 is fused into K2: the solver kernel draws its sample in registers
	gcr_status st = gcr_solve_minimal_generated(ctx, problem, sampler, first_index, count, seed, d_models, d_valid);
	if (st) return st;
	// K3 (+ K4's first stage inside its finalize kernel: best FP32 key per index class)
	unsigned long long *d_buckets = (unsigned long long *)(base + o_buckets);
	GCR_CHECK_CUDA(ctx, cudaMemsetAsync(d_buckets, 0, P * SEL_BUCKETS * sizeof(unsigned long long), ctx->stream));
	st = gcr_score_fused(ctx, problem, scoring, d_models, d_valid, (int)K, threshold, flags, d_scores, nullptr, d_buckets, index_base);
	if (st) return st;
	// K4: FP64 re-score of the near-best candidates -> winner (see k_select_exact)
	if (scoring == GCR_MAGSAC)
	{
		st = gcr_ensure_magsac_lut64(ctx);
		if (st) return st;
	}
	SelectArgs sa;
	sa.pts = ctx->d_pts64;
	sa.models = d_models;
	sa.lut64 = ctx->d_magsac_lut64;
	sa.buckets = d_buckets;
	sa.keys = (unsigned long long *)d_keys;
	sa.best_models = d_best_models;
	sa.best_scores = (double *)(base + o_best_scores);
	sa.best_counts = (uint32_t *)(base + o_best_counts);
	sa.n = ctx->n;
	sa.k = (int)K;
	sa.index_base = index_base;
	sa.thr = threshold;
	const int sel_threads = ctx->n_pairs < 2 * ctx->sm_count ? SEL_MAX_THREADS : SEL_THREADS;
#define GCR_SEL(PR)                                                                                          \
	do                                                                                                       \
	{                                                                                                        \
		if (scoring == GCR_MSAC)                                                                             \
			k_select_exact<PR, GCR_MSAC><<<ctx->n_pairs, sel_threads, 0, ctx->stream>>>(sa);                  \
		else                                                                                                 \
			k_select_exact<PR, GCR_MAGSAC><<<ctx->n_pairs, sel_threads, 0, ctx->stream>>>(sa);                \
	} while (0)
	if (problem == GCR_H) GCR_SEL(GCR_H);
	else if (problem == GCR_PNP) GCR_SEL(GCR_PNP);
	else GCR_SEL(GCR_F);
#undef GCR_SEL
	GCR_LAUNCHED(ctx);
	ctx->d_round_best_scores = sa.best_scores;
	ctx->d_round_best_counts = sa.best_counts;
	if (evals_out)
		*evals_out = (uint64_t)P * K * (uint64_t)ctx->n;
	return GCR_OK;
}

extern "C" gcr_status gcr_round_results(gcr_ctx *ctx, double *h_best_scores, uint32_t *h_best_counts)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!ctx->d_round_best_scores)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_round_results: no round has run on this context");
	gcr_status st = GCR_OK;
	if (h_best_scores)
		st = gcr_memcpy_d2h(ctx, h_best_scores, ctx->d_round_best_scores, (size_t)ctx->n_pairs * sizeof(double));
	if (!st && h_best_counts)
		st = gcr_memcpy_d2h(ctx, h_best_counts, ctx->d_round_best_counts, (size_t)ctx->n_pairs * sizeof(uint32_t));
	return st;
}
