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
extern "C" gcr_status gcr_set_tuning(gcr_ctx *ctx, int knob, int value)
{
	if (!ctx)
		return GCR_ERR_ARG;
	switch (knob)
	{
	case GCR_TUNE_K3_CTAS_PER_SM: ctx->tune_k3_ctas = value < 0 ? 0 : value; return GCR_OK;
	case GCR_TUNE_ROUND_PREFETCH: ctx->tune_round_prefetch = value != 0; return GCR_OK;
	case GCR_TUNE_SMEM_CARVEOUT: ctx->tune_carveout = value; return GCR_OK;
	default: GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_set_tuning: unknown knob %d", knob);
	}
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
// own error (~2e-5 relative for Sampson: the model coefficients themselves are rounded to FP32).  K3's finalize kernel
// leaves each pair's best FP32 key; k_select_exact then re-scores in FP64 -- reference operation order per point --
// EVERY slot whose FP32 score lies within a relative margin of that best (no bucketing: the candidate set is complete),
// and picks the winner by (FP64 score, lowest index): the strict '<' first-wins rule of GCRANSAC.h:440 /
// scoring_function.h:63-66 evaluated in the reference's precision.  With FP32 score errors of at most e, the true FP64
// winner W satisfies s32(W) >= s32(top) (1 - 2e), so the margin must be >= 2e: MSAC 2e-4 (measured e <= 2e-5), MAGSAC++
// 1e-3 (e <= 2e-4).  Every re-score also records |s32 - s64| / s64 (counters[4], the largest seen), so the tests can
// assert the premise instead of trusting it.
//
// Grid: (cx, n_pairs).  CTA (b, p) scans slots [b k / cx, (b + 1) k / cx) of pair p, keeps its own best and leaves it in
// partials[p][b]; the last CTA of a pair to finish reduces the cx partials by the same (score, index) rule -- the result
// does not depend on which CTA that is -- merges with the running winner record when asked to (chunks of one problem:
// the running best never leaves the device, and stays FP64), and writes every output.
#define SEL_MARGIN_MSAC 2e-4f
#define SEL_MARGIN_MAGSAC 1e-3f
#define SEL_THREADS 256
#define SEL_MAX_THREADS 1024 /* few pairs: big CTAs re-score a candidate over all points in ~1 us */

struct SelPartial
{
	double score;
	uint32_t slot;
	uint32_t count;
};

struct SelectArgs
{
	const double *pts; // [n_pairs][n][cols]
	const double *models; // [n_pairs][k][ML]
	const double *lut64;
	const float *scores32;               // [n_pairs][k] K3's FP32 scores
	const unsigned long long *pairbest;  // [n_pairs] best FP32 key of each pair (finalize)
	SelPartial *partials;                // [n_pairs][gridDim.x]
	uint32_t *done;                      // [n_pairs] CTAs finished (zeroed with pairbest)
	unsigned long long *keys;            // [n_pairs] legacy packed key of THIS round's winner, or null
	double *best_models;                 // [n_pairs][ML] or null
	double *best_scores;                 // [n_pairs] FP64 score of this round's winner
	uint32_t *best_counts;               // [n_pairs] its exact inlier count
	gcr_winner *winners;                 // [n_pairs] winner records, or null
	unsigned long long *counters;
	int n, k, accumulate;
	unsigned long long index_base;
	double thr;
	float margin;
};

template <int PROBLEM, int SCORING>
__global__ void __launch_bounds__(SEL_MAX_THREADS) k_select_exact(const SelectArgs a)
{
	constexpr int COLS = PROBLEM == GCR_PNP ? 5 : 4;
	constexpr int ML = PROBLEM == GCR_PNP ? 12 : 9;
	__shared__ uint32_t s_cand[SEL_MAX_THREADS];
	__shared__ uint32_t s_ncand;
	__shared__ double s_model[ML];
	__shared__ double warp_sum[SEL_MAX_THREADS / 32];
	__shared__ uint32_t warp_cnt[SEL_MAX_THREADS / 32], warp_slot[SEL_MAX_THREADS / 32];
	__shared__ double s_best_score;
	__shared__ uint32_t s_best_slot, s_best_count, s_last, s_take, s_rescored;
	__shared__ float s_max_err;
	const int pair = blockIdx.y, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = (int)blockDim.x >> 5;
	const unsigned long long top_key = a.pairbest[pair];
	if (tid == 0)
	{
		s_best_score = -1.0;
		s_best_slot = 0xFFFFFFFFu;
		s_best_count = 0;
		s_rescored = 0;
		s_max_err = 0.f;
	}
	const double *P = a.pts + (size_t)pair * a.n * COLS;
	// thresholds / MAGSAC constants exactly as the reference forms them (GCRANSAC.h:253-254, scoring_function.h:257-296)
	const double truncated_threshold = 3.0 / 2.0 * a.thr;
	const double T = truncated_threshold * truncated_threshold;
	const double gamma_k = 0.0036572608340910764;
	const double gamma_difference = 0.88622692545275801365 - gamma_k;
	const double one_over_sigma = 0.25 * 2.8284271247461903 / a.thr;
	const double weight_zero = one_over_sigma * gamma_difference;
	const double squared_sigma_max_2 = a.thr * a.thr * 2.0;
	if (top_key)
	{
		const float top = gcr_from_ordered_u32((uint32_t)(top_key >> 32));
		const float floor_score = top - a.margin * top;
		const int per = (a.k + (int)gridDim.x - 1) / (int)gridDim.x;
		const int lo = (int)blockIdx.x * per, hi = min(a.k, lo + per);
		const float *sc = a.scores32 + (size_t)pair * a.k;
		for (int base = lo; base < hi; base += (int)blockDim.x)
		{
			__syncthreads();
			if (tid == 0)
				s_ncand = 0;
			__syncthreads();
			const int slot = base + tid;
			if (slot < hi && sc[slot] >= floor_score)
				s_cand[atomicAdd(&s_ncand, 1u)] = (uint32_t)slot; // order is irrelevant: the rule below is symmetric
			__syncthreads();
			const int nc = (int)s_ncand;
			for (int c = 0; c < nc; ++c)
			{
				const uint32_t slot_c = s_cand[c];
				if (tid < ML)
					s_model[tid] = a.models[((size_t)pair * a.k + slot_c) * ML + tid];
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
					for (int w = 0; w < nwarps; ++w) { value += warp_sum[w]; inl += warp_cnt[w]; }
					const double score = SCORING == GCR_MSAC ? (inl == 0 ? 0.0 : (value + inl * T) / T) : value;
					if (score > s_best_score || (score == s_best_score && slot_c < s_best_slot))
					{
						s_best_score = score;
						s_best_slot = slot_c;
						s_best_count = inl;
					}
					++s_rescored;
					if (score > 0.0)
						s_max_err = fmaxf(s_max_err, (float)(fabs((double)sc[slot_c] - score) / score));
				}
				__syncthreads();
			}
		}
	}
	// ---- this CTA's best -> partials; the last CTA of the pair reduces them ----
	__syncthreads();
	if (tid == 0)
	{
		SelPartial sp;
		sp.score = s_best_score;
		sp.slot = s_best_slot;
		sp.count = s_best_count;
		a.partials[(size_t)pair * gridDim.x + blockIdx.x] = sp;
		if (s_rescored)
		{
			atomicAdd(&a.counters[3], (unsigned long long)s_rescored);
			atomicMax(&a.counters[4], (unsigned long long)__float_as_uint(s_max_err)); // non-negative floats order like their bits
		}
		__threadfence();
		s_last = atomicAdd(&a.done[pair], 1u) == gridDim.x - 1 ? 1u : 0u;
	}
	__syncthreads();
	if (!s_last)
		return;
	__threadfence();
	double bs = -1.0;
	uint32_t bslot = 0xFFFFFFFFu, bcnt = 0;
	for (int b = tid; b < (int)gridDim.x; b += (int)blockDim.x)
	{
		const SelPartial *sp = &a.partials[(size_t)pair * gridDim.x + b];
		const double s = __ldcg(&sp->score);
		const uint32_t sl = __ldcg(&sp->slot), c = __ldcg(&sp->count);
		if (s > bs || (s == bs && sl < bslot)) { bs = s; bslot = sl; bcnt = c; }
	}
	for (int o = 16; o > 0; o >>= 1)
	{
		const double s = __shfl_xor_sync(0xffffffffu, bs, o);
		const uint32_t sl = __shfl_xor_sync(0xffffffffu, bslot, o), c = __shfl_xor_sync(0xffffffffu, bcnt, o);
		if (s > bs || (s == bs && sl < bslot)) { bs = s; bslot = sl; bcnt = c; }
	}
	if (lane == 0) { warp_sum[warp] = bs; warp_slot[warp] = bslot; warp_cnt[warp] = bcnt; }
	__syncthreads();
	if (tid == 0)
	{
		for (int w = 1; w < nwarps; ++w)
			if (warp_sum[w] > bs || (warp_sum[w] == bs && warp_slot[w] < bslot)) { bs = warp_sum[w]; bslot = warp_slot[w]; bcnt = warp_cnt[w]; }
		const bool have = bslot != 0xFFFFFFFFu && bs > 0.0;
		const unsigned long long gidx = a.index_base + (unsigned long long)bslot;
		if (a.keys)
			a.keys[pair] = have ? (((unsigned long long)gcr_ordered_u32((float)bs) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)gidx)) : 0ull;
		a.best_scores[pair] = have ? bs : 0.0;
		a.best_counts[pair] = have ? bcnt : 0u;
		uint32_t take = have ? 1u : 0u; // does this round's winner replace the record?
		if (a.winners)
		{
			gcr_winner *w = &a.winners[pair];
			if (a.accumulate && w->valid)
			{
				// running best over the chunks of one problem: strict '<' on the FP64 score, ties to the lower global index
				if (!have || !(bs > w->score || (bs == w->score && gidx < w->index)))
					take = 0u;
			}
			else if (!have)
			{
				w->score = 0.0;
				w->index = 0ull;
				w->count = 0u;
				w->valid = 0u;
				w->reserved = 0ull;
			}
			if (take)
			{
				w->score = bs;
				w->index = gidx;
				w->count = bcnt;
				w->valid = 1u;
				w->reserved = 0ull;
			}
		}
		s_take = take;
		s_best_slot = have ? bslot : 0xFFFFFFFFu;
	}
	__syncthreads();
	const bool have = s_best_slot != 0xFFFFFFFFu;
	if (tid < 12)
	{
		const double v = have && tid < ML ? a.models[((size_t)pair * a.k + s_best_slot) * ML + tid] : 0.0;
		if (a.best_models && tid < ML)
			a.best_models[(size_t)pair * ML + tid] = v;
		if (a.winners && (s_take || !(a.accumulate && a.winners[pair].valid)))
			a.winners[pair].model[tid] = v;
	}
}

// ------------------------------------------------------------------------------------------------ gcr_round ---
// index_base: global slot index of (sample first_index, model 0) -- what the keys / winner records report.
static gcr_status round_impl(gcr_ctx *ctx, int problem, int scoring, int sampler, uint64_t first_index, int count,
                             uint64_t seed, double threshold, int flags, uint64_t index_base, uint64_t *d_keys,
                             double *d_best_models, gcr_winner *d_winners, int accumulate, uint64_t *evals_out)
{
	if (!ctx->d_pts64 || ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_round: no points uploaded");
	if (count <= 0 || (!d_keys && !d_winners))
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_round: bad arguments");
	const int m = gcr_sample_size(problem), mm = gcr_max_models(problem), ml = gcr_model_len(problem);
	if (m == 0)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_round: unknown problem %d", problem);
	const size_t P = (size_t)ctx->n_pairs, K = (size_t)count * mm;
	if (K >= (1ull << 31))
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_round: count * max_models must stay below 2^31 model slots per call (walk the samples in chunks)");
	// grid of the selection kernel: many CTAs per pair when there are few pairs
	const int sel_threads = ctx->n_pairs < 2 * ctx->sm_count ? SEL_MAX_THREADS : SEL_THREADS;
	int cx = (int)((K + sel_threads - 1) / sel_threads);
	const int cx_cap = ctx->n_pairs >= 2 * ctx->sm_count ? 1 : (2 * ctx->sm_count + ctx->n_pairs - 1) / ctx->n_pairs;
	if (cx > cx_cap) cx = cx_cap;
	if (cx < 1) cx = 1;
	// carve the round buffer
	size_t off = 0;
	auto take = [&](size_t bytes) {
		const size_t o = off;
		off += (bytes + 255) & ~(size_t)255;
		return o;
	};
	const size_t o_models = take(P * K * ml * sizeof(double));
	const size_t o_valid = take(P * K);
	const size_t o_scores = take(P * K * sizeof(float));
	const size_t o_pairbest = take(P * sizeof(unsigned long long) + P * sizeof(uint32_t)); // pairbest | done: one memset
	const size_t o_partials = take(P * cx * sizeof(SelPartial));
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
	double *d_models = (double *)(base + o_models);
	uint8_t *d_valid = (uint8_t *)(base + o_valid);
	float *d_scores = (float *)(base + o_scores);

	// K1 is fused into K2: the solver kernel draws its sample in registers
	gcr_status st = gcr_solve_minimal_generated(ctx, problem, sampler, first_index, count, seed, d_models, d_valid);
	if (st) return st;
	// K3 (+ K4's first stage inside its finalize kernel: the pair's best FP32 key)
	unsigned long long *d_pairbest = (unsigned long long *)(base + o_pairbest);
	GCR_CHECK_CUDA(ctx, cudaMemsetAsync(d_pairbest, 0, P * sizeof(unsigned long long) + P * sizeof(uint32_t), ctx->stream));
	st = gcr_score_fused(ctx, problem, scoring, d_models, d_valid, (int)K, threshold, flags, d_scores, nullptr, d_pairbest);
	if (st) return st;
	// K4: FP64 re-score of every near-best candidate -> winner (see k_select_exact)
	if (scoring == GCR_MAGSAC)
	{
		st = gcr_ensure_magsac_lut64(ctx);
		if (st) return st;
	}
	SelectArgs sa;
	sa.pts = ctx->d_pts64;
	sa.models = d_models;
	sa.lut64 = ctx->d_magsac_lut64;
	sa.scores32 = d_scores;
	sa.pairbest = d_pairbest;
	sa.partials = (SelPartial *)(base + o_partials);
	sa.done = (uint32_t *)(base + o_pairbest + P * sizeof(unsigned long long));
	sa.keys = (unsigned long long *)d_keys;
	sa.best_models = d_best_models;
	sa.best_scores = (double *)(base + o_best_scores);
	sa.best_counts = (uint32_t *)(base + o_best_counts);
	sa.winners = d_winners;
	sa.counters = ctx->d_counters;
	sa.n = ctx->n;
	sa.k = (int)K;
	sa.accumulate = accumulate;
	sa.index_base = index_base;
	sa.thr = threshold;
	sa.margin = scoring == GCR_MSAC ? SEL_MARGIN_MSAC : SEL_MARGIN_MAGSAC;
	const dim3 grid(cx, ctx->n_pairs);
#define GCR_SEL(PR)                                                                                          \
	do                                                                                                       \
	{                                                                                                        \
		if (scoring == GCR_MSAC)                                                                             \
			k_select_exact<PR, GCR_MSAC><<<grid, sel_threads, 0, ctx->stream>>>(sa);                          \
		else                                                                                                 \
			k_select_exact<PR, GCR_MAGSAC><<<grid, sel_threads, 0, ctx->stream>>>(sa);                        \
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

extern "C" gcr_status gcr_round(gcr_ctx *ctx, int problem, int scoring, int sampler, uint64_t first_index, int count,
                                uint64_t seed, double threshold, int flags, uint32_t index_base, uint64_t *d_keys,
                                double *d_best_models, uint64_t *evals_out)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!d_keys)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_round: d_keys is required");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	return round_impl(ctx, problem, scoring, sampler, first_index, count, seed, threshold, flags, (uint64_t)index_base, d_keys,
	                  d_best_models, nullptr, 0, evals_out);
}

extern "C" gcr_status gcr_round_winners(gcr_ctx *ctx, int problem, int scoring, int sampler, uint64_t first_index, int count,
                                        uint64_t seed, double threshold, int flags, uint64_t index_base, int accumulate,
                                        gcr_winner *d_winners, uint64_t *evals_out)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!d_winners)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_round_winners: d_winners is required");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	return round_impl(ctx, problem, scoring, sampler, first_index, count, seed, threshold, flags, index_base, nullptr, nullptr,
	                  d_winners, accumulate, evals_out);
}

extern "C" gcr_status gcr_round_results(gcr_ctx *ctx, double *h_best_scores, uint32_t *h_best_counts)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!ctx->d_round_best_scores)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_round_results: no round has run on this context");
	gcr_status st = GCR_OK;
	if (h_best_scores)
		st = gcr_memcpy_d2h_async(ctx, h_best_scores, ctx->d_round_best_scores, (size_t)ctx->n_pairs * sizeof(double));
	if (!st && h_best_counts)
		st = gcr_memcpy_d2h_async(ctx, h_best_counts, ctx->d_round_best_counts, (size_t)ctx->n_pairs * sizeof(uint32_t));
	const gcr_status sync_st = gcr_sync(ctx);
	return st ? st : sync_st;
}
