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
extern "C" gcr_status gcr_set_score_variant(gcr_ctx *ctx, int v)
{
	if (!ctx)
		return GCR_ERR_ARG;
	ctx->score_variant = v;
	return GCR_OK;
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

	gcr_status st = gcr_sample(ctx, sampler, m, first_index, count, seed, d_samples);
	if (st) return st;
	st = gcr_solve_minimal(ctx, problem, d_samples, count, d_models, d_valid);
	if (st) return st;
	st = gcr_score(ctx, problem, scoring, d_models, d_valid, (int)K, threshold, flags, d_scores, nullptr);
	if (st) return st;
	st = gcr_argmax(ctx, d_scores, (int)K, index_base, d_keys);
	if (st) return st;
	if (d_best_models)
	{
		k_gather_best<<<(ctx->n_pairs + 127) / 128, 128, 0, ctx->stream>>>((const unsigned long long *)d_keys, index_base,
		                                                                 d_models, (int)K, ml, d_best_models, ctx->n_pairs);
		GCR_LAUNCHED(ctx);
	}
	if (evals_out)
		*evals_out = (uint64_t)P * K * (uint64_t)ctx->n;
	return GCR_OK;
}
