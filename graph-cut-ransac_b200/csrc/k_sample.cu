// k_sample.cu -- K1: minimal-sample index sets from a counter-based generator.
//
// Replaces UniformSampler::sample (samplers/uniform_sampler.h:81-99) with
// UniformRandomGenerator::generateUniqueRandomSet (uniform_random_generator.h:90-103) and ProsacSampler::sample /
// incrementIterationNumber (samplers/prosac_sampler.h:65-88, 207-236).  The reference draws from a std::mt19937
// seeded by std::random_device (not reproducible, inherently sequential); here sample i is a pure function of
// (seed, i, pair) through Philox4x32-10, so any range of the sample stream can be generated in parallel, on any
// GPU, and replayed on the CPU by the oracle (gcro_sample_uniform / gcro_sample_prosac) bit for bit.
// One thread = one sample; integer work only.
#include "gcr_internal.cuh"

#include <math.h>
#include <cmath>
#include <vector>


// ps: optional per-pair additions to (first_index, seed) -- gcr_set_pair_streams, the batch engine's seats
struct PairStreams
{
	const uint64_t *first, *seed;
	const uint8_t *active;
};

__global__ void __launch_bounds__(256) k_sample_uniform(uint32_t *__restrict__ out, int count, int m, uint32_t n,
                                                         uint64_t first_index, uint64_t seed, const PairStreams ps)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	const int pair = blockIdx.y;
	if (ps.active && !ps.active[pair])
		return;
	if (ps.first) first_index += ps.first[pair];
	if (ps.seed) seed += ps.seed[pair];
	uint32_t v[8];
	gcr_gen_uniform(seed, pair, first_index + (uint64_t)i, m, n, v);
	uint32_t *o = out + ((size_t)pair * count + i) * m;
#pragma unroll
	for (int j = 0; j < 8; ++j)
		if (j < m)
			o[j] = v[j];
}

__global__ void __launch_bounds__(256) k_sample_prosac(uint32_t *__restrict__ out, int count, int m, uint32_t n,
                                                        uint64_t first_k, uint64_t seed,
                                                        const unsigned long long *__restrict__ growth,
                                                        unsigned long long convergence, const PairStreams ps)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	const int pair = blockIdx.y;
	if (ps.active && !ps.active[pair])
		return;
	if (ps.first) first_k += ps.first[pair];
	if (ps.seed) seed += ps.seed[pair];
	const unsigned long long k = first_k + (unsigned long long)i;
	uint32_t v[8];
	gcr_gen_prosac(seed, pair, k, m, n, growth, convergence, v);
	uint32_t *o = out + ((size_t)pair * count + i) * m;
#pragma unroll
	for (int j = 0; j < 8; ++j)
		if (j < m)
			o[j] = v[j];
}

// importance_sampler.h:126-137 (NG-RANSAC): m independent draws from the point-probability distribution
__global__ void __launch_bounds__(256) k_sample_importance(uint32_t *__restrict__ out, int count, int m, uint32_t n,
                                                            uint64_t first_index, uint64_t seed,
                                                            const double *__restrict__ cdf, const PairStreams ps)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	const int pair = blockIdx.y;
	if (ps.active && !ps.active[pair])
		return;
	if (ps.first) first_index += ps.first[pair];
	if (ps.seed) seed += ps.seed[pair];
	uint32_t v[8];
	gcr_gen_importance(seed, pair, first_index + (uint64_t)i, m, n, cdf, v);
	uint32_t *o = out + ((size_t)pair * count + i) * m;
#pragma unroll
	for (int j = 0; j < 8; ++j)
		if (j < m)
			o[j] = v[j];
}

// GCRANSAC.h:842-846: uniform draws of m (<= 49) distinct positions in a pool, mapped through the pool
__global__ void __launch_bounds__(128) k_sample_pool(const uint32_t *__restrict__ pool, uint32_t pool_size, int m,
                                                      uint64_t first_index, int count, uint64_t seed, uint32_t stream,
                                                      uint32_t *__restrict__ out)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= count)
		return;
	GcrDrawStream s(seed, first_index + (uint64_t)i, stream);
	uint32_t *o = out + (size_t)i * m;
	// draw-and-reject with the already chosen positions kept in the output row (positions first, mapped afterwards)
	for (int j = 0; j < m; ++j)
	{
		bool dup;
		uint32_t v;
		do
		{
			v = s.bounded(pool_size);
			dup = false;
			for (int q = 0; q < j; ++q)
				if (o[q] == v)
					dup = true;
		} while (dup);
		o[j] = v;
	}
	for (int j = 0; j < m; ++j)
		o[j] = pool[o[j]];
}

// the batch engine's LO sampler: request r draws `trials` samples of desc[r].sample_size positions from ITS pool
// (pools + desc[r].pool_offset, desc[r].pool_size entries) on ITS stream (seed, first_index, stream id), exactly the draws
// k_sample_pool makes for that request alone; sample t of request r goes to out + desc[r].out_offset + t * sample_size
__global__ void __launch_bounds__(128) k_sample_pools(const uint32_t *__restrict__ pools, const gcr_pool_request *__restrict__ desc,
                                                       int trials, uint32_t *__restrict__ out)
{
	const gcr_pool_request d = desc[blockIdx.y];
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= trials)
		return;
	const uint32_t *pool = pools + d.pool_offset;
	const int m = (int)d.sample_size;
	GcrDrawStream s(d.seed, d.first_index + (uint64_t)i, d.stream);
	uint32_t *o = out + d.out_offset + (size_t)i * m;
	for (int j = 0; j < m; ++j)
	{
		bool dup;
		uint32_t v;
		do
		{
			v = s.bounded(d.pool_size);
			dup = false;
			for (int q = 0; q < j; ++q)
				if (o[q] == v)
					dup = true;
		} while (dup);
		o[j] = v;
	}
	for (int j = 0; j < m; ++j)
		o[j] = pool[o[j]];
}

extern "C" gcr_status gcr_sample_from_pools(gcr_ctx *ctx, const uint32_t *d_pools, const gcr_pool_request *d_requests, int n_requests,
                                            int trials, uint32_t *d_out)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!d_pools || !d_requests || !d_out || n_requests <= 0 || trials <= 0)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_sample_from_pools: bad arguments");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	k_sample_pools<<<dim3((trials + 127) / 128, n_requests), 128, 0, ctx->stream>>>(d_pools, d_requests, trials, d_out);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}

extern "C" gcr_status gcr_set_active_pairs(gcr_ctx *ctx, const uint8_t *d_active)
{
	if (!ctx)
		return GCR_ERR_ARG;
	ctx->d_active_pairs = d_active;
	return GCR_OK;
}

extern "C" gcr_status gcr_set_pair_streams(gcr_ctx *ctx, const uint64_t *d_first_offsets, const uint64_t *d_seed_offsets)
{
	if (!ctx)
		return GCR_ERR_ARG;
	ctx->d_pair_first = d_first_offsets;
	ctx->d_pair_seed = d_seed_offsets;
	return GCR_OK;
}

// prosac_sampler.h:129-160 growth function, FP64 recurrence on the host exactly as the reference runs it
static void prosac_growth_host(uint32_t n, int m, uint64_t convergence_iters, std::vector<unsigned long long> &growth)
{
	const size_t sample_size = (size_t)m, point_number = n;
	growth.assign(n, 0);
	double T_n = (double)convergence_iters;
	for (size_t i = 0; i < sample_size; i++)
		T_n *= static_cast<double>(sample_size - i) / (point_number - i);
	size_t T_n_prime = 1;
	for (size_t i = 0; i < point_number; ++i)
	{
		if (i + 1 <= sample_size)
		{
			growth[i] = T_n_prime;
			continue;
		}
		double Tn_plus1 = static_cast<double>(i + 1) * T_n / (i + 1 - sample_size);
		growth[i] = T_n_prime + (unsigned int)ceil(Tn_plus1 - T_n);
		T_n = Tn_plus1;
		T_n_prime = growth[i];
	}
}

gcr_status gcr_ensure_prosac_growth(gcr_ctx *ctx, int m)
{
	if (ctx->d_prosac_growth && ctx->prosac_n == ctx->n && ctx->prosac_m == m)
		return GCR_OK;
	std::vector<unsigned long long> growth;
	prosac_growth_host((uint32_t)ctx->n, m, 100000, growth);
	GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	cudaFree(ctx->d_prosac_growth);
	ctx->d_prosac_growth = nullptr;
	GCR_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_prosac_growth, growth.size() * sizeof(unsigned long long)));
	GCR_CHECK_CUDA(ctx, cudaMemcpy(ctx->d_prosac_growth, growth.data(), growth.size() * sizeof(unsigned long long),
	                               cudaMemcpyHostToDevice));
	ctx->prosac_n = ctx->n;
	ctx->prosac_m = m;
	return GCR_OK;
}

extern "C" gcr_status gcr_sample(gcr_ctx *ctx, int sampler, int m, uint64_t first_index, int count, uint64_t seed,
                                 uint32_t *d_out)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_sample: no points uploaded (the pool is [0, n))");
	if (m < 1 || m > 8 || count <= 0 || !d_out)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_sample: need 1 <= m <= 8, count > 0");
	if (m > ctx->n)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_sample: sample size %d exceeds the %d points (uniform_sampler.h:87-88)", m, ctx->n);
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	const dim3 grid((count + 255) / 256, ctx->n_pairs);
	PairStreams ps;
	ps.first = ctx->d_pair_first;
	ps.seed = ctx->d_pair_seed;
	ps.active = ctx->d_active_pairs;
	if (sampler == GCR_UNIFORM)
		k_sample_uniform<<<grid, 256, 0, ctx->stream>>>(d_out, count, m, (uint32_t)ctx->n, first_index, seed, ps);
	else if (sampler == GCR_PROSAC)
	{
		if (first_index < 1 && !ps.first)
			GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_sample: PROSAC draws are numbered from k = 1");
		gcr_status st = gcr_ensure_prosac_growth(ctx, m);
		if (st)
			return st;
		k_sample_prosac<<<grid, 256, 0, ctx->stream>>>(d_out, count, m, (uint32_t)ctx->n, first_index, seed,
		                                                ctx->d_prosac_growth, 100000ull, ps);
	}
	else if (sampler == GCR_IMPORTANCE)
	{
		if (!ctx->d_importance_cdf || ctx->importance_n != ctx->n)
			GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_sample: GCR_IMPORTANCE needs gcr_set_point_probabilities for the %d points", ctx->n);
		k_sample_importance<<<grid, 256, 0, ctx->stream>>>(d_out, count, m, (uint32_t)ctx->n, first_index, seed,
		                                                    ctx->d_importance_cdf, ps);
	}
	else
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_sample: unknown sampler %d", sampler);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}

// The table std::discrete_distribution<int>::param_type::_M_initialize builds from the weights (libstdc++
// bits/random.tcc, the library the reference's ImportanceSampler is compiled against, importance_sampler.h:76-78):
// sum = accumulate(w), p_i = w_i / sum, cp = partial_sum(p), cp.back() = 1.  Sequential FP64 on the host, in that
// order, so that the device draws select exactly the indices the reference's distribution object would.
extern "C" gcr_status gcr_set_point_probabilities(gcr_ctx *ctx, const double *h_probabilities, int n)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!h_probabilities || n < 1)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_set_point_probabilities: need n >= 1 probabilities");
	if (n != ctx->n)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_set_point_probabilities: %d probabilities for %d points (importance_sampler.h:72-73)", n, ctx->n);
	double sum = 0.0;
	for (int i = 0; i < n; ++i)
	{
		if (!(h_probabilities[i] >= 0.0) || !std::isfinite(h_probabilities[i]))
			GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_set_point_probabilities: probability %d is negative or not finite", i);
		sum += h_probabilities[i];
	}
	if (!(sum > 0.0))
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_set_point_probabilities: the probabilities sum to zero");
	std::vector<double> cp((size_t)n);
	double run = 0.0;
	for (int i = 0; i < n; ++i)
	{
		run = (i == 0) ? h_probabilities[0] / sum : run + h_probabilities[i] / sum;
		cp[i] = run;
	}
	cp[n - 1] = 1.0;
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	if (ctx->importance_n != n)
	{
		GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // kernels still reading the old table
		cudaFree(ctx->d_importance_cdf);
		ctx->d_importance_cdf = nullptr;
		ctx->importance_n = 0;
		GCR_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_importance_cdf, (size_t)n * sizeof(double)));
	}
	// stream-ordered, through the context's pinned ring (a pageable copy would serialise the host threads of a
	// multi-context process)
	const gcr_status st = gcr_memcpy_h2d(ctx, ctx->d_importance_cdf, cp.data(), (size_t)n * sizeof(double));
	if (st != GCR_OK)
		return st;
	ctx->importance_n = n;
	return GCR_OK;
}

extern "C" gcr_status gcr_sample_from_pool(gcr_ctx *ctx, const uint32_t *d_pool, uint32_t pool_size, int m,
                                           uint64_t first_index, int count, uint64_t seed, uint32_t stream,
                                           uint32_t *d_out)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!d_pool || !d_out || m < 1 || count <= 0 || (uint32_t)m > pool_size)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_sample_from_pool: need 1 <= m <= pool_size, count > 0");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	k_sample_pool<<<(count + 127) / 128, 128, 0, ctx->stream>>>(d_pool, pool_size, m, first_index, count, seed, stream, d_out);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}
