// gcr_internal.cuh -- shared declarations of the sm_100a kernels behind include/gcr/gcr.h.
// Everything is compiled with -fmad=false: FP64 code then keeps the reference's separately-rounded mul/add
// (bit parity with the CPU build, which has no FMA contraction), and the FP32 hot loops state their FMAs
// explicitly (fmaf / __ffma2_rn).
#pragma once
#include <utility>
#include <vector>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "gcr/gcr.h"

struct gcr_ctx
{
	int device = 0;
	cudaStream_t stream = nullptr;
	cudaStream_t own_stream = nullptr;
	char err[512] = {0};
	int sm_count = 0;
	int sm_clock_khz = 0;
	size_t l2_bytes = 0;
	uint64_t launches = 0;

	// points
	int n = 0, cols = 0, n_pairs = 0;
	const double *d_pts64 = nullptr; // [n_pairs][n][cols]
	double *d_pts64_owned = nullptr;
	size_t pts64_capacity = 0;
	float *d_upload32 = nullptr; // staging of gcr_points_upload_f32
	size_t upload32_capacity = 0;
	// [n_pairs][n_pad][2] float4: (x1 x1 y1 y1)(x2 x2 y2 y2) operand-duplicated for the packed f32x2 scoring path
	// (cols == 4), or (u v 0 0)(X Y Z 0) (cols == 5)
	float4 *d_pts32 = nullptr;
	size_t pts32_capacity = 0;
	int n_pad = 0;              // n rounded up to the tile granule; padding points can never be inliers
	double *d_norm = nullptr;  // [n_pairs][8]
	unsigned long long *d_bbox = nullptr; // [n_pairs][cols][2] ordered-u64 min/max
	size_t norm_capacity = 0;

	// MAGSAC weight table (FP32): (Gamma(1.5, k/1000) - gamma_k) / (Gamma(1.5) - gamma_k), k = 0..5000
	float *d_magsac_lut = nullptr;
	double *d_magsac_lut64 = nullptr; // Gamma(1.5, k/1000), k = 0..5000 (FP64 exact path)

	// PROSAC growth function T'_n (prosac_sampler.h:129-160) for the current (n, m)
	unsigned long long *d_prosac_growth = nullptr;
	int prosac_n = 0, prosac_m = 0;
	// cumulative table of the importance sampler's std::discrete_distribution (importance_sampler.h:76-78), n entries
	double *d_importance_cdf = nullptr;
	int importance_n = 0;

	// per-round buffers of gcr_round (samples, models, valid flags, scores)
	void *d_round = nullptr;
	size_t round_capacity = 0;

	// FP64 score / exact inlier count of each pair's winner, written by the last gcr_round (inside d_round)
	double *d_round_best_scores = nullptr;
	uint32_t *d_round_best_counts = nullptr;

	// counters (device, u64): [0] valid models produced by K2, [1] FP64 re-check entries dropped (work list full),
	// [2] FP64 re-check entries queued, [3] K4 FP64 re-scores, [4] max relative FP32 score error seen by K4 (float bits)
	// -- since the last reset
	unsigned long long *d_counters = nullptr;
	// optional per-launch timing of the scoring kernel (bench.py's roofline leg)
	bool profile = false;
	static const int EV_RING = 128;
	cudaEvent_t ev_a[EV_RING] = {nullptr}, ev_b[EV_RING] = {nullptr};
	int ev_head = 0, ev_pending = 0; // ring of recorded-but-unread event pairs
	double score_ms = 0.0;
	uint64_t score_launches = 0;
	uint64_t score_evals = 0;

	unsigned long long *d_score_trace = nullptr; // gcr_set_score_trace: per-CTA timeline of the next K3 launches
	// workspace of gcr_score (compact model list, accumulators, FP64 work list)
	void *d_score_ws = nullptr;
	size_t score_ws_capacity = 0;
	void *d_score_amb = nullptr; // the FP64 re-check list of gcr_score (fixed capacity: allocated once, never with the workspace)

	// scratch
	void *d_scratch = nullptr;
	size_t scratch_capacity = 0;
	void *h_pinned = nullptr;
	size_t pinned_capacity = 0;

	// gcr_malloc / gcr_free: size-bucketed cache of device blocks owned by this context.  cudaMalloc / cudaFree
	// synchronise the whole device and serialise every host thread, so a driver that allocates its ~20 buffers per
	// image pair would stall all other contexts; freed blocks are kept and handed out again (reuse is ordered by
	// this context's stream, the only stream the block was used on).
	struct CachedBlock { void *p; size_t cap; bool used; };
	std::vector<CachedBlock> blocks;
	// small host <-> device copies go through this pinned bounce buffer (pageable cudaMemcpyAsync takes the driver's
	// global staging path); separate from h_pinned, which kernels' result read-backs use
	void *h_stage = nullptr;
	size_t stage_off = 0; // next free byte of the host -> device ring (first half of h_stage)
	// device -> host copies queued by gcr_memcpy_d2h_async: they land in the second half of h_stage and are handed to
	// their destinations by the next gcr_sync / gcr_memcpy_d2h (one stream synchronise for all of them)
	struct PendingD2H { size_t off; void *dst; size_t bytes; };
	std::vector<PendingD2H> pending_d2h;
	size_t d2h_off = 0;
	static const size_t STAGE_BYTES = 4 << 20;

	// gcr_set_active_pairs: the pairs the next sampling / solving calls work on (null = all)
	const uint8_t *d_active_pairs = nullptr;
	// gcr_set_pair_streams: per-pair additions to gcr_sample's (first_index, seed) (null = none)
	const uint64_t *d_pair_first = nullptr, *d_pair_seed = nullptr;

	// hypothesis-sharded mode (gcr_comm.cu): NCCL communicator of the ranks that share one problem
	void *nccl_comm = nullptr;
	int comm_rank = 0, comm_world = 1;
	gcr_winner *d_gather = nullptr; // [world][n_pairs] all-gather buffer (this rank's record accumulates in place)
	size_t gather_capacity = 0;     // in records

	double band_factor = 16.0; // FP32 error band in units of eps32 (normalised coordinates), see k_score.cu
	int score_variant = -1;    // -1 auto; else forces a kernel variant (bench / tests)
	std::vector<std::pair<const void *, int>> k3_occupancy; // scoring kernel -> resident CTAs per SM (queried once)
	int tune_k3_ctas = 0;      // GCR_TUNE_K3_CTAS_PER_SM
	int tune_round_prefetch = 1; // GCR_TUNE_ROUND_PREFETCH
	int tune_carveout = -1;      // GCR_TUNE_SMEM_CARVEOUT
};

#define GCR_CHECK_CUDA(ctx, expr)                                                                              \
	do                                                                                                         \
	{                                                                                                          \
		cudaError_t _e = (expr);                                                                               \
		if (_e != cudaSuccess)                                                                                 \
		{                                                                                                      \
			snprintf((ctx)->err, sizeof((ctx)->err), "%s:%d: %s -> %s", __FILE__, __LINE__, #expr,            \
			         cudaGetErrorString(_e));                                                                  \
			return GCR_ERR_CUDA;                                                                               \
		}                                                                                                      \
	} while (0)

#define GCR_FAIL(ctx, code, ...)                                                                               \
	do                                                                                                         \
	{                                                                                                          \
		snprintf((ctx)->err, sizeof((ctx)->err), __VA_ARGS__);                                                 \
		return (code);                                                                                         \
	} while (0)

#define GCR_LAUNCHED(ctx)                                                                                      \
	do                                                                                                         \
	{                                                                                                          \
		++(ctx)->launches;                                                                                     \
		GCR_CHECK_CUDA(ctx, cudaGetLastError());                                                               \
	} while (0)

gcr_status gcr_ensure_scratch(gcr_ctx *ctx, size_t bytes);
gcr_status gcr_ensure_pinned(gcr_ctx *ctx, size_t bytes);
gcr_status gcr_profile_begin(gcr_ctx *ctx);
gcr_status gcr_profile_end(gcr_ctx *ctx, uint64_t evals);
gcr_status gcr_ensure_prosac_growth(gcr_ctx *ctx, int m);
gcr_status gcr_ensure_magsac_lut64(gcr_ctx *ctx);
// gcr_score with K4's first stage fused into the finalize kernel: d_pairbest [n_pairs] (zeroed by the caller, may be
// null) receives each pair's best FP32 key (orderedBits(score) << 32 | ~slot)
gcr_status gcr_score_fused(gcr_ctx *ctx, int problem, int scoring, const double *d_models, const uint8_t *d_valid, int k,
                           double threshold, int flags, float *d_scores, uint32_t *d_inliers, unsigned long long *d_pairbest);
// K2 with the sample drawn inside the solver kernel (gcr_round's fused K1 + K2)
gcr_status gcr_solve_minimal_generated(gcr_ctx *ctx, int problem, int sampler, uint64_t first_index, int count, uint64_t seed,
                                       double *d_models, uint8_t *d_valid);

// K6 for GCR_PNP (k_refit_pnp.cu): normalised DLT + the LM of refine_pnp, one CTA per index set
gcr_status gcr_refit_pnp_launch(gcr_ctx *ctx, int pair, const uint32_t *d_index_sets, const uint32_t *d_set_offsets,
                                const uint32_t *d_set_sizes, int count, const double *d_weights, int weight_mode, double *d_models,
                                uint8_t *d_ok);

static inline int gcr_cols_of(int problem) { return problem == GCR_PNP ? 5 : 4; }

// ---- device helpers -------------------------------------------------------------------------------------------
#ifdef __CUDACC__

// Philox4x32-10 (same constants / round function as oracle gcro_philox4x32)
__device__ __forceinline__ void gcr_philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1, uint32_t out[4])
{
#pragma unroll
	for (int r = 0; r < 10; ++r)
	{
		const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
		const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
		const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
		c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct GcrDrawStream
{
	uint32_t c0, c1, block, stream, k0, k1, buf[4];
	int pos;
	__device__ GcrDrawStream(uint64_t seed, uint64_t index, uint32_t stream_)
	    : c0((uint32_t)index), c1((uint32_t)(index >> 32)), block(0), stream(stream_), k0((uint32_t)seed),
	      k1((uint32_t)(seed >> 32)), pos(4)
	{
	}
	__device__ uint32_t next()
	{
		if (pos == 4)
		{
			gcr_philox4x32(c0, c1, block, stream, k0, k1, buf);
			++block;
			pos = 0;
		}
		// avoid dynamic register indexing
		const uint32_t v = pos == 0 ? buf[0] : pos == 1 ? buf[1] : pos == 2 ? buf[2] : buf[3];
		++pos;
		return v;
	}
	__device__ uint32_t bounded(uint32_t range)
	{
		for (;;)
		{
			const uint32_t u = next();
			const uint32_t lo = u * range;
			if (lo < range)
			{
				const uint32_t thresh = (0u - range) % range;
				if (lo < thresh)
					continue;
			}
			return __umulhi(u, range);
		}
	}
	// m <= 8 distinct values in [0, range) by draw-and-reject, kept in draw order
	template <int MAXM>
	__device__ void unique_set(uint32_t *out, int count, uint32_t range)
	{
#pragma unroll
		for (int i = 0; i < MAXM; ++i)
		{
			if (i < count)
			{
				bool dup;
				uint32_t v;
				do
				{
					v = bounded(range);
					dup = false;
#pragma unroll
					for (int j = 0; j < MAXM; ++j)
						if (j < i && out[j] == v)
							dup = true;
				} while (dup);
				out[i] = v;
			}
		}
	}
};


// ---- minimal-sample generation (shared by K1's kernels and by K2's fused "generate in registers" path) --------
#define GCR_PAIR_SEED_STRIDE 0x9E3779B97F4A7C15ull

// where a solver kernel takes its sample from: an explicit index array, or the counter-based stream itself
struct GcrSampleSource
{
	const uint32_t *samples;            // [n_pairs][count][m], or nullptr = generate
	unsigned long long seed, first_index;
	int sampler;                        // GCR_UNIFORM / GCR_PROSAC / GCR_IMPORTANCE
	const unsigned long long *growth;   // PROSAC growth function
	const double *cdf;                  // importance sampler: cumulative probabilities (n entries)
	uint32_t n;                         // pool size = number of points
	const uint8_t *active;              // optional [n_pairs]: pairs with 0 are skipped (their valid flags are zeroed by the launcher)
};

// closed form of the subset-size recurrence of prosac_sampler.h:65-88 (see oracle gcro_prosac_subset_size)
__device__ __forceinline__ uint32_t gcr_prosac_subset_size(const unsigned long long *__restrict__ growth, uint32_t n, int m,
                                                           unsigned long long k)
{
	uint32_t lo = (uint32_t)m, hi = n;
	if (lo >= hi)
		return n < (uint32_t)m ? (uint32_t)m : hi;
	while (lo < hi)
	{
		const uint32_t mid = lo + (hi - lo) / 2;
		if (growth[mid - 1] >= k)
			hi = mid;
		else
			lo = mid + 1;
	}
	return lo;
}

// sample `index` of pair `pair`: v[0..m) (m <= 8)
__device__ __forceinline__ void gcr_gen_uniform(unsigned long long seed, int pair, unsigned long long index, int m, uint32_t n,
                                                uint32_t *v)
{
	GcrDrawStream s(seed + (unsigned long long)pair * GCR_PAIR_SEED_STRIDE, index, 0u);
	s.unique_set<8>(v, m, n);
}
__device__ __forceinline__ void gcr_gen_prosac(unsigned long long seed, int pair, unsigned long long k, int m, uint32_t n,
                                               const unsigned long long *__restrict__ growth, unsigned long long convergence,
                                               uint32_t *v)
{
	GcrDrawStream s(seed + (unsigned long long)pair * GCR_PAIR_SEED_STRIDE, k, 1u);
	if (k > convergence)
		s.unique_set<8>(v, m, n);
	else
	{
		const uint32_t subset = gcr_prosac_subset_size(growth, n, m, k);
		const uint32_t range = (k == 1) ? (uint32_t)m : subset - 1; // first-draw quirk, prosac_sampler.h:171-172
		s.unique_set<8>(v, m - 1, range);
#pragma unroll
		for (int j = 0; j < 8; ++j)
			if (j == m - 1)
				v[j] = subset - 1;
	}
}
// importance_sampler.h:134-135: m independent draws of std::discrete_distribution<int> (libstdc++: p =
// generate_canonical<double, 53>(urng) = (w0 + w1 * 2^32) / 2^64 in FP64, clamped below 1; index = lower_bound(cdf, p)).
// With replacement: a sample may contain a point twice, as in the reference.
__device__ __forceinline__ void gcr_gen_importance(unsigned long long seed, int pair, unsigned long long index, int m, uint32_t n,
                                                   const double *__restrict__ cdf, uint32_t *v)
{
	GcrDrawStream s(seed + (unsigned long long)pair * GCR_PAIR_SEED_STRIDE, index, 3u);
#pragma unroll
	for (int j = 0; j < 8; ++j)
	{
		if (j < m)
		{
			const uint32_t w0 = s.next(), w1 = s.next();
			double p = ((double)w0 + (double)w1 * 4294967296.0) / 18446744073709551616.0;
			if (p >= 1.0)
				p = 0.99999999999999988898; // nextafter(1, 0)
			uint32_t lo = 0, hi = n; // first i with cdf[i] >= p (cdf[n-1] = 1 > p)
			while (lo < hi)
			{
				const uint32_t mid = lo + (hi - lo) / 2;
				if (cdf[mid] < p)
					lo = mid + 1;
				else
					hi = mid;
			}
			v[j] = lo;
		}
	}
}
// the M indices of sample i of pair `pair` for a solver kernel
template <int M>
__device__ __forceinline__ void gcr_load_sample(const GcrSampleSource &src, int pair, int count, int i, uint32_t *sidx)
{
	if (src.samples)
	{
		const uint32_t *s = src.samples + ((size_t)pair * count + i) * M;
#pragma unroll
		for (int j = 0; j < M; ++j)
			sidx[j] = s[j];
		return;
	}
	uint32_t v[8];
	if (src.sampler == GCR_PROSAC)
		gcr_gen_prosac(src.seed, pair, src.first_index + (unsigned long long)i, M, src.n, src.growth, 100000ull, v);
	else if (src.sampler == GCR_IMPORTANCE)
		gcr_gen_importance(src.seed, pair, src.first_index + (unsigned long long)i, M, src.n, src.cdf, v);
	else
		gcr_gen_uniform(src.seed, pair, src.first_index + (unsigned long long)i, M, src.n, v);
#pragma unroll
	for (int j = 0; j < M; ++j)
		sidx[j] = v[j];
}

__device__ __forceinline__ unsigned long long gcr_ordered_u64(double x)
{
	const long long b = __double_as_longlong(x);
	return (unsigned long long)(b ^ ((b >> 63) | (long long)0x8000000000000000ull));
}
__device__ __forceinline__ double gcr_from_ordered_u64(unsigned long long k)
{
	const long long b = (k & 0x8000000000000000ull) ? (long long)(k ^ 0x8000000000000000ull) : (long long)~k;
	return __longlong_as_double(b);
}

// monotone map float -> uint32 (larger float => larger uint)
__host__ __device__ __forceinline__ uint32_t gcr_ordered_u32(float f)
{
#ifdef __CUDA_ARCH__
	const uint32_t b = __float_as_uint(f);
#else
	union { float f; uint32_t u; } cvt; cvt.f = f; const uint32_t b = cvt.u;
#endif
	return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float gcr_from_ordered_u32(uint32_t k)
{
	const uint32_t b = (k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k;
#ifdef __CUDA_ARCH__
	return __uint_as_float(b);
#else
	union { float f; uint32_t u; } cvt; cvt.u = b; return cvt.f;
#endif
}

// ---- FP64 residuals in the reference's operation order (no contraction: file is built with -fmad=false) -------
__device__ __forceinline__ double gcr_sqres_h64(const double *s, const double *H)
{
	const double x1 = s[0], y1 = s[1], x2 = s[2], y2 = s[3];
	const double t1 = H[0] * x1 + H[1] * y1 + H[2];
	const double t2 = H[3] * x1 + H[4] * y1 + H[5];
	const double t3 = H[6] * x1 + H[7] * y1 + H[8];
	const double d1 = x2 - (t1 / t3);
	const double d2 = y2 - (t2 / t3);
	return d1 * d1 + d2 * d2;
}
__device__ __forceinline__ double gcr_sqres_sampson64(const double *s, const double *F)
{
	const double x1 = s[0], y1 = s[1], x2 = s[2], y2 = s[3];
	const double rxc = F[0] * x2 + F[3] * y2 + F[6];
	const double ryc = F[1] * x2 + F[4] * y2 + F[7];
	const double rwc = F[2] * x2 + F[5] * y2 + F[8];
	const double r = (x1 * rxc + y1 * ryc + rwc);
	const double rx = F[0] * x1 + F[1] * y1 + F[2];
	const double ry = F[3] * x1 + F[4] * y1 + F[5];
	return r * r / (rxc * rxc + ryc * ryc + rx * rx + ry * ry);
}
__device__ __forceinline__ double gcr_sqres_pnp64(const double *s, const double *P)
{
	const double u = s[0], v = s[1], x = s[2], y = s[3], z = s[4];
	const double px = P[0] * x + P[1] * y + P[2] * z + P[3];
	const double py = P[4] * x + P[5] * y + P[6] * z + P[7];
	const double pz = P[8] * x + P[9] * y + P[10] * z + P[11];
	const double pu = px / pz, pv = py / pz;
	const double du = pu - u, dv = pv - v;
	return du * du + dv * dv;
}
template <int PROBLEM>
__device__ __forceinline__ double gcr_sqres64(const double *pt, const double *model)
{
	if (PROBLEM == GCR_H)
		return gcr_sqres_h64(pt, model);
	if (PROBLEM == GCR_PNP)
		return gcr_sqres_pnp64(pt, model);
	return gcr_sqres_sampson64(pt, model);
}

#endif // __CUDACC__
