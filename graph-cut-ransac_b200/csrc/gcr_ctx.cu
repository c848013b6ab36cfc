// gcr_ctx.cu -- context, device-memory helpers, point upload and the K0 "prepare points" kernels.
//
// Data layout in HBM (per context):
//   d_pts64   [n_pairs][n][cols]        FP64 row-major, exactly what the caller handed over (the reference's cv::Mat)
//   d_pts32   [n_pairs][n_pad][2]       float4 (x1' x1' y1' y1')(x2' x2' y2' y2'): centred + power-of-two scaled FP32
//                                       copy, operand-duplicated for the packed f32x2 scoring kernel
//             (PnP: (u' v' 0 0)(X' Y' Z' 0))
//   d_norm    [n_pairs][8]              FP64 normalisation (centres, scale) so models can be mapped alike
// n_pad = n rounded up to GCR_TILE_POINTS; padding rows are NaN so they can never be inliers.
#include "gcr_internal.cuh"

#include <math.h>
#include <new>
#include <string.h>

#define GCR_TILE_POINTS 256

extern "C" const char *gcr_version(void) { return "gcr-b200 0.1 (sm_100a)"; }

extern "C" gcr_status gcr_ctx_create(int device, gcr_ctx **out)
{
	if (!out)
		return GCR_ERR_ARG;
	*out = nullptr;
	int count = 0;
	cudaError_t e = cudaGetDeviceCount(&count);
	if (e != cudaSuccess || count == 0 || device < 0 || device >= count)
	{
		fprintf(stderr, "gcr_ctx_create: no usable CUDA device %d (%s); this library has no CPU fallback\n", device,
		        e == cudaSuccess ? "device index out of range" : cudaGetErrorString(e));
		return GCR_ERR_CUDA;
	}
	gcr_ctx *ctx = new (std::nothrow) gcr_ctx();
	if (!ctx)
		return GCR_ERR_NOMEM;
	ctx->device = device;
	// the context's stream runs at the HIGHEST priority, its side stream (gcr_round's look-ahead solver) at the lowest: the
	// block scheduler hands out a later grid's CTAs only after every CTA of the earlier grids of the same priority has been
	// placed, so a many-CTA solver grid in flight would otherwise hold back the small kernels of the scoring pipeline
	int prio_least = 0, prio_greatest = 0;
	if (cudaSetDevice(device) == cudaSuccess)
		cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
	if (const char *pe = getenv("GCR_STREAM_PRIORITY")) // experiments (tools/overlap_probe.py): 0 = least, 1 = greatest
		if (atoi(pe) == 0) prio_greatest = prio_least;
	if (cudaSetDevice(device) != cudaSuccess ||
	    cudaStreamCreateWithPriority(&ctx->own_stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess)
	{
		delete ctx;
		return GCR_ERR_CUDA;
	}
	ctx->stream = ctx->own_stream;
	cudaDeviceProp prop;
	cudaGetDeviceProperties(&prop, device);
	ctx->sm_count = prop.multiProcessorCount;
	ctx->l2_bytes = (size_t)prop.l2CacheSize;
	int khz = 0;
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, device);
	ctx->sm_clock_khz = khz;
	if (cudaMalloc(&ctx->d_counters, 8 * sizeof(unsigned long long)) != cudaSuccess ||
	    cudaMemset(ctx->d_counters, 0, 8 * sizeof(unsigned long long)) != cudaSuccess)
	{
		delete ctx;
		return GCR_ERR_CUDA;
	}
	*out = ctx;
	return GCR_OK;
}

extern "C" gcr_status gcr_ctx_destroy(gcr_ctx *ctx)
{
	if (!ctx)
		return GCR_OK;
	cudaSetDevice(ctx->device);
	cudaStreamSynchronize(ctx->stream);
	gcr_comm_destroy(ctx);
	cudaFree(ctx->d_gather);
	cudaFree(ctx->d_pts64_owned);
	cudaFree(ctx->d_upload32);
	cudaFree(ctx->d_pts32);
	cudaFree(ctx->d_score_ws);
	cudaFree(ctx->d_score_amb);
	cudaFree(ctx->d_norm);
	cudaFree(ctx->d_bbox);
	cudaFree(ctx->d_magsac_lut);
	cudaFree(ctx->d_magsac_lut64);
	cudaFree(ctx->d_scratch);
	cudaFree(ctx->d_prosac_growth);
	cudaFree(ctx->d_importance_cdf);
	cudaFree(ctx->d_round);
	cudaFree(ctx->d_counters);
	for (int i = 0; i < gcr_ctx::EV_RING; ++i)
	{
		if (ctx->ev_a[i]) cudaEventDestroy(ctx->ev_a[i]);
		if (ctx->ev_b[i]) cudaEventDestroy(ctx->ev_b[i]);
	}
	if (ctx->h_pinned)
		cudaFreeHost(ctx->h_pinned);
	if (ctx->h_stage)
		cudaFreeHost(ctx->h_stage);
	for (auto &b : ctx->blocks)
		cudaFree(b.p);
	cudaStreamDestroy(ctx->own_stream);
	delete ctx;
	return GCR_OK;
}

static gcr_status flush_d2h(gcr_ctx *ctx);
extern "C" gcr_status gcr_ctx_set_stream(gcr_ctx *ctx, void *cuda_stream)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (ctx->stage_off || !ctx->pending_d2h.empty())
	{
		gcr_status st = flush_d2h(ctx); // copies queued on the old stream are completed on it
		if (st)
			return st;
	}
	ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
	return GCR_OK;
}

extern "C" gcr_status gcr_sync(gcr_ctx *ctx)
{
	if (!ctx)
		return GCR_ERR_ARG;
	return flush_d2h(ctx); // synchronises the stream and delivers the copies queued by gcr_memcpy_d2h_async
}

extern "C" const char *gcr_last_error(const gcr_ctx *ctx) { return ctx ? ctx->err : "null context"; }

extern "C" gcr_status gcr_device_info(gcr_ctx *ctx, int *sm_count, int *sm_clock_khz, size_t *l2_bytes)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (sm_count) *sm_count = ctx->sm_count;
	if (sm_clock_khz) *sm_clock_khz = ctx->sm_clock_khz;
	if (l2_bytes) *l2_bytes = ctx->l2_bytes;
	return GCR_OK;
}

extern "C" uint64_t gcr_launch_count(const gcr_ctx *ctx) { return ctx ? ctx->launches : 0; }

// powers of two up to 256 MiB, 256 MiB granules above: a caller whose requests vary (index lists, staging areas of the
// multi-pair driver) settles on a handful of block sizes instead of one cudaMalloc -- a device-wide stall -- per size
static size_t bucket_of(size_t bytes)
{
	size_t b = 256;
	while (b < bytes && b < ((size_t)256 << 20))
		b <<= 1;
	if (b >= bytes)
		return b;
	return (bytes + ((size_t)256 << 20) - 1) & ~(((size_t)256 << 20) - 1);
}
extern "C" gcr_status gcr_malloc(gcr_ctx *ctx, size_t bytes, void **d_out)
{
	if (!ctx || !d_out)
		return GCR_ERR_ARG;
	const size_t cap = bucket_of(bytes ? bytes : 1);
	for (auto &b : ctx->blocks)
		if (!b.used && b.cap == cap)
		{
			b.used = true;
			*d_out = b.p;
			return GCR_OK;
		}
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	cudaError_t e = cudaMalloc(d_out, cap);
	if (e != cudaSuccess)
	{
		// out of memory: give the cached free blocks back and try once more
		cudaGetLastError();
		cudaStreamSynchronize(ctx->stream);
		for (size_t i = 0; i < ctx->blocks.size();)
			if (!ctx->blocks[i].used)
			{
				cudaFree(ctx->blocks[i].p);
				ctx->blocks[i] = ctx->blocks.back();
				ctx->blocks.pop_back();
			}
			else
				++i;
		GCR_CHECK_CUDA(ctx, cudaMalloc(d_out, cap));
	}
	ctx->blocks.push_back({*d_out, cap, true});
	return GCR_OK;
}
extern "C" gcr_status gcr_free(gcr_ctx *ctx, void *d_ptr)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!d_ptr)
		return GCR_OK;
	for (auto &b : ctx->blocks)
		if (b.p == d_ptr)
		{
			b.used = false; // kept for the next gcr_malloc of this size; released by gcr_ctx_destroy
			return GCR_OK;
		}
	GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_free: pointer was not allocated by gcr_malloc on this context");
}
static gcr_status ensure_stage(gcr_ctx *ctx)
{
	if (ctx->h_stage)
		return GCR_OK;
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	GCR_CHECK_CUDA(ctx, cudaMallocHost(&ctx->h_stage, gcr_ctx::STAGE_BYTES));
	return GCR_OK;
}
// Small host -> device copies: the source is copied into a ring inside the pinned bounce buffer and the transfer is
// queued WITHOUT a synchronise (the caller's buffer is free on return, the data is ordered on the context's stream like
// any kernel).  The ring occupies the first half of the buffer and restarts after any stream synchronise of this file;
// device -> host copies land in the second half and do synchronise (the host reads them).
extern "C" gcr_status gcr_memcpy_h2d(gcr_ctx *ctx, void *d_dst, const void *h_src, size_t bytes)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (bytes == 0)
		return GCR_OK;
	const size_t half = gcr_ctx::STAGE_BYTES / 2;
	if (bytes <= half && ensure_stage(ctx) == GCR_OK)
	{
		const size_t need = (bytes + 255) & ~(size_t)255;
		if (ctx->stage_off + need > half)
		{
			GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // the ring is full: wait for the queued copies
			ctx->stage_off = 0;
		}
		char *slot = (char *)ctx->h_stage + ctx->stage_off;
		memcpy(slot, h_src, bytes);
		GCR_CHECK_CUDA(ctx, cudaMemcpyAsync(d_dst, slot, bytes, cudaMemcpyHostToDevice, ctx->stream));
		ctx->stage_off += need;
		return GCR_OK;
	}
	GCR_CHECK_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
	GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	ctx->stage_off = 0;
	return GCR_OK;
}
// completes the queued device -> host copies: one synchronise, then the pieces go to their destinations
static gcr_status flush_d2h(gcr_ctx *ctx)
{
	GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	const size_t half = gcr_ctx::STAGE_BYTES / 2;
	for (const auto &p : ctx->pending_d2h)
		memcpy(p.dst, (char *)ctx->h_stage + half + p.off, p.bytes);
	ctx->pending_d2h.clear();
	ctx->d2h_off = 0;
	ctx->stage_off = 0; // everything queued before has completed
	return GCR_OK;
}

extern "C" gcr_status gcr_memcpy_d2h_async(gcr_ctx *ctx, void *h_dst, const void *d_src, size_t bytes)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (bytes == 0)
		return GCR_OK;
	const size_t half = gcr_ctx::STAGE_BYTES / 2;
	if (ensure_stage(ctx) != GCR_OK)
		return GCR_ERR_CUDA;
	// through the pinned buffer, in pieces if need be: a pageable destination would send the copy down the driver's own
	// staging path, which serialises the host threads of a multi-context process
	for (size_t done = 0; done < bytes;)
	{
		if (ctx->d2h_off >= half)
		{
			gcr_status st = flush_d2h(ctx);
			if (st)
				return st;
		}
		const size_t room = half - ctx->d2h_off;
		const size_t piece = bytes - done < room ? bytes - done : room;
		GCR_CHECK_CUDA(ctx, cudaMemcpyAsync((char *)ctx->h_stage + half + ctx->d2h_off, (const char *)d_src + done, piece,
		                                    cudaMemcpyDeviceToHost, ctx->stream));
		ctx->pending_d2h.push_back({ctx->d2h_off, (char *)h_dst + done, piece});
		ctx->d2h_off += (piece + 255) & ~(size_t)255;
		done += piece;
	}
	return GCR_OK;
}

extern "C" gcr_status gcr_memcpy_d2h(gcr_ctx *ctx, void *h_dst, const void *d_src, size_t bytes)
{
	if (!ctx)
		return GCR_ERR_ARG;
	gcr_status st = gcr_memcpy_d2h_async(ctx, h_dst, d_src, bytes);
	if (st)
		return st;
	return flush_d2h(ctx);
}
// ---- large transfers of the batch engine: caller-owned pinned host memory, no bounce buffer ----
extern "C" gcr_status gcr_host_alloc_pinned(gcr_ctx *ctx, size_t bytes, void **h_out)
{
	if (!ctx || !h_out)
		return GCR_ERR_ARG;
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	GCR_CHECK_CUDA(ctx, cudaMallocHost(h_out, bytes ? bytes : 1));
	return GCR_OK;
}
extern "C" gcr_status gcr_host_free_pinned(gcr_ctx *ctx, void *h_ptr)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (h_ptr)
		cudaFreeHost(h_ptr);
	return GCR_OK;
}
// queued on the context's stream; h_pinned must come from gcr_host_alloc_pinned.  The data is valid after gcr_sync.
extern "C" gcr_status gcr_memcpy_pinned_async(gcr_ctx *ctx, void *dst, const void *src, size_t bytes, int to_device)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (bytes == 0)
		return GCR_OK;
	GCR_CHECK_CUDA(ctx, cudaMemcpyAsync(dst, src, bytes, to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost, ctx->stream));
	return GCR_OK;
}
// d_dst[i] = row d_rows[i] of d_src (row_bytes each, a multiple of 4): the batch engine's "fetch these model slots"
__global__ void k_gather_rows(const uint32_t *__restrict__ src, uint32_t row_words, const uint64_t *__restrict__ rows, uint32_t count,
                              uint32_t *__restrict__ dst)
{
	const uint64_t total = (uint64_t)count * row_words;
	for (uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (uint64_t)gridDim.x * blockDim.x)
	{
		const uint64_t r = e / row_words, w = e % row_words;
		dst[e] = src[rows[r] * row_words + w];
	}
}
extern "C" gcr_status gcr_gather_rows(gcr_ctx *ctx, const void *d_src, uint32_t row_bytes, const uint64_t *d_rows, uint32_t count,
                                      void *d_dst)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!d_src || !d_rows || !d_dst || count == 0 || row_bytes == 0 || (row_bytes & 3))
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_gather_rows: bad arguments");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	const uint64_t total = (uint64_t)count * (row_bytes / 4);
	unsigned blocks = (unsigned)((total + 255) / 256);
	if (blocks > 4u * ctx->sm_count) blocks = 4u * ctx->sm_count;
	k_gather_rows<<<blocks, 256, 0, ctx->stream>>>((const uint32_t *)d_src, row_bytes / 4, d_rows, count, (uint32_t *)d_dst);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}

extern "C" gcr_status gcr_memset(gcr_ctx *ctx, void *d_dst, int value, size_t bytes)
{
	if (!ctx)
		return GCR_ERR_ARG;
	GCR_CHECK_CUDA(ctx, cudaMemsetAsync(d_dst, value, bytes, ctx->stream));
	return GCR_OK;
}

gcr_status gcr_ensure_scratch(gcr_ctx *ctx, size_t bytes)
{
	if (bytes <= ctx->scratch_capacity)
		return GCR_OK;
	GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	cudaFree(ctx->d_scratch);
	ctx->d_scratch = nullptr;
	ctx->scratch_capacity = 0;
	bytes = (bytes + (1u << 20)) & ~((size_t)(1u << 20) - 1);
	GCR_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_scratch, bytes));
	ctx->scratch_capacity = bytes;
	return GCR_OK;
}

gcr_status gcr_ensure_pinned(gcr_ctx *ctx, size_t bytes)
{
	if (bytes <= ctx->pinned_capacity)
		return GCR_OK;
	if (ctx->h_pinned)
		cudaFreeHost(ctx->h_pinned);
	ctx->h_pinned = nullptr;
	ctx->pinned_capacity = 0;
	bytes = (bytes + 4095) & ~(size_t)4095;
	GCR_CHECK_CUDA(ctx, cudaMallocHost(&ctx->h_pinned, bytes));
	ctx->pinned_capacity = bytes;
	return GCR_OK;
}

// ------------------------------------------------------------------------------- instrumentation helpers ---
static gcr_status profile_drain(gcr_ctx *ctx)
{
	while (ctx->ev_pending > 0)
	{
		const int slot = (ctx->ev_head - ctx->ev_pending + 2 * gcr_ctx::EV_RING) % gcr_ctx::EV_RING;
		GCR_CHECK_CUDA(ctx, cudaEventSynchronize(ctx->ev_b[slot]));
		float ms = 0.f;
		GCR_CHECK_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev_a[slot], ctx->ev_b[slot]));
		ctx->score_ms += ms;
		--ctx->ev_pending;
	}
	return GCR_OK;
}

gcr_status gcr_profile_begin(gcr_ctx *ctx)
{
	if (!ctx->profile)
		return GCR_OK;
	if (ctx->ev_pending == gcr_ctx::EV_RING)
	{
		gcr_status st = profile_drain(ctx);
		if (st)
			return st;
	}
	const int slot = ctx->ev_head;
	if (!ctx->ev_a[slot])
	{
		GCR_CHECK_CUDA(ctx, cudaEventCreate(&ctx->ev_a[slot]));
		GCR_CHECK_CUDA(ctx, cudaEventCreate(&ctx->ev_b[slot]));
	}
	GCR_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev_a[slot], ctx->stream));
	return GCR_OK;
}

gcr_status gcr_profile_end(gcr_ctx *ctx, uint64_t evals)
{
	if (!ctx->profile)
		return GCR_OK;
	GCR_CHECK_CUDA(ctx, cudaEventRecord(ctx->ev_b[ctx->ev_head], ctx->stream));
	ctx->ev_head = (ctx->ev_head + 1) % gcr_ctx::EV_RING;
	++ctx->ev_pending;
	++ctx->score_launches;
	ctx->score_evals += evals;
	return GCR_OK;
}

extern "C" gcr_status gcr_profile_enable(gcr_ctx *ctx, int on)
{
	if (!ctx)
		return GCR_ERR_ARG;
	gcr_status st = profile_drain(ctx);
	ctx->profile = on != 0;
	return st;
}

extern "C" gcr_status gcr_profile_read(gcr_ctx *ctx, double *score_ms, uint64_t *score_launches, uint64_t *score_evals)
{
	if (!ctx)
		return GCR_ERR_ARG;
	gcr_status st = profile_drain(ctx);
	if (st)
		return st;
	if (score_ms) *score_ms = ctx->score_ms;
	if (score_launches) *score_launches = ctx->score_launches;
	if (score_evals) *score_evals = ctx->score_evals;
	ctx->score_ms = 0.0;
	ctx->score_launches = 0;
	ctx->score_evals = 0;
	return GCR_OK;
}

extern "C" gcr_status gcr_counters_read(gcr_ctx *ctx, uint64_t out[8], int reset)
{
	if (!ctx || !out)
		return GCR_ERR_ARG;
	gcr_status st = gcr_memcpy_d2h_async(ctx, out, ctx->d_counters, 8 * sizeof(uint64_t));
	if (st)
		return st;
	if (reset)
		GCR_CHECK_CUDA(ctx, cudaMemsetAsync(ctx->d_counters, 0, 8 * sizeof(uint64_t), ctx->stream));
	return gcr_sync(ctx);
}

extern "C" gcr_status gcr_cache_stats(gcr_ctx *ctx, uint64_t *blocks, uint64_t *blocks_in_use, uint64_t *bytes)
{
	if (!ctx)
		return GCR_ERR_ARG;
	uint64_t used = 0, total = 0;
	for (const auto &b : ctx->blocks)
	{
		used += b.used ? 1 : 0;
		total += b.cap;
	}
	if (blocks) *blocks = ctx->blocks.size();
	if (blocks_in_use) *blocks_in_use = used;
	if (bytes) *bytes = total;
	return GCR_OK;
}

// ------------------------------------------------------------------------------------------------ K0 kernels ---
__global__ void k_bbox_init(unsigned long long *bbox, int entries)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i < entries)
		bbox[i] = (i & 1) ? 0ull : ~0ull; // [..][0] = min -> +max key, [..][1] = max -> 0
}

template <int COLS>
__global__ void __launch_bounds__(256) k_bbox(const double *__restrict__ pts, int n, unsigned long long *bbox)
{
	const int pair = blockIdx.y;
	const double *p = pts + (size_t)pair * n * COLS;
	double mn[COLS], mx[COLS];
#pragma unroll
	for (int c = 0; c < COLS; ++c)
	{
		mn[c] = INFINITY;
		mx[c] = -INFINITY;
	}
	for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
	{
#pragma unroll
		for (int c = 0; c < COLS; ++c)
		{
			const double v = p[(size_t)i * COLS + c];
			if (v < mn[c]) mn[c] = v;
			if (v > mx[c]) mx[c] = v;
		}
	}
#pragma unroll
	for (int c = 0; c < COLS; ++c)
	{
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
		{
			const double a = __shfl_xor_sync(0xffffffffu, mn[c], o), b = __shfl_xor_sync(0xffffffffu, mx[c], o);
			if (a < mn[c]) mn[c] = a;
			if (b > mx[c]) mx[c] = b;
		}
		if ((threadIdx.x & 31) == 0)
		{
			atomicMin(&bbox[((size_t)pair * COLS + c) * 2 + 0], gcr_ordered_u64(mn[c]));
			atomicMax(&bbox[((size_t)pair * COLS + c) * 2 + 1], gcr_ordered_u64(mx[c]));
		}
	}
}

__device__ __forceinline__ double pow2_scale(double half_extent)
{
	if (!(half_extent > 0.0) || isinf(half_extent))
		return 1.0;
	int e;
	frexp(half_extent, &e); // half_extent = f * 2^e, f in [0.5, 1)
	return ldexp(1.0, -e);
}

// centre + power-of-two scale of the primed frame from the bounding box
template <int COLS>
__device__ __forceinline__ void norm_from_bbox(double (&mn)[COLS], double (&mx)[COLS], double *o)
{
	for (int c = 0; c < COLS; ++c)
		if (!(mn[c] <= mx[c])) { mn[c] = 0.0; mx[c] = 0.0; }
	if (COLS == 4)
	{
		double h = 0.0;
		for (int c = 0; c < 4; ++c)
		{
			o[c] = 0.5 * (mn[c] + mx[c]);
			h = fmax(h, 0.5 * (mx[c] - mn[c]));
		}
		o[4] = pow2_scale(h);
		o[5] = o[6] = o[7] = 0.0;
	}
	else
	{
		o[0] = 0.5 * (mn[0] + mx[0]);
		o[1] = 0.5 * (mn[1] + mx[1]);
		o[2] = pow2_scale(fmax(0.5 * (mx[0] - mn[0]), 0.5 * (mx[1] - mn[1])));
		double h = 0.0;
		for (int c = 2; c < 5; ++c)
		{
			o[3 + c - 2] = 0.5 * (mn[c] + mx[c]);
			h = fmax(h, 0.5 * (mx[c] - mn[c]));
		}
		o[6] = pow2_scale(h);
		o[7] = 0.0;
	}
}

template <int COLS>
__global__ void k_norm(const unsigned long long *bbox, double *norm, int n_pairs)
{
	const int pair = blockIdx.x * blockDim.x + threadIdx.x;
	if (pair >= n_pairs)
		return;
	double mn[COLS], mx[COLS];
	for (int c = 0; c < COLS; ++c)
	{
		mn[c] = gcr_from_ordered_u64(bbox[((size_t)pair * COLS + c) * 2 + 0]);
		mx[c] = gcr_from_ordered_u64(bbox[((size_t)pair * COLS + c) * 2 + 1]);
	}
	norm_from_bbox<COLS>(mn, mx, norm + (size_t)pair * 8);
}

// the operand-duplicated FP32 copy of point i of a pair (or the NaN padding row)
template <int COLS>
__device__ __forceinline__ void convert_point(const double *__restrict__ pts_pair, const double *nm, int n, int i, float4 *__restrict__ out_pair)
{
	const float nanv = __int_as_float(0x7fc00000);
	if (COLS == 4)
	{
		float4 q = make_float4(nanv, nanv, nanv, nanv);
		if (i < n)
		{
			const double *p = pts_pair + (size_t)i * 4;
			const double s = nm[4];
			q.x = (float)((p[0] - nm[0]) * s);
			q.y = (float)((p[1] - nm[1]) * s);
			q.z = (float)((p[2] - nm[2]) * s);
			q.w = (float)((p[3] - nm[3]) * s);
		}
		out_pair[(size_t)i * 2 + 0] = make_float4(q.x, q.x, q.y, q.y);
		out_pair[(size_t)i * 2 + 1] = make_float4(q.z, q.z, q.w, q.w);
	}
	else
	{
		float4 a = make_float4(nanv, nanv, 0.f, 0.f), b = make_float4(nanv, nanv, nanv, 0.f);
		if (i < n)
		{
			const double *p = pts_pair + (size_t)i * 5;
			a.x = (float)((p[0] - nm[0]) * nm[2]);
			a.y = (float)((p[1] - nm[1]) * nm[2]);
			b.x = (float)((p[2] - nm[3]) * nm[6]);
			b.y = (float)((p[3] - nm[4]) * nm[6]);
			b.z = (float)((p[4] - nm[5]) * nm[6]);
		}
		out_pair[(size_t)i * 2 + 0] = a;
		out_pair[(size_t)i * 2 + 1] = b;
	}
}

// K0 in ONE launch for pairs of up to 32768 points: one CTA per pair finds the bounding box (min / max are exact and
// order-independent: the same box as the atomic version), derives the frame and converts the pair's points.  The four
// separate launches below cost ~12 us of launch latency on a 128 KB upload.
#define K0_FUSED_THREADS 1024
#define K0_FUSED_MAX_POINTS 32768
// seats != null (gcr_points_replace_pairs): CTA b first copies row b of `incoming` over the points of pair seats[b]
template <int COLS>
__global__ void __launch_bounds__(K0_FUSED_THREADS) k_prepare_fused(const double *__restrict__ pts, int n, int n_pad,
                                                                  double *__restrict__ norm, float4 *__restrict__ out32,
                                                                  const int32_t *__restrict__ seats = nullptr,
                                                                  const double *__restrict__ incoming = nullptr)
{
	__shared__ double s_mn[32][COLS], s_mx[32][COLS];
	__shared__ double s_nm[8];
	const int pair = seats ? seats[blockIdx.x] : (int)blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const double *p = pts + (size_t)pair * n * COLS;
	if (seats)
	{
		double *dst = const_cast<double *>(p);
		const double *src = incoming + (size_t)blockIdx.x * n * COLS;
		for (int e = tid; e < n * COLS; e += K0_FUSED_THREADS)
			dst[e] = src[e];
		__syncthreads(); // (each thread re-reads elements other threads wrote: block-level visibility)
	}
	double mn[COLS], mx[COLS];
#pragma unroll
	for (int c = 0; c < COLS; ++c)
	{
		mn[c] = INFINITY;
		mx[c] = -INFINITY;
	}
	for (int i = tid; i < n; i += K0_FUSED_THREADS)
#pragma unroll
		for (int c = 0; c < COLS; ++c)
		{
			const double v = p[(size_t)i * COLS + c];
			if (v < mn[c]) mn[c] = v;
			if (v > mx[c]) mx[c] = v;
		}
#pragma unroll
	for (int c = 0; c < COLS; ++c)
	{
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
		{
			const double a = __shfl_xor_sync(0xffffffffu, mn[c], o), b = __shfl_xor_sync(0xffffffffu, mx[c], o);
			if (a < mn[c]) mn[c] = a;
			if (b > mx[c]) mx[c] = b;
		}
		if (lane == 0)
		{
			s_mn[warp][c] = mn[c];
			s_mx[warp][c] = mx[c];
		}
	}
	__syncthreads();
	if (tid == 0)
	{
		for (int w = 1; w < K0_FUSED_THREADS / 32; ++w)
			for (int c = 0; c < COLS; ++c)
			{
				if (s_mn[w][c] < mn[c]) mn[c] = s_mn[w][c];
				if (s_mx[w][c] > mx[c]) mx[c] = s_mx[w][c];
			}
		double o[8];
		norm_from_bbox<COLS>(mn, mx, o);
		for (int j = 0; j < 8; ++j)
		{
			s_nm[j] = o[j];
			norm[(size_t)pair * 8 + j] = o[j];
		}
	}
	__syncthreads();
	float4 *out_pair = out32 + (size_t)pair * n_pad * 2;
	for (int i = tid; i < n_pad; i += K0_FUSED_THREADS)
		convert_point<COLS>(p, s_nm, n, i, out_pair);
}

template <int COLS>
__global__ void __launch_bounds__(256) k_convert(const double *__restrict__ pts, const double *__restrict__ norm,
                                                  int n, int n_pad, float4 *__restrict__ out32)
{
	const int pair = blockIdx.y;
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n_pad)
		return;
	convert_point<COLS>(pts + (size_t)pair * n * COLS, norm + (size_t)pair * 8, n, i, out32 + (size_t)pair * n_pad * 2);
}

static gcr_status prepare_points(gcr_ctx *ctx)
{
	const int n = ctx->n, cols = ctx->cols, n_pairs = ctx->n_pairs;
	ctx->n_pad = (n + GCR_TILE_POINTS - 1) / GCR_TILE_POINTS * GCR_TILE_POINTS;
	const size_t need32 = (size_t)n_pairs * ctx->n_pad * 2;
	if (need32 > ctx->pts32_capacity)
	{
		GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_pts32);
		ctx->d_pts32 = nullptr;
		ctx->pts32_capacity = 0;
		GCR_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_pts32, need32 * sizeof(float4)));
		ctx->pts32_capacity = need32;
	}
	if ((size_t)n_pairs > ctx->norm_capacity)
	{
		GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_norm);
		cudaFree(ctx->d_bbox);
		ctx->d_norm = nullptr;
		ctx->d_bbox = nullptr;
		GCR_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_norm, (size_t)n_pairs * 8 * sizeof(double)));
		GCR_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_bbox, (size_t)n_pairs * 5 * 2 * sizeof(unsigned long long)));
		ctx->norm_capacity = n_pairs;
	}
	if (n <= K0_FUSED_MAX_POINTS)
	{
		if (cols == 4)
			k_prepare_fused<4><<<n_pairs, K0_FUSED_THREADS, 0, ctx->stream>>>(ctx->d_pts64, n, ctx->n_pad, ctx->d_norm, ctx->d_pts32);
		else
			k_prepare_fused<5><<<n_pairs, K0_FUSED_THREADS, 0, ctx->stream>>>(ctx->d_pts64, n, ctx->n_pad, ctx->d_norm, ctx->d_pts32);
		GCR_LAUNCHED(ctx);
		return GCR_OK;
	}
	const int entries = n_pairs * cols * 2;
	k_bbox_init<<<(entries + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_bbox, entries);
	GCR_LAUNCHED(ctx);
	int gx = (n + 255) / 256;
	const int max_gx = n_pairs >= ctx->sm_count ? 1 : (4 * ctx->sm_count + n_pairs - 1) / n_pairs;
	if (gx > max_gx) gx = max_gx;
	if (gx < 1) gx = 1;
	const dim3 gb(gx, n_pairs), gc((ctx->n_pad + 255) / 256, n_pairs);
	if (cols == 4)
	{
		k_bbox<4><<<gb, 256, 0, ctx->stream>>>(ctx->d_pts64, n, ctx->d_bbox);
		GCR_LAUNCHED(ctx);
		k_norm<4><<<(n_pairs + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_bbox, ctx->d_norm, n_pairs);
		GCR_LAUNCHED(ctx);
		k_convert<4><<<gc, 256, 0, ctx->stream>>>(ctx->d_pts64, ctx->d_norm, n, ctx->n_pad, ctx->d_pts32);
		GCR_LAUNCHED(ctx);
	}
	else
	{
		k_bbox<5><<<gb, 256, 0, ctx->stream>>>(ctx->d_pts64, n, ctx->d_bbox);
		GCR_LAUNCHED(ctx);
		k_norm<5><<<(n_pairs + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_bbox, ctx->d_norm, n_pairs);
		GCR_LAUNCHED(ctx);
		k_convert<5><<<gc, 256, 0, ctx->stream>>>(ctx->d_pts64, ctx->d_norm, n, ctx->n_pad, ctx->d_pts32);
		GCR_LAUNCHED(ctx);
	}
	return GCR_OK;
}

static gcr_status check_shape(gcr_ctx *ctx, const void *p, int n, int cols, int n_pairs)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!p || n <= 0 || n_pairs <= 0 || (cols != 4 && cols != 5))
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_points: need n > 0, n_pairs > 0, cols in {4,5} (got n=%d cols=%d pairs=%d)", n, cols, n_pairs);
	if ((long long)n * n_pairs > (1ll << 31) - 1024)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_points: n * n_pairs too large");
	return GCR_OK;
}

extern "C" gcr_status gcr_points_upload(gcr_ctx *ctx, const double *h_points, int n, int cols, int n_pairs)
{
	gcr_status st = check_shape(ctx, h_points, n, cols, n_pairs);
	if (st)
		return st;
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	const size_t bytes = (size_t)n * cols * n_pairs * sizeof(double);
	if (bytes > ctx->pts64_capacity)
	{
		GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_pts64_owned);
		ctx->d_pts64_owned = nullptr;
		ctx->pts64_capacity = 0;
		GCR_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_pts64_owned, bytes));
		ctx->pts64_capacity = bytes;
	}
	cudaPointerAttributes attr;
	const bool pinned = cudaPointerGetAttributes(&attr, h_points) == cudaSuccess && attr.type == cudaMemoryTypeHost;
	cudaGetLastError();
	if (!pinned && bytes <= gcr_ctx::STAGE_BYTES && ensure_stage(ctx) == GCR_OK)
	{
		// pageable caller memory: bounce through the context's pinned buffer (a pinned source is copied asynchronously)
		if (ctx->stage_off || !ctx->pending_d2h.empty())
		{
			gcr_status fs = flush_d2h(ctx); // queued copies still use the buffer
			if (fs)
				return fs;
		}
		memcpy(ctx->h_stage, h_points, bytes);
		GCR_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->d_pts64_owned, ctx->h_stage, bytes, cudaMemcpyHostToDevice, ctx->stream));
		GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); // the bounce buffer is free again on return
		ctx->stage_off = 0;
	}
	else
		GCR_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->d_pts64_owned, h_points, bytes, cudaMemcpyHostToDevice, ctx->stream));
	ctx->d_pts64 = ctx->d_pts64_owned;
	ctx->n = n;
	ctx->cols = cols;
	ctx->n_pairs = n_pairs;
	return prepare_points(ctx);
}

// float32 host array -> the FP64 device copy (exact widening: what pybind's py::array_t<double> conversion does on the host
// for the float32 arrays the reference's own tests pass, tests/test_homography.py:52-60) -- half the bytes over PCIe
__global__ void __launch_bounds__(256) k_widen_f32(const float *__restrict__ in, double *__restrict__ out, size_t count)
{
	for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
		out[i] = (double)in[i];
}

extern "C" gcr_status gcr_points_upload_f32(gcr_ctx *ctx, const float *h_points, int n, int cols, int n_pairs)
{
	gcr_status st = check_shape(ctx, h_points, n, cols, n_pairs);
	if (st)
		return st;
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	const size_t count = (size_t)n * cols * n_pairs, bytes = count * sizeof(double), bytes32 = count * sizeof(float);
	if (bytes > ctx->pts64_capacity)
	{
		GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_pts64_owned);
		ctx->d_pts64_owned = nullptr;
		ctx->pts64_capacity = 0;
		GCR_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_pts64_owned, bytes));
		ctx->pts64_capacity = bytes;
	}
	if (bytes32 > ctx->upload32_capacity)
	{
		GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_upload32);
		ctx->d_upload32 = nullptr;
		ctx->upload32_capacity = 0;
		GCR_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_upload32, bytes32));
		ctx->upload32_capacity = bytes32;
	}
	cudaPointerAttributes attr;
	const bool pinned = cudaPointerGetAttributes(&attr, h_points) == cudaSuccess && attr.type == cudaMemoryTypeHost;
	cudaGetLastError();
	if (!pinned && bytes32 <= gcr_ctx::STAGE_BYTES && ensure_stage(ctx) == GCR_OK)
	{
		if (ctx->stage_off || !ctx->pending_d2h.empty())
		{
			gcr_status fs = flush_d2h(ctx);
			if (fs)
				return fs;
		}
		memcpy(ctx->h_stage, h_points, bytes32);
		GCR_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->d_upload32, ctx->h_stage, bytes32, cudaMemcpyHostToDevice, ctx->stream));
		GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		ctx->stage_off = 0;
	}
	else
		GCR_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->d_upload32, h_points, bytes32, cudaMemcpyHostToDevice, ctx->stream));
	size_t blocks = (count + 255) / 256;
	if (blocks > (size_t)ctx->sm_count * 16) blocks = (size_t)ctx->sm_count * 16;
	k_widen_f32<<<(unsigned)blocks, 256, 0, ctx->stream>>>(ctx->d_upload32, ctx->d_pts64_owned, count);
	GCR_LAUNCHED(ctx);
	ctx->d_pts64 = ctx->d_pts64_owned;
	ctx->n = n;
	ctx->cols = cols;
	ctx->n_pairs = n_pairs;
	return prepare_points(ctx);
}

// pairs h_pairs[0..count) of the uploaded context <- rows of h_points; frames and FP32 copies re-derived by the same kernel
// (same arithmetic, same reduction order) as the whole-context upload
extern "C" gcr_status gcr_points_replace_pairs(gcr_ctx *ctx, const int32_t *h_pairs, int count, const double *h_points)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!ctx->d_pts64_owned || ctx->d_pts64 != ctx->d_pts64_owned || ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_points_replace_pairs: needs points uploaded with gcr_points_upload");
	if (!h_pairs || !h_points || count <= 0 || count > ctx->n_pairs)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_points_replace_pairs: bad arguments");
	if (ctx->n > K0_FUSED_MAX_POINTS)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_points_replace_pairs: at most %d points per pair", K0_FUSED_MAX_POINTS);
	for (int r = 0; r < count; ++r)
		if (h_pairs[r] < 0 || h_pairs[r] >= ctx->n_pairs)
			GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_points_replace_pairs: pair %d out of range", h_pairs[r]);
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	const size_t row = (size_t)ctx->n * ctx->cols * sizeof(double), bytes = row * count, seat_bytes = ((size_t)count * 4 + 255) & ~(size_t)255;
	void *d_in = nullptr;
	gcr_status st = gcr_malloc(ctx, seat_bytes + bytes, &d_in);
	if (st)
		return st;
	// pageable sources: synchronous copies (the caller may reuse its buffers on return); stream-ordered after earlier work
	cudaError_t e = cudaMemcpyAsync(d_in, h_pairs, (size_t)count * 4, cudaMemcpyHostToDevice, ctx->stream);
	if (e == cudaSuccess)
		e = cudaMemcpyAsync((char *)d_in + seat_bytes, h_points, bytes, cudaMemcpyHostToDevice, ctx->stream);
	if (e == cudaSuccess)
	{
		const int32_t *d_seats = (const int32_t *)d_in;
		const double *d_rows = (const double *)((char *)d_in + seat_bytes);
		if (ctx->cols == 4)
			k_prepare_fused<4><<<count, K0_FUSED_THREADS, 0, ctx->stream>>>(ctx->d_pts64, ctx->n, ctx->n_pad, ctx->d_norm, ctx->d_pts32, d_seats, d_rows);
		else
			k_prepare_fused<5><<<count, K0_FUSED_THREADS, 0, ctx->stream>>>(ctx->d_pts64, ctx->n, ctx->n_pad, ctx->d_norm, ctx->d_pts32, d_seats, d_rows);
		++ctx->launches;
		e = cudaGetLastError();
	}
	// the block returns to the cache: the next user is ordered after this kernel on the context's stream
	gcr_free(ctx, d_in);
	GCR_CHECK_CUDA(ctx, e);
	return GCR_OK;
}

extern "C" gcr_status gcr_points_set_device(gcr_ctx *ctx, const double *d_points, int n, int cols, int n_pairs)
{
	gcr_status st = check_shape(ctx, d_points, n, cols, n_pairs);
	if (st)
		return st;
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	ctx->d_pts64 = d_points;
	ctx->n = n;
	ctx->cols = cols;
	ctx->n_pairs = n_pairs;
	return prepare_points(ctx);
}

extern "C" gcr_status gcr_points_get_norm(gcr_ctx *ctx, double *h_out)
{
	if (!ctx || !h_out)
		return GCR_ERR_ARG;
	if (!ctx->d_norm || ctx->n_pairs == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_points_get_norm: no points uploaded");
	return gcr_memcpy_d2h(ctx, h_out, ctx->d_norm, (size_t)ctx->n_pairs * 8 * sizeof(double));
}
