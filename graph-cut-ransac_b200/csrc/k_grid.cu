// k_grid.cu -- K7: the grid neighbourhood graph as an integer binning / sort problem.
//
// Replaces GridNeighborhoodGraph<4>::initialize (neighborhood/grid_neighborhood_graph.h:229-292): per point
// idx_d = floor(coord_d / cell_size_d) in FP64 (:251-252), key = sum idx_d * C^d (GridCell, :70-82); the reference
// then inserts every point into an unordered_map<key, vector<point>> in point order (:260-267), so each cell's list
// is ascending in the point index; neighbor_number = sum n (n - 1) / 2 over the filled cells (:283-284).
//
// Here: k_grid_keys (one thread per point; the divide + floor stay FP64 so keys are bit-exact), a stable LSD radix
// sort of (key, point index) -- 8-bit digits, only over the significant key bits; the payload starts as 0..n-1 and
// every pass is stable, so points stay ascending inside a cell -- then head flags + scan give the cell offsets, and
// an integer atomic accumulates neighbor_number.  HBM-bound integer work: 32 B/point read for the keys, then
// 12 B/point read + 12 B/point written per radix pass.
#include "gcr_internal.cuh"

#include <math.h>

#define GT 256 /* threads per CTA and elements per CTA tile */

__global__ void __launch_bounds__(GT) k_grid_keys(const double *__restrict__ pts, int n, double c0, double c1, double c2,
                                                  double c3, unsigned long long cells, unsigned long long *__restrict__ keys,
                                                  uint32_t *__restrict__ idx, unsigned long long *key_or)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	unsigned long long index = 0;
	if (i < n)
	{
	const double *p = pts + (size_t)i * 4;
	// (size_t) of a double like the reference's implicit conversion (:251-252)
	const unsigned long long i0 = (unsigned long long)(long long)floor(p[0] / c0);
	const unsigned long long i1 = (unsigned long long)(long long)floor(p[1] / c1);
	const unsigned long long i2 = (unsigned long long)(long long)floor(p[2] / c2);
	const unsigned long long i3 = (unsigned long long)(long long)floor(p[3] / c3);
	unsigned long long offset = 1;
	index += offset * i0; offset *= cells;
	index += offset * i1; offset *= cells;
	index += offset * i2; offset *= cells;
	index += offset * i3;
	keys[i] = index;
	idx[i] = (uint32_t)i;
	}
	// OR of all keys: tells the host how many radix passes the sort needs
	unsigned long long any = index;
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		any |= __shfl_xor_sync(0xffffffffu, any, o);
	if ((threadIdx.x & 31) == 0 && any)
		atomicOr(key_or, any);
}

// per-CTA digit histogram: hist[digit * n_blocks + block]
__global__ void __launch_bounds__(GT) k_radix_hist(const unsigned long long *__restrict__ keys, int n, int shift,
                                                   uint32_t *__restrict__ hist, int n_blocks)
{
	__shared__ uint32_t h[256];
	h[threadIdx.x] = 0;
	__syncthreads();
	const int i = blockIdx.x * GT + threadIdx.x;
	if (i < n)
		atomicAdd(&h[(uint32_t)(keys[i] >> shift) & 255u], 1u);
	__syncthreads();
	hist[(size_t)threadIdx.x * n_blocks + blockIdx.x] = h[threadIdx.x];
}

// exclusive scan of a u32 array in place (single CTA; arrays here are <= 256 * n_blocks entries)
__global__ void __launch_bounds__(1024) k_scan_exclusive(uint32_t *__restrict__ data, size_t len, uint32_t *total_out)
{
	__shared__ uint32_t warp_tot[32];
	__shared__ uint32_t carry;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	const size_t per = (len + 1023) / 1024;
	const size_t lo = (size_t)tid * per, hi = lo + per < len ? lo + per : len;
	uint32_t sum = 0;
	for (size_t i = lo; i < hi; ++i)
		sum += data[i];
	uint32_t x = sum;
#pragma unroll
	for (int o = 1; o < 32; o <<= 1)
	{
		const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
		if (lane >= o) x += y;
	}
	if (lane == 31)
		warp_tot[warp] = x;
	if (tid == 0)
		carry = 0;
	__syncthreads();
	uint32_t before = 0;
	for (int w = 0; w < warp; ++w)
		before += warp_tot[w];
	uint32_t run = before + x - sum;
	for (size_t i = lo; i < hi; ++i)
	{
		const uint32_t v = data[i];
		data[i] = run;
		run += v;
	}
	if (tid == 1023 && total_out)
		*total_out = before + x;
}

// stable scatter: rank inside the CTA = elements of the same digit at lower thread ids
__global__ void __launch_bounds__(GT) k_radix_scatter(const unsigned long long *__restrict__ keys_in,
                                                      const uint32_t *__restrict__ idx_in, int n, int shift,
                                                      const uint32_t *__restrict__ hist_scanned, int n_blocks,
                                                      unsigned long long *__restrict__ keys_out,
                                                      uint32_t *__restrict__ idx_out)
{
	__shared__ uint32_t warp_digit_count[GT / 32][256];
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	for (int j = tid; j < (GT / 32) * 256; j += GT)
		(&warp_digit_count[0][0])[j] = 0;
	__syncthreads();
	const int i = blockIdx.x * GT + tid;
	const bool ok = i < n;
	unsigned long long key = 0;
	uint32_t digit = 0;
	if (ok)
	{
		key = keys_in[i];
		digit = (uint32_t)(key >> shift) & 255u;
	}
	// peers in this warp with the same digit (inactive lanes use a digit no active lane can match)
	const uint32_t peers = __match_any_sync(0xffffffffu, ok ? digit : 0x100u + lane);
	const uint32_t rank_in_warp = __popc(peers & ((1u << lane) - 1u));
	if (ok && rank_in_warp == 0)
		warp_digit_count[warp][digit] = __popc(peers);
	__syncthreads();
	if (ok)
	{
		uint32_t before = 0;
		for (int w = 0; w < warp; ++w)
			before += warp_digit_count[w][digit];
		const uint32_t dst = hist_scanned[(size_t)digit * n_blocks + blockIdx.x] + before + rank_in_warp;
		keys_out[dst] = key;
		idx_out[dst] = idx_in[i];
	}
}

// head flags of the sorted key array: flag[i] = key[i] != key[i-1]
__global__ void __launch_bounds__(GT) k_grid_heads(const unsigned long long *__restrict__ sorted_keys, int n,
                                                   uint32_t *__restrict__ flags)
{
	const int i = blockIdx.x * GT + threadIdx.x;
	if (i < n)
		flags[i] = (i == 0 || sorted_keys[i] != sorted_keys[i - 1]) ? 1u : 0u;
}

// flags_scanned[i] = index of the cell point i starts (for heads): write the offsets, terminate with n
__global__ void __launch_bounds__(GT) k_grid_offsets(const unsigned long long *__restrict__ sorted_keys, int n,
                                                     const uint32_t *__restrict__ flags_scanned,
                                                     uint32_t *__restrict__ cell_offsets, const uint32_t *n_cells)
{
	const int i = blockIdx.x * GT + threadIdx.x;
	if (i < n && (i == 0 || sorted_keys[i] != sorted_keys[i - 1]))
		cell_offsets[flags_scanned[i]] = (uint32_t)i;
	if (i == 0)
		cell_offsets[*n_cells] = (uint32_t)n;
}

__global__ void __launch_bounds__(GT) k_grid_neighbors(const uint32_t *__restrict__ cell_offsets, const uint32_t *n_cells,
                                                       unsigned long long *neighbor_number)
{
	const uint32_t c = blockIdx.x * GT + threadIdx.x;
	unsigned long long mine = 0;
	if (c < *n_cells)
	{
		const unsigned long long sz = cell_offsets[c + 1] - cell_offsets[c];
		mine = sz * (sz - 1) / 2;
	}
#pragma unroll
	for (int o = 16; o > 0; o >>= 1)
		mine += __shfl_xor_sync(0xffffffffu, mine, o);
	if ((threadIdx.x & 31) == 0 && mine)
		atomicAdd(neighbor_number, mine);
}

extern "C" gcr_status gcr_grid_build(gcr_ctx *ctx, int pair, const double h_cell_sizes[4], uint64_t cells_per_axis,
                                     uint64_t *d_keys, uint32_t *d_sorted_idx, uint32_t *d_cell_offsets, uint32_t *h_n_cells,
                                     uint64_t *h_neighbor_number)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!ctx->d_pts64 || ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_grid_build: no points uploaded");
	if (ctx->cols != 4)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_grid_build: the grid graph is defined on 4-column correspondences");
	if (pair < 0 || pair >= ctx->n_pairs || !h_cell_sizes || cells_per_axis == 0 || !d_keys || !d_sorted_idx || !d_cell_offsets)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_grid_build: bad arguments");
	for (int d = 0; d < 4; ++d)
		if (!(h_cell_sizes[d] > 0.0))
			GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_grid_build: cell size %d must be positive", d);
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	const int n = ctx->n;
	const int n_blocks = (n + GT - 1) / GT;
	// scratch: keys_b | idx_a | idx_b | hist | flags | n_cells | neighbor_number
	size_t off = 0;
	auto take = [&](size_t bytes) {
		const size_t o = off;
		off += (bytes + 255) & ~(size_t)255;
		return o;
	};
	const size_t o_keys_a = take((size_t)n * 8), o_keys_b = take((size_t)n * 8), o_idx_b = take((size_t)n * 4);
	const size_t o_hist = take((size_t)256 * n_blocks * 4), o_flags = take((size_t)n * 4), o_small = take(64);
	gcr_status st = gcr_ensure_scratch(ctx, off);
	if (st)
		return st;
	char *base = (char *)ctx->d_scratch;
	unsigned long long *keys_a = (unsigned long long *)(base + o_keys_a), *keys_b = (unsigned long long *)(base + o_keys_b);
	uint32_t *idx_a = d_sorted_idx, *idx_b = (uint32_t *)(base + o_idx_b);
	uint32_t *hist = (uint32_t *)(base + o_hist), *flags = (uint32_t *)(base + o_flags);
	uint32_t *d_n_cells = (uint32_t *)(base + o_small);
	unsigned long long *d_nn = (unsigned long long *)(base + o_small + 8);
	GCR_CHECK_CUDA(ctx, cudaMemsetAsync(base + o_small, 0, 64, ctx->stream));

	const double *pts = ctx->d_pts64 + (size_t)pair * n * 4;
	k_grid_keys<<<n_blocks, GT, 0, ctx->stream>>>(pts, n, h_cell_sizes[0], h_cell_sizes[1], h_cell_sizes[2], h_cell_sizes[3],
	                                               cells_per_axis, (unsigned long long *)d_keys, idx_a,
	                                               (unsigned long long *)(base + o_small + 16));
	GCR_LAUNCHED(ctx);
	GCR_CHECK_CUDA(ctx, cudaMemcpyAsync(keys_a, d_keys, (size_t)n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
	// significant key bits decide the number of radix passes (keys of in-image points stay below C^4)
	st = gcr_ensure_pinned(ctx, 64);
	if (st)
		return st;
	GCR_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, base + o_small + 16, 8, cudaMemcpyDeviceToHost, ctx->stream));
	GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	int key_bits = 1;
	{
		const unsigned long long any = *(unsigned long long *)ctx->h_pinned;
		while (key_bits < 64 && (any >> key_bits) != 0)
			++key_bits;
	}
	const int passes = (key_bits + 7) / 8;
	unsigned long long *kin = keys_a, *kout = keys_b;
	uint32_t *iin = idx_a, *iout = idx_b;
	for (int p = 0; p < passes; ++p)
	{
		k_radix_hist<<<n_blocks, GT, 0, ctx->stream>>>(kin, n, 8 * p, hist, n_blocks);
		GCR_LAUNCHED(ctx);
		k_scan_exclusive<<<1, 1024, 0, ctx->stream>>>(hist, (size_t)256 * n_blocks, nullptr);
		GCR_LAUNCHED(ctx);
		k_radix_scatter<<<n_blocks, GT, 0, ctx->stream>>>(kin, iin, n, 8 * p, hist, n_blocks, kout, iout);
		GCR_LAUNCHED(ctx);
		unsigned long long *tk = kin; kin = kout; kout = tk;
		uint32_t *ti = iin; iin = iout; iout = ti;
	}
	if (iin != d_sorted_idx)
		GCR_CHECK_CUDA(ctx, cudaMemcpyAsync(d_sorted_idx, iin, (size_t)n * 4, cudaMemcpyDeviceToDevice, ctx->stream));
	k_grid_heads<<<n_blocks, GT, 0, ctx->stream>>>(kin, n, flags);
	GCR_LAUNCHED(ctx);
	k_scan_exclusive<<<1, 1024, 0, ctx->stream>>>(flags, (size_t)n, d_n_cells);
	GCR_LAUNCHED(ctx);
	k_grid_offsets<<<n_blocks, GT, 0, ctx->stream>>>(kin, n, flags, d_cell_offsets, d_n_cells);
	GCR_LAUNCHED(ctx);
	k_grid_neighbors<<<n_blocks, GT, 0, ctx->stream>>>(d_cell_offsets, d_n_cells, d_nn);
	GCR_LAUNCHED(ctx);
	GCR_CHECK_CUDA(ctx, cudaMemcpyAsync(ctx->h_pinned, base + o_small, 16, cudaMemcpyDeviceToHost, ctx->stream));
	GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
	if (h_n_cells)
		*h_n_cells = *(uint32_t *)ctx->h_pinned;
	if (h_neighbor_number)
		*h_neighbor_number = *(unsigned long long *)((char *)ctx->h_pinned + 8);
	return GCR_OK;
}


// --------------------------------------------------------------------------------- radius neighbourhood graph ----
// FlannNeighborhoodGraph::initialize (neighborhood/flann_neighborhood_graph.h:99-137): the points converted to FP32,
// every point's neighbours within `matching_radius` (L2 in the 4-D correspondence space), the point itself dropped.
// The reference asks OpenCV's FLANN for them (4 randomised kd-trees, 6 checks: an approximate, run-dependent answer);
// here the graph is the EXACT one, by brute force: N^2 distance evaluations are nothing for the sizes the graph-cut can
// handle at all (its edge bookkeeping is an N x N matrix, GCRANSAC.h:963).  One thread per point, the other points
// staged through shared memory; pass 1 counts, pass 2 writes the neighbours in ascending index order (CSR).
#define RG_THREADS 256
// COLS = 4 (correspondences) or 5 (2D-3D matches of find6DPose_: FlannNeighborhoodGraph over the whole 5-column row,
// py/gcransac_python.cpp:1090-1092)
struct RgPoint { float v[5]; };
template <bool FILL, int COLS>
__global__ void __launch_bounds__(RG_THREADS) k_radius_graph(const double *__restrict__ pts, int n, float r2,
                                                             uint32_t *__restrict__ counts, const uint32_t *__restrict__ offsets,
                                                             uint32_t *__restrict__ neighbors)
{
	__shared__ RgPoint tile[RG_THREADS];
	const int i = blockIdx.x * RG_THREADS + threadIdx.x;
	RgPoint me;
#pragma unroll
	for (int c = 0; c < 5; ++c) me.v[c] = 0.f;
	if (i < n)
	{
		const double *p = pts + (size_t)i * COLS;
#pragma unroll
		for (int c = 0; c < COLS; ++c) me.v[c] = (float)p[c]; // convertTo(CV_32F) (:115)
	}
	uint32_t found = 0;
	uint32_t *out = FILL && i < n ? neighbors + offsets[i] : nullptr;
	for (int j0 = 0; j0 < n; j0 += RG_THREADS)
	{
		const int j = j0 + threadIdx.x;
		if (j < n)
		{
			const double *p = pts + (size_t)j * COLS;
			RgPoint o;
#pragma unroll
			for (int c = 0; c < 5; ++c) o.v[c] = c < COLS ? (float)p[c < COLS ? c : 0] : 0.f;
			tile[threadIdx.x] = o;
		}
		__syncthreads();
		const int m = min(RG_THREADS, n - j0);
		if (i < n)
			for (int q = 0; q < m; ++q)
			{
				const RgPoint o = tile[q];
				const float dx = me.v[0] - o.v[0], dy = me.v[1] - o.v[1], dz = me.v[2] - o.v[2], dw = me.v[3] - o.v[3];
				float d2 = ((dx * dx + dy * dy) + dz * dz) + dw * dw;
				if (COLS == 5)
				{
					const float dv = me.v[4] - o.v[4];
					d2 = d2 + dv * dv;
				}
				if (d2 < r2 && j0 + q != i) // RadiusResultSet keeps dist < radius; the query point itself is dropped (:131-133)
				{
					if (FILL) out[found] = (uint32_t)(j0 + q);
					++found;
				}
			}
		__syncthreads();
	}
	if (!FILL && i < n)
		counts[i] = found;
}

extern "C" gcr_status gcr_radius_graph_count(gcr_ctx *ctx, int pair, double radius, uint32_t *d_counts)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!ctx->d_pts64 || ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_radius_graph_count: no points uploaded");
	if (pair < 0 || pair >= ctx->n_pairs || !d_counts || !(radius > 0.0))
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_radius_graph_count: bad arguments");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	const float r = (float)radius;
	if (ctx->cols == 5)
		k_radius_graph<false, 5><<<(ctx->n + RG_THREADS - 1) / RG_THREADS, RG_THREADS, 0, ctx->stream>>>(
		    ctx->d_pts64 + (size_t)pair * ctx->n * 5, ctx->n, r * r, d_counts, nullptr, nullptr);
	else
		k_radius_graph<false, 4><<<(ctx->n + RG_THREADS - 1) / RG_THREADS, RG_THREADS, 0, ctx->stream>>>(
		    ctx->d_pts64 + (size_t)pair * ctx->n * 4, ctx->n, r * r, d_counts, nullptr, nullptr);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}

extern "C" gcr_status gcr_radius_graph_fill(gcr_ctx *ctx, int pair, double radius, const uint32_t *d_offsets, uint32_t *d_neighbors)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!ctx->d_pts64 || ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_radius_graph_fill: no points uploaded");
	if (pair < 0 || pair >= ctx->n_pairs || !d_offsets || !d_neighbors || !(radius > 0.0))
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_radius_graph_fill: bad arguments");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	const float r = (float)radius;
	if (ctx->cols == 5)
		k_radius_graph<true, 5><<<(ctx->n + RG_THREADS - 1) / RG_THREADS, RG_THREADS, 0, ctx->stream>>>(
		    ctx->d_pts64 + (size_t)pair * ctx->n * 5, ctx->n, r * r, nullptr, d_offsets, d_neighbors);
	else
		k_radius_graph<true, 4><<<(ctx->n + RG_THREADS - 1) / RG_THREADS, RG_THREADS, 0, ctx->stream>>>(
		    ctx->d_pts64 + (size_t)pair * ctx->n * 4, ctx->n, r * r, nullptr, d_offsets, d_neighbors);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}

// ------------------------------------------------------------------- labeling(): edge list + energy terms on the GPU ----
// GCRANSAC::labeling (GCRANSAC.h:925-1020) walks point_idx ascending and, per point, its neighbour list ascending, adding
// one pairwise term per unordered pair -- it remembers the pairs already used in an N x N matrix (`used_edges`, :963), the
// reference's memory wall at N >> 10^4.  The pairs it ends up with are exactly (i, j > i) with j in i's list, in the order
// (i ascending, j ascending): built here from K7's sorted cells (or the radius graph's CSR) by count -> scan -> fill, once
// per run; per labeling call the unary terms and the pairwise weights come from the model's FP64 residuals in one pass.
__global__ void __launch_bounds__(GT) k_edge_count_grid(const uint32_t *__restrict__ sorted_idx, const uint32_t *__restrict__ cell_offsets,
                                                        uint32_t n_cells, int n, uint32_t *__restrict__ counts)
{
	const int q = blockIdx.x * GT + threadIdx.x;
	if (q >= n)
		return;
	uint32_t lo = 0, hi = n_cells; // cell_offsets[lo] <= q < cell_offsets[hi]
	while (hi - lo > 1)
	{
		const uint32_t mid = (lo + hi) >> 1;
		if (cell_offsets[mid] <= (uint32_t)q) lo = mid;
		else hi = mid;
	}
	// inside a cell the indices are ascending: the later entries of the cell are this point's partners j > i
	counts[sorted_idx[q]] = cell_offsets[lo + 1] - (uint32_t)q - 1u;
}

__global__ void __launch_bounds__(GT) k_edge_fill_grid(const uint32_t *__restrict__ sorted_idx, const uint32_t *__restrict__ cell_offsets,
                                                       uint32_t n_cells, int n, const uint32_t *__restrict__ edge_offsets,
                                                       uint32_t *__restrict__ edges)
{
	const int q = blockIdx.x * GT + threadIdx.x;
	if (q >= n)
		return;
	uint32_t lo = 0, hi = n_cells;
	while (hi - lo > 1)
	{
		const uint32_t mid = (lo + hi) >> 1;
		if (cell_offsets[mid] <= (uint32_t)q) lo = mid;
		else hi = mid;
	}
	const uint32_t i = sorted_idx[q], end = cell_offsets[lo + 1];
	uint32_t *out = edges + 2 * (size_t)edge_offsets[i];
	for (uint32_t t = (uint32_t)q + 1; t < end; ++t)
	{
		*out++ = i;
		*out++ = sorted_idx[t];
	}
}

// CSR neighbour lists (ascending): the partners j > i of point i
__global__ void __launch_bounds__(GT) k_edge_count_csr(const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ neighbors, int n,
                                                       uint32_t *__restrict__ counts)
{
	const int i = blockIdx.x * GT + threadIdx.x;
	if (i >= n)
		return;
	uint32_t c = 0;
	for (uint32_t q = offsets[i]; q < offsets[i + 1]; ++q)
		c += neighbors[q] > (uint32_t)i ? 1u : 0u;
	counts[i] = c;
}
__global__ void __launch_bounds__(GT) k_edge_fill_csr(const uint32_t *__restrict__ offsets, const uint32_t *__restrict__ neighbors, int n,
                                                      const uint32_t *__restrict__ edge_offsets, uint32_t *__restrict__ edges)
{
	const int i = blockIdx.x * GT + threadIdx.x;
	if (i >= n)
		return;
	uint32_t *out = edges + 2 * (size_t)edge_offsets[i];
	for (uint32_t q = offsets[i]; q < offsets[i + 1]; ++q)
		if (neighbors[q] > (uint32_t)i)
		{
			*out++ = (uint32_t)i;
			*out++ = neighbors[q];
		}
}

static gcr_status edges_common(gcr_ctx *ctx, const char *who, uint32_t *d_edge_offsets, uint32_t *d_edges, uint64_t capacity)
{
	if (!ctx->d_pts64 || ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "%s: no points uploaded", who);
	if (!d_edge_offsets || (!d_edges && capacity))
		GCR_FAIL(ctx, GCR_ERR_ARG, "%s: bad arguments", who);
	if (capacity >= (1ull << 31))
		GCR_FAIL(ctx, GCR_ERR_ARG, "%s: more than 2^31 edges (a neighbourhood this dense is not a sparse graph)", who);
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	return GCR_OK;
}

extern "C" gcr_status gcr_grid_edges(gcr_ctx *ctx, const uint32_t *d_sorted_idx, const uint32_t *d_cell_offsets, uint32_t n_cells,
                                     uint32_t *d_edge_offsets, uint32_t *d_edges, uint64_t capacity)
{
	if (!ctx)
		return GCR_ERR_ARG;
	gcr_status st = edges_common(ctx, "gcr_grid_edges", d_edge_offsets, d_edges, capacity);
	if (st)
		return st;
	if (!d_sorted_idx || !d_cell_offsets || n_cells == 0)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_grid_edges: bad arguments");
	const int n = ctx->n, nb = (n + GT - 1) / GT;
	GCR_CHECK_CUDA(ctx, cudaMemsetAsync(d_edge_offsets + n, 0, sizeof(uint32_t), ctx->stream));
	k_edge_count_grid<<<nb, GT, 0, ctx->stream>>>(d_sorted_idx, d_cell_offsets, n_cells, n, d_edge_offsets);
	GCR_LAUNCHED(ctx);
	k_scan_exclusive<<<1, 1024, 0, ctx->stream>>>(d_edge_offsets, (size_t)n + 1, nullptr);
	GCR_LAUNCHED(ctx);
	if (capacity)
	{
		k_edge_fill_grid<<<nb, GT, 0, ctx->stream>>>(d_sorted_idx, d_cell_offsets, n_cells, n, d_edge_offsets, d_edges);
		GCR_LAUNCHED(ctx);
	}
	return GCR_OK;
}

extern "C" gcr_status gcr_csr_edges(gcr_ctx *ctx, const uint32_t *d_offsets, const uint32_t *d_neighbors, uint32_t *d_edge_offsets,
                                    uint32_t *d_edges, uint64_t capacity)
{
	if (!ctx)
		return GCR_ERR_ARG;
	gcr_status st = edges_common(ctx, "gcr_csr_edges", d_edge_offsets, d_edges, capacity);
	if (st)
		return st;
	if (!d_offsets || (!d_neighbors && capacity))
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_csr_edges: bad arguments");
	const int n = ctx->n, nb = (n + GT - 1) / GT;
	GCR_CHECK_CUDA(ctx, cudaMemsetAsync(d_edge_offsets + n, 0, sizeof(uint32_t), ctx->stream));
	k_edge_count_csr<<<nb, GT, 0, ctx->stream>>>(d_offsets, d_neighbors, n, d_edge_offsets);
	GCR_LAUNCHED(ctx);
	k_scan_exclusive<<<1, 1024, 0, ctx->stream>>>(d_edge_offsets, (size_t)n + 1, nullptr);
	GCR_LAUNCHED(ctx);
	if (capacity)
	{
		k_edge_fill_csr<<<nb, GT, 0, ctx->stream>>>(d_offsets, d_neighbors, n, d_edge_offsets, d_edges);
		GCR_LAUNCHED(ctx);
	}
	return GCR_OK;
}

// unary terms: (to source, to sink) per point (GCRANSAC.h:933-957); dist = clamp(r^2 / T, 0, 1) kept for the pairwise pass
template <int PROBLEM>
__global__ void __launch_bounds__(GT) k_labeling_unary(const double *__restrict__ pts, int n, const double *__restrict__ model, double T,
                                                       double one_minus_lambda, double *__restrict__ dist, double *__restrict__ unary)
{
	constexpr int COLS = PROBLEM == GCR_PNP ? 5 : 4;
	constexpr int ML = PROBLEM == GCR_PNP ? 12 : 9;
	double M[ML];
	for (int j = 0; j < ML; ++j) M[j] = model[j];
	for (int i = blockIdx.x * GT + threadIdx.x; i < n; i += gridDim.x * GT)
	{
		double pt[COLS];
		for (int j = 0; j < COLS; ++j) pt[j] = pts[(size_t)i * COLS + j];
		const double r2 = gcr_sqres64<PROBLEM>(pt, M);
		const double d = r2 / T;
		// utils::clamp; a NaN residual counts as a full-cost outlier (the reference would feed NaN capacities to the max-flow)
		const double dc = d < 0.0 ? 0.0 : (d > 1.0 ? 1.0 : (d == d ? d : 1.0));
		dist[i] = dc;
		const double tmp_energy = 1 - dc;
		if (r2 <= T)
		{
			unary[2 * (size_t)i] = one_minus_lambda * tmp_energy;
			unary[2 * (size_t)i + 1] = 0.0;
		}
		else
		{
			unary[2 * (size_t)i] = 0.0;
			unary[2 * (size_t)i + 1] = one_minus_lambda * (1 - tmp_energy);
		}
	}
}
// pairwise weights e00 * lambda = 0.5 (dist_i + dist_j) * lambda (GCRANSAC.h:996-1003)
__global__ void __launch_bounds__(GT) k_labeling_pairwise(const uint32_t *__restrict__ edges, size_t n_edges, const double *__restrict__ dist,
                                                          double lambda, double *__restrict__ pairwise)
{
	for (size_t e = (size_t)blockIdx.x * GT + threadIdx.x; e < n_edges; e += (size_t)gridDim.x * GT)
	{
		const double e00 = 0.5 * (dist[edges[2 * e]] + dist[edges[2 * e + 1]]);
		pairwise[e] = e00 * lambda;
	}
}

extern "C" gcr_status gcr_labeling_terms(gcr_ctx *ctx, int problem, int pair, const double *d_model, double threshold, double lambda,
                                         const uint32_t *d_edges, uint64_t n_edges, double *d_unary, double *d_pairwise,
                                         double *d_dist)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!ctx->d_pts64 || ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_labeling_terms: no points uploaded");
	if (!d_model || !d_unary || !(threshold > 0.0) || pair < 0 || pair >= ctx->n_pairs || (n_edges && d_pairwise && !d_edges))
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_labeling_terms: bad arguments");
	if (gcr_cols_of(problem) != ctx->cols)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_labeling_terms: problem / point columns mismatch");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	const int n = ctx->n;
	double *dist = d_dist;
	if (!dist)
	{
		gcr_status st = gcr_ensure_scratch(ctx, (size_t)n * sizeof(double));
		if (st)
			return st;
		dist = (double *)ctx->d_scratch;
	}
	const double T = threshold * threshold * 9 / 4; // squared_truncated_threshold (GCRANSAC.h:929)
	const double *pts = ctx->d_pts64 + (size_t)pair * n * ctx->cols;
	int nb = (n + GT - 1) / GT;
	if (nb > 8 * ctx->sm_count) nb = 8 * ctx->sm_count;
	if (problem == GCR_H)
		k_labeling_unary<GCR_H><<<nb, GT, 0, ctx->stream>>>(pts, n, d_model, T, 1.0 - lambda, dist, d_unary);
	else if (problem == GCR_PNP)
		k_labeling_unary<GCR_PNP><<<nb, GT, 0, ctx->stream>>>(pts, n, d_model, T, 1.0 - lambda, dist, d_unary);
	else
		k_labeling_unary<GCR_F><<<nb, GT, 0, ctx->stream>>>(pts, n, d_model, T, 1.0 - lambda, dist, d_unary);
	GCR_LAUNCHED(ctx);
	if (n_edges && d_pairwise)
	{
		size_t eb = (n_edges + GT - 1) / GT;
		if (eb > (size_t)16 * ctx->sm_count) eb = (size_t)16 * ctx->sm_count;
		k_labeling_pairwise<<<(unsigned)eb, GT, 0, ctx->stream>>>(d_edges, (size_t)n_edges, dist, lambda, d_pairwise);
		GCR_LAUNCHED(ctx);
	}
	return GCR_OK;
}

// The batch engine's labeling terms: k models that belong to DIFFERENT pairs, one launch.  d_out [k][3 n]: per model the
// [n][2] unary capacities followed by the [n] clamped distances (what gcr_labeling_terms writes to d_unary and d_dist).
template <int PROBLEM>
__global__ void __launch_bounds__(GT) k_labeling_unary_multi(const double *__restrict__ pts_all, size_t pair_stride, int n,
                                                             const int32_t *__restrict__ pair_ids, const double *__restrict__ models,
                                                             double T, double one_minus_lambda, double *__restrict__ out)
{
	constexpr int ML = PROBLEM == GCR_PNP ? 12 : 9;
	const int m = blockIdx.y;
	double *o = out + (size_t)m * 3 * n;
	// (same arithmetic as k_labeling_unary: the single-model kernel with this model's pair, outputs and grid column)
	constexpr int COLS = PROBLEM == GCR_PNP ? 5 : 4;
	const double *pts = pts_all + (size_t)pair_ids[m] * pair_stride;
	double M[ML];
	for (int j = 0; j < ML; ++j) M[j] = models[(size_t)m * ML + j];
	double *unary = o, *dist = o + 2 * (size_t)n;
	for (int i = blockIdx.x * GT + threadIdx.x; i < n; i += gridDim.x * GT)
	{
		double pt[COLS];
		for (int j = 0; j < COLS; ++j) pt[j] = pts[(size_t)i * COLS + j];
		const double r2 = gcr_sqres64<PROBLEM>(pt, M);
		const double d = r2 / T;
		const double dc = d < 0.0 ? 0.0 : (d > 1.0 ? 1.0 : (d == d ? d : 1.0));
		dist[i] = dc;
		const double tmp_energy = 1 - dc;
		if (r2 <= T)
		{
			unary[2 * (size_t)i] = one_minus_lambda * tmp_energy;
			unary[2 * (size_t)i + 1] = 0.0;
		}
		else
		{
			unary[2 * (size_t)i] = 0.0;
			unary[2 * (size_t)i + 1] = one_minus_lambda * (1 - tmp_energy);
		}
	}
}

extern "C" gcr_status gcr_labeling_terms_multi(gcr_ctx *ctx, int problem, const int32_t *d_pair_ids, const double *d_models, int k,
                                               double threshold, double lambda, double *d_out)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!ctx->d_pts64 || ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_labeling_terms_multi: no points uploaded");
	if (!d_pair_ids || !d_models || !d_out || k <= 0 || !(threshold > 0.0) || gcr_cols_of(problem) != ctx->cols)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_labeling_terms_multi: bad arguments");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	const int n = ctx->n;
	const double T = threshold * threshold * 9 / 4;
	int nb = (n + GT - 1) / GT;
	if (nb > 8) nb = 8;
	const dim3 grid(nb, k);
	const size_t stride = (size_t)n * ctx->cols;
	if (problem == GCR_H)
		k_labeling_unary_multi<GCR_H><<<grid, GT, 0, ctx->stream>>>(ctx->d_pts64, stride, n, d_pair_ids, d_models, T, 1.0 - lambda, d_out);
	else if (problem == GCR_PNP)
		k_labeling_unary_multi<GCR_PNP><<<grid, GT, 0, ctx->stream>>>(ctx->d_pts64, stride, n, d_pair_ids, d_models, T, 1.0 - lambda, d_out);
	else
		k_labeling_unary_multi<GCR_F><<<grid, GT, 0, ctx->stream>>>(ctx->d_pts64, stride, n, d_pair_ids, d_models, T, 1.0 - lambda, d_out);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}
