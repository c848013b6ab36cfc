// k_solve.cu -- K2: batched minimal solvers, one hypothesis (sample) per thread, FP64, with the sample / model
// validity tests of the reference's estimators folded into the same kernel.
//
//   H  4-pt: RobustHomographyEstimator::isValidSample (estimators/homography_estimator.h:346-381) +
//            HomographyFourPointSolver::estimateMinimalModel (estimators/solver_homography_four_point.h:109-190) with
//            utils::gaussElimination<8> (math_utils.h:45-87) -- same operation order, no FMA contraction, so the
//            models are bit-identical to the reference's.
//   F  7-pt: FundamentalMatrixSevenPointSolver::estimateModel (estimators/solver_fundamental_matrix_seven_point.h
//            :91-291): full-pivot LU null space (Eigen FullPivLU::kernel semantics), cubic det(lambda F1 + (1-lambda)
//            F2) = 0 solved by bracketing at the derivative's roots + safeguarded Newton (only + - * / sqrt, so CPU
//            oracle and GPU agree bit for bit), roots in ascending order; then the oriented epipolar constraint of
//            FundamentalMatrixEstimator::estimateModel (estimators/fundamental_estimator.h:161-184, 737-800).
// The solvers are O(K) work against the O(K N) of scoring: they run from L2-resident FP64 points with
// per-thread local arrays; no attempt is made to put them on tensor cores.
#include "gcr_internal.cuh"

#include <math.h>

#include "solve_f7.cuh"
#include "solve_e5.cuh"
#include "solve_e5_group.cuh"
#include "solve_p3p.cuh"

extern "C" int gcr_max_models(int problem)
{
	switch (problem)
	{
	case GCR_H: return 1;
	case GCR_F: return 3;
	case GCR_E: return 10;
	case GCR_PNP: return 4;
	default: return 0;
	}
}
extern "C" int gcr_sample_size(int problem)
{
	switch (problem)
	{
	case GCR_H: return 4;
	case GCR_F: return 7;
	case GCR_E: return 5;
	case GCR_PNP: return 3;
	default: return 0;
	}
}
extern "C" int gcr_model_len(int problem) { return problem == GCR_PNP ? 12 : 9; }

// valid-model counter: one atomic per warp (threads that exited early simply do not take part)
__device__ __forceinline__ void count_valid(unsigned long long *counters, uint32_t mine)
{
	const uint32_t active = __activemask();
	const uint32_t total = __reduce_add_sync(active, mine); // the converged lanes of this warp
	if ((threadIdx.x & 31) == (uint32_t)(__ffs(active) - 1) && total)
		atomicAdd(&counters[0], (unsigned long long)total);
}

// ---------------------------------------------------------------------------------------------------- H4 ----
__device__ __forceinline__ void h_cross(double *r, const double *v1, const double *v2)
{
	r[0] = v1[1] - v2[1];
	r[1] = v2[0] - v1[0];
	r[2] = v1[0] * v2[1] - v1[1] * v2[0];
}

__device__ bool h_valid_sample(const double *a, const double *b, const double *c, const double *d)
{
	double p[3], q[3];
	h_cross(p, a, b);
	h_cross(q, a + 2, b + 2);
	if ((p[0] * c[0] + p[1] * c[1] + p[2]) * (q[0] * c[2] + q[1] * c[3] + q[2]) < 0)
		return false;
	if ((p[0] * d[0] + p[1] * d[1] + p[2]) * (q[0] * d[2] + q[1] * d[3] + q[2]) < 0)
		return false;
	h_cross(p, c, d);
	h_cross(q, c + 2, d + 2);
	if ((p[0] * a[0] + p[1] * a[1] + p[2]) * (q[0] * a[2] + q[1] * a[3] + q[2]) < 0)
		return false;
	if ((p[0] * b[0] + p[1] * b[1] + p[2]) * (q[0] * b[2] + q[1] * b[3] + q[2]) < 0)
		return false;
	return true;
}

// The oriented-constraint test rejects ~56 % of uniformly drawn 4-point samples (the "twisted quad" cases).  With one
// sample per thread those lanes would idle through the whole elimination, so a CTA first tests H4_ROUNDS samples per
// thread, packs the coordinates of the VALID ones into a shared-memory table (ballot + one shared atomic per warp; the
// packing order does not matter, every result goes back to its own sample's slot), and then the threads solve the
// packed samples: the FP64 work is issued for (nearly) full warps.
#define H4_THREADS 128
#define H4_ROUNDS 2
#define H4_SAMPLES (H4_THREADS * H4_ROUNDS)
__global__ void __launch_bounds__(H4_THREADS) k_solve_h4(const double *__restrict__ pts, int n,
                                                         const GcrSampleSource src, int count,
                                                         double *__restrict__ models, uint8_t *__restrict__ valid,
                                                         unsigned long long *__restrict__ counters)
{
	__shared__ double s_pt[16][H4_SAMPLES]; // coordinate e of packed sample q at [e][q]: conflict-free
	__shared__ int s_owner[H4_SAMPLES];     // sample index of packed sample q
	__shared__ int s_n;
	const int tid = threadIdx.x;
	const int pair = blockIdx.y;
	if (src.active && !src.active[pair])
		return; // a pair the batch engine has no request for (valid flags zeroed by the launcher)
	const double *P = pts + (size_t)pair * n * 4;
	if (tid == 0)
		s_n = 0;
	__syncthreads();
#pragma unroll 1
	for (int r = 0; r < H4_ROUNDS; ++r)
	{
		const int i = blockIdx.x * H4_SAMPLES + r * H4_THREADS + tid;
		double pt[4][4];
		bool ok = false;
		if (i < count)
		{
			uint32_t s[4];
			gcr_load_sample<4>(src, pair, count, i, s);
#pragma unroll
			for (int j = 0; j < 4; ++j)
			{
				const double *p = P + (size_t)s[j] * 4;
				pt[j][0] = p[0]; pt[j][1] = p[1]; pt[j][2] = p[2]; pt[j][3] = p[3];
			}
			ok = h_valid_sample(pt[0], pt[1], pt[2], pt[3]);
			if (!ok)
				valid[(size_t)pair * count + i] = 0; // (the model slot of a rejected sample is left as it is)
		}
		const uint32_t ballot = __ballot_sync(0xffffffffu, ok);
		int base = 0;
		if ((tid & 31) == 0 && ballot)
			base = atomicAdd(&s_n, __popc(ballot));
		base = __shfl_sync(0xffffffffu, base, 0);
		if (ok)
		{
			const int q = base + __popc(ballot & ((1u << (tid & 31)) - 1u));
			s_owner[q] = i;
#pragma unroll
			for (int j = 0; j < 4; ++j)
#pragma unroll
				for (int c = 0; c < 4; ++c)
					s_pt[4 * j + c][q] = pt[j][c];
		}
	}
	__syncthreads();
	const int n_valid = s_n;
#pragma unroll 1
	for (int q = tid; q < n_valid; q += H4_THREADS)
	{
	const int owner = s_owner[q];
	double *out = models + ((size_t)pair * count + owner) * 9;
	bool ok = true;
	double M[8][9];
	{
		int row = 0;
#pragma unroll
		for (int j = 0; j < 4; ++j)
		{
			const double x1 = s_pt[4 * j][q], y1 = s_pt[4 * j + 1][q], x2 = s_pt[4 * j + 2][q], y2 = s_pt[4 * j + 3][q];
			const double weight = 1.0;
			const double mwx1 = -weight * x1, mwy1 = -weight * y1, wx2 = weight * x2, wy2 = weight * y2;
			M[row][0] = mwx1; M[row][1] = mwy1; M[row][2] = -weight; M[row][3] = 0; M[row][4] = 0; M[row][5] = 0;
			M[row][6] = wx2 * x1; M[row][7] = wx2 * y1; M[row][8] = -wx2;
			++row;
			M[row][0] = 0; M[row][1] = 0; M[row][2] = 0; M[row][3] = mwx1; M[row][4] = mwy1; M[row][5] = -weight;
			M[row][6] = wy2 * x1; M[row][7] = wy2 * y1; M[row][8] = -wy2;
			++row;
		}
		// math_utils.h:54-62 "pivotisation"
		for (int a = 0; a < 8; a++)
			for (int k = a + 1; k < 8; k++)
				if (fabs(M[a][a]) < fabs(M[k][a]))
					for (int j = 0; j <= 8; j++)
					{
						const double temp = M[a][j];
						M[a][j] = M[k][j];
						M[k][j] = temp;
					}
		// :65-72 elimination
		for (int a = 0; a < 7; a++)
			for (int k = a + 1; k < 8; k++)
			{
				const double temp = M[k][a] / M[a][a];
				for (int j = 0; j <= 8; j++)
					M[k][j] = M[k][j] - temp * M[a][j];
			}
		// :75-86 back substitution
		double h[8];
		for (int a = 7; a >= 0; a--)
		{
			h[a] = M[a][8];
			for (int j = a + 1; j < 8; j++)
				h[a] = h[a] - M[a][j] * h[j];
			h[a] = h[a] / M[a][a];
		}
		for (int j = 0; j < 8; ++j)
			if (isnan(h[j]))
				ok = false;
		if (ok)
		{
			for (int j = 0; j < 8; ++j)
				out[j] = h[j];
			out[8] = 1.0;
		}
	}
	valid[(size_t)pair * count + owner] = ok ? 1 : 0;
	count_valid(counters, ok ? 1u : 0u);
	}
}

// ---------------------------------------------------------------------------------------------------- F7 ----
#ifndef F7_THREADS
#define F7_THREADS 64
#define F7_MIN_BLOCKS 6
#endif
#define F7_SMEM (63 * F7_THREADS * sizeof(double))

struct SamplePoints // sample point i, column c, straight from the (L2-resident) FP64 point array
{
	const double *P;
	const uint32_t *s;
	__device__ __forceinline__ double operator()(int i, int c) const { return P[(size_t)s[i] * 4 + c]; }
};

__global__ void __launch_bounds__(F7_THREADS, F7_MIN_BLOCKS) k_solve_f7(const double *__restrict__ pts, int n,
                                                         const GcrSampleSource src, int count,
                                                         double *__restrict__ models, uint8_t *__restrict__ valid,
                                                         unsigned long long *__restrict__ counters)
{
	extern __shared__ double f7_smem[];
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int pair = blockIdx.y;
	if (src.active && !src.active[pair])
		return; // a pair the batch engine has no request for (valid flags zeroed by the launcher)
	uint32_t n_ok = 0;
	if (i < count)
	{
		uint32_t sidx[7];
		gcr_load_sample<7>(src, pair, count, i, sidx);
		const SamplePoints pt{pts + (size_t)pair * n * 4, sidx};
		const gcr_f7::SmemMat A{f7_smem + threadIdx.x, F7_THREADS};
		double f1[9], f2[9], roots[3];
		const int nr = gcr_f7::solve(pt, A, f1, f2, roots);
		double *out = models + ((size_t)pair * count + i) * 27;
		uint8_t *v = valid + ((size_t)pair * count + i) * 3;
		// model slot = root index (ascending roots); a root the reference skips (|s| <= eps) leaves its slot invalid
#pragma unroll
		for (int mdl = 0; mdl < 3; ++mdl)
		{
			double F[9];
			bool ok = mdl < nr && gcr_f7::model(f1, f2, roots[mdl], F);
			if (ok)
				ok = gcr_f7::orientation_valid(F, pt, 7); // fundamental_estimator.h:175-180
			// (a slot whose flag is 0 keeps whatever it held: nobody reads it, and 89 % of the 7-point slots are such -- storing
			// zeros there was 57 MB of scattered writes per 262 144 samples, 7 % of a round)
			if (ok)
#pragma unroll
				for (int j = 0; j < 9; ++j)
					out[mdl * 9 + j] = F[j];
			v[mdl] = ok ? 1 : 0;
			n_ok += ok ? 1u : 0u;
		}
	}
	// valid-model counter: block reduction through shared memory (the matrix space is free again)
	__syncthreads();
	uint32_t *cnt = reinterpret_cast<uint32_t *>(f7_smem);
	if (threadIdx.x == 0)
		cnt[0] = 0;
	__syncthreads();
	if (n_ok)
		atomicAdd(&cnt[0], n_ok);
	__syncthreads();
	if (threadIdx.x == 0 && cnt[0])
		atomicAdd(&counters[0], (unsigned long long)cnt[0]);
}

// ---------------------------------------------------------------------------------------------------- E5 ----
// Ten lanes per sample, three samples per warp, the solver's working set (10 x 20 constraint matrix, Sturm chain,
// interval stack: ~5 KB per sample) in shared memory: solve_e5_group.cuh.  Lanes 30 / 31 of every warp idle.
#ifndef E5_MIN_BLOCKS
#define E5_MIN_BLOCKS 5 // 3.7 KB of shared memory per sample: five CTAs (20 warps) per SM if the registers allow (<= 102)
#endif
#ifndef E5_THREADS
#define E5_THREADS 128 // overridable at build time (GCR_NVCC_EXTRA=-DE5_THREADS=64 -DE5_MIN_BLOCKS=10) for launch-shape experiments
#endif
#define E5_SAMPLES_PER_CTA (3 * E5_THREADS / 32)
#define E5_SMEM (E5_SAMPLES_PER_CTA * gcr_e5g::SAMPLE_DOUBLES * sizeof(double))

struct GroupSync
{
	unsigned mask;      // the ten lanes of this group
	unsigned warp_mask; // every lane of the warp that runs a group
	int shift;          // first lane of the group
	__device__ __forceinline__ void operator()() const { __syncwarp(mask); }
	__device__ __forceinline__ bool any(bool x) const { return __any_sync(warp_mask, x) != 0; }
	__device__ __forceinline__ unsigned ballot(bool x) const { return (__ballot_sync(mask, x) >> shift) & 0x3ffu; }
};

__global__ void __launch_bounds__(E5_THREADS, E5_MIN_BLOCKS) k_solve_e5(const double *__restrict__ pts, int n,
                                                         const GcrSampleSource src, int count,
                                                         double *__restrict__ models, uint8_t *__restrict__ valid,
                                                         unsigned long long *__restrict__ counters)
{
	extern __shared__ double e5_smem[];
	__shared__ uint32_t cnt;
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	const int g = lane / 10, li = lane - 10 * g;
	const int pair = blockIdx.y;
	if (src.active && !src.active[pair])
		return; // a pair the batch engine has no request for (valid flags zeroed by the launcher)
	const int slot = warp * 3 + g;
	const int i = blockIdx.x * E5_SAMPLES_PER_CTA + slot;
	if (threadIdx.x == 0)
		cnt = 0;
	__syncthreads();
	const unsigned warp_mask = __ballot_sync(0xffffffffu, g < 3 && i < count);
	if (g < 3 && i < count)
	{
		double *S = e5_smem + (size_t)slot * gcr_e5g::SAMPLE_DOUBLES;
		const GroupSync sync{0x3ffu << (10 * g), warp_mask, 10 * g};
		uint32_t sidx[5];
		gcr_load_sample<5>(src, pair, count, i, sidx); // every lane draws the same five indices
		if (li < 5)
		{
			const double *p = pts + ((size_t)pair * n + sidx[li]) * 4;
			double *d = S + gcr_e5g::OFF_PT + 4 * li;
			d[0] = p[0]; d[1] = p[1]; d[2] = p[2]; d[3] = p[3];
		}
		sync();
		double *out = models + ((size_t)pair * count + i) * 90;
		uint8_t *v = valid + ((size_t)pair * count + i) * 10;
		const int n_ok = gcr_e5g::solve_group(li, S, sync, out, v);
		if (li == 0 && n_ok)
			atomicAdd(&cnt, (uint32_t)n_ok);
	}
	__syncthreads();
	if (threadIdx.x == 0 && cnt)
		atomicAdd(&counters[0], (unsigned long long)cnt);
}

// --------------------------------------------------------------------------------------------------- P3P ----
struct SamplePoints3
{
	const double *P; // rows of 5: u v X Y Z
	const uint32_t *s;
	__device__ __forceinline__ double operator()(int i, int c) const { return P[(size_t)s[i] * 5 + c]; }
};

__global__ void __launch_bounds__(128) k_solve_p3p(const double *__restrict__ pts, int n,
                                                   const GcrSampleSource src, int count,
                                                   double *__restrict__ models, uint8_t *__restrict__ valid,
                                                   unsigned long long *__restrict__ counters)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int pair = blockIdx.y;
	if (src.active && !src.active[pair])
		return; // a pair the batch engine has no request for (valid flags zeroed by the launcher)
	uint32_t n_ok = 0;
	if (i < count)
	{
		uint32_t sidx[3];
		gcr_load_sample<3>(src, pair, count, i, sidx);
		const SamplePoints3 pt{pts + (size_t)pair * n * 5, sidx};
		double P[48];
		const int nm = gcr_p3p::solve(pt, P);
		double *out = models + ((size_t)pair * count + i) * 48;
		uint8_t *v = valid + ((size_t)pair * count + i) * 4;
		for (int mdl = 0; mdl < 4; ++mdl)
		{
			bool ok = mdl < nm;
			if (ok)
				for (int j = 0; j < 12; ++j)
					if (!isfinite(P[12 * mdl + j])) ok = false; // PerspectiveNPointEstimator::estimateModel keeps what the solver returns
			if (ok)
				for (int j = 0; j < 12; ++j)
					out[12 * mdl + j] = P[12 * mdl + j];
			v[mdl] = ok ? 1 : 0;
			n_ok += ok ? 1u : 0u;
		}
	}
	__shared__ uint32_t cnt;
	if (threadIdx.x == 0)
		cnt = 0;
	__syncthreads();
	if (n_ok)
		atomicAdd(&cnt, n_ok);
	__syncthreads();
	if (threadIdx.x == 0 && cnt)
		atomicAdd(&counters[0], (unsigned long long)cnt);
}

// ------------------------------------------------------------------------------------------ plane + parallax ----
// FundamentalMatrixPlaneParallaxSolver::estimateModel (estimators/solver_fundamental_matrix_plane_and_parallax.h:109-160)
// + the oriented epipolar constraint on the two sample points (fundamental_estimator.h:161-184): DEGENSAC's inner
// minimal solver.  One sample (2 points) per thread; same statement order as oracle/gcr_oracle_f.cpp.
struct PlaneHomography { double h[9]; };

__device__ __forceinline__ void cross3(const double *a, const double *b, double *c)
{
	c[0] = a[1] * b[2] - a[2] * b[1];
	c[1] = a[2] * b[0] - a[0] * b[2];
	c[2] = a[0] * b[1] - a[1] * b[0];
}

__global__ void __launch_bounds__(128) k_solve_fpp(const double *__restrict__ pts, int n,
                                                   const uint32_t *__restrict__ samples, int count, const PlaneHomography H,
                                                   double *__restrict__ models, uint8_t *__restrict__ valid,
                                                   unsigned long long *__restrict__ counters)
{
	const int i = blockIdx.x * blockDim.x + threadIdx.x;
	const int pair = blockIdx.y;
	if (i >= count)
		return;
	const uint32_t *sidx = samples + ((size_t)pair * count + i) * 2;
	const SamplePoints pt{pts + (size_t)pair * n * 4, sidx};
	double line[2][3];
#pragma unroll
	for (int q = 0; q < 2; ++q)
	{
		const double src[3] = {pt(q, 0), pt(q, 1), 1.0}, dst[3] = {pt(q, 2), pt(q, 3), 1.0};
		double proj[3];
#pragma unroll
		for (int r = 0; r < 3; ++r)
			proj[r] = H.h[3 * r] * src[0] + H.h[3 * r + 1] * src[1] + H.h[3 * r + 2] * src[2];
		cross3(proj, dst, line[q]);
	}
	double e[3], F[9];
	cross3(line[0], line[1], e);
	bool ok = !(fabs(e[2]) < 2.220446049250313e-16);
	if (ok)
	{
		const double ex[9] = {0, -e[2], e[1], e[2], 0, -e[0], -e[1], e[0], 0};
#pragma unroll
		for (int r = 0; r < 3; ++r)
#pragma unroll
			for (int c = 0; c < 3; ++c)
				F[3 * r + c] = ex[3 * r] * H.h[c] + ex[3 * r + 1] * H.h[3 + c] + ex[3 * r + 2] * H.h[6 + c];
		ok = gcr_f7::orientation_valid(F, pt, 2);
	}
	double *out = models + ((size_t)pair * count + i) * 9;
	if (ok)
#pragma unroll
		for (int j = 0; j < 9; ++j)
			out[j] = F[j];
	valid[(size_t)pair * count + i] = ok ? 1 : 0;
	count_valid(counters, ok ? 1u : 0u);
}

extern "C" gcr_status gcr_solve_plane_parallax(gcr_ctx *ctx, const uint32_t *d_samples, int count, const double *h_homography,
                                               double *d_models, uint8_t *d_valid)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!ctx->d_pts64 || ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_solve_plane_parallax: no points uploaded");
	if (ctx->cols != 4 || !d_samples || !h_homography || !d_models || !d_valid || count <= 0)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_solve_plane_parallax: bad arguments");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	PlaneHomography H;
	for (int j = 0; j < 9; ++j)
		H.h[j] = h_homography[j];
	k_solve_fpp<<<dim3((count + 127) / 128, ctx->n_pairs), 128, 0, ctx->stream>>>(ctx->d_pts64, ctx->n, d_samples, count, H,
	                                                                             d_models, d_valid, ctx->d_counters);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}

static gcr_status solve_launch(gcr_ctx *ctx, int problem, const GcrSampleSource &src, int count, double *d_models,
                               uint8_t *d_valid)
{
	if (!ctx->d_pts64 || ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_solve_minimal: no points uploaded");
	if (problem < GCR_H || problem > GCR_PNP || gcr_cols_of(problem) != ctx->cols)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_solve_minimal: problem does not match the uploaded points");
	if (!d_models || !d_valid || count <= 0)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_solve_minimal: bad arguments");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	if (src.active) // skipped pairs keep no stale flags
		GCR_CHECK_CUDA(ctx, cudaMemsetAsync(d_valid, 0, (size_t)ctx->n_pairs * count * gcr_max_models(problem), ctx->stream));
	const dim3 grid((count + 127) / 128, ctx->n_pairs);
	switch (problem)
	{
	case GCR_H:
		k_solve_h4<<<dim3((count + H4_SAMPLES - 1) / H4_SAMPLES, ctx->n_pairs), H4_THREADS, 0, ctx->stream>>>(ctx->d_pts64, ctx->n, src, count, d_models, d_valid, ctx->d_counters);
		break;
	case GCR_F:
		GCR_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_solve_f7, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)F7_SMEM));
		if (ctx->tune_carveout >= 0)
			GCR_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_solve_f7, cudaFuncAttributePreferredSharedMemoryCarveout, ctx->tune_carveout));
		k_solve_f7<<<dim3((count + F7_THREADS - 1) / F7_THREADS, ctx->n_pairs), F7_THREADS, F7_SMEM, ctx->stream>>>(
		    ctx->d_pts64, ctx->n, src, count, d_models, d_valid, ctx->d_counters);
		break;
	case GCR_E:
		GCR_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_solve_e5, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)E5_SMEM));
		if (ctx->tune_carveout >= 0)
			GCR_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_solve_e5, cudaFuncAttributePreferredSharedMemoryCarveout, ctx->tune_carveout));
		k_solve_e5<<<dim3((count + E5_SAMPLES_PER_CTA - 1) / E5_SAMPLES_PER_CTA, ctx->n_pairs), E5_THREADS, E5_SMEM, ctx->stream>>>(
		    ctx->d_pts64, ctx->n, src, count, d_models, d_valid, ctx->d_counters);
		break;
	case GCR_PNP:
		k_solve_p3p<<<grid, 128, 0, ctx->stream>>>(ctx->d_pts64, ctx->n, src, count, d_models, d_valid, ctx->d_counters);
		break;
	default:
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_solve_minimal: unknown problem %d", problem);
	}
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}

extern "C" gcr_status gcr_solve_minimal(gcr_ctx *ctx, int problem, const uint32_t *d_samples, int count,
                                        double *d_models, uint8_t *d_valid)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!d_samples)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_solve_minimal: bad arguments");
	GcrSampleSource src;
	src.samples = d_samples;
	src.seed = src.first_index = 0;
	src.sampler = GCR_UNIFORM;
	src.growth = nullptr;
	src.cdf = nullptr;
	src.n = (uint32_t)ctx->n;
	src.active = ctx->d_active_pairs;
	return solve_launch(ctx, problem, src, count, d_models, d_valid);
}

// gcr_round's K1 + K2 in one kernel: the sample is drawn from the counter-based stream inside the solver kernel
// (bit-identical to gcr_sample's output for the same (seed, index, pair)) and never touches HBM.
gcr_status gcr_solve_minimal_generated(gcr_ctx *ctx, int problem, int sampler, uint64_t first_index, int count, uint64_t seed,
                                       double *d_models, uint8_t *d_valid)
{
	if (!ctx)
		return GCR_ERR_ARG;
	const int m = gcr_sample_size(problem);
	if (m == 0 || m > ctx->n)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_round: sample size %d exceeds the %d points (uniform_sampler.h:87-88)", m, ctx->n);
	GcrSampleSource src;
	src.samples = nullptr;
	src.seed = seed;
	src.first_index = first_index;
	src.sampler = sampler;
	src.growth = nullptr;
	src.cdf = nullptr;
	src.n = (uint32_t)ctx->n;
	src.active = ctx->d_active_pairs;
	if (sampler == GCR_PROSAC)
	{
		if (first_index < 1)
			GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_round: PROSAC draws are numbered from k = 1");
		gcr_status st = gcr_ensure_prosac_growth(ctx, m);
		if (st)
			return st;
		src.growth = ctx->d_prosac_growth;
	}
	else if (sampler == GCR_IMPORTANCE)
	{
		if (!ctx->d_importance_cdf || ctx->importance_n != ctx->n)
			GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_round: GCR_IMPORTANCE needs gcr_set_point_probabilities for the %d points", ctx->n);
		src.cdf = ctx->d_importance_cdf;
	}
	else if (sampler != GCR_UNIFORM)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_round: unknown sampler %d", sampler);
	return solve_launch(ctx, problem, src, count, d_models, d_valid);
}
