// k_score.cu -- K3: MSAC / MAGSAC++ scoring of K hypotheses against N correspondences (the kernel that carries the
// headline metric).  Replaces MSACScoringFunction::getScore (scoring_function.h:137-233) and
// MAGSACScoringFunction::getScore (:246-348) of the reference together with the estimators' squaredResidual
// (homography_estimator.h:181-199, fundamental_estimator.h:195-222, perspective_n_point_estimator.h:148-184).
//
// Mapping (B200-first, not a translation of the reference's per-model point loop):
//   * lane = hypothesis.  Each thread keeps R hypotheses in registers for the whole kernel (9 or 12 FP32 each,
//     mapped into the centred/scaled coordinate frame of the FP32 point copy); all 32 lanes of a warp read the SAME
//     point from shared memory (one broadcast LDS.128 per point), so there is no cross-lane reduction at all and no
//     bank conflict; per-hypothesis (score, inliers) accumulate in registers and are written once.
//   * point tiles (256 points, 4 KB / 8 KB) stream through a 4-stage shared-memory ring filled by 1-D bulk TMA
//     (cp.async.bulk ... mbarrier::complete_tx), issued by one elected thread; consumers wait on the stage's mbarrier.
//   * W = 2 variant: two hypotheses per thread are packed into f32x2 registers and evaluated with FFMA2 / FMUL2 /
//     FADD2 (sm_100 packed FP32), reading an operand-duplicated point copy, which halves the issue slots per
//     evaluation on the FMA pipe.
//   * arithmetic is FP32 FMA + one MUFU.RCP per evaluation (projective divide / Sampson denominator): FP32-pipe
//     bound, not a tensor-core contraction.
//   * GCR_SCORE_EXACT_COUNTS: an evaluation whose |r^2 - t^2| falls inside the FP32 error band is re-evaluated in
//     FP64 in the reference's operation order from the FP64 point/model copies (rare, divergent slow path), so inlier
//     counts equal the FP64 reference's.
//   * small K: the grid also splits the points into chunks (blockIdx.y) and a second tiny kernel adds the partials
//     in a fixed order (deterministic; no float atomics).
#include "gcr_internal.cuh"

#include <math.h>

#define TILE 256
#define STAGES 4
#define THREADS 256

// ------------------------------------------------------------------------------------------ mbarrier / TMA ----
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
	                 smem_u32(dst)),
	             "l"(src), "r"(bytes), "r"(smem_u32(bar))
	             : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity)
{
	const uint32_t addr = smem_u32(bar);
	uint32_t ok;
	do
	{
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(ok)
		             : "r"(addr), "r"(parity)
		             : "memory");
	} while (!ok);
}
__device__ __forceinline__ float rcp_approx(float x)
{
	float r;
	asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
	return r;
}

// ------------------------------------------------------------------------------- model -> normalised FP32 ----
// Points were mapped x' = s (x - c) per image (gcr_ctx.cu).  The same map applied to the model, in FP64, then one
// rounding to FP32.  Residuals in the primed frame are s^2 times the pixel-frame residuals.
template <int PROBLEM>
__device__ __forceinline__ void normalise_model(const double *__restrict__ M, const double *__restrict__ nm, float *out)
{
	if (PROBLEM == GCR_PNP)
	{
		// u' = su (u - cu), X' = sX (X - cX):  P' = Tu * [R | sX (R cX + t)]
		const double cu = nm[0], cv = nm[1], su = nm[2], cX = nm[3], cY = nm[4], cZ = nm[5], sX = nm[6];
		double A[12];
		for (int r = 0; r < 3; ++r)
		{
			A[4 * r + 0] = M[4 * r + 0];
			A[4 * r + 1] = M[4 * r + 1];
			A[4 * r + 2] = M[4 * r + 2];
			A[4 * r + 3] = sX * (M[4 * r + 0] * cX + M[4 * r + 1] * cY + M[4 * r + 2] * cZ + M[4 * r + 3]);
		}
		double B[12];
		double mx = 0.0;
		for (int c = 0; c < 4; ++c)
		{
			B[c] = su * (A[c] - cu * A[8 + c]);
			B[4 + c] = su * (A[4 + c] - cv * A[8 + c]);
			B[8 + c] = A[8 + c];
		}
		for (int i = 0; i < 12; ++i)
			mx = fmax(mx, fabs(B[i]));
		const double inv = mx > 0.0 ? 1.0 / mx : 1.0;
		for (int i = 0; i < 12; ++i)
			out[i] = (float)(B[i] * inv);
		return;
	}
	const double c1x = nm[0], c1y = nm[1], c2x = nm[2], c2y = nm[3], s = nm[4], is = 1.0 / nm[4];
	double A[9], B[9];
	for (int r = 0; r < 3; ++r) // A = M * T1^-1
	{
		A[3 * r + 0] = M[3 * r + 0] * is;
		A[3 * r + 1] = M[3 * r + 1] * is;
		A[3 * r + 2] = M[3 * r + 0] * c1x + M[3 * r + 1] * c1y + M[3 * r + 2];
	}
	if (PROBLEM == GCR_H) // H' = T2 * A
	{
		for (int c = 0; c < 3; ++c)
		{
			B[c] = s * (A[c] - c2x * A[6 + c]);
			B[3 + c] = s * (A[3 + c] - c2y * A[6 + c]);
			B[6 + c] = A[6 + c];
		}
	}
	else // F' = T2^-T * A
	{
		for (int c = 0; c < 3; ++c)
		{
			B[c] = A[c] * is;
			B[3 + c] = A[3 + c] * is;
			B[6 + c] = c2x * A[c] + c2y * A[3 + c] + A[6 + c];
		}
	}
	double mx = 0.0;
	for (int i = 0; i < 9; ++i)
		mx = fmax(mx, fabs(B[i]));
	const double inv = mx > 0.0 ? 1.0 / mx : 1.0;
	for (int i = 0; i < 9; ++i)
		out[i] = (float)(B[i] * inv);
}

// squared-threshold in the primed frame and the FP32 error band around it
struct Thresholds
{
	float negT;  // -(t'^2)
	float band;  // |u| < band  => FP64 re-check (u = r'^2 - t'^2)
	float c_idx; // MAGSAC: 1e4 / (2 t'^2)
	float invT;  // 1 / t'^2
	double T64;  // pixel-frame squared threshold, as the reference computes it
	double thr64;
	double s2;   // s^2 (frame scale)
};

template <int PROBLEM, int SCORING>
__device__ __forceinline__ Thresholds make_thresholds(const double *nm, double thr, float band_factor)
{
	Thresholds t;
	const double s = PROBLEM == GCR_PNP ? nm[2] : nm[4];
	double T;
	if (SCORING == GCR_MSAC)
	{
		const double truncated_threshold = 3.0 / 2.0 * thr; // GCRANSAC.h:253
		T = truncated_threshold * truncated_threshold;      // GCRANSAC.h:254
	}
	else
		T = thr * thr;
	t.T64 = T;
	t.thr64 = thr;
	t.s2 = s * s;
	const double Tp = T * s * s;
	t.negT = -(float)Tp;
	t.invT = (float)(1.0 / Tp);
	t.c_idx = (float)(1e4 / (2.0 * Tp));
	// |x'| <= 1 in the primed frame; delta_d ~ band_factor * eps32; |u| error ~ 2 t' delta_d + delta_d^2
	const double dd = band_factor * 1.1920929e-7;
	t.band = (float)(2.0 * sqrt(Tp) * dd + dd * dd);
	return t;
}

// ------------------------------------------------------------------------------------ FP64 slow-path fix-up ---
// Called for one (hypothesis, point) evaluation that landed inside the error band.  Removes the FP32 contribution
// and adds the FP64 one (reference operation order).
template <int PROBLEM, int SCORING>
__device__ __noinline__ void exact_fixup(const double *__restrict__ pts64, const double *__restrict__ model64,
                                         const Thresholds &th, const float *__restrict__ lut, float u, float w32,
                                         float &S, uint32_t &cnt)
{
	constexpr int COLS = PROBLEM == GCR_PNP ? 5 : 4;
	constexpr int ML = PROBLEM == GCR_PNP ? 12 : 9;
	double pt[COLS], M[ML];
	for (int i = 0; i < COLS; ++i)
		pt[i] = pts64[i];
	for (int i = 0; i < ML; ++i)
		M[i] = model64[i];
	const double r2 = gcr_sqres64<PROBLEM>(pt, M);
	if (SCORING == GCR_MSAC)
	{
		if (u < 0.f)
		{
			S -= u;
			cnt -= 1;
		}
		if (r2 < th.T64)
		{
			S += (float)((r2 - th.T64) * th.s2);
			cnt += 1;
		}
	}
	else
	{
		if (u <= 0.f)
		{
			S -= w32;
			cnt -= 1;
		}
		const double residual = sqrt(r2);
		if (!(residual > th.thr64))
		{
			const double squared_residual = residual * residual;
			long long x = llround(1e4 * squared_residual / (th.thr64 * th.thr64 * 2.0));
			if (x > 5000) x = 5000;
			S += residual < 2.220446049250313e-16 ? 1.0f : lut[x];
			cnt += 1;
		}
	}
}

// --------------------------------------------------------------------------------------------- the kernel ----
struct ScoreArgs
{
	const float4 *pts32;
	const double *pts64;
	const double *models;
	const uint8_t *valid;
	const double *norm;
	const float *lut;
	float *scores;
	uint32_t *inliers;
	float *part_scores;  // [n_chunks][n_pairs][k]
	uint32_t *part_inl;
	int n, n_pad, k, n_pairs;
	int tiles_per_chunk, n_chunks;
	double thr;
	float band_factor;
};

// W = 1: scalar FP32, point tile = float4 (x1 y1 x2 y2) [PnP: 2 x float4]
// W = 2: packed f32x2, point tile = 2 x float4 (x1 x1 y1 y1)(x2 x2 y2 y2); hypotheses 2r and 2r+1 share registers
template <int PROBLEM, int SCORING, int R, int W, bool EXACT>
__global__ void __launch_bounds__(THREADS, (R * (PROBLEM == GCR_PNP ? 12 : 9) > 40) ? 2 : 3) k_score(const ScoreArgs a)
{
	constexpr int ML = PROBLEM == GCR_PNP ? 12 : 9;
	constexpr int COLS = PROBLEM == GCR_PNP ? 5 : 4;
	constexpr int F4 = (W == 2 || PROBLEM == GCR_PNP) ? 2 : 1; // float4 per point in the tile
	constexpr uint32_t TILE_BYTES = TILE * F4 * sizeof(float4);
	static_assert(W == 1 || (R % 2 == 0 && PROBLEM != GCR_PNP), "packed variant needs an even R and 4-column data");

	extern __shared__ __align__(128) unsigned char smem_raw[];
	float4(*tiles)[TILE * F4] = reinterpret_cast<float4(*)[TILE * F4]>(smem_raw);
	uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem_raw + STAGES * TILE_BYTES);
	float *s_lut = reinterpret_cast<float *>(smem_raw + STAGES * TILE_BYTES + 64);

	const int pair = blockIdx.z, chunk = blockIdx.y, tid = threadIdx.x;
	const int tile0 = chunk * a.tiles_per_chunk;
	const int n_tiles_total = a.n_pad / TILE;
	int my_tiles = n_tiles_total - tile0;
	if (my_tiles > a.tiles_per_chunk) my_tiles = a.tiles_per_chunk;
	const float4 *src = a.pts32 + ((size_t)pair * a.n_pad + (size_t)tile0 * TILE) * F4;

	if (tid == 0)
	{
#pragma unroll
		for (int s = 0; s < STAGES; ++s)
			mbar_init(&full_bar[s], 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	if (SCORING == GCR_MAGSAC)
		for (int i = tid; i < 5001; i += THREADS)
			s_lut[i] = a.lut[i];
	__syncthreads();
	if (tid == 0)
	{
#pragma unroll
		for (int s = 0; s < STAGES; ++s)
			if (s < my_tiles)
			{
				mbar_expect_tx(&full_bar[s], TILE_BYTES);
				tma_load_1d(&tiles[s][0], src + (size_t)s * TILE * F4, TILE_BYTES, &full_bar[s]);
			}
	}

	// ---- prologue: this thread's R hypotheses, mapped to the primed frame ----
	const double *nm = a.norm + (size_t)pair * 8;
	const Thresholds th = make_thresholds<PROBLEM, SCORING>(nm, a.thr, a.band_factor);
	const int hyp_base = blockIdx.x * (THREADS * R);
	float m[R][ML];
	bool present[R];
#pragma unroll
	for (int r = 0; r < R; ++r)
	{
		const int h = hyp_base + r * THREADS + tid;
		present[r] = h < a.k && (a.valid == nullptr || a.valid[(size_t)pair * a.k + h] != 0);
		if (h < a.k)
			normalise_model<PROBLEM>(a.models + ((size_t)pair * a.k + h) * ML, nm, m[r]);
		else
		{
#pragma unroll
			for (int j = 0; j < ML; ++j)
				m[r][j] = __int_as_float(0x7fc00000);
		}
	}
	float S[R];
	uint32_t cnt[R];
#pragma unroll
	for (int r = 0; r < R; ++r)
	{
		S[r] = 0.f;
		cnt[r] = 0;
	}

	// ---- main loop over point tiles ----
	for (int t = 0; t < my_tiles; ++t)
	{
		const int stage = t % STAGES;
		mbar_wait(&full_bar[stage], (uint32_t)((t / STAGES) & 1));
		const float4 *tp = &tiles[stage][0];
#pragma unroll 2
		for (int i = 0; i < TILE; ++i)
		{
			float u[R], w[R];
			bool amb = false;
			if (W == 1)
			{
				float x1, y1, x2, y2, z3 = 0.f;
				if (PROBLEM == GCR_PNP)
				{
					const float4 qa = tp[2 * i], qb = tp[2 * i + 1];
					x1 = qa.x; y1 = qa.y; x2 = qb.x; y2 = qb.y; z3 = qb.z; // (u v) (X Y Z)
				}
				else
				{
					const float4 q = tp[i];
					x1 = q.x; y1 = q.y; x2 = q.z; y2 = q.w;
				}
#pragma unroll
				for (int r = 0; r < R; ++r)
				{
					const float *h = m[r];
					float r2;
					if (PROBLEM == GCR_H)
					{
						const float t1 = fmaf(h[0], x1, fmaf(h[1], y1, h[2]));
						const float t2 = fmaf(h[3], x1, fmaf(h[4], y1, h[5]));
						const float t3 = fmaf(h[6], x1, fmaf(h[7], y1, h[8]));
						const float inv = rcp_approx(t3);
						const float d1 = fmaf(t1, inv, -x2);
						const float d2 = fmaf(t2, inv, -y2);
						if (SCORING == GCR_MSAC)
							u[r] = fmaf(d2, d2, fmaf(d1, d1, th.negT));
						else
							r2 = fmaf(d2, d2, d1 * d1);
					}
					else if (PROBLEM == GCR_PNP)
					{
						const float px = fmaf(h[0], x2, fmaf(h[1], y2, fmaf(h[2], z3, h[3])));
						const float py = fmaf(h[4], x2, fmaf(h[5], y2, fmaf(h[6], z3, h[7])));
						const float pz = fmaf(h[8], x2, fmaf(h[9], y2, fmaf(h[10], z3, h[11])));
						const float inv = rcp_approx(pz);
						const float du = fmaf(px, inv, -x1);
						const float dv = fmaf(py, inv, -y1);
						if (SCORING == GCR_MSAC)
							u[r] = fmaf(dv, dv, fmaf(du, du, th.negT));
						else
							r2 = fmaf(dv, dv, du * du);
					}
					else
					{
						const float rxc = fmaf(h[0], x2, fmaf(h[3], y2, h[6]));
						const float ryc = fmaf(h[1], x2, fmaf(h[4], y2, h[7]));
						const float rwc = fmaf(h[2], x2, fmaf(h[5], y2, h[8]));
						const float rr = fmaf(x1, rxc, fmaf(y1, ryc, rwc));
						const float rx = fmaf(h[0], x1, fmaf(h[1], y1, h[2]));
						const float ry = fmaf(h[3], x1, fmaf(h[4], y1, h[5]));
						const float den = fmaf(ry, ry, fmaf(rx, rx, fmaf(ryc, ryc, rxc * rxc)));
						const float inv = rcp_approx(den);
						if (SCORING == GCR_MSAC)
							u[r] = fmaf(rr * rr, inv, th.negT);
						else
							r2 = rr * rr * inv;
					}
					if (SCORING == GCR_MSAC)
					{
						S[r] += fminf(u[r], 0.f);
						cnt[r] += __float_as_uint(u[r]) >> 31; // u < 0 (generated NaNs are the positive canonical NaN)
					}
					else
					{
						u[r] = r2 + th.negT;
						int idx = __float2int_rn(r2 * th.c_idx);
						idx = min(max(idx, 0), 5000);
						const float wv = s_lut[idx];
						const bool in = u[r] <= 0.f;
						w[r] = in ? wv : 0.f;
						S[r] += w[r];
						cnt[r] += in ? 1u : 0u;
					}
					if (EXACT)
						amb |= fabsf(u[r]) < th.band;
				}
			}
			else // W == 2: packed f32x2
			{
				const float4 qa = tp[2 * i], qb = tp[2 * i + 1];
				const float2 X1 = make_float2(qa.x, qa.y), Y1 = make_float2(qa.z, qa.w);
				const float2 X2 = make_float2(qb.x, qb.y), Y2 = make_float2(qb.z, qb.w);
				const float2 NEGT = make_float2(th.negT, th.negT);
#pragma unroll
				for (int r = 0; r < R; r += 2)
				{
					// hypotheses r and r+1 side by side
					float2 h[9];
#pragma unroll
					for (int j = 0; j < 9; ++j)
						h[j] = make_float2(m[r][j], m[r + 1][j]);
					float2 uu, r2v;
					if (PROBLEM == GCR_H)
					{
						const float2 t1 = __ffma2_rn(h[0], X1, __ffma2_rn(h[1], Y1, h[2]));
						const float2 t2 = __ffma2_rn(h[3], X1, __ffma2_rn(h[4], Y1, h[5]));
						const float2 t3 = __ffma2_rn(h[6], X1, __ffma2_rn(h[7], Y1, h[8]));
						const float2 inv = make_float2(rcp_approx(t3.x), rcp_approx(t3.y));
						// the duplicated copy stores x2, y2 (positive); d = x2 - t/t3 = fma(-t, inv, x2): use
						// nd = fma(t, inv, -x2) via the negated constants below
						const float2 d1 = __ffma2_rn(t1, inv, make_float2(-X2.x, -X2.y));
						const float2 d2 = __ffma2_rn(t2, inv, make_float2(-Y2.x, -Y2.y));
						if (SCORING == GCR_MSAC)
							uu = __ffma2_rn(d2, d2, __ffma2_rn(d1, d1, NEGT));
						else
							r2v = __ffma2_rn(d2, d2, __fmul2_rn(d1, d1));
					}
					else
					{
						const float2 rxc = __ffma2_rn(h[0], X2, __ffma2_rn(h[3], Y2, h[6]));
						const float2 ryc = __ffma2_rn(h[1], X2, __ffma2_rn(h[4], Y2, h[7]));
						const float2 rwc = __ffma2_rn(h[2], X2, __ffma2_rn(h[5], Y2, h[8]));
						const float2 rr = __ffma2_rn(X1, rxc, __ffma2_rn(Y1, ryc, rwc));
						const float2 rx = __ffma2_rn(h[0], X1, __ffma2_rn(h[1], Y1, h[2]));
						const float2 ry = __ffma2_rn(h[3], X1, __ffma2_rn(h[4], Y1, h[5]));
						const float2 den = __ffma2_rn(ry, ry, __ffma2_rn(rx, rx, __ffma2_rn(ryc, ryc, __fmul2_rn(rxc, rxc))));
						const float2 inv = make_float2(rcp_approx(den.x), rcp_approx(den.y));
						const float2 rr2 = __fmul2_rn(rr, rr);
						if (SCORING == GCR_MSAC)
							uu = __ffma2_rn(rr2, inv, NEGT);
						else
							r2v = __fmul2_rn(rr2, inv);
					}
					if (SCORING == GCR_MSAC)
					{
						u[r] = uu.x;
						u[r + 1] = uu.y;
						const float2 acc = __fadd2_rn(make_float2(S[r], S[r + 1]), make_float2(fminf(uu.x, 0.f), fminf(uu.y, 0.f)));
						S[r] = acc.x;
						S[r + 1] = acc.y;
						cnt[r] += __float_as_uint(uu.x) >> 31;
						cnt[r + 1] += __float_as_uint(uu.y) >> 31;
					}
					else
					{
						const float rs[2] = {r2v.x, r2v.y};
#pragma unroll
						for (int e = 0; e < 2; ++e)
						{
							u[r + e] = rs[e] + th.negT;
							int idx = __float2int_rn(rs[e] * th.c_idx);
							idx = min(max(idx, 0), 5000);
							const float wv = s_lut[idx];
							const bool in = u[r + e] <= 0.f;
							w[r + e] = in ? wv : 0.f;
							S[r + e] += w[r + e];
							cnt[r + e] += in ? 1u : 0u;
						}
					}
					if (EXACT)
						amb |= (fabsf(u[r]) < th.band) | (fabsf(u[r + 1]) < th.band);
				}
			}
			if (EXACT && amb)
			{
				const int pidx = (tile0 + t) * TILE + i;
#pragma unroll
				for (int r = 0; r < R; ++r)
				{
					const int h = hyp_base + r * THREADS + tid;
					if (fabsf(u[r]) < th.band && h < a.k && pidx < a.n)
						exact_fixup<PROBLEM, SCORING>(a.pts64 + ((size_t)pair * a.n + pidx) * COLS,
						                              a.models + ((size_t)pair * a.k + h) * ML, th,
						                              SCORING == GCR_MAGSAC ? s_lut : nullptr, u[r],
						                              SCORING == GCR_MAGSAC ? w[r] : 0.f, S[r], cnt[r]);
				}
			}
		}
		__syncthreads(); // everyone is done with this stage -> refill it
		if (tid == 0 && t + STAGES < my_tiles)
		{
			mbar_expect_tx(&full_bar[stage], TILE_BYTES);
			tma_load_1d(&tiles[stage][0], src + (size_t)(t + STAGES) * TILE * F4, TILE_BYTES, &full_bar[stage]);
		}
	}

	// ---- epilogue ----
#pragma unroll
	for (int r = 0; r < R; ++r)
	{
		const int h = hyp_base + r * THREADS + tid;
		if (h >= a.k)
			continue;
		if (a.n_chunks == 1)
		{
			float score;
			if (!present[r])
				score = -1.f, cnt[r] = 0;
			else if (SCORING == GCR_MSAC)
				score = cnt[r] == 0 ? 0.f : -S[r] * th.invT; // scoring_function.h:221-229
			else
				score = S[r];
			a.scores[(size_t)pair * a.k + h] = score;
			if (a.inliers)
				a.inliers[(size_t)pair * a.k + h] = cnt[r];
		}
		else
		{
			const size_t o = ((size_t)chunk * a.n_pairs + pair) * a.k + h;
			a.part_scores[o] = S[r];
			a.part_inl[o] = cnt[r];
		}
	}
}

template <int PROBLEM, int SCORING>
__global__ void k_score_reduce(const ScoreArgs a)
{
	const int pair = blockIdx.y;
	const int h = blockIdx.x * blockDim.x + threadIdx.x;
	if (h >= a.k)
		return;
	float S = 0.f;
	uint32_t c = 0;
	for (int ch = 0; ch < a.n_chunks; ++ch)
	{
		const size_t o = ((size_t)ch * a.n_pairs + pair) * a.k + h;
		S += a.part_scores[o];
		c += a.part_inl[o];
	}
	const Thresholds th = make_thresholds<PROBLEM, SCORING>(a.norm + (size_t)pair * 8, a.thr, a.band_factor);
	const bool present = a.valid == nullptr || a.valid[(size_t)pair * a.k + h] != 0;
	float score;
	if (!present)
		score = -1.f, c = 0;
	else if (SCORING == GCR_MSAC)
		score = c == 0 ? 0.f : -S * th.invT;
	else
		score = S;
	a.scores[(size_t)pair * a.k + h] = score;
	if (a.inliers)
		a.inliers[(size_t)pair * a.k + h] = c;
}

// ----------------------------------------------------------------------------------------------- dispatch ----
__global__ void k_fill_lut(float *lut)
{
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k > 5000)
		return;
	// Gamma(3/2, x) = sqrt(pi)/2 erfc(sqrt x) + sqrt(x) e^-x  (== gamma_values.cpp stored_gamma_values[k], x = k/1000)
	const double x = k / 1000.0, sx = sqrt(x);
	const double g = 0.88622692545275801365 * erfc(sx) + sx * exp(-x);
	const double gamma_k = 0.0036572608340910764;      // scoring_function.h:288
	const double gamma_value = 0.88622692545275801365; // tgamma(1.5), :285
	lut[k] = (float)((g - gamma_k) / (gamma_value - gamma_k));
}

static gcr_status ensure_lut(gcr_ctx *ctx)
{
	if (ctx->d_magsac_lut)
		return GCR_OK;
	GCR_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_magsac_lut, 5008 * sizeof(float)));
	k_fill_lut<<<(5001 + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_magsac_lut);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}

template <int PROBLEM, int SCORING, int R, int W>
static gcr_status launch_variant(gcr_ctx *ctx, ScoreArgs &a, bool exact)
{
	const int hyp_tiles = (a.k + THREADS * R - 1) / (THREADS * R);
	const int n_tiles = a.n_pad / TILE;
	// fill the machine: aim for >= 2 CTAs per SM slot; split the points when K alone cannot do it
	const long long ctas = (long long)hyp_tiles * a.n_pairs;
	const long long target = 2ll * 3 * ctx->sm_count;
	int n_chunks = 1;
	if (ctas < target)
	{
		n_chunks = (int)((target + ctas - 1) / ctas);
		if (n_chunks > n_tiles) n_chunks = n_tiles;
		if (n_chunks > 64) n_chunks = 64;
	}
	a.tiles_per_chunk = (n_tiles + n_chunks - 1) / n_chunks;
	a.n_chunks = (n_tiles + a.tiles_per_chunk - 1) / a.tiles_per_chunk;
	if (a.n_chunks > 1)
	{
		const size_t per = (size_t)a.n_chunks * a.n_pairs * a.k;
		gcr_status st = gcr_ensure_scratch(ctx, per * (sizeof(float) + sizeof(uint32_t)) + 256);
		if (st)
			return st;
		a.part_scores = (float *)ctx->d_scratch;
		a.part_inl = (uint32_t *)((char *)ctx->d_scratch + ((per * sizeof(float) + 255) & ~(size_t)255));
	}
	a.pts32 = (W == 2) ? ctx->d_pts32dup : ctx->d_pts32;
	const dim3 grid(hyp_tiles, a.n_chunks, a.n_pairs);
	constexpr int F4 = (W == 2 || PROBLEM == GCR_PNP) ? 2 : 1;
	const size_t smem = (size_t)STAGES * TILE * F4 * sizeof(float4) + 64 + (SCORING == GCR_MAGSAC ? 5008 * sizeof(float) : 0);
	gcr_status pst = gcr_profile_begin(ctx);
	if (pst)
		return pst;
	if (exact)
	{
		GCR_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_score<PROBLEM, SCORING, R, W, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		k_score<PROBLEM, SCORING, R, W, true><<<grid, THREADS, smem, ctx->stream>>>(a);
	}
	else
	{
		GCR_CHECK_CUDA(ctx, cudaFuncSetAttribute(k_score<PROBLEM, SCORING, R, W, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		k_score<PROBLEM, SCORING, R, W, false><<<grid, THREADS, smem, ctx->stream>>>(a);
	}
	GCR_LAUNCHED(ctx);
	pst = gcr_profile_end(ctx, (uint64_t)a.k * a.n_pairs * a.n);
	if (pst)
		return pst;
	if (a.n_chunks > 1)
	{
		const dim3 g2((a.k + 255) / 256, a.n_pairs);
		k_score_reduce<PROBLEM, SCORING><<<g2, 256, 0, ctx->stream>>>(a);
		GCR_LAUNCHED(ctx);
	}
	return GCR_OK;
}

template <int PROBLEM, int SCORING>
static gcr_status launch_problem(gcr_ctx *ctx, ScoreArgs &a, bool exact)
{
	// variants: 0 = scalar R4, 1 = scalar R8, 2 = packed R4, 3 = packed R8 (PnP: scalar only)
	int v = ctx->score_variant;
	if (PROBLEM == GCR_PNP)
		return launch_variant<PROBLEM, SCORING, 2, 1>(ctx, a, exact);
	if (v < 0)
		v = 2;
	switch (v)
	{
	case 0: return launch_variant<PROBLEM, SCORING, 4, 1>(ctx, a, exact);
	case 1: return launch_variant<PROBLEM, SCORING, 8, 1>(ctx, a, exact);
	case 3: return launch_variant < PROBLEM, SCORING, 8, PROBLEM == GCR_PNP ? 1 : 2 > (ctx, a, exact);
	default: return launch_variant < PROBLEM, SCORING, 4, PROBLEM == GCR_PNP ? 1 : 2 > (ctx, a, exact);
	}
}

extern "C" gcr_status gcr_score(gcr_ctx *ctx, int problem, int scoring, const double *d_models, const uint8_t *d_valid,
                                int k, double threshold, int flags, float *d_scores, uint32_t *d_inliers)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!ctx->d_pts32 || ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_score: no points uploaded");
	if (problem < GCR_H || problem > GCR_PNP || (scoring != GCR_MSAC && scoring != GCR_MAGSAC))
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_score: bad problem/scoring id");
	if (gcr_cols_of(problem) != ctx->cols)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_score: problem needs %d-column points, context holds %d", gcr_cols_of(problem), ctx->cols);
	if (!d_models || !d_scores || k <= 0 || !(threshold > 0.0))
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_score: need models, scores, k > 0 and threshold > 0");
	if (ctx->n_pairs > 65535)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_score: at most 65535 pairs per context");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	if (scoring == GCR_MAGSAC)
	{
		gcr_status st = ensure_lut(ctx);
		if (st)
			return st;
	}
	ScoreArgs a;
	a.pts32 = nullptr;
	a.pts64 = ctx->d_pts64;
	a.models = d_models;
	a.valid = d_valid;
	a.norm = ctx->d_norm;
	a.lut = ctx->d_magsac_lut;
	a.scores = d_scores;
	a.inliers = d_inliers;
	a.part_scores = nullptr;
	a.part_inl = nullptr;
	a.n = ctx->n;
	a.n_pad = ctx->n_pad;
	a.k = k;
	a.n_pairs = ctx->n_pairs;
	a.tiles_per_chunk = 0;
	a.n_chunks = 1;
	a.thr = threshold;
	a.band_factor = (float)ctx->band_factor;
	const bool exact = (flags & GCR_SCORE_EXACT_COUNTS) != 0;
#define GCR_DISPATCH(P)                                                                   \
	(scoring == GCR_MSAC ? launch_problem<P, GCR_MSAC>(ctx, a, exact) : launch_problem<P, GCR_MAGSAC>(ctx, a, exact))
	switch (problem)
	{
	case GCR_H: return GCR_DISPATCH(GCR_H);
	case GCR_PNP: return GCR_DISPATCH(GCR_PNP);
	default: return GCR_DISPATCH(GCR_F); // F and E share the Sampson residual
	}
#undef GCR_DISPATCH
}

extern "C" gcr_status gcr_score_host(gcr_ctx *ctx, int problem, int scoring, const double *h_models,
                                     const uint8_t *h_valid, int k, double threshold, int flags, float *h_scores,
                                     uint32_t *h_inliers)
{
	if (!ctx || !h_models || !h_scores || k <= 0)
		return GCR_ERR_ARG;
	const int ml = problem == GCR_PNP ? 12 : 9;
	const size_t total = (size_t)k * ctx->n_pairs;
	void *d_m = nullptr, *d_v = nullptr, *d_s = nullptr, *d_i = nullptr;
	gcr_status st = gcr_malloc(ctx, total * ml * sizeof(double), &d_m);
	if (!st) st = gcr_malloc(ctx, total, &d_v);
	if (!st) st = gcr_malloc(ctx, total * sizeof(float), &d_s);
	if (!st) st = gcr_malloc(ctx, total * sizeof(uint32_t), &d_i);
	if (!st) st = gcr_memcpy_h2d(ctx, d_m, h_models, total * ml * sizeof(double));
	if (!st && h_valid) st = gcr_memcpy_h2d(ctx, d_v, h_valid, total);
	if (!st)
		st = gcr_score(ctx, problem, scoring, (const double *)d_m, h_valid ? (const uint8_t *)d_v : nullptr, k, threshold,
		               flags, (float *)d_s, (uint32_t *)d_i);
	if (!st) st = gcr_memcpy_d2h(ctx, h_scores, d_s, total * sizeof(float));
	if (!st && h_inliers) st = gcr_memcpy_d2h(ctx, h_inliers, d_i, total * sizeof(uint32_t));
	cudaStreamSynchronize(ctx->stream);
	cudaFree(d_m);
	cudaFree(d_v);
	cudaFree(d_s);
	cudaFree(d_i);
	return st;
}
