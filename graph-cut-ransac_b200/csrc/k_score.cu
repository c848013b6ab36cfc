// k_score.cu -- K3: MSAC / MAGSAC++ scoring of K hypotheses against N correspondences (the kernel that carries the
// headline metric).  Replaces MSACScoringFunction::getScore (scoring_function.h:137-233) and
// MAGSACScoringFunction::getScore (:246-348) of the reference together with the estimators' squaredResidual
// (homography_estimator.h:181-199, fundamental_estimator.h:195-222, perspective_n_point_estimator.h:148-184).
//
// Pipeline of one gcr_score() call (all on the context's stream, no host synchronisation):
//   k_score_prepare   one thread per model slot: zero the slot's accumulators, append VALID slots to a per-pair compact
//                     list (warp-aggregated atomics) and map their FP64 model into the centred / scaled FP32 frame of
//                     the point copy.  Invalid slots (failed solver, oriented-constraint reject: ~90 % of 7-point
//                     slots) never reach the scoring kernel.
//                     The last CTA to finish also builds the per-pair prefix of hypothesis tiles.
//   k_score_main      the hot kernel.  One persistent CTA (128 threads) per SM slot; the work space (hypothesis tiles x
//                     point granules) is cut into equal contiguous slices, one per CTA ("stream-K"), so there is no
//                     tail wave whatever K, N and the pair count are.  lane = hypothesis: each thread keeps R
//                     hypotheses in registers for a whole segment, packed two-by-two into f32x2 registers and
//                     evaluated with FFMA2 / FMUL2 / FADD2
//                     (sm_100 packed FP32: half the issue slots per evaluation, which leaves the issue port to the
//                     MUFU.RCP, the min/compare/count work and the shared-memory loads).  All 32 lanes of a warp read
//                     the SAME point (broadcast LDS.128 of an operand-duplicated copy), so there is no cross-lane
//                     reduction and no bank conflict.  Point tiles (256 points, 8 KB) stream through a 4-stage
//                     shared-memory ring filled by 1-D bulk TMA (cp.async.bulk + mbarrier complete_tx) issued by one
//                     elected thread.  Per-hypothesis (score, inliers) accumulate in registers (FP32, flushed per
//                     tile into a second-level FP32 sum) and are added once per segment to the slot's accumulators
//                     as fixed-point int64 / uint32 atomics -- integer addition is associative, so the result does not
//                     depend on the order in which segments finish (deterministic, no float atomics).
//   k_score_fixup     GCR_SCORE_EXACT_COUNTS: evaluations whose |r^2 - t^2| fell inside the FP32 error band were left
//                     out by the hot kernel and appended to a work list; they are evaluated here in FP64 in the
//                     reference's operation order, so inlier counts equal the FP64 reference's.
//   k_score_finalize  accumulators -> float score / uint32 inliers per slot (-1 / 0 for invalid slots); inside gcr_round
//                     it also keeps the best key per index class (K4's first stage, see k_select.cu).
//
// Arithmetic is FP32 FMA + one MUFU.RCP per evaluation (projective divide / Sampson denominator): FP32-pipe bound,
// not a tensor-core contraction.
#include "gcr_internal.cuh"

#include <math.h>
#include <stddef.h>
#include <string.h>

#define TILE 256
#define STAGES 4
#define THREADS 128
#define WARPS (THREADS / 32)
#define MAX_TILES_PER_CHUNK 64 /* bounds one FP32 second-level sum to 64 tile sums */
#define AMB_CAPACITY (4u << 20)
#define FIX_ONE 4294967296.0 /* 2^32: fixed-point unit of the score accumulators */

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

// squared threshold in the primed frame and the FP32 error band around it
struct Thresholds
{
	float negT;  // -(t'^2)
	float band;  // |u| < band  => FP64 evaluation instead (u = r'^2 - t'^2)
	float c_idx; // MAGSAC: 1e4 / (2 t'^2)
	double scale; // FP32 partial sum -> fixed point: MSAC -2^32 / t'^2, MAGSAC 2^32
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
	const double Tp = T * s * s;
	t.negT = -(float)Tp;
	t.c_idx = (float)(1e4 / (2.0 * Tp));
	t.scale = SCORING == GCR_MSAC ? -FIX_ONE / Tp : FIX_ONE;
	// |x'| <= 1 in the primed frame; delta_d ~ band_factor * eps32; |u| error ~ 2 t' delta_d + delta_d^2
	const double dd = band_factor * 1.1920929e-7;
	t.band = (float)(2.0 * sqrt(Tp) * dd + dd * dd);
	return t;
}

// ------------------------------------------------------------------------------------------------ plumbing ----
struct ScorePlan
{
	uint32_t amb_count;   // entries appended to the FP64 work list
	uint32_t amb_dropped; // entries that did not fit (reported through gcr_counters_read)
	uint32_t blocks_done; // k_score_prepare: the last CTA to finish builds the hypothesis-tile prefix
	uint32_t pad;
};

struct ScoreArgs
{
	// inputs
	const float4 *pts32; // [n_pairs][n_pad][2] operand-duplicated (PnP: (u v 0 0)(X Y Z 0))
	const double *pts64;
	const double *models; // [n_pairs][k][ML]
	const uint8_t *valid; // [n_pairs][k] or null
	const double *norm;   // [n_pairs][8]
	const float *lut;     // MAGSAC FP32 weight table
	const double *lut64;  // MAGSAC Gamma(1.5, j/1000)
	// workspace
	uint32_t *counts;        // [n_pairs] valid models per pair
	uint32_t *prefix;        // [n_pairs + 1] hypothesis tiles before pair p
	ScorePlan *plan;
	uint32_t *slots;         // [n_pairs][k] compact list -> slot
	float *m32;              // [n_pairs][k][ML] normalised FP32 models, compact order
	unsigned long long *accS; // [n_pairs][k] fixed-point score accumulators (by slot)
	uint32_t *accC;          // [n_pairs][k] inlier counters (by slot)
	unsigned long long *amb; // [AMB_CAPACITY] (global slot << 32) | point index
	unsigned long long *counters;
	// outputs
	float *scores;
	uint32_t *inliers;
	unsigned long long *pairbest; // optional [n_pairs]: finalize also keeps the pair's best FP32 key (score << 32 | ~slot)
	int n, n_pad, k, n_pairs;
	int hyps_per_item, grid;
	double thr;
	float band_factor;
	int no_counts; // the main kernel ran without inlier counting: accC is all zero, finalize must not read it as "no inliers"
	unsigned long long *trace; // optional [grid][4]: SM id, first / last globaltimer tick, segments of each K3 CTA
};

// ------------------------------------------------------------------------------------------------- prepare ----
// exclusive prefix of hypothesis tiles over the pairs (prefix[n_pairs] = total), by one CTA of 256 threads
__device__ void plan_prefix(const ScoreArgs &a)
{
	__shared__ uint32_t warp_tot[8];
	__shared__ uint32_t carry;
	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
	if (tid == 0)
		carry = 0;
	__syncthreads();
	for (int base = 0; base < a.n_pairs; base += 256)
	{
		const int p = base + tid;
		const uint32_t tiles = p < a.n_pairs ? (__ldcg(&a.counts[p]) + a.hyps_per_item - 1) / a.hyps_per_item : 0u;
		uint32_t x = tiles;
#pragma unroll
		for (int o = 1; o < 32; o <<= 1)
		{
			const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
			if (lane >= o) x += y;
		}
		if (lane == 31)
			warp_tot[warp] = x;
		__syncthreads();
		uint32_t before = carry;
		for (int w = 0; w < warp; ++w)
			before += warp_tot[w];
		if (p < a.n_pairs)
			a.prefix[p] = before + x - tiles; // exclusive
		__syncthreads();
		if (tid == 255)
			carry = before + x;
		__syncthreads();
	}
	if (tid == 0)
		a.prefix[a.n_pairs] = carry;
}

template <int PROBLEM>
__global__ void __launch_bounds__(256) k_score_prepare(const ScoreArgs a)
{
	constexpr int ML = PROBLEM == GCR_PNP ? 12 : 9;
	__shared__ uint32_t s_last;
	const int pair = blockIdx.y;
	const int slot = blockIdx.x * blockDim.x + threadIdx.x;

	const size_t g = (size_t)pair * a.k + slot;
	const bool in_range = slot < a.k;
	const bool ok = in_range && (a.valid == nullptr || a.valid[g] != 0);
	if (ok) // (only the slots that will be scored: finalize does not read the accumulators of the others)
	{
		a.accS[g] = 0ull;
		a.accC[g] = 0u;
	}
	// positions in the pair's compact list: warp ballots -> one shared prefix per CTA -> ONE global atomic per CTA
	// (a per-warp atomic on the same counter serialises 24576 requests at the L2 for a 786432-slot pair)
	__shared__ uint32_t s_warp_base[8];
	__shared__ uint32_t s_cta_base;
	const uint32_t ballot = __ballot_sync(0xffffffffu, ok);
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	if (lane == 0)
		s_warp_base[warp] = (uint32_t)__popc(ballot);
	__syncthreads();
	if (threadIdx.x == 0)
	{
		uint32_t run = 0;
		for (int w = 0; w < 8; ++w)
		{
			const uint32_t c = s_warp_base[w];
			s_warp_base[w] = run;
			run += c;
		}
		s_cta_base = run ? atomicAdd(&a.counts[pair], run) : 0u;
	}
	__syncthreads();
	if (ok)
	{
		const uint32_t pos = s_cta_base + s_warp_base[warp] + __popc(ballot & ((1u << lane) - 1u));
		a.slots[(size_t)pair * a.k + pos] = (uint32_t)slot;
		float m[ML];
		normalise_model<PROBLEM>(a.models + g * ML, a.norm + (size_t)pair * 8, m);
		float *o = a.m32 + ((size_t)pair * a.k + pos) * ML;
#pragma unroll
		for (int j = 0; j < ML; ++j)
			o[j] = m[j];
	}
	// the last CTA to get here sees every pair's final count and builds the hypothesis-tile prefix (saves a launch)
	__threadfence();
	__syncthreads();
	if (threadIdx.x == 0)
		s_last = atomicAdd(&a.plan->blocks_done, 1u) == gridDim.x * gridDim.y - 1 ? 1u : 0u;
	__syncthreads();
	if (s_last)
		plan_prefix(a);
}

// ----------------------------------------------------------------------------------- per-evaluation maths ----
// packed f32x2 arithmetic, or (SC = true: experiment / fallback for toolchains without FFMA2) the same on two scalars
template <bool SC> __device__ __forceinline__ float2 pfma(float2 a, float2 b, float2 c)
{
	if (SC) return make_float2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y));
	return __ffma2_rn(a, b, c);
}
template <bool SC> __device__ __forceinline__ float2 pmul(float2 a, float2 b)
{
	if (SC) return make_float2(a.x * b.x, a.y * b.y);
	return __fmul2_rn(a, b);
}
template <bool SC> __device__ __forceinline__ float2 padd(float2 a, float2 b)
{
	if (SC) return make_float2(a.x + b.x, a.y + b.y);
	return __fadd2_rn(a, b);
}

// u = r'^2 - t'^2 (MSAC) or r'^2 (MAGSAC) for two hypotheses side by side; point operands are duplicated pairs
// (Tried for H and dropped: ONE MUFU.RCP of the product per packed pair, (1 / a, 1 / b) = (b, a) / (a b).  The XU pipe is not
// what binds -- every instruction costs the scheduler dispatch cycles, and 1 MUFU + 3 FMUL is more of them than 2 MUFU:
// h4k 0.571 -> 0.550 of FP32 peak.)
template <int PROBLEM, int SCORING, bool SC = false>
__device__ __forceinline__ float2 eval2(const float2 (&h)[9], const float2 X1, const float2 Y1, const float2 X2,
                                        const float2 Y2, const float2 NEGT)
{
	if (PROBLEM == GCR_H)
	{
		const float2 t1 = pfma<SC>(h[0], X1, pfma<SC>(h[1], Y1, h[2]));
		const float2 t2 = pfma<SC>(h[3], X1, pfma<SC>(h[4], Y1, h[5]));
		const float2 t3 = pfma<SC>(h[6], X1, pfma<SC>(h[7], Y1, h[8]));
		const float2 inv = make_float2(rcp_approx(t3.x), rcp_approx(t3.y));
		// the caller hands over -x2, -y2 (negated once per point), so d = t/t3 - x2 is one FMA
		const float2 d1 = pfma<SC>(t1, inv, X2);
		const float2 d2 = pfma<SC>(t2, inv, Y2);
		if (SCORING == GCR_MSAC)
			return pfma<SC>(d2, d2, pfma<SC>(d1, d1, NEGT));
		return pfma<SC>(d2, d2, pmul<SC>(d1, d1));
	}
	else
	{
		const float2 rxc = pfma<SC>(h[0], X2, pfma<SC>(h[3], Y2, h[6]));
		const float2 ryc = pfma<SC>(h[1], X2, pfma<SC>(h[4], Y2, h[7]));
		const float2 rwc = pfma<SC>(h[2], X2, pfma<SC>(h[5], Y2, h[8]));
		const float2 rr = pfma<SC>(X1, rxc, pfma<SC>(Y1, ryc, rwc));
		const float2 rx = pfma<SC>(h[0], X1, pfma<SC>(h[1], Y1, h[2]));
		const float2 ry = pfma<SC>(h[3], X1, pfma<SC>(h[4], Y1, h[5]));
		const float2 den = pfma<SC>(ry, ry, pfma<SC>(rx, rx, pfma<SC>(ryc, ryc, pmul<SC>(rxc, rxc))));
		const float2 inv = make_float2(rcp_approx(den.x), rcp_approx(den.y));
		const float2 rr2 = pmul<SC>(rr, rr);
		if (SCORING == GCR_MSAC)
			return pfma<SC>(rr2, inv, NEGT);
		return pmul<SC>(rr2, inv);
	}
}

// PnP, scalar: point = (u' v' . .)(X' Y' Z' .)
template <int SCORING>
__device__ __forceinline__ float eval_pnp(const float (&h)[12], const float4 qa, const float4 qb, const float negT)
{
	const float px = fmaf(h[0], qb.x, fmaf(h[1], qb.y, fmaf(h[2], qb.z, h[3])));
	const float py = fmaf(h[4], qb.x, fmaf(h[5], qb.y, fmaf(h[6], qb.z, h[7])));
	const float pz = fmaf(h[8], qb.x, fmaf(h[9], qb.y, fmaf(h[10], qb.z, h[11])));
	const float inv = rcp_approx(pz);
	const float du = fmaf(px, inv, -qa.x);
	const float dv = fmaf(py, inv, -qa.y);
	if (SCORING == GCR_MSAC)
		return fmaf(dv, dv, fmaf(du, du, negT));
	return fmaf(dv, dv, du * du);
}

// MAGSAC weight of one evaluation: LUT index with the reference's integer rounding (scoring_function.h:320-331)
__device__ __forceinline__ float magsac_weight(const float *s_lut, float r2, float c_idx)
{
	// round-to-nearest via the 1.5 * 2^23 trick (FMA pipe + one integer subtract instead of F2I on the XU pipe)
	const float t = fmaf(r2, c_idx, 12582912.0f);
	const uint32_t idx = min((uint32_t)(__float_as_int(t) - 0x4B400000), 5000u);
	return s_lut[idx];
}

__device__ __forceinline__ void amb_record(const ScoreArgs &a, size_t gslot, int pidx)
{
	const uint32_t pos = atomicAdd(&a.plan->amb_count, 1u);
	if (pos < AMB_CAPACITY)
		a.amb[pos] = ((unsigned long long)gslot << 32) | (unsigned long long)(uint32_t)pidx;
	else
		atomicAdd(&a.plan->amb_dropped, 1u);
}

// --------------------------------------------------------------------------------------------- the kernel ----
// Work decomposition ("stream-K" over hypothesis tiles x points): the work space is the concatenation, over all
// hypothesis tiles (THREADS * R hypotheses of one pair), of that pair's points in granules of GRAN; CTA b of G owns
// the contiguous slice [b W / G, (b + 1) W / G) of it and walks it as a few segments (hypothesis tile, point range).
// Every CTA gets the same amount of work to within one granule whatever K, N and the pair count are -- no tail wave
// and no scheduler atomics -- and partial sums meet in the integer accumulators.
#define GRAN 32

// COUNT = false: the per-model inlier counts are not wanted (gcr_round: the winner's exact count comes from K4) -- the
// hot loop sheds its two LEA.HI per pair of evaluations; every instruction of that loop costs the scheduler two
// dispatch cycles (packed FP32, half-rate ALU / XU pipes), so 8 of ~100 instructions are ~8 % of the kernel.
template <int PROBLEM, int SCORING, int R, bool EXACT, bool SC, bool COUNT = true>
__global__ void __launch_bounds__(THREADS, R >= 6 ? 4 : 5) k_score_main(const ScoreArgs a)
{
	constexpr int ML = PROBLEM == GCR_PNP ? 12 : 9;
	constexpr bool PACKED = PROBLEM != GCR_PNP;
	constexpr int NQ = PACKED ? R / 2 : R; // register groups per thread
	constexpr int HQ = PACKED ? 2 : 1;     // hypotheses per group
	constexpr int NQP = PACKED ? NQ : 1, NQS = PACKED ? 1 : NQ;
	static_assert(!PACKED || R % 2 == 0, "packed variant needs an even R");

	extern __shared__ __align__(128) unsigned char smem_raw[];
	float4(*tiles)[TILE * 2] = reinterpret_cast<float4(*)[TILE * 2]>(smem_raw);
	uint64_t *full_bar = reinterpret_cast<uint64_t *>(smem_raw + STAGES * TILE * 2 * sizeof(float4));
	float *s_lut = reinterpret_cast<float *>(smem_raw + STAGES * TILE * 2 * sizeof(float4) + 64);

	const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
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

	unsigned long long t_begin = 0, n_segments = 0;
	if (a.trace && tid == 0)
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
	const unsigned long long n_gran = (unsigned long long)(a.n_pad / GRAN);
	const unsigned long long W = (unsigned long long)a.prefix[a.n_pairs] * n_gran;
	// Static stream-K split.  (A dynamic chunk queue was tried after gcr_set_score_trace showed the 5 CTAs of an SM
	// finishing one after the other, 76 ... 269 us for equal slices -- the warp schedulers favour the older CTAs.  With
	// the queue every CTA ended within 10 % of the others and the kernel took exactly as long: the FMA pipe runs at
	// the same rate whether 5 warps per scheduler are resident or 1, so the staircase costs nothing.)
	unsigned long long pos = W * blockIdx.x / gridDim.x;
	const unsigned long long pos_end = W * (blockIdx.x + 1ull) / gridDim.x;
	uint32_t ring = 0; // ring tiles this CTA has consumed so far: stage = ring % STAGES, parity = (ring / STAGES) & 1
	const float qnan = __int_as_float(0x7fc00000);

	while (pos < pos_end)
	{
		// ---- segment: hypothesis tile ht, granules [g0, g1) of its pair's points ----
		const uint32_t ht = (uint32_t)(pos / n_gran);
		const uint32_t g0 = (uint32_t)(pos % n_gran);
		uint32_t g1 = (uint32_t)n_gran;
		if (pos_end - pos < (unsigned long long)(g1 - g0)) g1 = g0 + (uint32_t)(pos_end - pos);
		pos += g1 - g0;
		int lo = 0, hi = a.n_pairs; // prefix[lo] <= ht < prefix[hi]
		while (hi - lo > 1)
		{
			const int mid = (lo + hi) >> 1;
			if (a.prefix[mid] <= ht) lo = mid;
			else hi = mid;
		}
		const int pair = lo;
		const int count = (int)a.counts[pair];
		const int p0 = (int)g0 * GRAN, p1 = (int)g1 * GRAN;
		const int seg_tiles = (p1 - p0 + TILE - 1) / TILE;
		const float4 *src = a.pts32 + ((size_t)pair * a.n_pad + p0) * 2;
		if (tid == 0)
		{
#pragma unroll
			for (int s = 0; s < STAGES; ++s)
				if (s < seg_tiles)
				{
					const uint32_t st = (ring + s) % STAGES;
					const int np = min(TILE, p1 - p0 - s * TILE);
					mbar_expect_tx(&full_bar[st], (uint32_t)np * 32u);
					tma_load_1d(&tiles[st][0], src + (size_t)s * TILE * 2, (uint32_t)np * 32u, &full_bar[st]);
				}
		}

		// ---- prologue: this thread's R hypotheses (compact order; a warp owns 32 R consecutive ones, lane = hypothesis) ----
		// (Hoisting this into k_score_prepare -- one 24-byte record per pair, loaded here -- was measured and dropped: the
		// FP64 sqrt / divides overlap the first TMA tile in flight, while the dependent global load delayed every segment:
		// K3 0.251 -> 0.277 ms on configs[1], 0.972 -> 0.984 ms on the batched workload, identical inner loops.)
		const Thresholds th = make_thresholds<PROBLEM, SCORING>(a.norm + (size_t)pair * 8, a.thr, a.band_factor);
		const int hyp0 = (int)(ht - a.prefix[pair]) * (THREADS * R) + warp * (32 * R) + lane; // + 32 * (HQ * q + e)
		const float *mbase = a.m32 + (size_t)pair * a.k * ML;
		// warps whose block of hypotheses lies beyond the pair's list idle through the segment; inside an active warp every
		// register group is evaluated (absent hypotheses are NaN models and contribute nothing), which keeps all the
		// evaluation chains of one loop iteration in ONE basic block for the scheduler to interleave
		const bool warp_active = (hyp0 - lane) < count;

		float2 h2[NQP][9];
		float h1[NQS][12];
#pragma unroll
		for (int q = 0; q < NQ; ++q)
		{
			if (PACKED)
			{
				const int ja = hyp0 + 32 * (2 * q), jb = ja + 32;
#pragma unroll
				for (int j = 0; j < 9; ++j)
				{
					h2[q][j].x = ja < count ? mbase[(size_t)ja * 9 + j] : qnan;
					h2[q][j].y = jb < count ? mbase[(size_t)jb * 9 + j] : qnan;
				}
			}
			else
			{
				const int ja = hyp0 + 32 * q;
#pragma unroll
				for (int j = 0; j < 12; ++j)
					h1[q][j] = ja < count ? mbase[(size_t)ja * 12 + j] : qnan;
			}
		}
		float2 S2[NQP], T2[NQP]; // packed: second-level sums / per-tile sums
		float S1[NQS], T1[NQS];  // scalar (PnP)
		uint32_t cnt[R];
#pragma unroll
		for (int q = 0; q < NQP; ++q) S2[q] = make_float2(0.f, 0.f), T2[q] = make_float2(0.f, 0.f);
#pragma unroll
		for (int q = 0; q < NQS; ++q) S1[q] = 0.f, T1[q] = 0.f;
#pragma unroll
		for (int r = 0; r < R; ++r)
			cnt[r] = 0;
		const float2 NEGT = make_float2(th.negT, th.negT);
		// H evaluates d = t / t3 - x2: the second-image operands are negated once per point
		const float sg = PROBLEM == GCR_H ? -1.f : 1.f;

		// fixed-point flush of the register accumulators into the slot accumulators (segment end / every 64 tiles)
		auto flush = [&]() {
#pragma unroll
			for (int q = 0; q < NQ; ++q)
#pragma unroll
				for (int e = 0; e < HQ; ++e)
				{
					const int j = hyp0 + 32 * (HQ * q + e);
					if (j < count)
					{
						const float S = PACKED ? (e == 0 ? S2[q].x : S2[q].y) : S1[q];
						const size_t g = (size_t)pair * a.k + a.slots[(size_t)pair * a.k + j];
						const long long fx = __double2ll_rn((double)S * th.scale);
						if (fx != 0)
							atomicAdd(&a.accS[g], (unsigned long long)fx);
						if (cnt[HQ * q + e] != 0)
							atomicAdd(&a.accC[g], cnt[HQ * q + e]);
					}
					cnt[HQ * q + e] = 0;
				}
#pragma unroll
			for (int q = 0; q < NQP; ++q) S2[q] = make_float2(0.f, 0.f);
#pragma unroll
			for (int q = 0; q < NQS; ++q) S1[q] = 0.f;
		};

		// ---- main loop over the segment's ring tiles ----
		for (int t = 0; t < seg_tiles; ++t)
		{
			const uint32_t st = (ring + t) % STAGES;
			const int np = min(TILE, p1 - p0 - t * TILE); // multiple of GRAN
			mbar_wait(&full_bar[st], ((ring + t) / STAGES) & 1u);
			const float4 *tp = &tiles[st][0];
			if (warp_active)
			{
				// two points per iteration, the next pair of points prefetched into registers while these are evaluated
				float4 na0 = tp[0], nb0 = tp[1], na1 = tp[2], nb1 = tp[3];
				for (int i = 0; i < np; i += 2)
				{
					const float4 qa[2] = {na0, na1}, qb[2] = {nb0, nb1};
					const int inext = min(i + 2, np - 2);
					na0 = tp[2 * inext]; nb0 = tp[2 * inext + 1]; na1 = tp[2 * inext + 2]; nb1 = tp[2 * inext + 3];
					float mn = 3.0e38f; // min |u| over this thread's evaluations of the two points (EXACT)
#pragma unroll
					for (int pp = 0; pp < 2; ++pp)
					{
						if (PACKED)
						{
							const float2 X1 = make_float2(qa[pp].x, qa[pp].y), Y1 = make_float2(qa[pp].z, qa[pp].w);
							const float2 X2 = make_float2(sg * qb[pp].x, sg * qb[pp].y), Y2 = make_float2(sg * qb[pp].z, sg * qb[pp].w);
#pragma unroll
							for (int q = 0; q < NQP; ++q)
							{
								const float2 v = eval2<PROBLEM, SCORING, SC>(h2[q], X1, Y1, X2, Y2, NEGT);
								if (SCORING == GCR_MSAC)
								{
									T2[q] = padd<SC>(T2[q], make_float2(fminf(v.x, 0.f), fminf(v.y, 0.f)));
									if (COUNT)
									{
										cnt[2 * q] += __float_as_uint(v.x) >> 31; // u < 0 (generated NaNs are the positive canonical NaN)
										cnt[2 * q + 1] += __float_as_uint(v.y) >> 31;
									}
									if (EXACT) mn = fminf(mn, fminf(fabsf(v.x), fabsf(v.y)));
								}
								else
								{
									const float ux = v.x + th.negT, uy = v.y + th.negT;
									const float wx = magsac_weight(s_lut, v.x, th.c_idx), wy = magsac_weight(s_lut, v.y, th.c_idx);
									const bool inx = ux <= 0.f, iny = uy <= 0.f;
									T2[q] = padd<SC>(T2[q], make_float2(inx ? wx : 0.f, iny ? wy : 0.f));
									if (COUNT)
									{
										cnt[2 * q] += inx ? 1u : 0u;
										cnt[2 * q + 1] += iny ? 1u : 0u;
									}
									if (EXACT) mn = fminf(mn, fminf(fabsf(ux), fabsf(uy)));
								}
							}
						}
						else
						{
#pragma unroll
							for (int q = 0; q < NQS; ++q)
							{
								const float v = eval_pnp<SCORING>(h1[q], qa[pp], qb[pp], th.negT);
								if (SCORING == GCR_MSAC)
								{
									T1[q] += fminf(v, 0.f);
									if (COUNT) cnt[q] += __float_as_uint(v) >> 31;
									if (EXACT) mn = fminf(mn, fabsf(v));
								}
								else
								{
									const float u = v + th.negT;
									const float w = magsac_weight(s_lut, v, th.c_idx);
									const bool in = u <= 0.f;
									T1[q] += in ? w : 0.f;
									if (COUNT) cnt[q] += in ? 1u : 0u;
									if (EXACT) mn = fminf(mn, fabsf(u));
								}
							}
						}
					}
					if (EXACT && mn < th.band)
					{
						// rare: take the FP32 contribution of every in-band evaluation back out and queue it for FP64.
						// Recomputing v here gives the same bits as above (same instructions on the same operands).
#pragma unroll
						for (int pp = 0; pp < 2; ++pp)
						{
							const int pidx = p0 + t * TILE + i + pp;
							const float2 X1 = make_float2(qa[pp].x, qa[pp].y), Y1 = make_float2(qa[pp].z, qa[pp].w);
							const float2 X2 = make_float2(sg * qb[pp].x, sg * qb[pp].y), Y2 = make_float2(sg * qb[pp].z, sg * qb[pp].w);
#pragma unroll
							for (int q = 0; q < NQ; ++q)
							{
								float vv[2];
								if (PACKED)
								{
									const float2 v = eval2<PROBLEM, SCORING, SC>(h2[PACKED ? q : 0], X1, Y1, X2, Y2, NEGT);
									vv[0] = v.x;
									vv[1] = v.y;
								}
								else
									vv[0] = vv[1] = eval_pnp<SCORING>(h1[PACKED ? 0 : q], qa[pp], qb[pp], th.negT);
#pragma unroll
								for (int e = 0; e < HQ; ++e)
								{
									const float u = SCORING == GCR_MSAC ? vv[e] : vv[e] + th.negT;
									const int j = hyp0 + 32 * (HQ * q + e);
									if (fabsf(u) < th.band && j < count && pidx < a.n)
									{
										float back;
										uint32_t was_in;
										if (SCORING == GCR_MSAC)
										{
											back = fminf(u, 0.f);
											was_in = __float_as_uint(u) >> 31;
										}
										else
										{
											was_in = u <= 0.f ? 1u : 0u;
											back = was_in ? magsac_weight(s_lut, vv[e], th.c_idx) : 0.f;
										}
										if (PACKED)
										{
											if (e == 0) T2[PACKED ? q : 0].x -= back;
											else T2[PACKED ? q : 0].y -= back;
										}
										else
											T1[PACKED ? 0 : q] -= back;
										cnt[HQ * q + e] -= was_in;
										amb_record(a, (size_t)pair * a.k + a.slots[(size_t)pair * a.k + j], pidx);
									}
								}
							}
						}
					}
				}
				// second-level sum: one rounding per tile instead of one per point on the long accumulator
#pragma unroll
				for (int q = 0; q < NQP; ++q)
				{
					S2[q] = padd<SC>(S2[q], T2[q]);
					T2[q] = make_float2(0.f, 0.f);
				}
#pragma unroll
				for (int q = 0; q < NQS; ++q)
				{
					S1[q] += T1[q];
					T1[q] = 0.f;
				}
				if ((t & (MAX_TILES_PER_CHUNK - 1)) == MAX_TILES_PER_CHUNK - 1)
					flush(); // third level: bounds every FP32 sum to 64 tile sums whatever N is
			}
			__syncthreads(); // everyone is done with this stage -> refill it
			if (tid == 0 && t + STAGES < seg_tiles)
			{
				const int npn = min(TILE, p1 - p0 - (t + STAGES) * TILE);
				mbar_expect_tx(&full_bar[st], (uint32_t)npn * 32u);
				tma_load_1d(&tiles[st][0], src + (size_t)(t + STAGES) * TILE * 2, (uint32_t)npn * 32u, &full_bar[st]);
			}
		}
		ring += (uint32_t)seg_tiles;
		if (warp_active)
			flush();
		++n_segments;
	}
	if (a.trace && tid == 0)
	{
		unsigned long long t_end;
		uint32_t smid;
		asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
		asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
		a.trace[4 * blockIdx.x + 0] = smid;
		a.trace[4 * blockIdx.x + 1] = t_begin;
		a.trace[4 * blockIdx.x + 2] = t_end;
		a.trace[4 * blockIdx.x + 3] = n_segments;
	}
}

// ------------------------------------------------------------------------------------ FP64 slow-path fix-up ---
template <int PROBLEM, int SCORING>
__global__ void __launch_bounds__(256) k_score_fixup(const ScoreArgs a)
{
	constexpr int COLS = PROBLEM == GCR_PNP ? 5 : 4;
	constexpr int ML = PROBLEM == GCR_PNP ? 12 : 9;
	uint32_t total = a.plan->amb_count;
	if (total > AMB_CAPACITY) total = AMB_CAPACITY;
	for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < total; e += gridDim.x * blockDim.x)
	{
		const unsigned long long ent = a.amb[e];
		const size_t g = (size_t)(ent >> 32);
		const uint32_t pidx = (uint32_t)(ent & 0xffffffffull);
		const size_t pair = g / (size_t)a.k;
		double pt[COLS], M[ML];
		const double *p = a.pts64 + (pair * a.n + pidx) * COLS;
		for (int i = 0; i < COLS; ++i) pt[i] = p[i];
		for (int i = 0; i < ML; ++i) M[i] = a.models[g * ML + i];
		const double r2 = gcr_sqres64<PROBLEM>(pt, M);
		if (SCORING == GCR_MSAC)
		{
			const double truncated_threshold = 3.0 / 2.0 * a.thr;
			const double T = truncated_threshold * truncated_threshold;
			if (r2 < T) // scoring_function.h:163
			{
				atomicAdd(&a.accS[g], (unsigned long long)__double2ll_rn((T - r2) / T * FIX_ONE));
				atomicAdd(&a.accC[g], 1u);
			}
		}
		else
		{
			const double residual = sqrt(r2);
			if (!(residual > a.thr)) // scoring_function.h:304
			{
				double w = 1.0; // weight / weight_zero
				if (!(residual < 2.220446049250313e-16))
				{
					const double squared_residual = residual * residual;
					unsigned long long x = (unsigned long long)round(1e4 * squared_residual / (a.thr * a.thr * 2.0));
					if (x > 5000) x = 5000;
					const double gamma_k = 0.0036572608340910764;
					w = (a.lut64[x] - gamma_k) / (0.88622692545275801365 - gamma_k);
				}
				atomicAdd(&a.accS[g], (unsigned long long)__double2ll_rn(w * FIX_ONE));
				atomicAdd(&a.accC[g], 1u);
			}
		}
	}
}

// ------------------------------------------------------------------------------------------------ finalize ----
template <int SCORING>
__global__ void __launch_bounds__(256) k_score_finalize(const ScoreArgs a)
{
	__shared__ unsigned long long s_warp_best[8];
	const int pair = blockIdx.y;
	const int slot = blockIdx.x * blockDim.x + threadIdx.x;
	if (slot == 0 && pair == 0 && a.plan->amb_count)
	{
		atomicAdd(&a.counters[1], (unsigned long long)a.plan->amb_dropped);
		atomicAdd(&a.counters[2], (unsigned long long)a.plan->amb_count);
	}
	unsigned long long key = 0ull;
	if (slot < a.k)
	{
		const size_t g = (size_t)pair * a.k + slot;
		const bool present = a.valid == nullptr || a.valid[g] != 0;
		const uint32_t c = present ? a.accC[g] : 0u;
		float score;
		if (!present)
			score = -1.f;
		else if (SCORING == GCR_MSAC && c == 0 && !a.no_counts)
			score = 0.f; // scoring_function.h:221-222 (without counts: the sum over no inliers is 0 as well)
		else
			score = (float)((double)(long long)a.accS[g] * (1.0 / FIX_ONE));
		a.scores[g] = score;
		if (a.inliers)
			a.inliers[g] = present ? c : 0u;
		if (score > 0.f)
			key = ((unsigned long long)gcr_ordered_u32(score) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)slot);
	}
	if (a.pairbest) // K4's first stage, fused: the pair's best FP32 key (the anchor of k_select_exact's candidate margin)
	{
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
		{
			const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
			if (other > key) key = other;
		}
		if ((threadIdx.x & 31) == 0)
			s_warp_best[threadIdx.x >> 5] = key;
		__syncthreads();
		if (threadIdx.x == 0)
		{
			for (int w = 1; w < 8; ++w)
				if (s_warp_best[w] > key) key = s_warp_best[w];
			// most CTAs cannot raise the maximum: look before the (contended) atomic
			if (key > __ldcg(&a.pairbest[pair]))
				atomicMax(&a.pairbest[pair], key);
		}
	}
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
	gcr_status st = gcr_ensure_magsac_lut64(ctx);
	if (st || ctx->d_magsac_lut)
		return st;
	GCR_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_magsac_lut, 5008 * sizeof(float)));
	k_fill_lut<<<(5001 + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_magsac_lut);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}

// workspace carve-up (one allocation, grown on demand)
static gcr_status ensure_score_ws(gcr_ctx *ctx, ScoreArgs &a, int ml)
{
	const size_t PK = (size_t)a.n_pairs * a.k;
	size_t off = 0;
	auto take = [&](size_t bytes) {
		const size_t o = off;
		off += (bytes + 255) & ~(size_t)255;
		return o;
	};
	const size_t o_head = take(sizeof(ScorePlan) + (size_t)a.n_pairs * sizeof(uint32_t)); // plan | counts: one memset
	const size_t o_prefix = take(((size_t)a.n_pairs + 1) * sizeof(uint32_t));

	const size_t o_slots = take(PK * sizeof(uint32_t));
	const size_t o_m32 = take(PK * ml * sizeof(float));
	const size_t o_accS = take(PK * sizeof(unsigned long long));
	const size_t o_accC = take(PK * sizeof(uint32_t));
	// The workspace grows geometrically (at least 2 x, from 4 MiB) and the 32 MiB re-check list lives in a block of its own
	// that is allocated once: cudaFree / cudaMalloc stall every stream of the process, and a host driver's batches come
	// in a few sizes per run (1 x / 4 x / 16 x the configured one) -- one reallocation per size and context showed up as
	// 70 ms stalls of every thread of a multi-context pipeline.
	if (!ctx->d_score_amb)
		GCR_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_score_amb, (size_t)AMB_CAPACITY * sizeof(unsigned long long)));
	if (off > ctx->score_ws_capacity)
	{
		size_t want = ctx->score_ws_capacity ? 2 * ctx->score_ws_capacity : ((size_t)4 << 20);
		while (want < off)
			want *= 2;
		GCR_CHECK_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
		cudaFree(ctx->d_score_ws);
		ctx->d_score_ws = nullptr;
		ctx->score_ws_capacity = 0;
		GCR_CHECK_CUDA(ctx, cudaMalloc(&ctx->d_score_ws, want));
		ctx->score_ws_capacity = want;
	}
	char *base = (char *)ctx->d_score_ws;
	a.plan = (ScorePlan *)(base + o_head);
	a.counts = (uint32_t *)(base + o_head + sizeof(ScorePlan));
	a.prefix = (uint32_t *)(base + o_prefix);

	a.slots = (uint32_t *)(base + o_slots);
	a.m32 = (float *)(base + o_m32);
	a.accS = (unsigned long long *)(base + o_accS);
	a.accC = (uint32_t *)(base + o_accC);
	a.amb = (unsigned long long *)ctx->d_score_amb;
	GCR_CHECK_CUDA(ctx, cudaMemsetAsync(base + o_head, 0, sizeof(ScorePlan) + (size_t)a.n_pairs * sizeof(uint32_t), ctx->stream));
	return GCR_OK;
}

template <int PROBLEM, int SCORING, int R, bool SC = false>
static gcr_status launch_variant(gcr_ctx *ctx, ScoreArgs &a, bool exact)
{
	constexpr int ML = PROBLEM == GCR_PNP ? 12 : 9;
	gcr_status st = ensure_score_ws(ctx, a, ML);
	if (st)
		return st;
	a.hyps_per_item = THREADS * R;
	const size_t smem = (size_t)STAGES * TILE * 2 * sizeof(float4) + 64 + (SCORING == GCR_MAGSAC ? 5008 * sizeof(float) : 0);
	a.no_counts = !exact && a.inliers == nullptr ? 1 : 0;
	auto kern = exact ? k_score_main<PROBLEM, SCORING, R, true, SC>
	                  : (a.no_counts ? k_score_main<PROBLEM, SCORING, R, false, SC, false> : k_score_main<PROBLEM, SCORING, R, false, SC>);
	// the shared-memory opt-in and the occupancy query are per kernel, not per call (two driver calls on a path the
	// pipelines take ~50 times per image pair)
	int per_sm = 0;
	for (const auto &e : ctx->k3_occupancy)
		if (e.first == (const void *)kern)
			per_sm = e.second;
	if (per_sm == 0)
	{
		GCR_CHECK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
		if (ctx->tune_carveout >= 0)
			GCR_CHECK_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, ctx->tune_carveout));
		GCR_CHECK_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem));
		if (per_sm < 1)
			GCR_FAIL(ctx, GCR_ERR_CUDA, "gcr_score: scoring kernel does not fit on an SM");
		ctx->k3_occupancy.push_back({(const void *)kern, per_sm});
	}
	if (ctx->tune_k3_ctas > 0 && per_sm > ctx->tune_k3_ctas)
		per_sm = ctx->tune_k3_ctas;
	a.grid = per_sm * ctx->sm_count;

	const dim3 gslots((a.k + 255) / 256, a.n_pairs);
	k_score_prepare<PROBLEM><<<gslots, 256, 0, ctx->stream>>>(a); // + the hypothesis-tile prefix, by its last CTA
	GCR_LAUNCHED(ctx);
	st = gcr_profile_begin(ctx);
	if (st)
		return st;
	kern<<<a.grid, THREADS, smem, ctx->stream>>>(a);
	GCR_LAUNCHED(ctx);
	st = gcr_profile_end(ctx, (uint64_t)a.k * a.n_pairs * a.n);
	if (st)
		return st;
	if (exact)
	{
		k_score_fixup<PROBLEM, SCORING><<<2 * ctx->sm_count, 256, 0, ctx->stream>>>(a);
		GCR_LAUNCHED(ctx);
	}
	k_score_finalize<SCORING><<<gslots, 256, 0, ctx->stream>>>(a);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}

template <int PROBLEM, int SCORING>
static gcr_status launch_problem(gcr_ctx *ctx, ScoreArgs &a, bool exact)
{
	// variants: hypotheses per thread (PnP is scalar: 2 or 4; the packed problems take 4, 6 or 8)
	const int v = ctx->score_variant;
	if constexpr (PROBLEM == GCR_PNP)
		return v == 1 ? launch_variant<PROBLEM, SCORING, 4>(ctx, a, exact) : launch_variant<PROBLEM, SCORING, 2>(ctx, a, exact);
	else
	{
		switch (v)
		{
		case 0: return launch_variant<PROBLEM, SCORING, 4>(ctx, a, exact);
		case 1: return launch_variant<PROBLEM, SCORING, 6>(ctx, a, exact);
		case 2: return launch_variant<PROBLEM, SCORING, 8>(ctx, a, exact);
		case 3: return launch_variant<PROBLEM, SCORING, 4, true>(ctx, a, exact);
		default: return launch_variant<PROBLEM, SCORING, 4>(ctx, a, exact);
		}
	}
}

extern "C" gcr_status gcr_score(gcr_ctx *ctx, int problem, int scoring, const double *d_models, const uint8_t *d_valid,
                                int k, double threshold, int flags, float *d_scores, uint32_t *d_inliers)
{
	return gcr_score_fused(ctx, problem, scoring, d_models, d_valid, k, threshold, flags, d_scores, d_inliers, nullptr);
}

gcr_status gcr_score_fused(gcr_ctx *ctx, int problem, int scoring, const double *d_models, const uint8_t *d_valid, int k,
                           double threshold, int flags, float *d_scores, uint32_t *d_inliers, unsigned long long *d_pairbest)
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
	if ((unsigned long long)ctx->n_pairs * (unsigned long long)k >= (1ull << 32))
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_score: n_pairs * k must stay below 2^32 model slots per call");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	if (scoring == GCR_MAGSAC)
	{
		gcr_status st = ensure_lut(ctx);
		if (st)
			return st;
	}
	ScoreArgs a;
	memset(&a, 0, sizeof(a));
	a.pts32 = ctx->d_pts32;
	a.pts64 = ctx->d_pts64;
	a.models = d_models;
	a.valid = d_valid;
	a.norm = ctx->d_norm;
	a.lut = ctx->d_magsac_lut;
	a.lut64 = ctx->d_magsac_lut64;
	a.counters = ctx->d_counters;
	a.scores = d_scores;
	a.inliers = d_inliers;
	a.pairbest = d_pairbest;
	a.n = ctx->n;
	a.n_pad = ctx->n_pad;
	a.k = k;
	a.n_pairs = ctx->n_pairs;
	a.thr = threshold;
	a.band_factor = (float)ctx->band_factor;
	a.trace = ctx->d_score_trace;
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

// FP64 re-check evaluations the LAST gcr_score call of this context could not queue (work list full): queued read-back,
// delivered by the next gcr_sync.  Non-zero means that call's GCR_SCORE_EXACT_COUNTS counts are not exact.
extern "C" gcr_status gcr_score_dropped_async(gcr_ctx *ctx, uint32_t *h_dropped)
{
	if (!ctx || !h_dropped)
		return GCR_ERR_ARG;
	if (!ctx->d_score_ws)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_score_dropped_async: gcr_score has not run on this context");
	return gcr_memcpy_d2h_async(ctx, h_dropped, (const char *)ctx->d_score_ws + offsetof(ScorePlan, amb_dropped), sizeof(uint32_t));
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
	if (!st) st = gcr_memcpy_d2h_async(ctx, h_scores, d_s, total * sizeof(float));
	if (!st && h_inliers) st = gcr_memcpy_d2h_async(ctx, h_inliers, d_i, total * sizeof(uint32_t));
	const gcr_status sync_st = gcr_sync(ctx); // one synchronise delivers both read-backs
	if (!st) st = sync_st;
	// back to the context's block cache on every path (never a raw cudaFree: the cache entry would dangle)
	gcr_free(ctx, d_m);
	gcr_free(ctx, d_v);
	gcr_free(ctx, d_s);
	gcr_free(ctx, d_i);
	return st;
}
