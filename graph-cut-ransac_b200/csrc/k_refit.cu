// k_refit.cu -- K6: weighted least-squares refits on index sets (the local-optimisation inner step, the IRLS rounds
// and the final fit of GCRANSAC::run).
//
//   H: RobustHomographyEstimator::estimateModelNonminimal (estimators/homography_estimator.h:136-173) with
//      normalizePoints (:213-310) and HomographyFourPointSolver::estimateNonMinimalModel
//      (estimators/solver_homography_four_point.h:192-264; n == 4 falls back to the Gauss path :109-190).
//   F: FundamentalMatrixEstimator::estimateModelNonminimal (estimators/fundamental_estimator.h:574-618) with
//      normalizePoints (:637-734) and FundamentalMatrixEightPointSolver::estimateModel
//      (estimators/solver_fundamental_matrix_eight_point.h:93-182).
//
//   E: EssentialMatrixEstimator::estimateModelNonminimal (estimators/essential_estimator.h:306-336) with
//      EssentialMatrixFivePointSteweniusSolver::estimateModel
//      (estimators/solver_essential_matrix_five_point_stewenius.h:114-231): up to 10 models per set.
//
// Mapping: one CTA per index set (the LO step hands over <= lo_number sets of <= 7 m indices; IRLS / final fit one set
// of all inliers).  Threads stride over the set: Hartley normalisation (centroid, mean distance) and the normal
// equations are three block reductions in FP64 with a fixed reduction tree (deterministic), then one thread solves.
//   H: the reference solves the 2n x 8 weighted system with Eigen's colPivHouseholderQr; here A^T W^2 A h = A^T W^2 b
//      (36 + 8 accumulators) + Cholesky.  With Hartley-normalised data the two agree to ~kappa^2 eps << 1e-5
//      (tests/test_gpu_refit.py measures it); a non-positive pivot reports ok = 0.
//   F: the reference forms A^T A (9 x 9) itself and takes the last column of FullPivHouseholderQR(A^T A).matrixQ();
//      the same 45 accumulators, then the same full-pivot Householder QR in one thread (the result is NOT the
//      smallest eigenvector when A^T A has full rank, so the decomposition itself is reproduced, not replaced).
//      No rank-2 enforcement, as in the reference.
//   E: weighted epipolar rows on the K-normalised points (no Hartley step), A^T A (45 accumulators), then in one
//      thread the reference's ColPivHouseholderQR(A^T A).matrixQ().rightCols<4>() as the 4-D "null space" (the
//      decomposition is reproduced: for noisy data that span is NOT the four smallest eigenvectors), the GrevLex
//      trace / determinant constraints, the full-pivot 10 x 10 elimination, the action matrix and its real
//      eigen-pairs (solve_e_stewenius.cuh; Hessenberg + Francis QR stands in for Eigen::EigenSolver).
//      n == 5 takes the LU null space like the reference (:175-180).
// Weight indexing: GCR_WEIGHT_POSITIONAL reads weights[i] with i the position inside the set, which is
// what the reference's H / 8-point solvers do; GCR_WEIGHT_BY_POINT reads weights[index].
#include "gcr_internal.cuh"
#include "solve_e_stewenius.cuh"

#include <float.h>
#include <math.h>

#define RF_THREADS 128
#define RF_WARPS (RF_THREADS / 32)
// A call with MANY sets is the inner RANSAC of the local optimisation: `lo_number` samples of at most 7 m points each
// (GCRANSAC.h:842-866).  One warp per set: the serial parts of a fit (decompositions, LM steps) keep one thread busy either
// way, and a 128-thread CTA idling around it holds four times the registers -- with sixteen pairs in flight the 50-set
// launches queued for SM slots.  A call with ONE set (all inliers: IRLS, the final fit) keeps the wide shapes.
#define RF_SMALL_THREADS 32
#ifndef GCR_SVD3_MAX_SWEEPS
#define GCR_SVD3_MAX_SWEEPS 100
#endif
// Two columns count as orthogonal when |a_p . a_q| <= 4 eps |a_p| |a_q| (Eigen's JacobiSVD stops at 2 eps of the largest
// diagonal entry).  The 1e-17 of round 1 lies BELOW the rounding noise of the dot product itself, so the sweep loop never
// saw a sweep without a rotation and always ran its 100 sweeps of noise-level rotations: 107 of the 195 us of a
// bundle-adjustment fit (tools/refit_bench.py).  oracle/linalg.h carries the same tolerance.
#define GCR_SVD3_TOL 8.881784197001252e-16
#ifndef GCR_BA_MAX_IT
#define GCR_BA_MAX_IT 25 /* max_iterations of lm_F_impl (relative_pose/bundle.cpp:253-339); overridable for timing experiments */
#endif
#define RF_MAX_THREADS 512 /* the single-set launches of the bundle adjustment (all inliers of a run) use 512 threads */
#define RF_MAX_WARPS (RF_MAX_THREADS / 32)

struct RefitArgs
{
	const double *pts; // [n][4] of the selected pair
	const uint32_t *sets, *offsets, *sizes;
	const double *weights;
	int weight_mode;
	double *models; // [count][R][9], R = gcr_max_refit_models(problem)
	uint8_t *ok;    // [count][R]
	int ba_frame;   // F only: 0 = linear 8-point fit; GCR_BA_HARTLEY / GCR_BA_PIXEL = + LM bundle adjustment in that frame
	// gcr_refit_pairs: set s belongs to pair s -- its indices start at sets + s * set_stride (offsets == null) and point into
	// pts + s * pair_stride
	size_t pair_stride, set_stride;
	// gcr_refit_multi (the batch engine): set s belongs to pair pair_of_set[s], reads its weights from weights +
	// weight_offsets[s] and writes its models to output slot out_slots[s] (all three optional)
	const int32_t *pair_of_set;
	const uint32_t *weight_offsets;
	const uint32_t *out_slots;
};

// fixed-tree block sum of NV values per thread; result valid in out[] for every thread after the call
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double *smem /* RF_WARPS * NV */, double *out)
{
	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
	for (int j = 0; j < NV; ++j)
	{
		double x = v[j];
#pragma unroll
		for (int o = 16; o > 0; o >>= 1)
			x += __shfl_xor_sync(0xffffffffu, x, o);
		if (lane == 0)
			smem[warp * NV + j] = x;
	}
	__syncthreads();
	const int n_warps = (int)blockDim.x >> 5;
	for (int j = threadIdx.x; j < NV; j += (int)blockDim.x)
	{
		double s = 0.0;
		for (int w = 0; w < n_warps; ++w)
			s += smem[w * NV + j];
		out[j] = s;
	}
	__syncthreads();
}

__device__ __forceinline__ void mul3(const double *A, const double *B, double *C)
{
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
			C[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
}

// math_utils.h:45-87 on an 8 x 9 augmented system (same as k_solve.cu's minimal path)
__device__ bool gauss8(double M[8][9], double *h)
{
	for (int a = 0; a < 8; a++)
		for (int k = a + 1; k < 8; k++)
			if (fabs(M[a][a]) < fabs(M[k][a]))
				for (int j = 0; j <= 8; j++)
				{
					const double t = M[a][j];
					M[a][j] = M[k][j];
					M[k][j] = t;
				}
	for (int a = 0; a < 7; a++)
		for (int k = a + 1; k < 8; k++)
		{
			const double t = M[k][a] / M[a][a];
			for (int j = 0; j <= 8; j++)
				M[k][j] = M[k][j] - t * M[a][j];
		}
	for (int a = 7; a >= 0; a--)
	{
		h[a] = M[a][8];
		for (int j = a + 1; j < 8; j++)
			h[a] = h[a] - M[a][j] * h[j];
		h[a] = h[a] / M[a][a];
	}
	for (int j = 0; j < 8; ++j)
		if (isnan(h[j]))
			return false;
	return true;
}

// Eigen FullPivHouseholderQR(M).matrixQ().rightCols<1>() for a 9 x 9 row-major M (destroyed)
// __noinline__: inlined into the bundle-adjustment instantiations of k_refit, nvcc 12.9 -O3 produced a kernel whose
// Householder vectors were garbage for a correct A^T A (observed on B200; the stand-alone function is correct)
__device__ __noinline__ void fullpiv_householder_q_lastcol9(double *M, double *q)
{
	const int n = 9;
	int rowswap[9];
	double ess[9][8];
	double taus[9];
	int steps = n;
	for (int k = 0; k < n; ++k)
	{
		int pr = k, pc = k;
		double biggest = 0.0;
		for (int j = k; j < n; ++j)
			for (int i = k; i < n; ++i)
			{
				const double a = fabs(M[i * n + j]);
				if (a > biggest) { biggest = a; pr = i; pc = j; }
			}
		if (biggest == 0.0)
		{
			steps = k;
			break;
		}
		rowswap[k] = pr;
		if (pr != k)
			for (int j = 0; j < n; ++j) { const double t = M[k * n + j]; M[k * n + j] = M[pr * n + j]; M[pr * n + j] = t; }
		if (pc != k)
			for (int i = 0; i < n; ++i) { const double t = M[i * n + k]; M[i * n + k] = M[i * n + pc]; M[i * n + pc] = t; }
		double tail2 = 0.0;
		for (int i = k + 1; i < n; ++i) tail2 += M[i * n + k] * M[i * n + k];
		const double c0 = M[k * n + k];
		double tau = 0.0, beta = c0;
		for (int i = 0; i < 8; ++i) ess[k][i] = 0.0;
		if (tail2 > DBL_MIN)
		{
			beta = sqrt(c0 * c0 + tail2);
			if (c0 >= 0) beta = -beta;
			for (int i = k + 1; i < n; ++i) ess[k][i - k - 1] = M[i * n + k] / (c0 - beta);
			tau = (beta - c0) / beta;
		}
		taus[k] = tau;
		for (int j = k + 1; j < n; ++j)
		{
			double dot = M[k * n + j];
			for (int i = k + 1; i < n; ++i) dot += ess[k][i - k - 1] * M[i * n + j];
			dot *= tau;
			M[k * n + j] -= dot;
			for (int i = k + 1; i < n; ++i) M[i * n + j] -= dot * ess[k][i - k - 1];
		}
		M[k * n + k] = beta;
		for (int i = k + 1; i < n; ++i) M[i * n + k] = 0.0;
	}
	for (int i = 0; i < n; ++i) q[i] = 0.0;
	q[n - 1] = 1.0;
	for (int k = steps - 1; k >= 0; --k)
	{
		double dot = q[k];
		for (int i = k + 1; i < n; ++i) dot += ess[k][i - k - 1] * q[i];
		dot *= taus[k];
		q[k] -= dot;
		for (int i = k + 1; i < n; ++i) q[i] -= dot * ess[k][i - k - 1];
		if (rowswap[k] != k) { const double t = q[k]; q[k] = q[rowswap[k]]; q[rowswap[k]] = t; }
	}
}

// ---- LM bundle adjustment of F (relative_pose/bundle.cpp:253-339 lm_F_impl, relative_pose/jacobian_impl.h:472-606) ----
// One CTA per index set: the point sums (cost; 28 + 7 entries of J^T J, J^T r) are strided over the threads and meet in
// the fixed-tree block_sum; thread 0 does the 7 x 7 damped Cholesky step and the Lie update of U, V.  Same statements
// as oracle/gcr_oracle_f.cpp (gcro_refine_f_lm), summation order aside.
__device__ void jacobi_svd3(const double *A_in, double *U, double *s, double *V)
{
	// one-sided Jacobi, columns of A rotated until orthogonal: A = U diag(s) V^T, s descending (oracle/linalg.h jacobi_svd)
	double A[9], Vt[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
	for (int i = 0; i < 9; ++i) A[i] = A_in[i];
	for (int sweep = 0; sweep < GCR_SVD3_MAX_SWEEPS; ++sweep)
	{
		bool rotated = false;
		for (int p = 0; p < 2; ++p)
			for (int q = p + 1; q < 3; ++q)
			{
				double alpha = 0, beta = 0, gamma = 0;
				for (int i = 0; i < 3; ++i)
				{
					const double ap = A[3 * i + p], aq = A[3 * i + q];
					alpha += ap * ap; beta += aq * aq; gamma += ap * aq;
				}
				if (fabs(gamma) <= GCR_SVD3_TOL * sqrt(alpha * beta) || gamma == 0.0)
					continue;
				rotated = true;
				const double zeta = (beta - alpha) / (2.0 * gamma);
				const double t = (zeta >= 0 ? 1.0 : -1.0) / (fabs(zeta) + sqrt(1.0 + zeta * zeta));
				const double c = 1.0 / sqrt(1.0 + t * t), sn = c * t;
				for (int i = 0; i < 3; ++i)
				{
					const double ap = A[3 * i + p], aq = A[3 * i + q];
					A[3 * i + p] = c * ap - sn * aq;
					A[3 * i + q] = sn * ap + c * aq;
					const double vp = Vt[3 * i + p], vq = Vt[3 * i + q];
					Vt[3 * i + p] = c * vp - sn * vq;
					Vt[3 * i + q] = sn * vp + c * vq;
				}
			}
		if (!rotated)
			break;
	}
	double sv[3];
	int order[3] = {0, 1, 2};
	for (int j = 0; j < 3; ++j)
		sv[j] = sqrt(A[j] * A[j] + A[3 + j] * A[3 + j] + A[6 + j] * A[6 + j]);
	for (int a = 0; a < 2; ++a) // stable descending sort of three
		for (int b = 0; b < 2 - a; ++b)
			if (sv[order[b]] < sv[order[b + 1]]) { const int t = order[b]; order[b] = order[b + 1]; order[b + 1] = t; }
	for (int c = 0; c < 3; ++c)
	{
		const int o = order[c];
		s[c] = sv[o];
		for (int i = 0; i < 3; ++i)
		{
			U[3 * i + c] = sv[o] > 0 ? A[3 * i + o] / sv[o] : 0.0;
			V[3 * i + c] = Vt[3 * i + o];
		}
	}
}

__device__ __forceinline__ void factorized_f(const double *U, const double *V, double sigma, double *F)
{
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c)
			F[3 * r + c] = U[3 * r] * V[3 * c] + sigma * U[3 * r + 1] * V[3 * c + 1];
}

// M <- M + (sin(t) [w]x + (1 - cos(t)) [w]x^2) M, t = |w| (bundle.cpp:290-316)
__device__ void lie_update3(const double *w3, const double *M, double *out)
{
	double w[3] = {w3[0], w3[1], w3[2]};
	const double theta = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
	for (int i = 0; i < 3; ++i) w[i] /= theta;
	const double a = sin(theta), b = cos(theta);
	const double sw[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
	double sw2[9], R[9], RM[9];
	mul3(sw, sw, sw2);
	for (int i = 0; i < 9; ++i) R[i] = a * sw[i] + (1 - b) * sw2[i];
	mul3(R, M, RM);
	for (int i = 0; i < 9; ++i) out[i] = M[i] + RM[i];
}

struct BaFrame { double msx, msy, mdx, mdy, rs, rd; };

__device__ __forceinline__ double ba_weight_of(const RefitArgs &a, const uint32_t *idx, int i)
{
	return !a.weights ? 1.0 : a.weight_mode == GCR_WEIGHT_BY_POINT ? a.weights[idx[i]] : a.weights[i];
}

// refines sh->F in place (shared memory); all threads of the CTA call it
struct BaShared
{
	double F[9], U[9], V[9], Un[9], Vn[9], Fn[9], JtJ[49], Jtr[7], sigma, sigman, lambda;
	int flag; // 0 go on, 1 stop
};

// One pass over the set for the model (F, U, V): the 28 + 7 sums of J^T J / J^T r AND the cost, 36 values through one
// block_sum.  The LM loop needs the cost of a candidate to accept or reject it and, if it accepts, the Jacobian sums at
// that same candidate for the next step: evaluated together, an accepted step costs one pass and one reduction instead
// of two (a rejected one wastes the Jacobian arithmetic, no barrier).  Every sum is formed by the statements, in the
// order, of the two separate passes it replaces (a cost pass and a Jacobian pass per round): the same bits.
__device__ void ba_pass(const RefitArgs &a, const uint32_t *idx, int n, const BaFrame &fr, const double *F, const double *U, const double *V,
                        double *red, double *tot)
{
	const int tid = threadIdx.x;
	double ds[9];
	for (int r = 0; r < 3; ++r)
		for (int c = 0; c < 3; ++c)
			ds[3 * r + c] = U[3 * r + 1] * V[3 * c + 1];
#define F_(r, c) F[3 * (r) + (c)]
#define D_(r, c) ds[3 * (r) + (c)]
	const double dFdp[9][7] = {
	    {0, F_(2, 0), -F_(1, 0), 0, F_(0, 2), -F_(0, 1), D_(0, 0)},
	    {-F_(2, 0), 0, F_(0, 0), 0, F_(1, 2), -F_(1, 1), D_(1, 0)},
	    {F_(1, 0), -F_(0, 0), 0, 0, F_(2, 2), -F_(2, 1), D_(2, 0)},
	    {0, F_(2, 1), -F_(1, 1), -F_(0, 2), 0, F_(0, 0), D_(0, 1)},
	    {-F_(2, 1), 0, F_(0, 1), -F_(1, 2), 0, F_(1, 0), D_(1, 1)},
	    {F_(1, 1), -F_(0, 1), 0, -F_(2, 2), 0, F_(2, 0), D_(2, 1)},
	    {0, F_(2, 2), -F_(1, 2), F_(0, 1), -F_(0, 0), 0, D_(0, 2)},
	    {-F_(2, 2), 0, F_(0, 2), F_(1, 1), -F_(1, 0), 0, D_(1, 2)},
	    {F_(1, 2), -F_(0, 2), 0, F_(2, 1), -F_(2, 0), 0, D_(2, 2)}};
#undef F_
#undef D_
	double acc[36]; // lower triangle of J^T J row by row (28), J^T r (7), the cost
#pragma unroll
	for (int j = 0; j < 36; ++j) acc[j] = 0.0;
	for (int i = tid; i < n; i += (int)blockDim.x)
	{
		const double *p = a.pts + (size_t)idx[i] * 4;
		const double x1 = (p[0] - fr.msx) * fr.rs, y1 = (p[1] - fr.msy) * fr.rs, x2 = (p[2] - fr.mdx) * fr.rd, y2 = (p[3] - fr.mdy) * fr.rd;
		const double Fx0 = F[0] * x1 + F[1] * y1 + F[2], Fx1 = F[3] * x1 + F[4] * y1 + F[5], Fx2 = F[6] * x1 + F[7] * y1 + F[8];
		const double C = x2 * Fx0 + y2 * Fx1 + Fx2;
		const double JC[4] = {F[0] * x2 + F[3] * y2 + F[6], F[1] * x2 + F[4] * y2 + F[7], Fx0, Fx1};
		{
			// the cost term of this point: w_k * min(r_k^2, 1), r_k^2 = C^2 / |J_C|^2 (bundle.cpp's TruncatedLoss(1))
			const double nJc_sq = (Fx0 * Fx0 + Fx1 * Fx1) + (JC[0] * JC[0] + JC[1] * JC[1]);
			const double r2 = (C * C) / nJc_sq;
			const double loss = fmin(r2, 1.0);
			acc[35] += a.weights ? ba_weight_of(a, idx, i) * loss : loss;
		}
		const double nJ_C = sqrt(JC[0] * JC[0] + JC[1] * JC[1] + JC[2] * JC[2] + JC[3] * JC[3]);
		const double inv_nJ_C = 1.0 / nJ_C;
		const double r = C * inv_nJ_C;
		double weight = ((r * r) < 1.0 ? 1.0 : 0.0) / n;
		if (a.weights)
			weight = ba_weight_of(a, idx, i) * weight;
		if (weight == 0.0)
			continue;
		double dF[9] = {x1 * x2, x1 * y2, x1, y1 * x2, y1 * y2, y1, x2, y2, 1.0};
		const double s = C * inv_nJ_C * inv_nJ_C;
		dF[0] -= s * (JC[2] * x1 + JC[0] * x2);
		dF[1] -= s * (JC[3] * x1 + JC[0] * y2);
		dF[2] -= s * (JC[0]);
		dF[3] -= s * (JC[2] * y1 + JC[1] * x2);
		dF[4] -= s * (JC[3] * y1 + JC[1] * y2);
		dF[5] -= s * (JC[1]);
		dF[6] -= s * (JC[2]);
		dF[7] -= s * (JC[3]);
#pragma unroll
		for (int q = 0; q < 9; ++q) dF[q] *= inv_nJ_C;
		double J[7];
#pragma unroll
		for (int j = 0; j < 7; ++j)
		{
			double t = 0.0;
#pragma unroll
			for (int q = 0; q < 9; ++q) t += dF[q] * dFdp[q][j];
			J[j] = t;
		}
		int q = 0;
#pragma unroll
		for (int r7 = 0; r7 < 7; ++r7)
#pragma unroll
			for (int c7 = 0; c7 <= r7; ++c7)
				acc[q++] += weight * (J[r7] * J[c7]);
#pragma unroll
		for (int r7 = 0; r7 < 7; ++r7)
			acc[28 + r7] += weight * C * inv_nJ_C * J[r7];
	}
	block_sum<36>(acc, red, tot);
}

// thread 0: the sums of a pass become the normal equations; gradient_tol (bundle.cpp: |J^T r| < 1e-8) sets the stop flag
__device__ void ba_take_sums(BaShared *sh, const double *tot)
{
	int q = 0;
	for (int r7 = 0; r7 < 7; ++r7)
		for (int c7 = 0; c7 <= r7; ++c7)
			sh->JtJ[7 * r7 + c7] = tot[q++];
	double g = 0.0;
	for (int r7 = 0; r7 < 7; ++r7)
	{
		sh->Jtr[r7] = tot[28 + r7];
		g += tot[28 + r7] * tot[28 + r7];
	}
	if (sqrt(g) < 1e-8) // gradient_tol
		sh->flag = 1;
}

__device__ void ba_refine_f(const RefitArgs &a, const uint32_t *idx, int n, const BaFrame &fr, BaShared *sh, double *red, double *tot,
                            int max_iterations)
{
	const int tid = threadIdx.x;
	if (tid == 0)
	{
		double s[3];
		jacobi_svd3(sh->F, sh->U, s, sh->V); // FactorizedFundamentalMatrix(F) (jacobian_impl.h:474-480)
		sh->sigma = s[1] / s[0];
		factorized_f(sh->U, sh->V, sh->sigma, sh->F);
		sh->lambda = 1e-3;
		sh->flag = 0;
	}
	__syncthreads();
	ba_pass(a, idx, n, fr, sh->F, sh->U, sh->V, red, tot);
	double cost = tot[35];
	if (tid == 0)
		ba_take_sums(sh, tot);
	__syncthreads();
	if (sh->flag)
		return;
	for (int iter = 0; iter < max_iterations; ++iter)
	{
		if (tid == 0)
		{
			for (int k = 0; k < 7; ++k) sh->JtJ[7 * k + k] += sh->lambda;
			// sol = -(L L^T)^-1 Jtr from the lower triangle; NaNs propagate if it is not positive definite (Eigen's llt().solve)
			double L[49], y[7], sol[7];
			for (int j = 0; j < 7; ++j)
			{
				double d = sh->JtJ[7 * j + j];
				for (int k = 0; k < j; ++k) d -= L[7 * j + k] * L[7 * j + k];
				d = sqrt(d);
				L[7 * j + j] = d;
				for (int i = j + 1; i < 7; ++i)
				{
					double t = sh->JtJ[7 * i + j];
					for (int k = 0; k < j; ++k) t -= L[7 * i + k] * L[7 * j + k];
					L[7 * i + j] = t / d;
				}
			}
			for (int i = 0; i < 7; ++i)
			{
				double t = sh->Jtr[i];
				for (int k = 0; k < i; ++k) t -= L[7 * i + k] * y[k];
				y[i] = t / L[7 * i + i];
			}
			for (int i = 6; i >= 0; --i)
			{
				double t = y[i];
				for (int k = i + 1; k < 7; ++k) t -= L[7 * k + i] * sol[k];
				sol[i] = t / L[7 * i + i];
			}
			double sn = 0.0;
			for (int i = 0; i < 7; ++i)
			{
				sol[i] = -sol[i];
				sn += sol[i] * sol[i];
			}
			if (sqrt(sn) < 1e-8) // step_tol
				sh->flag = 1;
			else
			{
				lie_update3(sol, sh->U, sh->Un);
				lie_update3(sol + 3, sh->V, sh->Vn);
				sh->sigman = sh->sigma + sol[6];
				factorized_f(sh->Un, sh->Vn, sh->sigman, sh->Fn);
			}
		}
		__syncthreads();
		if (sh->flag)
			break;
		// the candidate's cost and, should it be accepted, its normal equations: one pass
		ba_pass(a, idx, n, fr, sh->Fn, sh->Un, sh->Vn, red, tot);
		const double cost_new = tot[35];
		const bool accept = cost_new < cost; // the same value in every thread
		if (tid == 0)
		{
			if (accept)
			{
				for (int i = 0; i < 9; ++i) { sh->U[i] = sh->Un[i]; sh->V[i] = sh->Vn[i]; sh->F[i] = sh->Fn[i]; }
				sh->sigma = sh->sigman;
				sh->lambda /= 10;
				ba_take_sums(sh, tot); // (the gradient test the next round would start with)
			}
			else
			{
				for (int k = 0; k < 7; ++k) sh->JtJ[7 * k + k] -= sh->lambda;
				sh->lambda *= 10;
			}
		}
		if (accept)
			cost = cost_new;
		__syncthreads();
		if (sh->flag)
			break;
	}
	__syncthreads();
}

template <int PROBLEM, int BA = 0>
__global__ void __launch_bounds__(BA != 0 ? RF_MAX_THREADS : RF_THREADS) k_refit(const RefitArgs a_in)
{
	RefitArgs a = a_in;
	a.pts += (size_t)(a_in.pair_of_set ? a_in.pair_of_set[blockIdx.x] : (int)blockIdx.x) * a_in.pair_stride;
	if (a.weights && a_in.weight_offsets)
		a.weights += a_in.weight_offsets[blockIdx.x];
	const size_t oslot = a_in.out_slots ? a_in.out_slots[blockIdx.x] : blockIdx.x;
	constexpr int NV = PROBLEM == GCR_H ? 44 : 45;
	__shared__ double red[(BA != 0 ? RF_MAX_WARPS : RF_WARPS) * 45];
	__shared__ double tot[45];
	const int set = blockIdx.x, tid = threadIdx.x;
	const uint32_t *idx = a.offsets ? a.sets + a.offsets[set] : a.sets + (size_t)set * a.set_stride;
	const int n = (int)a.sizes[set];
	const int min_n = PROBLEM == GCR_H ? 4 : 8;
	if (n < min_n)
	{
		if (tid == 0)
		{
			for (int j = 0; j < 9; ++j) a.models[oslot * 9 + j] = 0.0;
			a.ok[oslot] = 0;
		}
		return;
	}
	constexpr bool pixel_frame = PROBLEM == GCR_F && BA == GCR_BA_PIXEL; // the post-run polish fits in pixels
	// ---- normalizePoints: centroids ----
	if (!pixel_frame)
	{
		double v[4] = {0, 0, 0, 0};
		for (int i = tid; i < n; i += (int)blockDim.x)
		{
			const double *p = a.pts + (size_t)idx[i] * 4;
			v[0] += p[0]; v[1] += p[1]; v[2] += p[2]; v[3] += p[3];
		}
		block_sum<4>(v, red, tot);
	}
	const double msx = pixel_frame ? 0.0 : tot[0] / n, msy = pixel_frame ? 0.0 : tot[1] / n, mdx = pixel_frame ? 0.0 : tot[2] / n,
	             mdy = pixel_frame ? 0.0 : tot[3] / n;
	__syncthreads();
	// ---- mean distances ----
	if (!pixel_frame)
	{
		double v[2] = {0, 0};
		for (int i = tid; i < n; i += (int)blockDim.x)
		{
			const double *p = a.pts + (size_t)idx[i] * 4;
			const double dx1 = msx - p[0], dy1 = msy - p[1], dx2 = mdx - p[2], dy2 = mdy - p[3];
			v[0] += sqrt(dx1 * dx1 + dy1 * dy1);
			v[1] += sqrt(dx2 * dx2 + dy2 * dy2);
		}
		block_sum<2>(v, red, tot);
	}
	const double rs = pixel_frame ? 1.0 : 1.4142135623730951 / (tot[0] / n), rd = pixel_frame ? 1.0 : 1.4142135623730951 / (tot[1] / n);
	__syncthreads();
	const double T1[9] = {rs, 0, -rs * msx, 0, rs, -rs * msy, 0, 0, 1};
	const double T2[9] = {rd, 0, -rd * mdx, 0, rd, -rd * mdy, 0, 0, 1};

	if (PROBLEM == GCR_H && n == 4)
	{
		if (tid == 0)
		{
			double M[8][9];
			int row = 0;
			for (int i = 0; i < 4; ++i)
			{
				const double *p = a.pts + (size_t)idx[i] * 4;
				const double x1 = (p[0] - msx) * rs, y1 = (p[1] - msy) * rs, x2 = (p[2] - mdx) * rd, y2 = (p[3] - mdy) * rd;
				const double weight = !a.weights ? 1.0 : a.weight_mode == GCR_WEIGHT_BY_POINT ? a.weights[idx[i]] : a.weights[i];
				const double mwx1 = -weight * x1, mwy1 = -weight * y1, wx2 = weight * x2, wy2 = weight * y2;
				M[row][0] = mwx1; M[row][1] = mwy1; M[row][2] = -weight; M[row][3] = 0; M[row][4] = 0; M[row][5] = 0;
				M[row][6] = wx2 * x1; M[row][7] = wx2 * y1; M[row][8] = -wx2;
				++row;
				M[row][0] = 0; M[row][1] = 0; M[row][2] = 0; M[row][3] = mwx1; M[row][4] = mwy1; M[row][5] = -weight;
				M[row][6] = wy2 * x1; M[row][7] = wy2 * y1; M[row][8] = -wy2;
				++row;
			}
			double Hn[9];
			const bool ok = gauss8(M, Hn);
			Hn[8] = 1.0;
			double T2i[9] = {1.0 / rd, 0, mdx, 0, 1.0 / rd, mdy, 0, 0, 1}, tmp[9], H[9];
			mul3(T2i, Hn, tmp);
			mul3(tmp, T1, H);
			for (int j = 0; j < 9; ++j) a.models[oslot * 9 + j] = ok ? H[j] : 0.0;
			a.ok[oslot] = ok ? 1 : 0;
		}
		return;
	}

	// ---- normal equations ----
	double acc[NV];
#pragma unroll
	for (int j = 0; j < NV; ++j) acc[j] = 0.0;
	for (int i = tid; i < n; i += (int)blockDim.x)
	{
		const double *p = a.pts + (size_t)idx[i] * 4;
		const double x1 = (p[0] - msx) * rs, y1 = (p[1] - msy) * rs, x2 = (p[2] - mdx) * rd, y2 = (p[3] - mdy) * rd;
		const double w = !a.weights ? 1.0 : a.weight_mode == GCR_WEIGHT_BY_POINT ? a.weights[idx[i]] : a.weights[i];
		if (PROBLEM == GCR_H)
		{
			// rows w * [-x1 -y1 -1 0 0 0 x2x1 x2y1 | -x2], w * [0 0 0 -x1 -y1 -1 y2x1 y2y1 | -y2]
			const double r1[9] = {-w * x1, -w * y1, -w, 0, 0, 0, w * x2 * x1, w * x2 * y1, -w * x2};
			const double r2[9] = {0, 0, 0, -w * x1, -w * y1, -w, w * y2 * x1, w * y2 * y1, -w * y2};
			int q = 0;
#pragma unroll
			for (int r = 0; r < 8; ++r)
#pragma unroll
				for (int c = r; c < 8; ++c)
					acc[q++] += r1[r] * r1[c] + r2[r] * r2[c];
#pragma unroll
			for (int r = 0; r < 8; ++r)
				acc[36 + r] += r1[r] * r1[8] + r2[r] * r2[8];
		}
		else
		{
			const double r[9] = {w * x2 * x1, w * x2 * y1, w * x2, w * y2 * x1, w * y2 * y1, w * y2, w * x1, w * y1, w};
			int q = 0;
#pragma unroll
			for (int i2 = 0; i2 < 9; ++i2)
#pragma unroll
				for (int c = i2; c < 9; ++c)
					acc[q++] += r[i2] * r[c];
		}
	}
	block_sum<NV>(acc, red, tot);
	if (PROBLEM == GCR_F && BA != 0)
	{
		// FundamentalMatrixBundleAdjustmentSolver::estimateModel (solver_fundamental_matrix_bundle_adjustment.h:113-177):
		// the 8-point model initialises the LM refinement, in the same frame
		__shared__ BaShared ba;
		if (tid == 0)
		{
			double AtA[81];
			int q = 0;
			for (int r = 0; r < 9; ++r)
				for (int c = r; c < 9; ++c)
				{
					AtA[9 * r + c] = tot[q];
					AtA[9 * c + r] = tot[q];
					++q;
				}
#ifdef GCR_BA_DEBUG
			if (blockIdx.x == 0) printf("AtA %g %g %g tot44 %g n %d rs %g\n", AtA[0], AtA[1], AtA[80], tot[44], n, rs);
#endif
			double f[9];
			fullpiv_householder_q_lastcol9(AtA, f);
			for (int i = 0; i < 9; ++i) ba.F[i] = f[i];
#ifdef GCR_BA_DEBUG
			if (blockIdx.x == 0) printf("f %g %g %g %g\n", ba.F[0], ba.F[1], ba.F[4], ba.F[8]);
#endif
		}
		__syncthreads();
		const BaFrame fr = {msx, msy, mdx, mdy, rs, rd};
		ba_refine_f(a, idx, n, fr, &ba, red, tot, GCR_BA_MAX_IT);
		if (tid != 0)
			return;
		double out[9];
		bool ok = true;
		if (pixel_frame)
			for (int i = 0; i < 9; ++i) out[i] = ba.F[i];
		else
		{
			const double T2t[9] = {T2[0], T2[3], T2[6], T2[1], T2[4], T2[7], T2[2], T2[5], T2[8]};
			double tmp[9];
			mul3(T2t, ba.F, tmp);
			mul3(tmp, T1, out);
			double nrm = 0.0;
			for (int i = 0; i < 9; ++i) nrm += out[i] * out[i];
			nrm = sqrt(nrm);
			for (int i = 0; i < 9; ++i) out[i] /= nrm;
			if (out[8] < 0)
				for (int i = 0; i < 9; ++i) out[i] = -out[i];
		}
		for (int i = 0; i < 9; ++i)
			if (isnan(out[i])) ok = false;
		for (int j = 0; j < 9; ++j) a.models[oslot * 9 + j] = ok ? out[j] : 0.0;
		a.ok[oslot] = ok ? 1 : 0;
		return;
	}
	if (tid != 0)
		return;

	double out[9];
	bool ok = true;
	if (PROBLEM == GCR_H)
	{
		double L[8][8], b[8];
		int q = 0;
		for (int r = 0; r < 8; ++r)
			for (int c = r; c < 8; ++c)
			{
				L[c][r] = tot[q];
				L[r][c] = tot[q];
				++q;
			}
		for (int r = 0; r < 8; ++r) b[r] = tot[36 + r];
		// Cholesky L L^T (lower, in place)
		for (int j = 0; j < 8 && ok; ++j)
		{
			double d = L[j][j];
			for (int k = 0; k < j; ++k) d -= L[j][k] * L[j][k];
			if (!(d > 0.0)) { ok = false; break; }
			d = sqrt(d);
			L[j][j] = d;
			for (int i = j + 1; i < 8; ++i)
			{
				double s = L[i][j];
				for (int k = 0; k < j; ++k) s -= L[i][k] * L[j][k];
				L[i][j] = s / d;
			}
		}
		double Hn[9];
		if (ok)
		{
			double y[8];
			for (int i = 0; i < 8; ++i)
			{
				double s = b[i];
				for (int k = 0; k < i; ++k) s -= L[i][k] * y[k];
				y[i] = s / L[i][i];
			}
			for (int i = 7; i >= 0; --i)
			{
				double s = y[i];
				for (int k = i + 1; k < 8; ++k) s -= L[k][i] * Hn[k];
				Hn[i] = s / L[i][i];
			}
			Hn[8] = 1.0;
			for (int i = 0; i < 8; ++i)
				if (isnan(Hn[i])) ok = false;
		}
		if (ok)
		{
			// homography_estimator.h:168-171: H = T2^-1 * Hn * T1
			const double T2i[9] = {1.0 / rd, 0, mdx, 0, 1.0 / rd, mdy, 0, 0, 1};
			double tmp[9];
			mul3(T2i, Hn, tmp);
			mul3(tmp, T1, out);
		}
	}
	else
	{
		double AtA[81], f[9];
		int q = 0;
		for (int r = 0; r < 9; ++r)
			for (int c = r; c < 9; ++c)
			{
				AtA[9 * r + c] = tot[q];
				AtA[9 * c + r] = tot[q];
				++q;
			}
		fullpiv_householder_q_lastcol9(AtA, f);
		// fundamental_estimator.h:606-616: F = T2^T Fn T1, Frobenius-normalise, sign so that F(2,2) >= 0
		const double T2t[9] = {T2[0], T2[3], T2[6], T2[1], T2[4], T2[7], T2[2], T2[5], T2[8]};
		double tmp[9];
		mul3(T2t, f, tmp);
		mul3(tmp, T1, out);
		double nrm = 0.0;
		for (int i = 0; i < 9; ++i) nrm += out[i] * out[i];
		nrm = sqrt(nrm);
		for (int i = 0; i < 9; ++i) out[i] /= nrm;
		if (out[8] < 0)
			for (int i = 0; i < 9; ++i) out[i] = -out[i];
		for (int i = 0; i < 9; ++i)
			if (isnan(out[i])) ok = false;
	}
	for (int j = 0; j < 9; ++j) a.models[oslot * 9 + j] = ok ? out[j] : 0.0;
	a.ok[oslot] = ok ? 1 : 0;
}

// Eigen ColPivHouseholderQR(M).matrixQ().rightCols<4>() of a 9 x 9 row-major M (destroyed): N[b][j] = Q(j, 5 + b).
// Same pivoting rule as Eigen 3.4 computeInPlace (first column of largest down-dated norm, LAPACK-style down-dating,
// recomputation below sqrt(eps)) and makeHouseholder's (beta, tau, essential) convention.
// __noinline__ ON PURPOSE: inlined into k_refit_e's lane-parallel form, nvcc 12.9 gave `updated[9]` the stack slot of the
// caller's live AtA[81] (SASS: both at R1 + 0 -- the column norms overwrote row 0 of the matrix, wrong pivots, a basis
// that is not the null space; tools/refit_e_dump.py + scripts/gpu_refit_e_ab.sh found it).  As a real call its locals get
// a frame of their own.
__device__ __noinline__ void colpiv_householder_q_last4_9(double *M, double (&N)[4][9])
{
	const int n = 9;
	double updated[9], direct[9], tau[9];
	for (int j = 0; j < n; ++j)
	{
		double s = 0.0;
		for (int i = 0; i < n; ++i) s += M[i * n + j] * M[i * n + j];
		updated[j] = direct[j] = sqrt(s);
	}
	const double downdate_threshold = 1.4901161193847656e-08; // sqrt(eps)
	for (int k = 0; k < n; ++k)
	{
		int big = k;
		for (int j = k + 1; j < n; ++j)
			if (updated[j] > updated[big]) big = j;
		if (big != k)
		{
			for (int i = 0; i < n; ++i) { const double t = M[i * n + k]; M[i * n + k] = M[i * n + big]; M[i * n + big] = t; }
			double t = updated[k]; updated[k] = updated[big]; updated[big] = t;
			t = direct[k]; direct[k] = direct[big]; direct[big] = t;
		}
		double tail = 0.0;
		for (int i = k + 1; i < n; ++i) tail += M[i * n + k] * M[i * n + k];
		const double c0 = M[k * n + k];
		double beta;
		if (tail <= DBL_MIN)
		{
			tau[k] = 0.0;
			beta = c0;
			for (int i = k + 1; i < n; ++i) M[i * n + k] = 0.0;
		}
		else
		{
			beta = sqrt(c0 * c0 + tail);
			if (c0 >= 0) beta = -beta;
			for (int i = k + 1; i < n; ++i) M[i * n + k] /= (c0 - beta);
			tau[k] = (beta - c0) / beta;
		}
		M[k * n + k] = beta;
		for (int j = k + 1; j < n; ++j)
		{
			double dot = M[k * n + j];
			for (int i = k + 1; i < n; ++i) dot += M[i * n + k] * M[i * n + j];
			dot *= tau[k];
			M[k * n + j] -= dot;
			for (int i = k + 1; i < n; ++i) M[i * n + j] -= dot * M[i * n + k];
		}
		for (int j = k + 1; j < n; ++j)
			if (updated[j] != 0.0)
			{
				double temp = fabs(M[k * n + j]) / updated[j];
				temp = (1.0 + temp) * (1.0 - temp);
				if (temp < 0) temp = 0;
				const double r = updated[j] / direct[j];
				const double temp2 = temp * r * r;
				if (temp2 <= downdate_threshold)
				{
					double s = 0.0;
					for (int i = k + 1; i < n; ++i) s += M[i * n + j] * M[i * n + j];
					direct[j] = sqrt(s);
					updated[j] = direct[j];
				}
				else
					updated[j] *= sqrt(temp);
			}
	}
	// columns 5..8 of Q = H_0 ... H_8 applied to the unit vectors e_5..e_8
	for (int b = 0; b < 4; ++b)
	{
		double q[9];
		for (int i = 0; i < n; ++i) q[i] = i == 5 + b ? 1.0 : 0.0;
		for (int k = n - 1; k >= 0; --k)
		{
			double dot = q[k];
			for (int i = k + 1; i < n; ++i) dot += M[i * n + k] * q[i];
			dot *= tau[k];
			q[k] -= dot;
			for (int i = k + 1; i < n; ++i) q[i] -= dot * M[i * n + k];
		}
		for (int i = 0; i < n; ++i) N[b][i] = q[i];
	}
}

#define RF_E_MODELS 10

__global__ void __launch_bounds__(RF_THREADS) k_refit_e(const RefitArgs a_in)
{
	RefitArgs a = a_in;
	a.pts += (size_t)(a_in.pair_of_set ? a_in.pair_of_set[blockIdx.x] : (int)blockIdx.x) * a_in.pair_stride;
	if (a.weights && a_in.weight_offsets)
		a.weights += a_in.weight_offsets[blockIdx.x];
	const size_t oslot = a_in.out_slots ? a_in.out_slots[blockIdx.x] : blockIdx.x;
	__shared__ double red[RF_WARPS * 45];
	__shared__ double tot[45];
	const int set = blockIdx.x, tid = threadIdx.x;
	const uint32_t *idx = a.sets + a.offsets[set];
	const int n = (int)a.sizes[set];
	double *models = a.models + oslot * RF_E_MODELS * 9;
	uint8_t *ok = a.ok + oslot * RF_E_MODELS;
	if (n < 5) // nonMinimalSampleSize() (essential_estimator.h:315-316)
	{
		for (int j = tid; j < RF_E_MODELS * 9; j += (int)blockDim.x) models[j] = 0.0;
		if (tid < RF_E_MODELS) ok[tid] = 0;
		return;
	}
	__shared__ double s_N[4][9], s_Am[100], s_w[20];
	__shared__ int s_slot[RF_E_MODELS];
	double N[4][9];
	bool have_basis = true;
	if (n == 5)
	{
		// stewenius.h:175-180: LU kernel of the (weighted) 5 x 9 matrix
		if (tid == 0)
		{
			double lu[5][9];
			for (int i = 0; i < 5; ++i)
			{
				const double *p = a.pts + (size_t)idx[i] * 4;
				const double w = !a.weights ? 1.0 : a.weight_mode == GCR_WEIGHT_BY_POINT ? a.weights[idx[i]] : a.weights[i];
				const double wx0 = w * p[0], wx1 = w * p[2], wy0 = w * p[1], wy1 = w * p[3];
				lu[i][0] = wx0 * p[2]; lu[i][1] = wx0 * p[3]; lu[i][2] = wx0; lu[i][3] = wy0 * p[2]; lu[i][4] = wy0 * p[3];
				lu[i][5] = wy0; lu[i][6] = wx1; lu[i][7] = wy1; lu[i][8] = w;
			}
			have_basis = gcr_e5::null_space_lu(lu, N);
		}
	}
	else
	{
		double acc[45];
#pragma unroll
		for (int j = 0; j < 45; ++j) acc[j] = 0.0;
		for (int i = tid; i < n; i += (int)blockDim.x)
		{
			const double *p = a.pts + (size_t)idx[i] * 4;
			const double w = !a.weights ? 1.0 : a.weight_mode == GCR_WEIGHT_BY_POINT ? a.weights[idx[i]] : a.weights[i];
			const double wx0 = w * p[0], wx1 = w * p[2], wy0 = w * p[1], wy1 = w * p[3];
			const double r[9] = {wx0 * p[2], wx0 * p[3], wx0, wy0 * p[2], wy0 * p[3], wy0, wx1, wy1, w};
			int q = 0;
#pragma unroll
			for (int i2 = 0; i2 < 9; ++i2)
#pragma unroll
				for (int c = i2; c < 9; ++c)
					acc[q++] += r[i2] * r[c];
		}
		block_sum<45>(acc, red, tot);
		if (tid == 0)
		{
			double AtA[81];
			int q = 0;
			for (int r = 0; r < 9; ++r)
				for (int c = r; c < 9; ++c)
				{
					AtA[9 * r + c] = tot[q];
					AtA[9 * c + r] = tot[q];
					++q;
				}
			colpiv_householder_q_last4_9(AtA, N);
		}
	}
	// Thread 0: basis -> action matrix + eigenvalues (serial: 10 x 10 LU, Hessenberg QR).  Then one lane per eigenvalue:
	// the inverse iterations are independent, model m = the m-th REAL eigenvalue in the order of the eigenvalues.
	if (tid == 0)
	{
		int n_eig = 0;
		if (have_basis)
		{
			bool finite = true;
			for (int b = 0; b < 4; ++b)
				for (int j = 0; j < 9; ++j)
					if (!isfinite(N[b][j])) finite = false;
			if (finite)
				n_eig = gcr_e_nm::action_matrix(N, s_Am, s_w, s_w + 10);
		}
		int count = 0;
		for (int e = 0; e < RF_E_MODELS; ++e)
			s_slot[e] = e < n_eig && s_w[10 + e] == 0.0 ? count++ : -1;
		for (int b = 0; b < 4; ++b)
			for (int j = 0; j < 9; ++j)
				s_N[b][j] = N[b][j];
		// slots beyond the number of real eigenvalues
		for (int m = count; m < RF_E_MODELS; ++m)
		{
			for (int i = 0; i < 9; ++i) models[9 * m + i] = 0.0;
			ok[m] = 0;
		}
	}
	__syncthreads();
	if (tid >= RF_E_MODELS || s_slot[tid] < 0)
		return;
	const int m = s_slot[tid];
	double E[9];
	gcr_e_nm::model_of_eigenvalue(s_N, s_Am, s_w[tid], E);
	// essential_estimator.h:327-333: Frobenius-normalise, sign so that E(2,2) >= 0
	bool good = true;
	double out[9], nrm = 0.0;
	for (int i = 0; i < 9; ++i) nrm += E[i] * E[i];
	nrm = sqrt(nrm);
	for (int i = 0; i < 9; ++i) out[i] = E[i] / nrm;
	if (out[8] < 0)
		for (int i = 0; i < 9; ++i) out[i] = -out[i];
	for (int i = 0; i < 9; ++i)
		if (isnan(out[i])) good = false;
	for (int i = 0; i < 9; ++i) models[9 * m + i] = good ? out[i] : 0.0;
	ok[m] = good ? 1 : 0;
}

extern "C" int gcr_max_refit_models(int problem)
{
	return problem == GCR_E ? RF_E_MODELS : (problem == GCR_H || problem == GCR_F || problem == GCR_PNP) ? 1 : 0;
}

static gcr_status refit_launch(gcr_ctx *ctx, int problem, int pair, const uint32_t *d_index_sets,
                               const uint32_t *d_set_offsets, const uint32_t *d_set_sizes, int count,
                               const double *d_weights, int weight_mode, int ba_frame, double *d_models, uint8_t *d_ok);

extern "C" gcr_status gcr_refit(gcr_ctx *ctx, int problem, int pair, const uint32_t *d_index_sets,
                                const uint32_t *d_set_offsets, const uint32_t *d_set_sizes, int count,
                                const double *d_weights, int weight_mode, double *d_models, uint8_t *d_ok)
{
	return refit_launch(ctx, problem, pair, d_index_sets, d_set_offsets, d_set_sizes, count, d_weights, weight_mode, 0, d_models, d_ok);
}

extern "C" gcr_status gcr_refit_f_ba(gcr_ctx *ctx, int pair, const uint32_t *d_index_sets, const uint32_t *d_set_offsets,
                                     const uint32_t *d_set_sizes, int count, const double *d_weights, int weight_mode,
                                     int frame, double *d_models, uint8_t *d_ok)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (frame != GCR_BA_HARTLEY && frame != GCR_BA_PIXEL)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_refit_f_ba: frame must be GCR_BA_HARTLEY or GCR_BA_PIXEL");
	return refit_launch(ctx, GCR_F, pair, d_index_sets, d_set_offsets, d_set_sizes, count, d_weights, weight_mode, frame, d_models, d_ok);
}

static gcr_status refit_launch(gcr_ctx *ctx, int problem, int pair, const uint32_t *d_index_sets,
                               const uint32_t *d_set_offsets, const uint32_t *d_set_sizes, int count,
                               const double *d_weights, int weight_mode, int ba_frame, double *d_models, uint8_t *d_ok)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!ctx->d_pts64 || ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_refit: no points uploaded");
	if (problem == GCR_PNP ? ctx->cols != 5 : (ctx->cols != 4 || (problem != GCR_H && problem != GCR_F && problem != GCR_E)))
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_refit: non-minimal refits exist for H, F (8-point), E (Stewenius) on 4-column points and PnP (DLT + LM) on 5-column points");
	if (pair < 0 || pair >= ctx->n_pairs || count <= 0 || !d_index_sets || !d_set_offsets || !d_set_sizes || !d_models || !d_ok)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_refit: bad arguments");
	if (d_weights && weight_mode != GCR_WEIGHT_POSITIONAL && weight_mode != GCR_WEIGHT_BY_POINT)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_refit: weights given but weight_mode is %d", weight_mode);
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	if (problem == GCR_PNP)
		return gcr_refit_pnp_launch(ctx, pair, d_index_sets, d_set_offsets, d_set_sizes, count, d_weights, weight_mode, d_models, d_ok);
	RefitArgs a;
	a.pts = ctx->d_pts64 + (size_t)pair * ctx->n * 4;
	a.sets = d_index_sets;
	a.offsets = d_set_offsets;
	a.sizes = d_set_sizes;
	a.weights = d_weights;
	a.weight_mode = weight_mode;
	a.models = d_models;
	a.ok = d_ok;
	a.ba_frame = ba_frame;
	a.pair_stride = a.set_stride = 0;
	a.pair_of_set = nullptr;
	a.weight_offsets = nullptr;
	a.out_slots = nullptr;
	const int threads = count == 1 ? RF_THREADS : RF_SMALL_THREADS, ba_threads = count == 1 ? RF_MAX_THREADS : RF_SMALL_THREADS;
	if (problem == GCR_H)
		k_refit<GCR_H><<<count, threads, 0, ctx->stream>>>(a);
	else if (problem == GCR_E)
		k_refit_e<<<count, threads, 0, ctx->stream>>>(a);
	else if (ba_frame == GCR_BA_HARTLEY)
		k_refit<GCR_F, GCR_BA_HARTLEY><<<count, ba_threads, 0, ctx->stream>>>(a);
	else if (ba_frame == GCR_BA_PIXEL)
		k_refit<GCR_F, GCR_BA_PIXEL><<<count, ba_threads, 0, ctx->stream>>>(a);
	else
		k_refit<GCR_F><<<count, threads, 0, ctx->stream>>>(a);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}

// One non-minimal fit per pair of the context, all pairs in one launch: set p = the first d_sizes[p] entries of
// d_index_sets[p][n] (e.g. the inlier lists gcr_score_exact_pairs wrote), fitted to pair p's points.  H and F (8-point).
extern "C" gcr_status gcr_refit_pairs(gcr_ctx *ctx, int problem, const uint32_t *d_index_sets, const uint32_t *d_sizes,
                                      double *d_models, uint8_t *d_ok)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!ctx->d_pts64 || ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_refit_pairs: no points uploaded");
	if (ctx->cols != 4 || (problem != GCR_H && problem != GCR_F) || !d_index_sets || !d_sizes || !d_models || !d_ok)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_refit_pairs: H or F (8-point) on 4-column points");
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	RefitArgs a;
	a.pts = ctx->d_pts64;
	a.sets = d_index_sets;
	a.offsets = nullptr;
	a.sizes = d_sizes;
	a.weights = nullptr;
	a.weight_mode = GCR_WEIGHT_NONE;
	a.models = d_models;
	a.ok = d_ok;
	a.ba_frame = 0;
	a.pair_stride = (size_t)ctx->n * 4;
	a.pair_of_set = nullptr;
	a.weight_offsets = nullptr;
	a.out_slots = nullptr;
	a.set_stride = (size_t)ctx->n;
	if (problem == GCR_H)
		k_refit<GCR_H><<<ctx->n_pairs, RF_THREADS, 0, ctx->stream>>>(a);
	else
		k_refit<GCR_F><<<ctx->n_pairs, RF_THREADS, 0, ctx->stream>>>(a);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}

// The batch engine's refit: `count` index sets that belong to DIFFERENT pairs of the context, one launch.  Set s: pair
// d_pair_of_set[s], indices d_index_sets + d_set_offsets[s] (d_set_sizes[s] of them), weights (when d_weights is given) at
// d_weights + d_weight_offsets[s] read in `weight_mode`, models written to slot d_out_slots[s] of d_models / d_ok (slot x
// R models each; d_out_slots == NULL: slot s).  ba_frame as gcr_refit_f_ba (F only; 0 = the linear fit); wide != 0: the
// sets are whole inlier sets (one per request), fitted with the CTA shape of a single-set gcr_refit / gcr_refit_f_ba call;
// wide == 0: the sets are the samples of local-optimisation rounds, one warp each like a many-set call.
extern "C" gcr_status gcr_refit_multi(gcr_ctx *ctx, int problem, int ba_frame, const int32_t *d_pair_of_set, const uint32_t *d_index_sets,
                                      const uint32_t *d_set_offsets, const uint32_t *d_set_sizes, int count, const double *d_weights,
                                      const uint32_t *d_weight_offsets, int weight_mode, const uint32_t *d_out_slots, int wide,
                                      double *d_models, uint8_t *d_ok)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!ctx->d_pts64 || ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_refit_multi: no points uploaded");
	if (ctx->cols != 4 || (problem != GCR_H && problem != GCR_F && problem != GCR_E))
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_refit_multi: H, F or E on 4-column points");
	if (count <= 0 || !d_pair_of_set || !d_index_sets || !d_set_offsets || !d_set_sizes || !d_models || !d_ok)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_refit_multi: bad arguments");
	if (ba_frame && (problem != GCR_F || (ba_frame != GCR_BA_HARTLEY && ba_frame != GCR_BA_PIXEL)))
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_refit_multi: ba_frame is for F (GCR_BA_HARTLEY / GCR_BA_PIXEL)");
	if (d_weights && weight_mode != GCR_WEIGHT_POSITIONAL && weight_mode != GCR_WEIGHT_BY_POINT)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_refit_multi: weights given but weight_mode is %d", weight_mode);
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	RefitArgs a;
	a.pts = ctx->d_pts64;
	a.sets = d_index_sets;
	a.offsets = d_set_offsets;
	a.sizes = d_set_sizes;
	a.weights = d_weights;
	a.weight_mode = weight_mode;
	a.models = d_models;
	a.ok = d_ok;
	a.ba_frame = ba_frame;
	a.pair_stride = (size_t)ctx->n * 4;
	a.set_stride = 0;
	a.pair_of_set = d_pair_of_set;
	a.weight_offsets = d_weight_offsets;
	a.out_slots = d_out_slots;
	// `wide`: the CTA shapes gcr_refit / gcr_refit_f_ba give a call with ONE set (all inliers of a run) -- the block sums
	// associate by CTA shape, so a batched single-set fit uses the shape the unbatched call would; else the one-warp shape
	// of a many-set call
	const int threads = wide ? RF_THREADS : RF_SMALL_THREADS, ba_threads = wide ? RF_MAX_THREADS : RF_SMALL_THREADS;
	if (problem == GCR_H)
		k_refit<GCR_H><<<count, threads, 0, ctx->stream>>>(a);
	else if (problem == GCR_E)
		k_refit_e<<<count, threads, 0, ctx->stream>>>(a);
	else if (ba_frame == GCR_BA_HARTLEY)
		k_refit<GCR_F, GCR_BA_HARTLEY><<<count, ba_threads, 0, ctx->stream>>>(a);
	else if (ba_frame == GCR_BA_PIXEL)
		k_refit<GCR_F, GCR_BA_PIXEL><<<count, ba_threads, 0, ctx->stream>>>(a);
	else
		k_refit<GCR_F><<<count, threads, 0, ctx->stream>>>(a);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}
