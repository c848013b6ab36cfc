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

struct RefitArgs
{
	const double *pts; // [n][4] of the selected pair
	const uint32_t *sets, *offsets, *sizes;
	const double *weights;
	int weight_mode;
	double *models; // [count][R][9], R = gcr_max_refit_models(problem)
	uint8_t *ok;    // [count][R]
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
	for (int j = threadIdx.x; j < NV; j += RF_THREADS)
	{
		double s = 0.0;
#pragma unroll
		for (int w = 0; w < RF_WARPS; ++w)
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
__device__ void fullpiv_householder_q_lastcol9(double *M, double *q)
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

template <int PROBLEM>
__global__ void __launch_bounds__(RF_THREADS) k_refit(const RefitArgs a)
{
	constexpr int NV = PROBLEM == GCR_H ? 44 : 45;
	__shared__ double red[RF_WARPS * 45];
	__shared__ double tot[45];
	const int set = blockIdx.x, tid = threadIdx.x;
	const uint32_t *idx = a.sets + a.offsets[set];
	const int n = (int)a.sizes[set];
	const int min_n = PROBLEM == GCR_H ? 4 : 8;
	if (n < min_n)
	{
		if (tid == 0)
		{
			for (int j = 0; j < 9; ++j) a.models[(size_t)set * 9 + j] = 0.0;
			a.ok[set] = 0;
		}
		return;
	}
	// ---- normalizePoints: centroids ----
	{
		double v[4] = {0, 0, 0, 0};
		for (int i = tid; i < n; i += RF_THREADS)
		{
			const double *p = a.pts + (size_t)idx[i] * 4;
			v[0] += p[0]; v[1] += p[1]; v[2] += p[2]; v[3] += p[3];
		}
		block_sum<4>(v, red, tot);
	}
	const double msx = tot[0] / n, msy = tot[1] / n, mdx = tot[2] / n, mdy = tot[3] / n;
	__syncthreads();
	// ---- mean distances ----
	{
		double v[2] = {0, 0};
		for (int i = tid; i < n; i += RF_THREADS)
		{
			const double *p = a.pts + (size_t)idx[i] * 4;
			const double dx1 = msx - p[0], dy1 = msy - p[1], dx2 = mdx - p[2], dy2 = mdy - p[3];
			v[0] += sqrt(dx1 * dx1 + dy1 * dy1);
			v[1] += sqrt(dx2 * dx2 + dy2 * dy2);
		}
		block_sum<2>(v, red, tot);
	}
	const double rs = 1.4142135623730951 / (tot[0] / n), rd = 1.4142135623730951 / (tot[1] / n);
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
			for (int j = 0; j < 9; ++j) a.models[(size_t)set * 9 + j] = ok ? H[j] : 0.0;
			a.ok[set] = ok ? 1 : 0;
		}
		return;
	}

	// ---- normal equations ----
	double acc[NV];
#pragma unroll
	for (int j = 0; j < NV; ++j) acc[j] = 0.0;
	for (int i = tid; i < n; i += RF_THREADS)
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
	for (int j = 0; j < 9; ++j) a.models[(size_t)set * 9 + j] = ok ? out[j] : 0.0;
	a.ok[set] = ok ? 1 : 0;
}

// Eigen ColPivHouseholderQR(M).matrixQ().rightCols<4>() of a 9 x 9 row-major M (destroyed): N[b][j] = Q(j, 5 + b).
// Same pivoting rule as Eigen 3.4 computeInPlace (first column of largest down-dated norm, LAPACK-style down-dating,
// recomputation below sqrt(eps)) and makeHouseholder's (beta, tau, essential) convention.
__device__ void colpiv_householder_q_last4_9(double *M, double (&N)[4][9])
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

__global__ void __launch_bounds__(RF_THREADS) k_refit_e(const RefitArgs a)
{
	__shared__ double red[RF_WARPS * 45];
	__shared__ double tot[45];
	const int set = blockIdx.x, tid = threadIdx.x;
	const uint32_t *idx = a.sets + a.offsets[set];
	const int n = (int)a.sizes[set];
	double *models = a.models + (size_t)set * RF_E_MODELS * 9;
	uint8_t *ok = a.ok + (size_t)set * RF_E_MODELS;
	if (n < 5) // nonMinimalSampleSize() (essential_estimator.h:315-316)
	{
		for (int j = tid; j < RF_E_MODELS * 9; j += RF_THREADS) models[j] = 0.0;
		if (tid < RF_E_MODELS) ok[tid] = 0;
		return;
	}
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
		for (int i = tid; i < n; i += RF_THREADS)
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
	if (tid != 0)
		return;
	double E[RF_E_MODELS * 9];
	int count = 0;
	if (have_basis)
	{
		bool finite = true;
		for (int b = 0; b < 4; ++b)
			for (int j = 0; j < 9; ++j)
				if (!isfinite(N[b][j])) finite = false;
		if (finite)
			count = gcr_e_nm::solve_from_basis(N, E);
	}
	// essential_estimator.h:327-333: Frobenius-normalise, sign so that E(2,2) >= 0
	for (int m = 0; m < RF_E_MODELS; ++m)
	{
		bool good = m < count;
		double out[9];
		if (good)
		{
			double nrm = 0.0;
			for (int i = 0; i < 9; ++i) nrm += E[9 * m + i] * E[9 * m + i];
			nrm = sqrt(nrm);
			for (int i = 0; i < 9; ++i) out[i] = E[9 * m + i] / nrm;
			if (out[8] < 0)
				for (int i = 0; i < 9; ++i) out[i] = -out[i];
			for (int i = 0; i < 9; ++i)
				if (isnan(out[i])) good = false;
		}
		for (int i = 0; i < 9; ++i) models[9 * m + i] = good ? out[i] : 0.0;
		ok[m] = good ? 1 : 0;
	}
}

extern "C" int gcr_max_refit_models(int problem)
{
	return problem == GCR_E ? RF_E_MODELS : (problem == GCR_H || problem == GCR_F) ? 1 : 0;
}

extern "C" gcr_status gcr_refit(gcr_ctx *ctx, int problem, int pair, const uint32_t *d_index_sets,
                                const uint32_t *d_set_offsets, const uint32_t *d_set_sizes, int count,
                                const double *d_weights, int weight_mode, double *d_models, uint8_t *d_ok)
{
	if (!ctx)
		return GCR_ERR_ARG;
	if (!ctx->d_pts64 || ctx->n == 0)
		GCR_FAIL(ctx, GCR_ERR_STATE, "gcr_refit: no points uploaded");
	if (ctx->cols != 4 || (problem != GCR_H && problem != GCR_F && problem != GCR_E))
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_refit: non-minimal refits exist for H, F (8-point) and E (Stewenius) on 4-column points");
	if (pair < 0 || pair >= ctx->n_pairs || count <= 0 || !d_index_sets || !d_set_offsets || !d_set_sizes || !d_models || !d_ok)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_refit: bad arguments");
	if (d_weights && weight_mode != GCR_WEIGHT_POSITIONAL && weight_mode != GCR_WEIGHT_BY_POINT)
		GCR_FAIL(ctx, GCR_ERR_ARG, "gcr_refit: weights given but weight_mode is %d", weight_mode);
	GCR_CHECK_CUDA(ctx, cudaSetDevice(ctx->device));
	RefitArgs a;
	a.pts = ctx->d_pts64 + (size_t)pair * ctx->n * 4;
	a.sets = d_index_sets;
	a.offsets = d_set_offsets;
	a.sizes = d_set_sizes;
	a.weights = d_weights;
	a.weight_mode = weight_mode;
	a.models = d_models;
	a.ok = d_ok;
	if (problem == GCR_H)
		k_refit<GCR_H><<<count, RF_THREADS, 0, ctx->stream>>>(a);
	else if (problem == GCR_E)
		k_refit_e<<<count, RF_THREADS, 0, ctx->stream>>>(a);
	else
		k_refit<GCR_F><<<count, RF_THREADS, 0, ctx->stream>>>(a);
	GCR_LAUNCHED(ctx);
	return GCR_OK;
}
