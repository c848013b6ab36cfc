// oracle/gcr_oracle_f.cpp -- TEST INFRASTRUCTURE ONLY (see gcr_oracle.h).
// Fundamental-matrix path: 7-point minimal solver, oriented epipolar constraint, normalised 8-point refit and the
// symmetric-epipolar validity test, restated from the reference (file:line beside each function).
// [pinned:source]: the reference versions call Eigen (FullPivLU::kernel, PolynomialSolver, FullPivHouseholderQR), which is
// neither vendored in /root/reference nor installed here; the restatement follows Eigen 3.4's documented semantics
// (SURVEY.md A.2), is cross-checked against numpy in tests/test_oracle_solvers.py and held against the reference's own
// source compiled verbatim over oracle/shim_solvers in tests/test_reference_solver_sources.py (7-point: same model sets to
// 1e-7, the order of the cubic's roots aside; 8-point refit 1e-9; the default estimator's BA refit and the polish bit for bit).
#include "gcr_oracle.h"
#include "linalg.h"

#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace
{
	// [same statements as graph-cut-ransac_b200/csrc/solve_f7.cuh]
	// Real roots of c0 + c1 x + c2 x^2 + c3 x^3 (c3 != 0), ascending; returns their number (1 or 3).
	// Closed form (Cardano for one real root, the trigonometric form for three) followed by three Newton steps on the
	// monic cubic: fixed work per call, no data-dependent iteration count (the earlier bracketing + safeguarded Newton
	// version spent ~9000 instructions per sample walking in from the Cauchy bound).  oracle/gcr_oracle_f.cpp runs the
	// same statements; libm's cbrt / acos / cos differ from CUDA's by <= 2 ulp, which the Newton steps absorb except
	// on near-multiple roots (tests compare at 1e-6, the north star allows 1e-5).
	int cubic_roots(double c0, double c1, double c2, double c3, double *roots)
	{
		const double a = c2 / c3, b = c1 / c3, c = c0 / c3;
		const double a3 = a / 3.0;
		const double p = b - a * a3;                              // depressed cubic t^3 + p t + q, x = t - a / 3
		const double q = 2.0 * a3 * a3 * a3 - a3 * b + c;
		const double hq = 0.5 * q, tp = p / 3.0;
		const double disc = hq * hq + tp * tp * tp;
		int n;
		if (disc > 0.0 || !(tp < 0.0))
		{
			const double s = std::sqrt(std::fmax(disc, 0.0));
			const double u = -std::copysign(std::cbrt(std::fabs(hq) + s), hq);   // the larger-magnitude cube root: no cancellation
			const double t = u != 0.0 ? u - tp / u : 0.0;
			roots[0] = t - a3;
			n = 1;
		}
		else
		{
			const double m = 2.0 * std::sqrt(-tp);
			double arg = -hq / (-tp * std::sqrt(-tp));                  // std::cos(3 theta)
			arg = std::fmin(1.0, std::fmax(-1.0, arg));
			const double theta = std::acos(arg) / 3.0;
			const double two_pi_3 = 2.0943951023931954923;
			roots[0] = m * std::cos(theta + two_pi_3) - a3;             // ascending for theta in [0, pi / 3]
			roots[1] = m * std::cos(theta - two_pi_3) - a3;
			roots[2] = m * std::cos(theta) - a3;
			n = 3;
		}
		for (int r = 0; r < 3; ++r)
		{
			if (r < n)
			{
				double x = roots[r];
				for (int it = 0; it < 3; ++it)
				{
					const double f = ((x + a) * x + b) * x + c;
					const double d = (3.0 * x + 2.0 * a) * x + b;
					if (d != 0.0 && std::isfinite(f / d))
						x -= f / d;
				}
				roots[r] = x;
			}
		}
		if (n == 3) // Newton can reorder near-coincident roots
		{
			if (roots[0] > roots[1]) { const double t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
			if (roots[1] > roots[2]) { const double t = roots[1]; roots[1] = roots[2]; roots[2] = t; }
			if (roots[0] > roots[1]) { const double t = roots[0]; roots[0] = roots[1]; roots[1] = t; }
		}
		return n;
	}
}

extern "C" {

// solver_fundamental_matrix_seven_point.h:91-291, minimal branch (:159-170), weights_ == nullptr
int gcro_solve_f7(const double *pts, const uint32_t *sample, double *F27)
{
	const double eps = std::numeric_limits<double>::epsilon();
	double A[7 * 9];
	for (int i = 0; i < 7; ++i)
	{
		const double *p = pts + 4 * (size_t)sample[i];
		const double x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
		double *r = A + 9 * i;
		r[0] = x1 * x0; r[1] = x1 * y0; r[2] = x1; r[3] = y1 * x0; r[4] = y1 * y0; r[5] = y1; r[6] = x0; r[7] = y0; r[8] = 1;
	}
	double ker[9 * 2];
	const int dim = gcro::fullpiv_lu_kernel(A, 7, 9, ker, 2);
	if (dim != 2) // :162-163
		return 0;
	double f1[9], f2[9];
	for (int j = 0; j < 9; ++j)
	{
		f1[j] = ker[2 * j + 0];
		f2[j] = ker[2 * j + 1];
	}
	for (int j = 0; j < 9; ++j) // :190
		f1[j] -= f2[j];

	double c[4], t0, t1, t2;
	t0 = f2[4] * f2[8] - f2[5] * f2[7];
	t1 = f2[3] * f2[8] - f2[5] * f2[6];
	t2 = f2[3] * f2[7] - f2[4] * f2[6];
	c[0] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2;
	c[1] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2 -
	       f1[3] * (f2[1] * f2[8] - f2[2] * f2[7]) +
	       f1[4] * (f2[0] * f2[8] - f2[2] * f2[6]) -
	       f1[5] * (f2[0] * f2[7] - f2[1] * f2[6]) +
	       f1[6] * (f2[1] * f2[5] - f2[2] * f2[4]) -
	       f1[7] * (f2[0] * f2[5] - f2[2] * f2[3]) +
	       f1[8] * (f2[0] * f2[4] - f2[1] * f2[3]);
	t0 = f1[4] * f1[8] - f1[5] * f1[7];
	t1 = f1[3] * f1[8] - f1[5] * f1[6];
	t2 = f1[3] * f1[7] - f1[4] * f1[6];
	c[2] = f2[0] * t0 - f2[1] * t1 + f2[2] * t2 -
	       f2[3] * (f1[1] * f1[8] - f1[2] * f1[7]) +
	       f2[4] * (f1[0] * f1[8] - f1[2] * f1[6]) -
	       f2[5] * (f1[0] * f1[7] - f1[1] * f1[6]) +
	       f2[6] * (f1[1] * f1[5] - f1[2] * f1[4]) -
	       f2[7] * (f1[0] * f1[5] - f1[2] * f1[3]) +
	       f2[8] * (f1[0] * f1[4] - f1[1] * f1[3]);
	c[3] = f1[0] * t0 - f1[1] * t1 + f1[2] * t2;

	if (std::fabs(c[0] + c[1] + c[2] + c[3]) < 1e-9 || std::fabs(c[0]) < eps || std::fabs(c[1]) < eps ||
	    std::fabs(c[2]) < eps || std::fabs(c[3]) < eps) // :246-251
		return 0;

	// :254-260 Eigen::PolynomialSolver<double,3>::realRoots -- restated as a direct real-root finder, ascending order
	double roots[3];
	const int nr = cubic_roots(c[0], c[1], c[2], c[3], roots);
	int nm = 0;
	for (int r = 0; r < nr; ++r)
	{
		double lambda = roots[r], mu = 1.;
		const double s = f1[8] * roots[r] + f2[8];
		if (std::fabs(s) > eps) // :273
		{
			mu = 1.0f / s;
			lambda *= mu;
			double *o = F27 + 9 * nm;
			for (int j = 0; j < 8; ++j)
				o[j] = f1[j] * lambda + f2[j] * mu;
			o[8] = 1.0;
			++nm;
		}
	}
	return nm;
}

// fundamental_estimator.h:737-800 (getEpipole, getOrientationSignum, isOrientationValid)
int gcro_f_orientation_valid(const double *F, const double *pts, const uint32_t *sample, int m)
{
	const double epsilon = 1.9984e-15;
	double e[3];
	gcro::cross3(F + 0, F + 6, e); // row(0).cross(row(2))
	bool big = false;
	for (int i = 0; i < 3; i++)
		if ((e[i] > epsilon) || (e[i] < -epsilon))
			big = true;
	if (!big)
		gcro::cross3(F + 3, F + 6, e);
	double signum2 = 0.0;
	for (int i = 0; i < m; ++i)
	{
		const double *p = pts + 4 * (size_t)(sample ? sample[i] : (uint32_t)i);
		const double s1 = F[0] * p[2] + F[3] * p[3] + F[6], s2 = e[1] - e[2] * p[1];
		const double signum1 = s1 * s2;
		if (i == 0)
			signum2 = signum1;
		else if (signum2 * signum1 < 0)
			return 0;
	}
	return 1;
}

// fundamental_estimator.h:161-184
int gcro_estimate_f7(const double *pts, const uint32_t *sample, double *F27)
{
	double all[27];
	const int n = gcro_solve_f7(pts, sample, all);
	int kept = 0;
	for (int i = 0; i < n; ++i)
		if (gcro_f_orientation_valid(all + 9 * i, pts, sample, 7))
		{
			std::memcpy(F27 + 9 * kept, all + 9 * i, 9 * sizeof(double));
			++kept;
		}
	return kept;
}

// FullPivHouseholderQR(M).matrixQ().rightCols<1>() for a square n x n matrix M (solver_fundamental_matrix_eight_point.h
// :170-174).  Eigen: at step k the largest |entry| of the remaining corner is brought to (k,k) by a row and a column
// transposition, then a Householder reflection zeroes the column below it; matrixQ() = product over k of
// (row transposition_k * H_k).  Its last column is Q e_{n-1}.
static void fullpiv_householder_q_lastcol(const double *M_in, int n, double *q)
{
	std::vector<double> M(M_in, M_in + (size_t)n * n);
	std::vector<int> rowswap(n);
	std::vector<std::vector<double>> vs(n); // essential parts
	std::vector<double> taus(n, 0.0);
	int steps = n;
	for (int k = 0; k < n; ++k)
	{
		int pr = k, pc = k;
		double biggest = 0.0;
		for (int j = k; j < n; ++j)
			for (int i = k; i < n; ++i)
			{
				const double a = std::fabs(M[(size_t)i * n + j]);
				if (a > biggest) { biggest = a; pr = i; pc = j; }
			}
		if (biggest == 0.0) // Eigen: corner exactly zero -> remaining reflections are identity
		{
			steps = k;
			break;
		}
		rowswap[k] = pr;
		if (pr != k)
			for (int j = 0; j < n; ++j) std::swap(M[(size_t)k * n + j], M[(size_t)pr * n + j]);
		if (pc != k)
			for (int i = 0; i < n; ++i) std::swap(M[(size_t)i * n + k], M[(size_t)i * n + pc]);
		// makeHouseholderInPlace on column k, rows k..n-1: H = I - tau v v^T, v = [1; essential]
		double tail2 = 0.0;
		for (int i = k + 1; i < n; ++i) tail2 += M[(size_t)i * n + k] * M[(size_t)i * n + k];
		const double c0 = M[(size_t)k * n + k];
		std::vector<double> ess(n - k - 1, 0.0);
		double tau = 0.0, beta = c0;
		if (tail2 > std::numeric_limits<double>::min())
		{
			beta = std::sqrt(c0 * c0 + tail2);
			if (c0 >= 0) beta = -beta;
			for (int i = k + 1; i < n; ++i) ess[i - k - 1] = M[(size_t)i * n + k] / (c0 - beta);
			tau = (beta - c0) / beta;
		}
		vs[k] = ess;
		taus[k] = tau;
		// apply H to the remaining columns (k+1..n-1) from the left
		for (int j = k + 1; j < n; ++j)
		{
			double dot = M[(size_t)k * n + j];
			for (int i = k + 1; i < n; ++i) dot += ess[i - k - 1] * M[(size_t)i * n + j];
			dot *= tau;
			M[(size_t)k * n + j] -= dot;
			for (int i = k + 1; i < n; ++i) M[(size_t)i * n + j] -= dot * ess[i - k - 1];
		}
		M[(size_t)k * n + k] = beta;
		for (int i = k + 1; i < n; ++i) M[(size_t)i * n + k] = 0.0;
	}
	// q = Q e_{n-1} with Q = T_0 H_0 T_1 H_1 ... : apply in reverse order
	for (int i = 0; i < n; ++i) q[i] = 0.0;
	q[n - 1] = 1.0;
	for (int k = steps - 1; k >= 0; --k)
	{
		// H_k
		double dot = q[k];
		for (int i = k + 1; i < n; ++i) dot += vs[k][i - k - 1] * q[i];
		dot *= taus[k];
		q[k] -= dot;
		for (int i = k + 1; i < n; ++i) q[i] -= dot * vs[k][i - k - 1];
		// T_k
		if (rowswap[k] != k) std::swap(q[k], q[rowswap[k]]);
	}
}

// exported for oracle/shim_solvers (the reference's eight-point solver compiled verbatim calls the same routine)
void gcro_fullpiv_householder_q_lastcol(const double *M_in, int n, double *q) { fullpiv_householder_q_lastcol(M_in, n, q); }

// fundamental_estimator.h:574-618 (+ normalizePoints :637-734) + solver_fundamental_matrix_eight_point.h:93-182.
// The 8-point solver receives the normalised copy with sample_ == nullptr, so it reads weights_[i] by POSITION (:113-118).
// FundamentalMatrixEightPointSolver::estimateModel (solver_fundamental_matrix_eight_point.h:93-182) on n points given
// in the working frame (rows x0 y0 x1 y1), positional weights: the null vector of A^T A as a 3 x 3 matrix
static void f8_in_frame(const double *np, int n, const double *weights, double *f)
{
	std::vector<double> A((size_t)n * 9);
	for (int i = 0; i < n; ++i)
	{
		const double *p = np + 4 * (size_t)i;
		const double x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
		double *r = A.data() + 9 * (size_t)i;
		if (!weights)
		{
			r[0] = x1 * x0; r[1] = x1 * y0; r[2] = x1; r[3] = y1 * x0; r[4] = y1 * y0; r[5] = y1; r[6] = x0; r[7] = y0; r[8] = 1;
		}
		else
		{
			const double weight = weights[i];
			const double wx0 = weight * x0, wy0 = weight * y0, wx1 = weight * x1, wy1 = weight * y1;
			r[0] = wx1 * x0; r[1] = wx1 * y0; r[2] = wx1; r[3] = wy1 * x0; r[4] = wy1 * y0; r[5] = wy1; r[6] = wx0; r[7] = wy0; r[8] = weight;
		}
	}
	// coefficients.transpose() * coefficients (:171)
	double AtA[81];
	for (int a = 0; a < 9; ++a)
		for (int b = 0; b < 9; ++b)
		{
			double s = 0.0;
			for (int i = 0; i < n; ++i)
				s += A[9 * (size_t)i + a] * A[9 * (size_t)i + b];
			AtA[9 * a + b] = s;
		}
	fullpiv_householder_q_lastcol(AtA, 9, f);
}

// fundamental_estimator.h:606-616: F = T2^T * Fn * T1, Frobenius-normalise, sign so that F(2,2) >= 0
static void f_denormalise(const double *f, const double *T1, const double *T2, double *F9)
{
	double T2t[9], tmp[9], F[9];
	gcro::transpose3(T2, T2t);
	gcro::mul3(T2t, f, tmp);
	gcro::mul3(tmp, T1, F);
	double nrm = 0.0;
	for (int i = 0; i < 9; ++i) nrm += F[i] * F[i];
	nrm = std::sqrt(nrm);
	for (int i = 0; i < 9; ++i) F[i] /= nrm;
	if (F[8] < 0)
		for (int i = 0; i < 9; ++i) F[i] = -F[i];
	std::memcpy(F9, F, sizeof(F));
}

int gcro_refit_f8(const double *pts, const uint32_t *sample, int n, const double *weights, double *F9)
{
	if (n < 8)
		return 0;
	std::vector<double> np((size_t)n * 4);
	double T1[9], T2[9], f[9];
	gcro_normalize_points(pts, 4, sample, n, np.data(), T1, T2);
	f8_in_frame(np.data(), n, weights, f);
	f_denormalise(f, T1, T2, F9);
	return 1;
}

// ---- LM bundle adjustment of F (relative_pose/bundle.cpp:253-339 lm_F_impl, relative_pose/jacobian_impl.h:472-606) ----
// F = U diag(1, sigma, 0) V^T (FactorizedFundamentalMatrix: 7 parameters, rank 2 by construction), Sampson residual,
// TRUNCATED loss with scale 1 (solver_fundamental_matrix_bundle_adjustment.h:56-62, bundle.h:106-118), lambda0 = 1e-3,
// x / : 10, gradient / step tolerance 1e-8, <= 25 iterations, positional weights (jacobian_impl.h:525-528, 573-577).
namespace
{
	struct FactorizedF
	{
		double U[9], V[9], sigma;
		void F(double *out) const // :481-483
		{
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
					out[3 * r + c] = U[3 * r] * V[3 * c] + sigma * U[3 * r + 1] * V[3 * c + 1];
		}
	};

	double lm_f_cost(const double *np, int n, const double *weights, const FactorizedF &FF)
	{
		double F[9];
		FF.F(F);
		double cost = 0.0;
		for (int k = 0; k < n; ++k)
		{
			const double *p = np + 4 * (size_t)k;
			const double x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
			const double Fx0 = F[0] * x1 + F[1] * y1 + F[2], Fx1 = F[3] * x1 + F[4] * y1 + F[5], Fx2 = F[6] * x1 + F[7] * y1 + F[8];
			const double C = x2 * Fx0 + y2 * Fx1 + Fx2;
			const double Ft0 = F[0] * x2 + F[3] * y2 + F[6], Ft1 = F[1] * x2 + F[4] * y2 + F[7];
			const double nJc_sq = (Fx0 * Fx0 + Fx1 * Fx1) + (Ft0 * Ft0 + Ft1 * Ft1);
			const double r2 = (C * C) / nJc_sq;
			const double loss = std::min(r2, 1.0);
			cost += weights ? weights[k] * loss : loss;
		}
		return cost;
	}

	void lm_f_accumulate(const double *np, int n, const double *weights, const FactorizedF &FF, double *JtJ /*7x7 lower*/,
	                     double *Jtr)
	{
		double F[9];
		FF.F(F);
		double ds[9]; // d_sigma = U.col(1) * V.col(1)^T
		for (int r = 0; r < 3; ++r)
			for (int c = 0; c < 3; ++c)
				ds[3 * r + c] = FF.U[3 * r + 1] * FF.V[3 * c + 1];
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
		for (int k = 0; k < n; ++k)
		{
			const double *p = np + 4 * (size_t)k;
			const double x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
			const double Fx0 = F[0] * x1 + F[1] * y1 + F[2], Fx1 = F[3] * x1 + F[4] * y1 + F[5], Fx2 = F[6] * x1 + F[7] * y1 + F[8];
			const double C = x2 * Fx0 + y2 * Fx1 + Fx2;
			// J_C = [F(:, 0:1)^T x2h ; F(0:1, :) x1h]
			const double JC[4] = {F[0] * x2 + F[3] * y2 + F[6], F[1] * x2 + F[4] * y2 + F[7], Fx0, Fx1};
			const double nJ_C = std::sqrt(JC[0] * JC[0] + JC[1] * JC[1] + JC[2] * JC[2] + JC[3] * JC[3]);
			const double inv_nJ_C = 1.0 / nJ_C;
			const double r = C * inv_nJ_C;
			double weight = ((r * r) < 1.0 ? 1.0 : 0.0) / n;
			if (weights)
				weight = weights[k] * weight;
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
			for (int i = 0; i < 9; ++i) dF[i] *= inv_nJ_C;
			double J[7];
			for (int j = 0; j < 7; ++j)
			{
				double acc = 0.0;
				for (int i = 0; i < 9; ++i) acc += dF[i] * dFdp[i][j];
				J[j] = acc;
			}
			for (int i = 0; i < 7; ++i)
			{
				Jtr[i] += weight * C * inv_nJ_C * J[i];
				for (int j = 0; j <= i; ++j)
					JtJ[7 * i + j] += weight * (J[i] * J[j]);
			}
		}
	}

	// x = -(L L^T)^-1 b from the lower triangle of A (7 x 7); NaNs propagate when A is not positive definite, as with
	// Eigen's unchecked llt().solve
	void llt_solve7_neg(const double *A, const double *b, double *x)
	{
		double L[49];
		for (int j = 0; j < 7; ++j)
		{
			double d = A[7 * j + j];
			for (int k = 0; k < j; ++k) d -= L[7 * j + k] * L[7 * j + k];
			d = std::sqrt(d);
			L[7 * j + j] = d;
			for (int i = j + 1; i < 7; ++i)
			{
				double s = A[7 * i + j];
				for (int k = 0; k < j; ++k) s -= L[7 * i + k] * L[7 * j + k];
				L[7 * i + j] = s / d;
			}
		}
		double y[7];
		for (int i = 0; i < 7; ++i)
		{
			double s = b[i];
			for (int k = 0; k < i; ++k) s -= L[7 * i + k] * y[k];
			y[i] = s / L[7 * i + i];
		}
		for (int i = 6; i >= 0; --i)
		{
			double s = y[i];
			for (int k = i + 1; k < 7; ++k) s -= L[7 * k + i] * x[k];
			x[i] = s / L[7 * i + i];
		}
		for (int i = 0; i < 7; ++i) x[i] = -x[i];
	}

	// M <- M + (a [w]x + (1 - b) [w]x^2) M   (:290-316)
	void lie_update(const double *w3, const double *M, double *out)
	{
		double w[3] = {w3[0], w3[1], w3[2]};
		const double theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
		for (int i = 0; i < 3; ++i) w[i] /= theta;
		const double a = std::sin(theta), b = std::cos(theta);
		const double sw[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
		double sw2[9], R[9], RM[9];
		gcro::mul3(sw, sw, sw2);
		for (int i = 0; i < 9; ++i) R[i] = a * sw[i] + (1 - b) * sw2[i];
		gcro::mul3(R, M, RM);
		for (int i = 0; i < 9; ++i) out[i] = M[i] + RM[i];
	}
}

// refine_fundamental (bundle.cpp:454-499) with TruncatedLoss(1.0): F9 in / out, points in the working frame
int gcro_refine_f_lm(const double *np, int n, const double *weights, double *F9, int max_iterations)
{
	FactorizedF F;
	{
		double s[3], Ucols[9], Vcols[9];
		gcro::jacobi_svd(F9, 3, 3, Ucols, s, Vcols); // FactorizedFundamentalMatrix(F): JacobiSVD, sigma = s1 / s0 (:474-480)
		std::memcpy(F.U, Ucols, sizeof(Ucols));
		std::memcpy(F.V, Vcols, sizeof(Vcols));
		F.sigma = s[1] / s[0];
	}
	double lambda = 1e-3;
	double cost = lm_f_cost(np, n, weights, F);
	bool recompute_jac = true;
	double JtJ[49], Jtr[7];
	int iter;
	for (iter = 0; iter < max_iterations; ++iter)
	{
		if (recompute_jac)
		{
			for (int i = 0; i < 49; ++i) JtJ[i] = 0.0;
			for (int i = 0; i < 7; ++i) Jtr[i] = 0.0;
			lm_f_accumulate(np, n, weights, F, JtJ, Jtr);
			double g = 0.0;
			for (int i = 0; i < 7; ++i) g += Jtr[i] * Jtr[i];
			if (std::sqrt(g) < 1e-8)
				break;
		}
		for (int k = 0; k < 7; ++k) JtJ[7 * k + k] += lambda;
		double sol[7];
		llt_solve7_neg(JtJ, Jtr, sol);
		double sn = 0.0;
		for (int i = 0; i < 7; ++i) sn += sol[i] * sol[i];
		if (std::sqrt(sn) < 1e-8)
			break;
		FactorizedF Fn;
		lie_update(sol, F.U, Fn.U);
		lie_update(sol + 3, F.V, Fn.V);
		Fn.sigma = F.sigma + sol[6];
		const double cost_new = lm_f_cost(np, n, weights, Fn);
		if (cost_new < cost)
		{
			F = Fn;
			lambda /= 10;
			cost = cost_new;
			recompute_jac = true;
		}
		else
		{
			for (int k = 0; k < 7; ++k) JtJ[7 * k + k] -= lambda;
			lambda *= 10;
			recompute_jac = false;
		}
	}
	F.F(F9);
	return iter;
}

// FundamentalMatrixEstimator<SevenPoint, FundamentalMatrixBundleAdjustmentSolver>::estimateModelNonminimal
// (fundamental_estimator.h:574-618; the reference's DEFAULT F estimator, types.h:78-80): Hartley normalisation, the
// 8-point initialisation and the LM refinement both in the normalised frame
// (solver_fundamental_matrix_bundle_adjustment.h:113-177), then T2^T F T1.  n == 7 (the 7-point initialisation with up
// to three models) is not restated: returns 0.
int gcro_refit_f_ba(const double *pts, const uint32_t *sample, int n, const double *weights, double *F9)
{
	if (n < 8)
		return 0;
	std::vector<double> np((size_t)n * 4);
	double T1[9], T2[9], f[9];
	gcro_normalize_points(pts, 4, sample, n, np.data(), T1, T2);
	f8_in_frame(np.data(), n, weights, f);
	gcro_refine_f_lm(np.data(), n, weights, f, 25);
	f_denormalise(f, T1, T2, F9);
	return 1;
}

// The post-run polish of findFundamentalMatrix_ (py/gcransac_python.cpp:1851-1858): the BA solver called directly on
// the pixel coordinates of the inliers -- un-normalised 8-point initialisation, LM in pixels.  Output not rescaled.
int gcro_polish_f_ba(const double *pts, const uint32_t *inliers, int n, const double *weights, double *F9)
{
	if (n < 8)
		return 0;
	std::vector<double> sel((size_t)n * 4);
	for (int i = 0; i < n; ++i)
		std::memcpy(sel.data() + 4 * (size_t)i, pts + 4 * (size_t)inliers[i], 4 * sizeof(double));
	f8_in_frame(sel.data(), n, weights, F9);
	gcro_refine_f_lm(sel.data(), n, weights, F9, 25);
	return 1;
}

// ---- DEGENSAC (fundamental_estimator.h:341-572) -------------------------------------------------------------------
// First half, :341-500: is the minimal sample H-degenerate?  The epipole of the second image comes from
// JacobiSVD(F).matrixU().col(2) / U(2,2) (:365-372) -- here the right null vector of F^T from a one-sided Jacobi SVD
// (V is orthonormal by construction; U's column for the vanishing singular value is not).  For each of the five point
// triplets (:349-355) the plane homography H = A - e (M^-1 b)^T with A = [e]_x F (:380-442) is tested on the other
// sample points with the forward transfer error (:445-483); >= 5 consistent points of 7 make the sample H-degenerate.
int gcro_degensac_sample_homography(const double *F9, const double *pts, const uint32_t *sample, double sq_h_threshold,
                                    double *H9)
{
	static const int triplets[15] = {0, 1, 2, 3, 4, 5, 0, 1, 6, 3, 4, 6, 2, 5, 6};
	double Ft[9], U[9], sv[3], V[9];
	gcro::transpose3(F9, Ft);
	gcro::jacobi_svd(Ft, 3, 3, U, sv, V);
	const double e[3] = {V[2] / V[8], V[5] / V[8], V[8] / V[8]};
	const double ex[9] = {0, -e[2], e[1], e[2], 0, -e[0], -e[1], e[0], 0};
	double A[9];
	gcro::mul3(ex, F9, A);
	for (int t = 0; t < 5; ++t)
	{
		uint32_t idx[3];
		double p1[3][3], p2[3][3], b[3];
		for (int q = 0; q < 3; ++q)
		{
			idx[q] = sample[triplets[3 * t + q]];
			const double *p = pts + 4 * (size_t)idx[q];
			p1[q][0] = p[0]; p1[q][1] = p[1]; p1[q][2] = 1;
			p2[q][0] = p[2]; p2[q][1] = p[3]; p2[q][2] = 1;
			double pxe[3], Ap[3], pxAp[3];
			gcro::cross3(p2[q], e, pxe);
			for (int r = 0; r < 3; ++r)
				Ap[r] = A[3 * r] * p1[q][0] + A[3 * r + 1] * p1[q][1] + A[3 * r + 2] * p1[q][2];
			gcro::cross3(p2[q], Ap, pxAp);
			b[q] = (pxAp[0] * pxe[0] + pxAp[1] * pxe[1] + pxAp[2] * pxe[2]) / (pxe[0] * pxe[0] + pxe[1] * pxe[1] + pxe[2] * pxe[2]);
		}
		const double M[9] = {p1[0][0], p1[0][1], p1[0][2], p1[1][0], p1[1][1], p1[1][2], p1[2][0], p1[2][1], p1[2][2]};
		double Mi[9], Mb[3], H[9];
		gcro::inv3(M, Mi);
		for (int r = 0; r < 3; ++r)
			Mb[r] = Mi[3 * r] * b[0] + Mi[3 * r + 1] * b[1] + Mi[3 * r + 2] * b[2];
		for (int r = 0; r < 3; ++r)
			for (int c = 0; c < 3; ++c)
				H[3 * r + c] = A[3 * r + c] - e[r] * Mb[c];
		int inlier_number = 3;
		for (int i = 0; i < 7; ++i)
		{
			const uint32_t id = sample[i];
			if (id == idx[0] || id == idx[1] || id == idx[2])
				continue;
			const double *p = pts + 4 * (size_t)id;
			const double t1 = H[0] * p[0] + H[1] * p[1] + H[2], t2 = H[3] * p[0] + H[4] * p[1] + H[5],
			             t3 = H[6] * p[0] + H[7] * p[1] + H[8];
			const double d1 = p[2] - (t1 / t3), d2 = p[3] - (t2 / t3);
			if (d1 * d1 + d2 * d2 < sq_h_threshold)
				++inlier_number;
		}
		if (inlier_number >= 5)
		{
			std::memcpy(H9, H, sizeof(H));
			return 1;
		}
	}
	return 0;
}

// FundamentalMatrixPlaneParallaxSolver::estimateModel (solver_fundamental_matrix_plane_and_parallax.h:109-160): the
// epipole is the intersection of the two lines joining H x1 and x2; F = [e]_x H.
int gcro_solve_f_plane_parallax(const double *pts, const uint32_t *sample, const double *H9, double *F9)
{
	double line[2][3];
	for (int q = 0; q < 2; ++q)
	{
		const double *p = pts + 4 * (size_t)sample[q];
		const double src[3] = {p[0], p[1], 1}, dst[3] = {p[2], p[3], 1};
		double proj[3];
		for (int r = 0; r < 3; ++r)
			proj[r] = H9[3 * r] * src[0] + H9[3 * r + 1] * src[1] + H9[3 * r + 2] * src[2];
		gcro::cross3(proj, dst, line[q]);
	}
	double e[3];
	gcro::cross3(line[0], line[1], e);
	if (std::fabs(e[2]) < 2.220446049250313e-16)
		return 0;
	const double ex[9] = {0, -e[2], e[1], e[2], 0, -e[0], -e[1], e[0], 0};
	gcro::mul3(ex, H9, F9);
	return 1;
}

// FundamentalMatrixEstimator<PlaneParallax, EightPoint>::estimateModel (fundamental_estimator.h:161-184): the solver +
// the oriented epipolar constraint on the two sample points
int gcro_estimate_f_plane_parallax(const double *pts, const uint32_t *sample, const double *H9, double *F9)
{
	if (!gcro_solve_f_plane_parallax(pts, sample, H9, F9))
		return 0;
	return gcro_f_orientation_valid(F9, pts, sample, 2) ? 1 : 0;
}

// fundamental_estimator.h:287-318 (and essential_estimator.h:279-305): the symmetric-epipolar part of isValidModel
int gcro_f_valid_model_ratio(const double *F9, const double *pts, const uint32_t *inl, int n_inl, int m,
                             double truncated_thr, double min_ratio)
{
	size_t inlier_number = 0;
	const double a = (double)m, b = n_inl * min_ratio;
	const size_t minimum_inlier_number = (size_t)(a > b ? a : b); // const size_t x = MAX(sample_size, inliers_.size() * ratio): truncated
	const double squared_threshold = truncated_thr * truncated_thr;
	for (int i = 0; i < n_inl; ++i)
		if (gcro_sqres_symepi(pts + 4 * (size_t)inl[i], F9) < squared_threshold)
			if (++inlier_number >= minimum_inlier_number)
				return 1;
	return 0;
}

} // extern "C"
