// solve_f7.cuh -- device 7-point fundamental solver (one thread per sample, FP64).
// Follows FundamentalMatrixSevenPointSolver::estimateModel (estimators/solver_fundamental_matrix_seven_point.h:91-291)
// for the minimal case (sample_number_ == 7, no weights) and the oriented epipolar constraint of
// estimators/fundamental_estimator.h:737-800.  Differences from the reference, by design:
//   * the null space comes from a full-pivot LU written out here with Eigen 3.4 FullPivLU::kernel() semantics
//     (column-major pivot search, rank threshold eps * 7 * max pivot, free variables set to unit vectors);
//   * the cubic is solved by bracketing at the derivative's roots + safeguarded Newton instead of the companion-matrix
//     eigen-solve of Eigen::PolynomialSolver; real roots come out in ASCENDING order (Eigen's order is unspecified).
#pragma once

namespace gcr_f7
{
	struct Cubic
	{
		double a, b, c; // monic x^3 + a x^2 + b x + c
		__device__ __forceinline__ double p(double x) const { return ((x + a) * x + b) * x + c; }
		__device__ __forceinline__ double dp(double x) const { return (3.0 * x + 2.0 * a) * x + b; }

		// root inside [lo, hi] where p changes sign; Newton with bisection safeguard ("rtsafe")
		__device__ double bracket(double lo, double hi) const
		{
			const double fl = p(lo), fh = p(hi);
			if (fl == 0.0) return lo;
			if (fh == 0.0) return hi;
			double xl, xh;
			if (fl < 0.0) { xl = lo; xh = hi; }
			else { xl = hi; xh = lo; }
			double rts = 0.5 * (lo + hi);
			double dxold = fabs(hi - lo), dx = dxold;
			double f = p(rts), df = dp(rts);
			for (int it = 0; it < 200; ++it)
			{
				if ((((rts - xh) * df - f) * ((rts - xl) * df - f) > 0.0) || (fabs(2.0 * f) > fabs(dxold * df)))
				{
					dxold = dx;
					dx = 0.5 * (xh - xl);
					rts = xl + dx;
					if (xl == rts) return rts;
				}
				else
				{
					dxold = dx;
					dx = f / df;
					const double temp = rts;
					rts -= dx;
					if (temp == rts) return rts;
				}
				if (fabs(dx) <= 4.440892098500626e-16 * fabs(rts)) return rts;
				f = p(rts);
				df = dp(rts);
				if (f < 0.0) xl = rts;
				else xh = rts;
			}
			return rts;
		}
	};

	// real roots of c0 + c1 x + c2 x^2 + c3 x^3 (c3 != 0), ascending; returns their number (1..3)
	__device__ inline int cubic_roots(double c0, double c1, double c2, double c3, double *roots)
	{
		Cubic q;
		q.a = c2 / c3;
		q.b = c1 / c3;
		q.c = c0 / c3;
		const double R = 1.0 + fmax(fabs(q.a), fmax(fabs(q.b), fabs(q.c))); // Cauchy bound
		const double disc = q.a * q.a - 3.0 * q.b;
		int n = 0;
		if (!(disc > 0.0))
		{
			roots[n++] = q.bracket(-R, R);
			return n;
		}
		const double sq = sqrt(disc);
		const double x1 = (-q.a - sq) / 3.0, x2 = (-q.a + sq) / 3.0; // local max at x1 < local min at x2
		const double p1 = q.p(x1), p2 = q.p(x2);
		if (p1 >= 0.0)
			roots[n++] = (p1 == 0.0) ? x1 : q.bracket(-R, x1);
		if (p1 > 0.0 && p2 < 0.0)
			roots[n++] = q.bracket(x1, x2);
		if (p2 <= 0.0)
			roots[n++] = (p2 == 0.0) ? x2 : q.bracket(x2, R);
		return n;
	}

	// pt: 7 x 4 (x0 y0 x1 y1), F: up to 3 x 9; returns the number of models
	__device__ inline int solve(const double *pt, double *F)
	{
		const double eps = 2.220446049250313e-16;
		double A[7][9];
		for (int i = 0; i < 7; ++i)
		{
			const double x0 = pt[4 * i], y0 = pt[4 * i + 1], x1 = pt[4 * i + 2], y1 = pt[4 * i + 3];
			A[i][0] = x1 * x0; A[i][1] = x1 * y0; A[i][2] = x1;
			A[i][3] = y1 * x0; A[i][4] = y1 * y0; A[i][5] = y1;
			A[i][6] = x0; A[i][7] = y0; A[i][8] = 1;
		}
		int colperm[9];
		for (int j = 0; j < 9; ++j)
			colperm[j] = j;
		double maxpivot = 0.0;
		int nonzero = 7;
		for (int k = 0; k < 7; ++k)
		{
			int pr = k, pc = k;
			double biggest = 0.0;
			for (int j = k; j < 9; ++j) // column-major search like Eigen's maxCoeff on a column-major block
				for (int i = k; i < 7; ++i)
				{
					const double a = fabs(A[i][j]);
					if (a > biggest) { biggest = a; pr = i; pc = j; }
				}
			if (biggest == 0.0) { nonzero = k; break; }
			if (biggest > maxpivot) maxpivot = biggest;
			if (pr != k)
				for (int j = 0; j < 9; ++j) { const double t = A[k][j]; A[k][j] = A[pr][j]; A[pr][j] = t; }
			if (pc != k)
			{
				for (int i = 0; i < 7; ++i) { const double t = A[i][k]; A[i][k] = A[i][pc]; A[i][pc] = t; }
				const int t = colperm[k]; colperm[k] = colperm[pc]; colperm[pc] = t;
			}
			const double piv = A[k][k];
			for (int i = k + 1; i < 7; ++i)
			{
				const double f = A[i][k] / piv;
				A[i][k] = f;
				for (int j = k + 1; j < 9; ++j)
					A[i][j] -= f * A[k][j];
			}
		}
		// dimensionOfKernel() == 2  <=>  all 7 pivots above eps * 7 * maxpivot (:161-163)
		const double premult = maxpivot * eps * 7.0;
		int rank = 0;
		for (int k = 0; k < nonzero; ++k)
			if (fabs(A[k][k]) > premult) ++rank;
		if (rank != 7)
			return 0;
		double f1[9], f2[9];
		for (int c = 0; c < 2; ++c)
		{
			double x[7];
			for (int i = 6; i >= 0; --i)
			{
				double s = -A[i][7 + c];
				for (int j = i + 1; j < 7; ++j)
					s -= A[i][j] * x[j];
				x[i] = s / A[i][i];
			}
			double *f = c == 0 ? f1 : f2;
			for (int i = 0; i < 7; ++i)
				f[colperm[i]] = x[i];
			f[colperm[7 + c]] = 1.0;
			f[colperm[7 + (1 - c)]] = 0.0;
		}
		for (int j = 0; j < 9; ++j) // :190 f1 -= f2
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

		if (fabs(c[0] + c[1] + c[2] + c[3]) < 1e-9 || fabs(c[0]) < eps || fabs(c[1]) < eps || fabs(c[2]) < eps ||
		    fabs(c[3]) < eps) // :246-251
			return 0;

		double roots[3];
		const int nr = cubic_roots(c[0], c[1], c[2], c[3], roots);
		int nm = 0;
		for (int r = 0; r < nr; ++r)
		{
			double lambda = roots[r], mu = 1.;
			const double s = f1[8] * roots[r] + f2[8];
			if (fabs(s) > eps) // :273
			{
				mu = 1.0 / s;
				lambda *= mu;
				double *o = F + 9 * nm;
				for (int j = 0; j < 8; ++j)
					o[j] = f1[j] * lambda + f2[j] * mu;
				o[8] = 1.0;
				++nm;
			}
		}
		return nm;
	}

	// fundamental_estimator.h:737-800 (sample given as m rows of x0 y0 x1 y1)
	__device__ inline bool orientation_valid(const double *F, const double *pt, int m)
	{
		const double epsilon = 1.9984e-15;
		double e[3];
		e[0] = F[1] * F[8] - F[2] * F[7]; // row0 x row2
		e[1] = F[2] * F[6] - F[0] * F[8];
		e[2] = F[0] * F[7] - F[1] * F[6];
		bool big = false;
		for (int i = 0; i < 3; ++i)
			if (e[i] > epsilon || e[i] < -epsilon)
				big = true;
		if (!big)
		{
			e[0] = F[4] * F[8] - F[5] * F[7]; // row1 x row2
			e[1] = F[5] * F[6] - F[3] * F[8];
			e[2] = F[3] * F[7] - F[4] * F[6];
		}
		double sig2 = 0.0;
		for (int i = 0; i < m; ++i)
		{
			const double *p = pt + 4 * i;
			const double s1 = F[0] * p[2] + F[3] * p[3] + F[6];
			const double s2 = e[1] - e[2] * p[1];
			const double sig = s1 * s2;
			if (i == 0)
				sig2 = sig;
			else if (sig2 * sig < 0)
				return false;
		}
		return true;
	}
}
