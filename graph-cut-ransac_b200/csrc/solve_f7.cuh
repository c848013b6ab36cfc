// solve_f7.cuh -- device 7-point fundamental solver (one thread per sample, FP64).
// Follows FundamentalMatrixSevenPointSolver::estimateModel (estimators/solver_fundamental_matrix_seven_point.h:91-291)
// for the minimal case (sample_number_ == 7, no weights) and the oriented epipolar constraint of
// estimators/fundamental_estimator.h:737-800.  Differences from the reference, by design:
//   * the null space comes from a full-pivot LU written out here with Eigen 3.4 FullPivLU::kernel() semantics
//     (column-major pivot search, rank threshold eps * 7 * max pivot, free variables set to unit vectors);
//   * the cubic is solved in closed form (Cardano / trigonometric) + three Newton steps instead of the companion-matrix
//     eigen-solve of Eigen::PolynomialSolver; real roots come out in ASCENDING order (Eigen's order is unspecified).
#pragma once

namespace gcr_f7
{
	// Real roots of c0 + c1 x + c2 x^2 + c3 x^3 (c3 != 0), ascending; returns their number (1 or 3).
	// Closed form (Cardano for one real root, the trigonometric form for three) followed by three Newton steps on the
	// monic cubic: fixed work per call, no data-dependent iteration count (the earlier bracketing + safeguarded Newton
	// version spent ~9000 instructions per sample walking in from the Cauchy bound).  oracle/gcr_oracle_f.cpp runs the
	// same statements; libm's cbrt / acos / cos differ from CUDA's by <= 2 ulp, which the Newton steps absorb except
	// on near-multiple roots (tests compare at 1e-6, the north star allows 1e-5).
	__device__ inline int cubic_roots(double c0, double c1, double c2, double c3, double *roots)
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
			const double s = sqrt(fmax(disc, 0.0));
			const double u = -copysign(cbrt(fabs(hq) + s), hq);   // the larger-magnitude cube root: no cancellation
			const double t = u != 0.0 ? u - tp / u : 0.0;
			roots[0] = t - a3;
			n = 1;
		}
		else
		{
			const double m = 2.0 * sqrt(-tp);
			double arg = -hq / (-tp * sqrt(-tp));                  // cos(3 theta)
			arg = fmin(1.0, fmax(-1.0, arg));
			const double theta = acos(arg) / 3.0;
			const double two_pi_3 = 2.0943951023931954923;
			roots[0] = m * cos(theta + two_pi_3) - a3;             // ascending for theta in [0, pi / 3]
			roots[1] = m * cos(theta - two_pi_3) - a3;
			roots[2] = m * cos(theta) - a3;
			n = 3;
		}
		for (int r = 0; r < 3; ++r)
		{
			if (r < n)
			{
				double x = roots[r];
#pragma unroll
				for (int it = 0; it < 3; ++it)
				{
					const double f = ((x + a) * x + b) * x + c;
					const double d = (3.0 * x + 2.0 * a) * x + b;
					if (d != 0.0 && isfinite(f / d))
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

	// The 7 x 9 epipolar matrix lives in SHARED memory, one column of doubles per thread (element e of thread t at
	// base[e * stride + t]): the full-pivot LU indexes rows / columns by data-dependent pivots, which in registers
	// would mean local memory (measured: 984 B stack per thread and 0.45 ms for 262144 samples; HBM/L2-bound spills).
	// 63 doubles per thread; the two null vectors overlay rows 0..1 of the matrix once both back-substitutions are done.
	struct SmemMat
	{
		double *base; // already offset by the thread index
		int stride;   // threads per CTA
		__device__ __forceinline__ double &operator()(int i, int j) const { return base[(i * 9 + j) * stride]; }
		__device__ __forceinline__ double &f(int which, int j) const { return base[(which * 9 + j) * stride]; }
	};

	// pt(i, c): sample point i, column c (x0 y0 x1 y1).  Returns the number of real roots (0..3, ascending) of
	// det(lambda f1 + (1 - lambda) f2) = 0 together with the two null vectors (f1 already holds f1 - f2, :190);
	// model(f1, f2, root, F) turns one root into a matrix.
	template <class PointFn>
	__device__ inline int solve(const PointFn &pt, const SmemMat &A, double *f1, double *f2, double *roots)
	{
		const double eps = 2.220446049250313e-16;
		// The full-pivot search of step k + 1 is folded into the elimination pass of step k (the updated entries are in
		// registers when they are stored), one independent running maximum per COLUMN and a strict '>' merge of the
		// columns in ascending order: the same element Eigen's column-major maxCoeff visitor picks, without the serial
		// 196-element compare chain (and its 196 shared-memory reloads) of a separate search.
		double biggest = 0.0;
		int pr = 0, pc = 0;
		{
			double colmax[9];
			int colarg[9];
#pragma unroll
			for (int j = 0; j < 9; ++j) { colmax[j] = 0.0; colarg[j] = 0; }
#pragma unroll
			for (int i = 0; i < 7; ++i)
			{
				const double x0 = pt(i, 0), y0 = pt(i, 1), x1 = pt(i, 2), y1 = pt(i, 3);
				double row[9];
				row[0] = x1 * x0; row[1] = x1 * y0; row[2] = x1;
				row[3] = y1 * x0; row[4] = y1 * y0; row[5] = y1;
				row[6] = x0; row[7] = y0; row[8] = 1;
#pragma unroll
				for (int j = 0; j < 9; ++j)
				{
					A(i, j) = row[j];
					const double a = fabs(row[j]);
					if (a > colmax[j]) { colmax[j] = a; colarg[j] = i; }
				}
			}
#pragma unroll
			for (int j = 0; j < 9; ++j)
				if (colmax[j] > biggest) { biggest = colmax[j]; pr = colarg[j]; pc = j; }
		}
		unsigned long long colperm = 0x876543210ull; // nibble j = original column now at position j
		double maxpivot = 0.0;
		double diag[7];
#pragma unroll
		for (int k = 0; k < 7; ++k)
		{
			if (biggest == 0.0) // FullPivLU stops here: fewer than 7 non-zero pivots, dimensionOfKernel() != 2
				return 0;
			if (biggest > maxpivot) maxpivot = biggest;
			// rows k <-> pr on the columns still in play (the L part is never read: kernel() only needs U and the
			// column permutation); unconditional, a swap with itself is harmless
			{
				double rk[9], rp[9];
#pragma unroll
				for (int j = k; j < 9; ++j) { rk[j] = A(k, j); rp[j] = A(pr, j); }
#pragma unroll
				for (int j = k; j < 9; ++j) { A(pr, j) = rk[j]; }
#pragma unroll
				for (int j = k; j < 9; ++j) { A(k, j) = rp[j]; }
			}
			// columns k <-> pc on every row (rows < k hold U)
			double fac[7];
			{
				double ck[7], cp[7];
#pragma unroll
				for (int i = 0; i < 7; ++i) { ck[i] = A(i, k); cp[i] = A(i, pc); }
#pragma unroll
				for (int i = 0; i < 7; ++i) { A(i, pc) = ck[i]; }
#pragma unroll
				for (int i = 0; i <= k; ++i) { A(i, k) = cp[i]; } // below the diagonal: L, never read again
				const unsigned long long a = (colperm >> (4 * k)) & 15ull, b = (colperm >> (4 * pc)) & 15ull;
				colperm &= ~((15ull << (4 * k)) | (15ull << (4 * pc)));
				colperm |= (b << (4 * k)) | (a << (4 * pc));
				diag[k] = cp[k]; // the pivot
#pragma unroll
				for (int i = k + 1; i < 7; ++i)
					fac[i] = cp[i] / cp[k];
			}
			biggest = 0.0;
			pr = pc = k + 1;
#pragma unroll
			for (int j = k + 1; j < 9; ++j)
			{
				const double r = A(k, j);
				double cm = 0.0;
				int ci = k + 1;
#pragma unroll
				for (int i = k + 1; i < 7; ++i)
				{
					const double v = A(i, j) - fac[i] * r;
					A(i, j) = v;
					const double a = fabs(v);
					if (a > cm) { cm = a; ci = i; }
				}
				if (cm > biggest) { biggest = cm; pr = ci; pc = j; }
			}
		}
		// dimensionOfKernel() == 2  <=>  all 7 pivots above eps * 7 * maxpivot (:161-163)
		const double premult = maxpivot * eps * 7.0;
		int rank = 0;
#pragma unroll
		for (int k = 0; k < 7; ++k)
			if (fabs(diag[k]) > premult) ++rank;
		if (rank != 7)
			return 0;
		double x[2][7];
#pragma unroll
		for (int c = 0; c < 2; ++c)
		{
#pragma unroll
			for (int i = 6; i >= 0; --i)
			{
				double s = -A(i, 7 + c);
#pragma unroll
				for (int j = i + 1; j < 7; ++j)
					s -= A(i, j) * x[c][j];
				x[c][i] = s / diag[i];
			}
		}
		// the matrix is dead from here on: its first two rows become f1 / f2 (un-permuted)
#pragma unroll
		for (int c = 0; c < 2; ++c)
		{
#pragma unroll
			for (int i = 0; i < 7; ++i)
				A.f(c, (int)((colperm >> (4 * i)) & 15ull)) = x[c][i];
			A.f(c, (int)((colperm >> (4 * (7 + c))) & 15ull)) = 1.0;
			A.f(c, (int)((colperm >> (4 * (7 + (1 - c)))) & 15ull)) = 0.0;
		}
#pragma unroll
		for (int j = 0; j < 9; ++j)
		{
			f2[j] = A.f(1, j);
			f1[j] = A.f(0, j) - f2[j]; // :190 f1 -= f2
		}

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

		return cubic_roots(c[0], c[1], c[2], c[3], roots);
	}

	// :267-288: F = lambda f1 + mu f2 scaled so that F33 = 1; false if |s| <= eps (the reference skips the root)
	__device__ __forceinline__ bool model(const double *f1, const double *f2, double root, double *F)
	{
		const double eps = 2.220446049250313e-16;
		double lambda = root, mu = 1.;
		const double s = f1[8] * root + f2[8];
		if (!(fabs(s) > eps)) // :273
			return false;
		mu = 1.0 / s;
		lambda *= mu;
#pragma unroll
		for (int j = 0; j < 8; ++j)
			F[j] = f1[j] * lambda + f2[j] * mu;
		F[8] = 1.0;
		return true;
	}

	// fundamental_estimator.h:737-800 (sample given as m rows of x0 y0 x1 y1)
	template <class PointFn>
	__device__ inline bool orientation_valid(const double *F, const PointFn &pt, int m)
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
			const double s1 = F[0] * pt(i, 2) + F[3] * pt(i, 3) + F[6];
			const double s2 = e[1] - e[2] * pt(i, 1);
			const double sig = s1 * s2;
			if (i == 0)
				sig2 = sig;
			else if (sig2 * sig < 0)
				return false;
		}
		return true;
	}
}
