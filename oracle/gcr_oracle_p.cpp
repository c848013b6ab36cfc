// oracle/gcr_oracle_p.cpp -- TEST INFRASTRUCTURE ONLY (see gcr_oracle.h).
// Absolute pose from three 2D-3D correspondences: the Lambda-Twist P3P of the reference
// (estimators/solver_p3p.h:177-385; helpers computeEig3x3known0 :108-142, refineLambda :145-175).
// [pinned:source]: the reference version is written on Eigen vectors / matrices (absent here); this restatement uses plain
// arrays, is checked in tests/test_oracle_solvers.py (exact reprojection of the three sample points, rotation
// orthonormality, ground-truth pose among the solutions) and agrees BIT FOR BIT with the reference's own source compiled
// verbatim over oracle/shim_solvers (tests/test_reference_solver_sources.py: 300 samples, count, order, coefficients).
#include "gcr_oracle.h"

#include <cmath>
#include <cstring>

namespace
{
	inline void cross(const double *a, const double *b, double *c)
	{
		c[0] = a[1] * b[2] - a[2] * b[1];
		c[1] = a[2] * b[0] - a[0] * b[2];
		c[2] = a[0] * b[1] - a[1] * b[0];
	}
	inline double dot(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }

	// inverse of the 3 x 3 matrix whose COLUMNS are c0, c1, c2; row-major result (cofactors / determinant)
	void inverse_from_columns(const double *c0, const double *c1, const double *c2, double *inv)
	{
		const double m[9] = {c0[0], c1[0], c2[0], c0[1], c1[1], c2[1], c0[2], c1[2], c2[2]};
		const double det = m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
		const double id = 1.0 / det;
		inv[0] = (m[4] * m[8] - m[5] * m[7]) * id; inv[1] = (m[2] * m[7] - m[1] * m[8]) * id; inv[2] = (m[1] * m[5] - m[2] * m[4]) * id;
		inv[3] = (m[5] * m[6] - m[3] * m[8]) * id; inv[4] = (m[0] * m[8] - m[2] * m[6]) * id; inv[5] = (m[2] * m[3] - m[0] * m[5]) * id;
		inv[6] = (m[3] * m[7] - m[4] * m[6]) * id; inv[7] = (m[1] * m[6] - m[0] * m[7]) * id; inv[8] = (m[0] * m[4] - m[1] * m[3]) * id;
	}

	// :108-142: two eigenvectors / eigenvalues of a symmetric 3 x 3 matrix with one zero eigenvalue
	void eig3x3_known0(const double M[3][3], double E[3][2], double &sig1, double &sig2)
	{
		const double p1 = -M[0][0] - M[1][1] - M[2][2];
		const double p0 = -M[0][1] * M[0][1] - M[0][2] * M[0][2] - M[1][2] * M[1][2] + M[0][0] * (M[1][1] + M[2][2]) + M[1][1] * M[2][2];
		const double disc = std::sqrt(p1 * p1 / 4.0 - p0);
		const double tmp = -p1 / 2.0;
		sig1 = tmp + disc;
		sig2 = tmp - disc;
		if (std::fabs(sig1) < std::fabs(sig2))
		{
			const double t = sig1; sig1 = sig2; sig2 = t;
		}
		const double sig[2] = {sig1, sig2};
		for (int k = 0; k < 2; ++k)
		{
			const double s = sig[k];
			const double c = s * s + M[0][0] * M[1][1] - s * (M[0][0] + M[1][1]) - M[0][1] * M[0][1];
			const double a1 = (s * M[0][2] + M[0][1] * M[1][2] - M[0][2] * M[1][1]) / c;
			const double a2 = (s * M[1][2] + M[0][1] * M[0][2] - M[0][0] * M[1][2]) / c;
			const double n = 1.0 / std::sqrt(1 + a1 * a1 + a2 * a2);
			E[0][k] = a1 * n; E[1][k] = a2 * n; E[2][k] = n;
		}
	}

	// :145-175 Gauss-Newton on the three distance equations
	void refine_lambda(double &l1, double &l2, double &l3, double a12, double a13, double a23, double b12, double b13, double b23)
	{
		for (int iter = 0; iter < 5; ++iter)
		{
			const double r1 = (l1 * l1 - 2.0 * l1 * l2 * b12 + l2 * l2 - a12);
			const double r2 = (l1 * l1 - 2.0 * l1 * l3 * b13 + l3 * l3 - a13);
			const double r3 = (l2 * l2 - 2.0 * l2 * l3 * b23 + l3 * l3 - a23);
			if (std::fabs(r1) + std::fabs(r2) + std::fabs(r3) < 1e-10)
				return;
			const double x11 = l1 - l2 * b12, x12 = l2 - l1 * b12, x21 = l1 - l3 * b13, x23 = l3 - l1 * b13, x32 = l2 - l3 * b23, x33 = l3 - l2 * b23;
			const double detJ = 0.5 / (x11 * x23 * x32 + x12 * x21 * x33);
			l1 += (-x23 * x32 * r1 - x12 * x33 * r2 + x12 * x23 * r3) * detJ;
			l2 += (-x21 * x33 * r1 + x11 * x33 * r2 - x11 * x23 * r3) * detJ;
			l3 += (x21 * x32 * r1 - x11 * x32 * r2 - x12 * x21 * r3) * detJ;
		}
	}
}

extern "C" {

// solver_p3p.h:177-385; pts5 rows = (u, v, X, Y, Z) with (u, v) K-normalised; returns #models, 3 x 4 row-major each
int gcro_solve_p3p(const double *pts5, const uint32_t *sample, double *P48)
{
	double y[3][3], X[3][3];
	for (int i = 0; i < 3; ++i)
	{
		const double *p = pts5 + 5 * (size_t)(sample ? sample[i] : (uint32_t)i);
		const double nrm = std::sqrt(p[0] * p[0] + p[1] * p[1] + 1.0);
		y[i][0] = p[0] / nrm; y[i][1] = p[1] / nrm; y[i][2] = 1.0 / nrm;
		X[i][0] = p[2]; X[i][1] = p[3]; X[i][2] = p[4];
	}
	double dX12[3], dX13[3], dX23[3];
	for (int k = 0; k < 3; ++k)
	{
		dX12[k] = X[0][k] - X[1][k];
		dX13[k] = X[0][k] - X[2][k];
		dX23[k] = X[1][k] - X[2][k];
	}
	const double a12 = dot(dX12, dX12), b12 = dot(y[0], y[1]);
	const double a13 = dot(dX13, dX13), b13 = dot(y[0], y[2]);
	const double a23 = dot(dX23, dX23), b23 = dot(y[1], y[2]);
	const double a23b12 = a23 * b12, a12b23 = a12 * b23, a23b13 = a23 * b13, a13b23 = a13 * b23;
	// :225-226 (both symmetric)
	const double D1[3][3] = {{a23, -a23b12, 0.0}, {-a23b12, a23 - a12, a12b23}, {0.0, a12b23, -a12}};
	const double D2[3][3] = {{a23, 0.0, -a23b13}, {0.0, -a13, a13b23}, {-a23b13, a13b23, a23 - a13}};
	// DXk: columns = cross products of Dk's columns (:228-230)
	double DX1[3][3], DX2[3][3];
	for (int which = 0; which < 2; ++which)
	{
		const double(*D)[3] = which == 0 ? D1 : D2;
		double(*DX)[3] = which == 0 ? DX1 : DX2;
		double col[3][3];
		for (int c = 0; c < 3; ++c)
			for (int r = 0; r < 3; ++r)
				col[c][r] = D[r][c];
		double x0[3], x1[3], x2[3];
		cross(col[1], col[2], x0);
		cross(col[2], col[0], x1);
		cross(col[0], col[1], x2);
		for (int r = 0; r < 3; ++r)
		{
			DX[r][0] = x0[r]; DX[r][1] = x1[r]; DX[r][2] = x2[r];
		}
	}
	// :234-237 p(gamma) = det(D1 + gamma D2)
	double c3 = D2[0][0] * DX2[0][0] + D2[1][0] * DX2[1][0] + D2[2][0] * DX2[2][0];
	double c2 = 0.0, c1 = 0.0;
	for (int c = 0; c < 3; ++c) // column-major traversal like Eigen's array sum
		for (int r = 0; r < 3; ++r)
		{
			c2 += D1[r][c] * DX2[r][c];
			c1 += D2[r][c] * DX1[r][c];
		}
	double c0 = D1[0][0] * DX1[0][0] + D1[1][0] * DX1[1][0] + D1[2][0] * DX1[2][0];
	const double c3inv = 1.0 / c3;
	c2 *= c3inv; c1 *= c3inv; c0 *= c3inv;
	// :245-258 one real root of the monic cubic
	double a = c1 - c2 * c2 / 3.0;
	double b = (2.0 * c2 * c2 * c2 - 9.0 * c2 * c1) / 27.0 + c0;
	double c = b * b / 4.0 + a * a * a / 27.0;
	double gamma;
	if (c > 0)
	{
		c = std::sqrt(c);
		b *= -0.5;
		gamma = std::cbrt(b + c) + std::cbrt(b - c) - c2 / 3.0;
	}
	else
	{
		c = 3.0 * b / (2.0 * a) * std::sqrt(-3.0 / a);
		gamma = 2.0 * std::sqrt(-a / 3.0) * std::cos(std::acos(c) / 3.0) - c2 / 3.0;
	}
	// :261-263 one Newton step
	const double f = gamma * gamma * gamma + c2 * gamma * gamma + c1 * gamma + c0;
	const double df = 3.0 * gamma * gamma + 2.0 * c2 * gamma + c1;
	gamma = gamma - f / df;

	double D0[3][3];
	for (int r = 0; r < 3; ++r)
		for (int cc = 0; cc < 3; ++cc)
			D0[r][cc] = D1[r][cc] + gamma * D2[r][cc];
	double E[3][2], sig1, sig2;
	eig3x3_known0(D0, E, sig1, sig2);
	double s = std::sqrt(-sig2 / sig1);

	double XX[9], xc[3];
	cross(dX12, dX13, xc);
	inverse_from_columns(dX12, dX13, xc, XX);

	const double TOL_DOUBLE_ROOT = 1e-12;
	int n_models = 0;
	for (int s_flip = 0; s_flip < 2; ++s_flip, s = -s)
	{
		const double u1 = E[0][0] - s * E[0][1], u2 = E[1][0] - s * E[1][1], u3 = E[2][0] - s * E[2][1];
		const bool switch_12 = std::fabs(u1) < std::fabs(u2);
		double qa, qb, qc, w0, w1;
		if (switch_12)
		{
			w0 = -u1 / u2;
			w1 = -u3 / u2;
			qa = -a13 * w1 * w1 + 2 * a13b23 * w1 - a13 + a23;
			qb = 2 * a13b23 * w0 - 2 * a23b13 - 2 * a13 * w0 * w1;
			qc = -a13 * w0 * w0 + a23;
		}
		else
		{
			w0 = -u2 / u1;
			w1 = -u3 / u1;
			qa = (a13 - a12) * w1 * w1 + 2.0 * a12 * b13 * w1 - a12;
			qb = -2.0 * a13 * b12 * w1 + 2.0 * a12 * b13 * w0 - 2.0 * w0 * w1 * (a12 - a13);
			qc = (a13 - a12) * w0 * w0 - 2.0 * a13 * b12 * w0 + a13;
		}
		const double b2m4ac = qb * qb - 4.0 * qa * qc;
		if (b2m4ac < -TOL_DOUBLE_ROOT)
			continue;
		const double sq = std::sqrt(b2m4ac > 0.0 ? b2m4ac : 0.0);
		double tau = (qb > 0) ? (2.0 * qc) / (-qb - sq) : (2.0 * qc) / (-qb + sq);
		for (int tau_flip = 0; tau_flip < 2; ++tau_flip, tau = qc / (qa * tau))
		{
			if (tau > 0)
			{
				double l1, l2, l3;
				bool ok = true;
				if (switch_12)
				{
					l1 = std::sqrt(a13 / (tau * (tau - 2.0 * b13) + 1.0));
					l3 = tau * l1;
					l2 = w0 * l1 + w1 * l3;
					if (l2 < 0) ok = false;
				}
				else
				{
					l2 = std::sqrt(a23 / (tau * (tau - 2.0 * b23) + 1.0));
					l3 = tau * l2;
					l1 = w0 * l2 + w1 * l3;
					if (l1 < 0) ok = false;
				}
				if (!ok)
					continue; // (the reference's `continue` also skips the double-root break below)
				refine_lambda(l1, l2, l3, a12, a13, a23, b12, b13, b23);
				double v1[3], v2[3], v3[3];
				for (int k = 0; k < 3; ++k)
				{
					v1[k] = l1 * y[0][k] - l2 * y[1][k];
					v2[k] = l1 * y[0][k] - l3 * y[2][k];
				}
				cross(v1, v2, v3);
				// R = YY * XX with YY = [v1 v2 v3] as columns
				double R[9];
				for (int r = 0; r < 3; ++r)
					for (int cc = 0; cc < 3; ++cc)
						R[3 * r + cc] = v1[r] * XX[cc] + v2[r] * XX[3 + cc] + v3[r] * XX[6 + cc];
				double *P = P48 + 12 * n_models;
				for (int r = 0; r < 3; ++r)
				{
					P[4 * r + 0] = R[3 * r + 0]; P[4 * r + 1] = R[3 * r + 1]; P[4 * r + 2] = R[3 * r + 2];
					P[4 * r + 3] = l1 * y[0][r] - (R[3 * r] * X[0][0] + R[3 * r + 1] * X[0][1] + R[3 * r + 2] * X[0][2]);
				}
				++n_models;
			}
			if (b2m4ac < TOL_DOUBLE_ROOT)
				break;
		}
	}
	return n_models;
}

} // extern "C"
