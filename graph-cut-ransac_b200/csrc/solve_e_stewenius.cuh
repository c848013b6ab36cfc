// solve_e_stewenius.cuh -- device back end of the non-minimal essential-matrix refit (one thread per index set, FP64).
// Follows EssentialMatrixFivePointSteweniusSolver::estimateModel
// (estimators/solver_essential_matrix_five_point_stewenius.h:189-228) from the 4-D "null space" on:
//   10 x 20 trace / determinant constraints in GrevLex order (:361-390; built with the 5-point kernel's product tables
//   and permuted) -> FullPivLU solve of the left 10 x 10 block into the right one (:198-199) -> 10 x 10 action matrix
//   (:201-209) -> its real eigenvalues (the reference calls Eigen::EigenSolver; here Householder reduction to
//   Hessenberg form + Francis double-shift QR, the EISPACK orthes / hqr pair Eigen's RealSchur derives from) ->
//   eigenvector of each real eigenvalue (inverse iteration on A - lambda I) -> E = nullSpace * v.tail<4>() (:216-228).
// The degree-10 polynomial route of the minimal solver (solve_e5.cuh) finds the same solutions but loses ~1e-5 on
// badly conditioned sets; the action matrix keeps the refit within the 1e-5 model tolerance of the reference's own
// algorithm (tests/test_gpu_refit_grid.py).
#pragma once
#include "solve_e5.cuh"

namespace gcr_e_nm
{
	// column c of the GrevLex order (x^3 x^2y xy^2 y^3 x^2z xyz y^2z xz^2 yz^2 z^3 | x^2 xy y^2 xz yz z^2 x y z 1) in the
	// 5-point kernel's order (x^3 y^3 x^2y xy^2 x^2z x^2 y^2z y^2 xyz xy | xz^2 xz x yz^2 yz y z^3 z^2 z 1)
	__device__ __constant__ const unsigned char GREVLEX_TO_E5[20] = {0, 2, 3, 1, 4, 8, 6, 10, 13, 16, 5, 9, 7, 11, 14, 17, 12, 15, 18, 19};

	__device__ __forceinline__ double sign_of(double x, double y) { return y >= 0 ? fabs(x) : -fabs(x); }

	// eigenvalues of the 10 x 10 row-major matrix a (destroyed); false if the QR iteration does not converge
	__device__ inline bool eigenvalues10(double *a, double *wr, double *wi)
	{
		const int n = 10;
#define A_(i, j) a[(i) * n + (j)]
		for (int m = 1; m < n - 1; ++m)
		{
			double scale = 0.0;
			for (int i = m; i < n; ++i) scale += fabs(A_(i, m - 1));
			if (scale == 0.0) continue;
			double ort[10];
			double h = 0.0;
			for (int i = n - 1; i >= m; --i)
			{
				ort[i] = A_(i, m - 1) / scale;
				h += ort[i] * ort[i];
			}
			double g = sqrt(h);
			if (ort[m] > 0) g = -g;
			h -= ort[m] * g;
			ort[m] -= g;
			for (int j = m; j < n; ++j)
			{
				double f = 0.0;
				for (int i = n - 1; i >= m; --i) f += ort[i] * A_(i, j);
				f /= h;
				for (int i = m; i < n; ++i) A_(i, j) -= f * ort[i];
			}
			for (int i = 0; i < n; ++i)
			{
				double f = 0.0;
				for (int j = n - 1; j >= m; --j) f += ort[j] * A_(i, j);
				f /= h;
				for (int j = m; j < n; ++j) A_(i, j) -= f * ort[j];
			}
			A_(m, m - 1) = scale * g;
			for (int i = m + 1; i < n; ++i) A_(i, m - 1) = 0.0;
		}
		double anorm = 0.0;
		for (int i = 0; i < n; ++i)
			for (int j = (i > 0 ? i - 1 : 0); j < n; ++j)
				anorm += fabs(A_(i, j));
		int nn = n - 1;
		double t = 0.0;
		while (nn >= 0)
		{
			int its = 0, l;
			do
			{
				for (l = nn; l >= 1; --l)
				{
					double s = fabs(A_(l - 1, l - 1)) + fabs(A_(l, l));
					if (s == 0.0) s = anorm;
					if (fabs(A_(l, l - 1)) + s == s)
					{
						A_(l, l - 1) = 0.0;
						break;
					}
				}
				double x = A_(nn, nn);
				if (l == nn)
				{
					wr[nn] = x + t;
					wi[nn--] = 0.0;
				}
				else
				{
					double y = A_(nn - 1, nn - 1);
					double w = A_(nn, nn - 1) * A_(nn - 1, nn);
					if (l == nn - 1)
					{
						double p = 0.5 * (y - x);
						double q = p * p + w;
						double z = sqrt(fabs(q));
						x += t;
						if (q >= 0.0)
						{
							z = p + sign_of(z, p);
							wr[nn - 1] = wr[nn] = x + z;
							if (z != 0.0) wr[nn] = x - w / z;
							wi[nn - 1] = wi[nn] = 0.0;
						}
						else
						{
							wr[nn - 1] = wr[nn] = x + p;
							wi[nn - 1] = -(wi[nn] = z);
						}
						nn -= 2;
					}
					else
					{
						if (its == 60) return false;
						if (its == 10 || its == 20)
						{
							t += x;
							for (int i = 0; i <= nn; ++i) A_(i, i) -= x;
							const double s = fabs(A_(nn, nn - 1)) + fabs(A_(nn - 1, nn - 2));
							y = x = 0.75 * s;
							w = -0.4375 * s * s;
						}
						++its;
						int m;
						double p = 0, q = 0, r = 0, z;
						for (m = nn - 2; m >= l; --m)
						{
							z = A_(m, m);
							r = x - z;
							double s = y - z;
							p = (r * s - w) / A_(m + 1, m) + A_(m, m + 1);
							q = A_(m + 1, m + 1) - z - r - s;
							r = A_(m + 2, m + 1);
							s = fabs(p) + fabs(q) + fabs(r);
							p /= s; q /= s; r /= s;
							if (m == l) break;
							const double u = fabs(A_(m, m - 1)) * (fabs(q) + fabs(r));
							const double v = fabs(p) * (fabs(A_(m - 1, m - 1)) + fabs(z) + fabs(A_(m + 1, m + 1)));
							if (u + v == v) break;
						}
						for (int i = m + 2; i <= nn; ++i)
						{
							A_(i, i - 2) = 0.0;
							if (i != m + 2) A_(i, i - 3) = 0.0;
						}
						for (int k = m; k <= nn - 1; ++k)
						{
							if (k != m)
							{
								p = A_(k, k - 1);
								q = A_(k + 1, k - 1);
								r = 0.0;
								if (k != nn - 1) r = A_(k + 2, k - 1);
								if ((x = fabs(p) + fabs(q) + fabs(r)) != 0.0)
								{
									p /= x; q /= x; r /= x;
								}
							}
							const double s = sign_of(sqrt(p * p + q * q + r * r), p);
							if (s != 0.0)
							{
								if (k == m)
								{
									if (l != m) A_(k, k - 1) = -A_(k, k - 1);
								}
								else
									A_(k, k - 1) = -s * x;
								p += s;
								x = p / s;
								y = q / s;
								z = r / s;
								q /= p;
								r /= p;
								for (int j = k; j <= nn; ++j)
								{
									p = A_(k, j) + q * A_(k + 1, j);
									if (k != nn - 1)
									{
										p += r * A_(k + 2, j);
										A_(k + 2, j) -= p * z;
									}
									A_(k + 1, j) -= p * y;
									A_(k, j) -= p * x;
								}
								const int mmin = nn < k + 3 ? nn : k + 3;
								for (int i = l; i <= mmin; ++i)
								{
									p = x * A_(i, k) + y * A_(i, k + 1);
									if (k != nn - 1)
									{
										p += z * A_(i, k + 2);
										A_(i, k + 2) -= p * r;
									}
									A_(i, k + 1) -= p * q;
									A_(i, k) -= p;
								}
							}
						}
					}
				}
			} while (l < nn - 1);
		}
#undef A_
		return true;
	}

	// unit eigenvector of the real eigenvalue lambda of the 10 x 10 matrix Am: inverse iteration on Am - lambda I
	// (partial-pivot LU, pivots below eps |A| replaced; the first pass solves U x = 1)
	__device__ inline void eigenvector10(const double *Am, double lambda, double *v)
	{
		const int n = 10;
		double M[100];
		double anorm = 0.0;
		for (int i = 0; i < 100; ++i)
		{
			M[i] = Am[i];
			anorm = fmax(anorm, fabs(Am[i]));
		}
		for (int i = 0; i < n; ++i) M[i * n + i] -= lambda;
		const double tiny = 2.220446049250313e-16 * (anorm > 0 ? anorm : 1.0);
		int piv[10];
		for (int k = 0; k < n; ++k)
		{
			int p = k;
			for (int i = k + 1; i < n; ++i)
				if (fabs(M[i * n + k]) > fabs(M[p * n + k])) p = i;
			piv[k] = p;
			if (p != k)
				for (int j = 0; j < n; ++j) { const double t = M[k * n + j]; M[k * n + j] = M[p * n + j]; M[p * n + j] = t; }
			if (fabs(M[k * n + k]) < tiny) M[k * n + k] = tiny;
			for (int i = k + 1; i < n; ++i)
			{
				const double f = M[i * n + k] / M[k * n + k];
				M[i * n + k] = f;
				for (int j = k + 1; j < n; ++j) M[i * n + j] -= f * M[k * n + j];
			}
		}
		for (int i = 0; i < n; ++i) v[i] = 1.0;
		for (int iter = 0; iter < 3; ++iter)
		{
			if (iter > 0)
				for (int k = 0; k < n; ++k)
				{
					const double t = v[k]; v[k] = v[piv[k]]; v[piv[k]] = t;
					for (int i = k + 1; i < n; ++i) v[i] -= M[i * n + k] * v[k];
				}
			for (int i = n - 1; i >= 0; --i)
			{
				double s = v[i];
				for (int j = i + 1; j < n; ++j) s -= M[i * n + j] * v[j];
				v[i] = s / M[i * n + i];
			}
			double nrm = 0.0;
			for (int i = 0; i < n; ++i) nrm += v[i] * v[i];
			nrm = sqrt(nrm);
			for (int i = 0; i < n; ++i) v[i] /= nrm;
		}
	}

	// Part 1 (one thread).  N[b][j]: coefficient j (column-major E.data() index) of basis vector b -> the 10 x 10 action
	// matrix Am and its eigenvalues (wr, wi).  Returns the number of eigenvalues to look at (10), 0 when the system is
	// singular / not finite / the eigenvalue iteration does not converge.
	__device__ inline int action_matrix(const double (&N)[4][9], double *Am, double *wr, double *wi)
	{
		// ---- constraints in the 5-point kernel's monomial order, rows 0..8 = (E E^T - 1/2 tr) E, row 9 = det ----
		double C5[10][20];
		for (int i = 0; i < 10; ++i)
			for (int j = 0; j < 20; ++j)
				C5[i][j] = 0.0;
		{
			double Ee[3][3][4];
			for (int i = 0; i < 3; ++i)
				for (int j = 0; j < 3; ++j)
					for (int b = 0; b < 4; ++b)
						Ee[i][j][b] = N[b][3 * j + i];
			double EET[3][3][10];
			for (int i = 0; i < 3; ++i)
				for (int j = 0; j < 3; ++j)
				{
					for (int m = 0; m < 10; ++m) EET[i][j][m] = 0.0;
					for (int k = 0; k < 3; ++k)
						gcr_e5::mul11(Ee[i][k], Ee[j][k], EET[i][j], 1.0);
				}
			for (int m = 0; m < 10; ++m)
			{
				const double t = 0.5 * (EET[0][0][m] + EET[1][1][m] + EET[2][2][m]);
				EET[0][0][m] -= t; EET[1][1][m] -= t; EET[2][2][m] -= t;
			}
			for (int i = 0; i < 3; ++i)
				for (int j = 0; j < 3; ++j)
					for (int k = 0; k < 3; ++k)
						gcr_e5::mul21(EET[i][k], Ee[k][j], C5[3 * i + j]);
			double d[10];
			for (int m = 0; m < 10; ++m) d[m] = 0.0;
			gcr_e5::mul11(Ee[0][1], Ee[1][2], d, 1.0); gcr_e5::mul11(Ee[0][2], Ee[1][1], d, -1.0); gcr_e5::mul21(d, Ee[2][0], C5[9]);
			for (int m = 0; m < 10; ++m) d[m] = 0.0;
			gcr_e5::mul11(Ee[0][2], Ee[1][0], d, 1.0); gcr_e5::mul11(Ee[0][0], Ee[1][2], d, -1.0); gcr_e5::mul21(d, Ee[2][1], C5[9]);
			for (int m = 0; m < 10; ++m) d[m] = 0.0;
			gcr_e5::mul11(Ee[0][0], Ee[1][1], d, 1.0); gcr_e5::mul11(Ee[0][1], Ee[1][0], d, -1.0); gcr_e5::mul21(d, Ee[2][2], C5[9]);
		}
		// ---- GrevLex columns; the reference's trace rows carry a factor 2 (eet = 2 E E^T, :368) ----
		double A[100], B[100];
		for (int i = 0; i < 10; ++i)
		{
			const double f = i < 9 ? 2.0 : 1.0;
			for (int j = 0; j < 10; ++j)
			{
				A[10 * i + j] = f * C5[i][GREVLEX_TO_E5[j]];
				B[10 * i + j] = f * C5[i][GREVLEX_TO_E5[10 + j]];
			}
		}
		// ---- FullPivLU solve A X = B (:198-199) ----
		int cp[10];
		for (int j = 0; j < 10; ++j) cp[j] = j;
		for (int k = 0; k < 10; ++k)
		{
			int pr = k, pc = k;
			double big = 0.0;
			for (int j = k; j < 10; ++j)
				for (int i = k; i < 10; ++i)
					if (fabs(A[10 * i + j]) > big) { big = fabs(A[10 * i + j]); pr = i; pc = j; }
			if (big == 0.0)
				return 0;
			if (pr != k)
				for (int j = 0; j < 10; ++j)
				{
					double t = A[10 * k + j]; A[10 * k + j] = A[10 * pr + j]; A[10 * pr + j] = t;
					t = B[10 * k + j]; B[10 * k + j] = B[10 * pr + j]; B[10 * pr + j] = t;
				}
			if (pc != k)
			{
				for (int i = 0; i < 10; ++i) { const double t = A[10 * i + k]; A[10 * i + k] = A[10 * i + pc]; A[10 * i + pc] = t; }
				const int t = cp[k]; cp[k] = cp[pc]; cp[pc] = t;
			}
			const double piv = A[10 * k + k];
			for (int i = k + 1; i < 10; ++i)
			{
				const double f = A[10 * i + k] / piv;
				if (f == 0.0) continue;
				for (int j = k + 1; j < 10; ++j) A[10 * i + j] -= f * A[10 * k + j];
				for (int j = 0; j < 10; ++j) B[10 * i + j] -= f * B[10 * k + j];
			}
		}
		for (int c = 0; c < 10; ++c)
			for (int i = 9; i >= 0; --i)
			{
				double s = B[10 * i + c];
				for (int j = i + 1; j < 10; ++j) s -= A[10 * i + j] * B[10 * j + c];
				B[10 * i + c] = s / A[10 * i + i];
			}
		// B now holds Y with X[cp[j]] = Y[j]
		// ---- action matrix (:201-209): rows 0,1,2,4,5,7 of X, then the four -1 entries ----
		for (int i = 0; i < 100; ++i) Am[i] = 0.0;
		const int src_row[6] = {0, 1, 2, 4, 5, 7};
		for (int j = 0; j < 10; ++j)
			for (int r = 0; r < 6; ++r)
				if (cp[j] == src_row[r])
					for (int c = 0; c < 10; ++c) Am[10 * r + c] = B[10 * j + c];
		Am[10 * 6 + 0] = -1.0;
		Am[10 * 7 + 1] = -1.0;
		Am[10 * 8 + 3] = -1.0;
		Am[10 * 9 + 6] = -1.0;
		for (int i = 0; i < 100; ++i)
			if (!isfinite(Am[i]))
				return 0;
		for (int i = 0; i < 100; ++i) A[i] = Am[i];
		if (!eigenvalues10(A, wr, wi))
			return 0;
		return 10;
	}

	// Part 2 (any thread, one REAL eigenvalue): its eigenvector -> the model E (row-major, un-normalised).  The eigenvalues
	// are independent of each other: k_refit_e gives each one a lane of its own.
	__device__ inline void model_of_eigenvalue(const double (&N)[4][9], const double *Am, double lambda, double *E)
	{
		double v[10];
		eigenvector10(Am, lambda, v);
		for (int k = 0; k < 9; ++k)
			E[3 * (k % 3) + k / 3] = N[0][k] * v[6] + N[1][k] * v[7] + N[2][k] * v[8] + N[3][k] * v[9];
	}

	// the two parts in sequence: E_out: up to 10 x 9 row-major, un-normalised, in the order of the eigenvalues
	__device__ inline int solve_from_basis(const double (&N)[4][9], double *E_out)
	{
		double Am[100], wr[10], wi[10];
		if (!action_matrix(N, Am, wr, wi))
			return 0;
		int n_models = 0;
		for (int e = 0; e < 10; ++e)
		{
			if (wi[e] != 0.0)
				continue;
			model_of_eigenvalue(N, Am, wr[e], E_out + 9 * n_models);
			++n_models;
		}
		return n_models;
	}
}
