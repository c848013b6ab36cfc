// solve_e5.cuh -- device 5-point essential-matrix solver (one thread per sample, FP64).
// Follows EssentialMatrixFivePointNisterSolver::estimateModel (estimators/solver_essential_matrix_five_point_nister.h
// :263-615), the reference's default minimal E solver (types.h:108-110), minimal branch:
//   5 x 9 epipolar matrix (:282-319) -> 4-D null space with Eigen FullPivLU::kernel() semantics (:324-330: complete
//   pivoting, free variables set to unit vectors -- NOT an orthonormal basis, exactly like the reference) ->
//   10 x 20 trace / determinant constraints (:204-261) -> partial-pivot LU solve of the left 10 x 10 block into the
//   right one (:350) -> three polynomial rows in z (:353-366) -> degree-10 determinant (:410-570; the reference
//   writes the expansion out, here it is three nested polynomial products) -> real roots by Sturm isolation + safeguarded
//   Newton (stands in for poselib::sturm::bisect_sturm<10>, maths/sturm_polynomial_solver.h:223-264; ascending roots)
//   -> back substitution per root (:583-612).
// Arithmetic order matches oracle/gcr_oracle_e.cpp statement by statement (built with -fmad=false), so the CPU oracle
// and this kernel agree to the last bit even on the ill-conditioned samples the non-orthonormal basis produces.
#pragma once

namespace gcr_e5
{
	__device__ __constant__ const unsigned char IDX11[4][4] = {{0, 1, 2, 3}, {1, 4, 5, 6}, {2, 5, 7, 8}, {3, 6, 8, 9}};
	__device__ __constant__ const unsigned char IDX21[10][4] = {{0, 2, 4, 5}, {2, 3, 8, 9}, {4, 8, 10, 11}, {5, 9, 11, 12}, {3, 1, 6, 7},
	                                                           {8, 6, 13, 14}, {9, 7, 14, 15}, {10, 13, 16, 17}, {11, 14, 17, 18}, {12, 15, 18, 19}};

	// c (deg 2, 10 monomials) += sign * a (deg 1) * b (deg 1)
	__device__ __forceinline__ void mul11(const double *a, const double *b, double *c, double sign)
	{
		for (int i = 0; i < 4; ++i)
			for (int j = 0; j < 4; ++j)
				c[IDX11[i][j]] += sign * a[i] * b[j];
	}
	// c (deg 3, 20 monomials) += a (deg 2) * b (deg 1)
	__device__ __forceinline__ void mul21(const double *a, const double *b, double *c)
	{
		for (int i = 0; i < 10; ++i)
			for (int j = 0; j < 4; ++j)
				c[IDX21[i][j]] += a[i] * b[j];
	}
	__device__ __forceinline__ void conv(const double *a, int da, const double *b, int db, double *c)
	{
		for (int i = 0; i <= da + db; ++i) c[i] = 0.0;
		for (int i = 0; i <= da; ++i)
			for (int j = 0; j <= db; ++j)
				c[i + j] += a[i] * b[j];
	}
	__device__ __forceinline__ double horner(const double *c, int deg, double x)
	{
		double v = c[deg];
		for (int i = deg - 1; i >= 0; --i)
			v = v * x + c[i];
		return v;
	}

	// ---- Sturm chain root finder (same algorithm as oracle/sturm.h) ----
	struct Chain
	{
		double p[11][11];
		int deg[11];
		int len;
	};

	__device__ inline void chain_build(const double *f, int n, Chain &ch)
	{
		for (int i = 0; i <= n; ++i)
			ch.p[0][i] = f[i];
		ch.deg[0] = n;
		for (int i = 0; i < n; ++i)
			ch.p[1][i] = f[i + 1] * (double)(i + 1);
		ch.deg[1] = n - 1;
		ch.len = 2;
		while (ch.deg[ch.len - 1] > 0 && ch.len <= 10)
		{
			const double *a = ch.p[ch.len - 2];
			const double *b = ch.p[ch.len - 1];
			const int da = ch.deg[ch.len - 2], db = ch.deg[ch.len - 1];
			double r[11];
			for (int i = 0; i <= da; ++i)
				r[i] = a[i];
			for (int k = da; k >= db; --k)
			{
				const double q = r[k] / b[db];
				for (int i = 0; i <= db; ++i)
					r[k - db + i] -= q * b[i];
				r[k] = 0.0;
			}
			int dr = db - 1;
			while (dr >= 0 && r[dr] == 0.0)
				--dr;
			if (dr < 0)
				break;
			double mx = 0.0;
			for (int i = 0; i <= dr; ++i)
				mx = fmax(mx, fabs(r[i]));
			const double s = -1.0 / mx;
			for (int i = 0; i <= dr; ++i)
				ch.p[ch.len][i] = r[i] * s;
			ch.deg[ch.len] = dr;
			++ch.len;
		}
	}

	__device__ inline int sign_changes(const Chain &ch, double x)
	{
		int changes = 0, last = 0;
		for (int k = 0; k < ch.len; ++k)
		{
			const double v = horner(ch.p[k], ch.deg[k], x);
			const int s = v > 0.0 ? 1 : (v < 0.0 ? -1 : 0);
			if (s != 0)
			{
				if (last != 0 && s != last)
					++changes;
				last = s;
			}
		}
		return changes;
	}

	__device__ inline double polish(const double *f, const double *df, int n, double lo, double hi)
	{
		const double fl = horner(f, n, lo), fh = horner(f, n, hi);
		if (fl == 0.0) return lo;
		if (fh == 0.0) return hi;
		double xl, xh;
		if (fl < 0.0) { xl = lo; xh = hi; }
		else { xl = hi; xh = lo; }
		double rts = 0.5 * (lo + hi);
		double dxold = fabs(hi - lo), dx = dxold;
		double v = horner(f, n, rts), d = horner(df, n - 1, rts);
		for (int it = 0; it < 200; ++it)
		{
			if ((((rts - xh) * d - v) * ((rts - xl) * d - v) > 0.0) || (fabs(2.0 * v) > fabs(dxold * d)))
			{
				dxold = dx;
				dx = 0.5 * (xh - xl);
				rts = xl + dx;
				if (xl == rts) return rts;
			}
			else
			{
				dxold = dx;
				dx = v / d;
				const double t = rts;
				rts -= dx;
				if (t == rts) return rts;
			}
			if (fabs(dx) <= 4.440892098500626e-16 * fabs(rts)) return rts;
			v = horner(f, n, rts);
			d = horner(df, n - 1, rts);
			if (v < 0.0) xl = rts;
			else xh = rts;
		}
		return rts;
	}

	struct Iv
	{
		double a, b;
		int sa, sb, depth;
	};

	// real roots of c[0] + ... + c[10] x^10 (c[10] != 0), ascending; returns their number
	__device__ inline int real_roots10(const double *coeffs, double *roots)
	{
		const int n = 10;
		if (coeffs[n] == 0.0)
			return 0;
		double f[11], df[11];
		const double inv = 1.0 / coeffs[n];
		double mx = 0.0;
		for (int i = 0; i < n; ++i)
		{
			f[i] = coeffs[i] * inv;
			mx = fmax(mx, fabs(f[i]));
		}
		f[n] = 1.0;
		for (int i = 0; i < n; ++i)
			df[i] = f[i + 1] * (double)(i + 1);
		const double R = 1.0 + mx;
		Chain ch;
		chain_build(f, n, ch);
		Iv stack[72]; // depth <= 60, one pending sibling per level + the pair pushed last
		int top = 0, n_roots = 0;
		const int s_lo = sign_changes(ch, -R), s_hi = sign_changes(ch, R);
		if (s_lo - s_hi <= 0)
			return 0;
		stack[top++] = {-R, R, s_lo, s_hi, 0};
		while (top > 0 && n_roots < n)
		{
			const Iv iv = stack[--top];
			const int cnt = iv.sa - iv.sb;
			if (cnt <= 0)
				continue;
			if (cnt == 1)
			{
				const double fa = horner(f, n, iv.a), fb = horner(f, n, iv.b);
				if ((fa < 0.0) != (fb < 0.0) || fa == 0.0 || fb == 0.0)
				{
					roots[n_roots++] = polish(f, df, n, iv.a, iv.b);
					continue;
				}
			}
			if (iv.depth >= 60 || !(iv.b - iv.a > 1e-15 * (fabs(iv.a) + fabs(iv.b))))
				continue;
			const double c = 0.5 * (iv.a + iv.b);
			const int sc = sign_changes(ch, c);
			stack[top++] = {c, iv.b, sc, iv.sb, iv.depth + 1};
			stack[top++] = {iv.a, c, iv.sa, sc, iv.depth + 1};
		}
		return n_roots;
	}

	// 4-D null space of the 5 x 9 epipolar matrix `lu` (destroyed) with Eigen FullPivLU::kernel() semantics; N[b][j] =
	// coefficient j of basis vector b.  false if the rank is not 5 (dimensionOfKernel() != 4, :326-328).
	__device__ inline bool null_space_lu(double (&lu)[5][9], double (&N)[4][9])
	{
		const double eps = 2.220446049250313e-16;
		int colperm[9];
		for (int j = 0; j < 9; ++j)
			colperm[j] = j;
		double maxpivot = 0.0;
		int nonzero = 5;
		for (int k = 0; k < 5; ++k)
		{
			int pr = k, pc = k;
			double biggest = 0.0;
			for (int j = k; j < 9; ++j)
				for (int i = k; i < 5; ++i)
				{
					const double a = fabs(lu[i][j]);
					if (a > biggest) { biggest = a; pr = i; pc = j; }
				}
			if (biggest == 0.0) { nonzero = k; break; }
			if (biggest > maxpivot) maxpivot = biggest;
			if (pr != k)
				for (int j = 0; j < 9; ++j) { const double t = lu[k][j]; lu[k][j] = lu[pr][j]; lu[pr][j] = t; }
			if (pc != k)
			{
				for (int i = 0; i < 5; ++i) { const double t = lu[i][k]; lu[i][k] = lu[i][pc]; lu[i][pc] = t; }
				const int t = colperm[k]; colperm[k] = colperm[pc]; colperm[pc] = t;
			}
			const double piv = lu[k][k];
			for (int i = k + 1; i < 5; ++i)
			{
				const double f = lu[i][k] / piv;
				lu[i][k] = f;
				for (int j = k + 1; j < 9; ++j)
					lu[i][j] -= f * lu[k][j];
			}
		}
		const double premult = maxpivot * eps * 5.0;
		int rank = 0;
		for (int k = 0; k < nonzero; ++k)
			if (fabs(lu[k][k]) > premult) ++rank;
		if (rank != 5)
			return false;
		// kernel basis: N[b][j], b = 0..3
		for (int c = 0; c < 4; ++c)
		{
			double x[5];
			for (int i = 4; i >= 0; --i)
			{
				double s = -lu[i][5 + c];
				for (int j = i + 1; j < 5; ++j)
					s -= lu[i][j] * x[j];
				x[i] = s / lu[i][i];
			}
			for (int i = 0; i < 5; ++i)
				N[c][colperm[i]] = x[i];
			for (int q = 0; q < 4; ++q)
				N[c][colperm[5 + q]] = q == c ? 1.0 : 0.0;
		}
		return true;
	}

	// Everything after the null space (:204-261, :350-612): E = x N0 + y N1 + z N2 + N3 -> up to 10 models (row-major).
	// Also the back end of the non-minimal E refit (k_refit.cu), whose basis comes from a pivoted QR of A^T A.
	__device__ inline int solve_from_basis(const double (&N)[4][9], double *E_out)
	{
		// ---- constraints ----
		double C[10][20];
		for (int i = 0; i < 10; ++i)
			for (int j = 0; j < 20; ++j)
				C[i][j] = 0.0;
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
						mul11(Ee[i][k], Ee[j][k], EET[i][j], 1.0);
				}
			for (int m = 0; m < 10; ++m)
			{
				const double t = 0.5 * (EET[0][0][m] + EET[1][1][m] + EET[2][2][m]);
				EET[0][0][m] -= t; EET[1][1][m] -= t; EET[2][2][m] -= t;
			}
			for (int i = 0; i < 3; ++i)
				for (int j = 0; j < 3; ++j)
					for (int k = 0; k < 3; ++k)
						mul21(EET[i][k], Ee[k][j], C[3 * i + j]);
			double d[10];
			for (int m = 0; m < 10; ++m) d[m] = 0.0;
			mul11(Ee[0][1], Ee[1][2], d, 1.0); mul11(Ee[0][2], Ee[1][1], d, -1.0); mul21(d, Ee[2][0], C[9]);
			for (int m = 0; m < 10; ++m) d[m] = 0.0;
			mul11(Ee[0][2], Ee[1][0], d, 1.0); mul11(Ee[0][0], Ee[1][2], d, -1.0); mul21(d, Ee[2][1], C[9]);
			for (int m = 0; m < 10; ++m) d[m] = 0.0;
			mul11(Ee[0][0], Ee[1][1], d, 1.0); mul11(Ee[0][1], Ee[1][0], d, -1.0); mul21(d, Ee[2][2], C[9]);
		}
		// ---- partial-pivot LU: C[:, 10:20] <- C[:, 0:10]^-1 C[:, 10:20] (in place; rows get swapped together) ----
		for (int k = 0; k < 10; ++k)
		{
			int p = k;
			double big = fabs(C[k][k]);
			for (int i = k + 1; i < 10; ++i)
				if (fabs(C[i][k]) > big) { big = fabs(C[i][k]); p = i; }
			if (p != k)
				for (int j = 0; j < 20; ++j) { const double t = C[k][j]; C[k][j] = C[p][j]; C[p][j] = t; }
			const double piv = C[k][k];
			for (int i = k + 1; i < 10; ++i)
			{
				const double f = C[i][k] / piv;
				if (f == 0.0) continue;
				for (int j = k + 1; j < 10; ++j) C[i][j] -= f * C[k][j];
				for (int j = 10; j < 20; ++j) C[i][j] -= f * C[k][j];
			}
		}
		// back substitution, only rows 4..9 of X are needed but every row feeds the ones above it: solve all
		for (int c = 10; c < 20; ++c)
			for (int i = 9; i >= 0; --i)
			{
				double s = C[i][c];
				for (int j = i + 1; j < 10; ++j) s -= C[i][j] * C[j][c];
				C[i][c] = s / C[i][i];
			}
		// ---- polynomial rows in z ----
		double px[3][4], py[3][4], pc[3][5];
		for (int i = 0; i < 3; ++i)
		{
			const double *ra = &C[4 + 2 * i][10], *rb = &C[5 + 2 * i][10];
			px[i][3] = -rb[0]; px[i][2] = ra[0] - rb[1]; px[i][1] = ra[1] - rb[2]; px[i][0] = ra[2];
			py[i][3] = -rb[3]; py[i][2] = ra[3] - rb[4]; py[i][1] = ra[4] - rb[5]; py[i][0] = ra[5];
			pc[i][4] = -rb[6]; pc[i][3] = ra[6] - rb[7]; pc[i][2] = ra[7] - rb[8]; pc[i][1] = ra[8] - rb[9]; pc[i][0] = ra[9];
		}
		double c[11];
		for (int i = 0; i <= 10; ++i) c[i] = 0.0;
		const int perm[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {0, 2, 1}, {1, 0, 2}, {2, 1, 0}};
		for (int p = 0; p < 6; ++p)
		{
			double t6[7], t10[11];
			conv(px[perm[p][0]], 3, py[perm[p][1]], 3, t6);
			conv(t6, 6, pc[perm[p][2]], 4, t10);
			const double sign = p < 3 ? 1.0 : -1.0;
			for (int i = 0; i <= 10; ++i) c[i] += sign * t10[i];
		}
		double roots[10];
		const int n_sols = real_roots10(c, roots);
		int n_models = 0;
		for (int s = 0; s < n_sols; ++s)
		{
			const double z = roots[s];
			double B[3][2], b[3];
			for (int i = 0; i < 3; ++i)
			{
				B[i][0] = horner(px[i], 3, z);
				B[i][1] = horner(py[i], 3, z);
				b[i] = horner(pc[i], 4, z);
			}
			const double det = B[0][0] * B[1][1] - B[0][1] * B[1][0];
			double xz0 = (B[1][1] * b[0] - B[0][1] * b[1]) / det;
			double xz1 = (-B[1][0] * b[0] + B[0][0] * b[1]) / det;
			if (!(fabs(B[2][0] * xz0 + B[2][1] * xz1 - b[2]) <= 1e-6))
			{
				double M[3][3] = {{B[0][0], B[0][1], b[0]}, {B[1][0], B[1][1], b[1]}, {B[2][0], B[2][1], b[2]}};
				int cp[2] = {0, 1};
				for (int k = 0; k < 2; ++k)
				{
					int pr = k, pcq = k;
					double big = 0.0;
					for (int j = k; j < 2; ++j)
						for (int i = k; i < 3; ++i)
							if (fabs(M[i][j]) > big) { big = fabs(M[i][j]); pr = i; pcq = j; }
					if (pr != k) for (int j = 0; j < 3; ++j) { const double t = M[k][j]; M[k][j] = M[pr][j]; M[pr][j] = t; }
					if (pcq != k)
					{
						for (int i = 0; i < 3; ++i) { const double t = M[i][k]; M[i][k] = M[i][pcq]; M[i][pcq] = t; }
						const int t = cp[k]; cp[k] = cp[pcq]; cp[pcq] = t;
					}
					for (int i = k + 1; i < 3; ++i)
					{
						const double f = M[i][k] / M[k][k];
						for (int j = k; j < 3; ++j) M[i][j] -= f * M[k][j];
					}
				}
				double v[2];
				v[1] = M[1][2] / M[1][1];
				v[0] = (M[0][2] - M[0][1] * v[1]) / M[0][0];
				double sol[2];
				sol[cp[0]] = v[0];
				sol[cp[1]] = v[1];
				xz0 = sol[0];
				xz1 = sol[1];
			}
			const double x = -xz0, y = -xz1;
			const double inv = 1.0 / sqrt(x * x + y * y + z * z + 1.0);
			double e[9];
			bool finite = true;
			for (int j = 0; j < 9; ++j)
			{
				e[j] = (N[0][j] * x + N[1][j] * y + N[2][j] * z + N[3][j]) * inv;
				if (!isfinite(e[j])) finite = false;
			}
			if (!finite)
				continue;
			double *E = E_out + 9 * n_models;
			for (int r = 0; r < 3; ++r)
				for (int cc = 0; cc < 3; ++cc)
					E[3 * r + cc] = e[3 * cc + r];
			++n_models;
		}
		return n_models;
	}
	// pt(i, c): sample point i (x0 y0 x1 y1, K-normalised); E: up to 10 x 9 row-major; returns the number of models
	template <class PointFn>
	__device__ inline int solve(const PointFn &pt, double *E_out)
	{
		// ---- 5 x 9 epipolar rows and the complete-pivoting LU ----
		double lu[5][9];
		for (int i = 0; i < 5; ++i)
		{
			const double x0 = pt(i, 0), y0 = pt(i, 1), x1 = pt(i, 2), y1 = pt(i, 3);
			lu[i][0] = x0 * x1; lu[i][1] = x0 * y1; lu[i][2] = x0; lu[i][3] = y0 * x1; lu[i][4] = y0 * y1; lu[i][5] = y0;
			lu[i][6] = x1; lu[i][7] = y1; lu[i][8] = 1.0;
		}
		double N[4][9];
		if (!null_space_lu(lu, N))
			return 0;
		return solve_from_basis(N, E_out);
	}
}
