// oracle/gcr_oracle_e.cpp -- TEST INFRASTRUCTURE ONLY (see gcr_oracle.h).
// Essential-matrix path: Nister 5-point minimal solver (the reference's default minimal E solver, types.h:108-110),
// the oriented epipolar constraint of EssentialMatrixEstimator::estimateModel, and normalizeCorrespondences.
// [pinned:source]: the reference version calls Eigen (FullPivLU::kernel, partialPivLu), absent here; this restatement
// follows the same construction (file:line beside each step), is cross-checked against numpy in
// tests/test_oracle_solvers.py (epipolar / trace / determinant constraints, ground-truth recovery, root count against
// numpy.roots of the same degree-10 polynomial) and held against the reference's own source compiled verbatim over
// oracle/shim_solvers in tests/test_reference_solver_sources.py: identical up to the real-root finder (with the reference's
// own maths/sturm_polynomial_solver.h plugged in: same count, same order, converged roots to 2e-8), where this one parts
// ways on purpose -- it returns a superset of the reference's converged roots (the reference stops bisecting at depth 30
// and accepts a root after ten Newton steps whatever its residual: ~8 % of its models miss the essential-matrix
// constraints by more than 1e-9, ~4 % of the samples lose valid roots).
#include "gcr_oracle.h"
#include "linalg.h"
#include "sturm.h"

#include <cmath>
#include <cstring>
#include <vector>

// Test hook (tests/test_reference_solver_sources.py): another real-root finder for the degree-10 polynomial, ascending
// coefficients in, roots out, returns their number.  oracle/ref_solvers_glue.cpp plugs in the reference's own
// maths/sturm_polynomial_solver.h to show that the root finder is the ONLY place where the restated five-point solver and
// the reference's source part ways (by design: the reference stops bisecting at depth 30 and drops roots it has not
// separated by then, and accepts a root after ten Newton steps whatever its residual).  Null = the oracle's own (sturm.h).
int (*gcro_e5_root_finder)(const double *c11, double *roots10) = nullptr;

namespace
{
	// Monomials of degree <= 3 in (x, y, z), in the column order of the reference's 10 x 20 constraint matrix
	// (solver_essential_matrix_five_point_nister.h:158-160):
	// x^3 y^3 x^2y xy^2 x^2z x^2 y^2z y^2 xyz xy | xz^2 xz x yz^2 yz y z^3 z^2 z 1
	const int MONO3[20][3] = {{3, 0, 0}, {0, 3, 0}, {2, 1, 0}, {1, 2, 0}, {2, 0, 1}, {2, 0, 0}, {0, 2, 1}, {0, 2, 0},
	                          {1, 1, 1}, {1, 1, 0}, {1, 0, 2}, {1, 0, 1}, {1, 0, 0}, {0, 1, 2}, {0, 1, 1}, {0, 1, 0},
	                          {0, 0, 3}, {0, 0, 2}, {0, 0, 1}, {0, 0, 0}};
	// degree <= 2 (:106): x^2 xy xz x y^2 yz y z^2 z 1
	const int MONO2[10][3] = {{2, 0, 0}, {1, 1, 0}, {1, 0, 1}, {1, 0, 0}, {0, 2, 0}, {0, 1, 1}, {0, 1, 0}, {0, 0, 2}, {0, 0, 1}, {0, 0, 0}};
	// degree <= 1: x y z 1
	const int MONO1[4][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}};

	int find_mono(const int (*tab)[3], int n, int ex, int ey, int ez)
	{
		for (int i = 0; i < n; ++i)
			if (tab[i][0] == ex && tab[i][1] == ey && tab[i][2] == ez)
				return i;
		return -1;
	}
	// c (deg 2) += sign * a (deg 1) * b (deg 1)
	void mul11(const double *a, const double *b, double *c, double sign)
	{
		for (int i = 0; i < 4; ++i)
			for (int j = 0; j < 4; ++j)
				c[find_mono(MONO2, 10, MONO1[i][0] + MONO1[j][0], MONO1[i][1] + MONO1[j][1], MONO1[i][2] + MONO1[j][2])] += sign * a[i] * b[j];
	}
	// c (deg 3) += a (deg 2) * b (deg 1)
	void mul21(const double *a, const double *b, double *c)
	{
		for (int i = 0; i < 10; ++i)
			for (int j = 0; j < 4; ++j)
				c[find_mono(MONO3, 20, MONO2[i][0] + MONO1[j][0], MONO2[i][1] + MONO1[j][1], MONO2[i][2] + MONO1[j][2])] += a[i] * b[j];
	}
	// polynomial product in one variable, ascending coefficients
	void conv(const double *a, int da, const double *b, int db, double *c)
	{
		for (int i = 0; i <= da + db; ++i) c[i] = 0.0;
		for (int i = 0; i <= da; ++i)
			for (int j = 0; j <= db; ++j)
				c[i + j] += a[i] * b[j];
	}
}

extern "C" {

// solver_essential_matrix_five_point_nister.h:263-615, minimal branch (sample_number_ == 5, weights_ == nullptr)
int gcro_solve_e5_nister(const double *pts, const uint32_t *sample, double *E90)
{
	// :282-319 epipolar rows [x0x1 x0y1 x0 y0x1 y0y1 y0 x1 y1 1]
	double A[5 * 9];
	for (int i = 0; i < 5; ++i)
	{
		const double *p = pts + 4 * (size_t)(sample ? sample[i] : i);
		const double x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
		double *r = A + 9 * i;
		r[0] = x0 * x1; r[1] = x0 * y1; r[2] = x0; r[3] = y0 * x1; r[4] = y0 * y1; r[5] = y0; r[6] = x1; r[7] = y1; r[8] = 1.0;
	}
	// :324-330 4-D null space (FullPivLU::kernel); N[b][j]: basis vector b, coefficient j
	double ker[9 * 4];
	if (gcro::fullpiv_lu_kernel(A, 5, 9, ker, 4) != 4)
		return 0;
	double N[4][9];
	for (int j = 0; j < 9; ++j)
		for (int b = 0; b < 4; ++b)
			N[b][j] = ker[j * 4 + b];
	// entry (i, j) of E = x N0 + y N1 + z N2 + N3 as a degree-1 polynomial [x y z 1]; coefficient index 3 j + i (:208-209)
	double Ee[3][3][4];
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
			for (int b = 0; b < 4; ++b)
				Ee[i][j][b] = N[b][3 * j + i];

	// :204-261 constraint rows: 9 x (E E^T - 1/2 trace(E E^T) I) E = 0, then det E = 0 (row 9)
	double C[10][20];
	std::memset(C, 0, sizeof(C));
	double EET[3][3][10];
	std::memset(EET, 0, sizeof(EET));
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
			for (int k = 0; k < 3; ++k)
				mul11(Ee[i][k], Ee[j][k], EET[i][j], 1.0);
	for (int m = 0; m < 10; ++m)
	{
		const double t = 0.5 * (EET[0][0][m] + EET[1][1][m] + EET[2][2][m]);
		EET[0][0][m] -= t; EET[1][1][m] -= t; EET[2][2][m] -= t;
	}
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
			for (int k = 0; k < 3; ++k)
				mul21(EET[i][k], Ee[k][j], C[3 * i + j]);
	{
		double d[10];
		// det E = sum over rows of the third row entry times the 2 x 2 minors of the first two rows
		std::memset(d, 0, sizeof(d)); mul11(Ee[0][1], Ee[1][2], d, 1.0); mul11(Ee[0][2], Ee[1][1], d, -1.0); mul21(d, Ee[2][0], C[9]);
		std::memset(d, 0, sizeof(d)); mul11(Ee[0][2], Ee[1][0], d, 1.0); mul11(Ee[0][0], Ee[1][2], d, -1.0); mul21(d, Ee[2][1], C[9]);
		std::memset(d, 0, sizeof(d)); mul11(Ee[0][0], Ee[1][1], d, 1.0); mul11(Ee[0][1], Ee[1][0], d, -1.0); mul21(d, Ee[2][2], C[9]);
	}
	// :350 right block <- left^-1 * right (partialPivLu)
	double L[100], Rm[100], X[100];
	for (int i = 0; i < 10; ++i)
		for (int j = 0; j < 10; ++j)
		{
			L[10 * i + j] = C[i][j];
			Rm[10 * i + j] = C[i][10 + j];
		}
	if (!gcro::partialpiv_lu_solve(L, 10, Rm, 10, X))
		return 0;
	// :353-366 eliminate x z^2.., y z^2..: polynomials in z.  px_i, py_i (deg 3), pc_i (deg 4), ascending coefficients.
	double px[3][4], py[3][4], pc[3][5];
	for (int i = 0; i < 3; ++i)
	{
		const double *ra = X + 10 * (4 + 2 * i), *rb = X + 10 * (5 + 2 * i);
		// columns: 0 xz^2, 1 xz, 2 x | 3 yz^2, 4 yz, 5 y | 6 z^3, 7 z^2, 8 z, 9 1;  row_a - z * row_b
		px[i][3] = -rb[0]; px[i][2] = ra[0] - rb[1]; px[i][1] = ra[1] - rb[2]; px[i][0] = ra[2];
		py[i][3] = -rb[3]; py[i][2] = ra[3] - rb[4]; py[i][1] = ra[4] - rb[5]; py[i][0] = ra[5];
		pc[i][4] = -rb[6]; pc[i][3] = ra[6] - rb[7]; pc[i][2] = ra[7] - rb[8]; pc[i][1] = ra[8] - rb[9]; pc[i][0] = ra[9];
	}
	// :410-570 degree-10 determinant of [px py pc] (expanded there; polynomial products here)
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
	// :574 real roots
	double roots[10];
	const int n_sols = gcro_e5_root_finder ? gcro_e5_root_finder(c, roots) : gcro_sturm::real_roots(c, 10, roots);
	// :583-612 back substitution
	int n_models = 0;
	for (int s = 0; s < n_sols; ++s)
	{
		const double z = roots[s];
		double B[3][2], b[3];
		for (int i = 0; i < 3; ++i)
		{
			B[i][0] = gcro_sturm::horner(px[i], 3, z);
			B[i][1] = gcro_sturm::horner(py[i], 3, z);
			b[i] = gcro_sturm::horner(pc[i], 4, z);
		}
		const double det = B[0][0] * B[1][1] - B[0][1] * B[1][0];
		double xz0 = (B[1][1] * b[0] - B[0][1] * b[1]) / det;
		double xz1 = (-B[1][0] * b[0] + B[0][0] * b[1]) / det;
		if (!(std::fabs(B[2][0] * xz0 + B[2][1] * xz1 - b[2]) <= 1e-6))
		{
			// :599-601 fullPivLu().solve on the 3 x 2 system: eliminate with complete pivoting, solve the 2 pivot rows
			double M[3][3] = {{B[0][0], B[0][1], b[0]}, {B[1][0], B[1][1], b[1]}, {B[2][0], B[2][1], b[2]}};
			int cp[2] = {0, 1};
			for (int k = 0; k < 2; ++k)
			{
				int pr = k, pcq = k;
				double big = 0.0;
				for (int j = k; j < 2; ++j)
					for (int i = k; i < 3; ++i)
						if (std::fabs(M[i][j]) > big) { big = std::fabs(M[i][j]); pr = i; pcq = j; }
				if (pr != k) for (int j = 0; j < 3; ++j) std::swap(M[k][j], M[pr][j]);
				if (pcq != k) { for (int i = 0; i < 3; ++i) std::swap(M[i][k], M[i][pcq]); std::swap(cp[k], cp[pcq]); }
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
		const double inv = 1.0 / std::sqrt(x * x + y * y + z * z + 1.0);
		double e[9];
		bool finite = true;
		for (int j = 0; j < 9; ++j)
		{
			e[j] = (N[0][j] * x + N[1][j] * y + N[2][j] * z + N[3][j]) * inv;
			if (!std::isfinite(e[j])) finite = false;
		}
		if (!finite)
			continue;
		double *E = E90 + 9 * n_models;
		for (int r = 0; r < 3; ++r)
			for (int cc = 0; cc < 3; ++cc)
				E[3 * r + cc] = e[3 * cc + r]; // e maps onto column-major E.data() (:581, 604)
		++n_models;
	}
	return n_models;
}

// ---- Stewenius non-minimal solver ------------------------------------------------------------------------------------
namespace
{
	// GrevLex orders of solver_essential_matrix_five_point_stewenius.h:232-318
	// degree <= 2: x^2 xy y^2 xz yz z^2 x y z 1
	const int S_MONO2[10][3] = {{2, 0, 0}, {1, 1, 0}, {0, 2, 0}, {1, 0, 1}, {0, 1, 1}, {0, 0, 2}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}};
	// degree <= 3: x^3 x^2y xy^2 y^3 x^2z xyz y^2z xz^2 yz^2 z^3 x^2 xy y^2 xz yz z^2 x y z 1
	const int S_MONO3[20][3] = {{3, 0, 0}, {2, 1, 0}, {1, 2, 0}, {0, 3, 0}, {2, 0, 1}, {1, 1, 1}, {0, 2, 1}, {1, 0, 2}, {0, 1, 2}, {0, 0, 3},
	                            {2, 0, 0}, {1, 1, 0}, {0, 2, 0}, {1, 0, 1}, {0, 1, 1}, {0, 0, 2}, {1, 0, 0}, {0, 1, 0}, {0, 0, 1}, {0, 0, 0}};
	// c (deg 2, GrevLex) += s * a * b  (multiplyDegOnePoly, :235-260)
	void s_mul11(const double *a, const double *b, double *c, double s)
	{
		for (int i = 0; i < 4; ++i)
			for (int j = 0; j < 4; ++j)
				c[find_mono(S_MONO2, 10, MONO1[i][0] + MONO1[j][0], MONO1[i][1] + MONO1[j][1], MONO1[i][2] + MONO1[j][2])] += s * a[i] * b[j];
	}
	// c (deg 3, GrevLex) += s * a (deg 2) * b (deg 1)  (multiplyDegTwoDegOnePoly, :264-309)
	void s_mul21(const double *a, const double *b, double *c, double s)
	{
		for (int i = 0; i < 10; ++i)
			for (int j = 0; j < 4; ++j)
				c[find_mono(S_MONO3, 20, S_MONO2[i][0] + MONO1[j][0], S_MONO2[i][1] + MONO1[j][1], S_MONO2[i][2] + MONO1[j][2])] += s * a[i] * b[j];
	}
}

extern "C" {

// solver_essential_matrix_five_point_stewenius.h:114-231.  weights (optional) are read by POINT index
// (weights_[sample_[i]], :143) -- unlike the H / 8-point solvers.  Models come out un-normalised, in the order of the
// action matrix' real eigenvalues as our QR iteration deflates them (Eigen's order is equally an artefact of its
// Schur form; callers compare score-wise, GCRANSAC.h:866-890).
int gcro_solve_e5_stewenius(const double *pts, const uint32_t *sample, int n, const double *weights, double *E90)
{
	if (n < 5)
		return 0;
	// :127-168 weighted epipolar rows, and their normal matrix (:182-183)
	std::vector<double> A((size_t)n * 9);
	for (int i = 0; i < n; ++i)
	{
		const size_t idx = sample ? sample[i] : (size_t)i;
		const double *p = pts + 4 * idx;
		const double w = weights ? weights[idx] : 1.0;
		const double x0 = p[0], y0 = p[1], x1 = p[2], y1 = p[3];
		const double wx0 = w * x0, wx1 = w * x1, wy0 = w * y0, wy1 = w * y1;
		double *r = &A[(size_t)i * 9];
		r[0] = wx0 * x1; r[1] = wx0 * y1; r[2] = wx0; r[3] = wy0 * x1; r[4] = wy0 * y1; r[5] = wy0; r[6] = wx1; r[7] = wy1; r[8] = w;
	}
	double N[9][4]; // nullSpace (9 x 4)
	if (n == 5)
	{
		double ker[9 * 4];
		if (gcro::fullpiv_lu_kernel(A.data(), 5, 9, ker, 4) != 4) // :176-180
			return 0;
		for (int j = 0; j < 9; ++j)
			for (int b = 0; b < 4; ++b)
				N[j][b] = ker[j * 4 + b];
	}
	else
	{
		double AtA[81], Q[81];
		for (int a = 0; a < 9; ++a)
			for (int b = 0; b < 9; ++b)
			{
				double s = 0;
				for (int i = 0; i < n; ++i)
					s += A[(size_t)i * 9 + a] * A[(size_t)i * 9 + b];
				AtA[9 * a + b] = s;
			}
		gcro::colpiv_householder_q(AtA, 9, Q); // :182-187 ColPivHouseholderQR(A^T A).matrixQ().rightCols<4>()
		for (int j = 0; j < 9; ++j)
			for (int b = 0; b < 4; ++b)
				N[j][b] = Q[9 * j + 5 + b];
	}
	// :189-192 E(i, j) as a polynomial in [x y z 1]: row i + 3 j of the null space
	const double *Ee[3][3];
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
			Ee[i][j] = N[i + 3 * j];
	// :195, :361-390 trace constraint rows 0..8, determinant row 9
	double C[10][20];
	std::memset(C, 0, sizeof(C));
	double eet[3][3][10], trace[10];
	std::memset(eet, 0, sizeof(eet));
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
			for (int k = 0; k < 3; ++k)
				s_mul11(Ee[i][k], Ee[j][k], eet[i][j], 2.0);
	for (int m = 0; m < 10; ++m)
		trace[m] = eet[0][0][m] + eet[1][1][m] + eet[2][2][m];
	for (int i = 0; i < 3; ++i)
		for (int j = 0; j < 3; ++j)
		{
			for (int k = 0; k < 3; ++k)
				s_mul21(eet[i][k], Ee[k][j], C[3 * i + j], 1.0);
			s_mul21(trace, Ee[i][j], C[3 * i + j], -0.5);
		}
	{
		double d[10];
		std::memset(d, 0, sizeof(d)); s_mul11(Ee[0][1], Ee[1][2], d, 1.0); s_mul11(Ee[0][2], Ee[1][1], d, -1.0); s_mul21(d, Ee[2][0], C[9], 1.0);
		std::memset(d, 0, sizeof(d)); s_mul11(Ee[0][2], Ee[1][0], d, 1.0); s_mul11(Ee[0][0], Ee[1][2], d, -1.0); s_mul21(d, Ee[2][1], C[9], 1.0);
		std::memset(d, 0, sizeof(d)); s_mul11(Ee[0][0], Ee[1][1], d, 1.0); s_mul11(Ee[0][1], Ee[1][0], d, -1.0); s_mul21(d, Ee[2][2], C[9], 1.0);
	}
	// :198-199 FullPivLU of the left block, solve for the right block
	double L[100], R[100], X[100];
	for (int i = 0; i < 10; ++i)
		for (int j = 0; j < 10; ++j)
		{
			L[10 * i + j] = C[i][j];
			R[10 * i + j] = C[i][10 + j];
		}
	if (!gcro::fullpiv_lu_solve(L, 10, R, 10, X))
		return 0;
	// :201-209 action matrix
	double Am[100];
	std::memset(Am, 0, sizeof(Am));
	const int src_row[6] = {0, 1, 2, 4, 5, 7};
	for (int r = 0; r < 6; ++r)
		for (int j = 0; j < 10; ++j)
			Am[10 * r + j] = X[10 * src_row[r] + j];
	Am[10 * 6 + 0] = -1.0;
	Am[10 * 7 + 1] = -1.0;
	Am[10 * 8 + 3] = -1.0;
	Am[10 * 9 + 6] = -1.0;
	// :211-228 real eigenvalues -> eigenvector tail (x y z 1) -> E
	double wr[10], wi[10];
	for (int i = 0; i < 100; ++i)
		if (!std::isfinite(Am[i]))
			return 0;
	if (!gcro::real_eigenvalues(Am, 10, wr, wi))
		return 0;
	int n_models = 0;
	for (int e = 0; e < 10; ++e)
	{
		if (wi[e] != 0.0)
			continue;
		double v[10];
		gcro::eigenvector_of(Am, 10, wr[e], v);
		double *E = E90 + 9 * n_models;
		for (int k = 0; k < 9; ++k) // E.data()[k] (column-major) = nullSpace.row(k) . v.tail<4>()
		{
			const double val = N[k][0] * v[6] + N[k][1] * v[7] + N[k][2] * v[8] + N[k][3] * v[9];
			E[3 * (k % 3) + k / 3] = val;
		}
		++n_models;
	}
	return n_models;
}

// essential_estimator.h:306-336 estimateModelNonminimal: Stewenius + Frobenius normalisation + sign (E33 >= 0)
int gcro_refit_e(const double *pts, const uint32_t *sample, int n, const double *weights, double *E90)
{
	if (n < 5) // nonMinimalSampleSize()
		return 0;
	const int count = gcro_solve_e5_stewenius(pts, sample, n, weights, E90);
	for (int m = 0; m < count; ++m)
	{
		double *E = E90 + 9 * m;
		double s = 0;
		for (int k = 0; k < 9; ++k) s += E[k] * E[k];
		s = std::sqrt(s);
		for (int k = 0; k < 9; ++k) E[k] /= s;
		if (E[8] < 0)
			for (int k = 0; k < 9; ++k) E[k] = -E[k];
	}
	return count;
}

/* test hooks for the dense kernels of linalg.h (pinned against numpy in tests/test_oracle_solvers.py) */
void gcro_test_colpiv_q(const double *A, int n, double *Q) { gcro::colpiv_householder_q(A, n, Q); }
int gcro_test_eigenvalues(const double *A, int n, double *wr, double *wi) { return gcro::real_eigenvalues(A, n, wr, wi) ? 1 : 0; }
void gcro_test_eigenvector(const double *A, int n, double lambda, double *v) { gcro::eigenvector_of(A, n, lambda, v); }

} // extern "C"

// essential_estimator.h:338-401 (same arithmetic as fundamental_estimator.h:737-800)
static int e_orientation_valid(const double *E, const double *pts, const uint32_t *sample, int m)
{
	return gcro_f_orientation_valid(E, pts, sample, m);
}

// essential_estimator.h:152-175: Nister + orientation filter (models erased from the back; order of the kept ones stays)
int gcro_estimate_e5(const double *pts, const uint32_t *sample, double *E90)
{
	double all[90];
	const int n = gcro_solve_e5_nister(pts, sample, all);
	int kept = 0;
	for (int i = 0; i < n; ++i)
		if (e_orientation_valid(all + 9 * i, pts, sample, 5))
		{
			std::memcpy(E90 + 9 * kept, all + 9 * i, 9 * sizeof(double));
			++kept;
		}
	return kept;
}

// utils.h:65-110
void gcro_normalize_correspondences(const double *pts, int n, const double *K1, const double *K2, double *out)
{
	double K1i[9], K2i[9];
	gcro::inv3(K1, K1i);
	gcro::inv3(K2, K2i);
	for (int r = 0; r < n; ++r)
	{
		const double x0 = pts[4 * (size_t)r], y0 = pts[4 * (size_t)r + 1], x1 = pts[4 * (size_t)r + 2], y1 = pts[4 * (size_t)r + 3];
		out[4 * (size_t)r + 0] = K1i[0] * x0 + K1i[1] * y0 + K1i[2] * 1.0;
		out[4 * (size_t)r + 1] = K1i[3] * x0 + K1i[4] * y0 + K1i[5] * 1.0;
		out[4 * (size_t)r + 2] = K2i[0] * x1 + K2i[1] * y1 + K2i[2] * 1.0;
		out[4 * (size_t)r + 3] = K2i[3] * x1 + K2i[4] * y1 + K2i[5] * 1.0;
	}
}

} // extern "C"
