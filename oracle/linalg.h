// oracle/linalg.h -- TEST INFRASTRUCTURE ONLY.
// Small dense FP64 kernels that stand in for the Eigen 3.4 decompositions the reference's solvers call
// (Eigen is neither vendored in /root/reference nor installed here; SURVEY.md 8c, Appendix A.2 list the
// call sites and the semantics mirrored below).  Row-major storage everywhere.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace gcro
{
	inline double det3(const double *M)
	{
		return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) +
		       M[2] * (M[3] * M[7] - M[4] * M[6]);
	}

	// Eigen Matrix3d::inverse(): cofactor matrix / determinant
	inline void inv3(const double *M, double *R)
	{
		const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
		const double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
		const double id = 1.0 / det;
		R[0] = c00 * id; R[1] = (M[2] * M[7] - M[1] * M[8]) * id; R[2] = (M[1] * M[5] - M[2] * M[4]) * id;
		R[3] = c01 * id; R[4] = (M[0] * M[8] - M[2] * M[6]) * id; R[5] = (M[2] * M[3] - M[0] * M[5]) * id;
		R[6] = c02 * id; R[7] = (M[1] * M[6] - M[0] * M[7]) * id; R[8] = (M[0] * M[4] - M[1] * M[3]) * id;
	}

	inline void mul3(const double *A, const double *B, double *C)
	{
		double t[9];
		for (int i = 0; i < 3; ++i)
			for (int j = 0; j < 3; ++j)
				t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
		std::memcpy(C, t, sizeof(t));
	}

	inline void transpose3(const double *A, double *T)
	{
		double t[9] = {A[0], A[3], A[6], A[1], A[4], A[7], A[2], A[5], A[8]};
		std::memcpy(T, t, sizeof(t));
	}

	inline void cross3(const double *a, const double *b, double *c)
	{
		const double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
		c[0] = x; c[1] = y; c[2] = z;
	}

	// ---- Householder QR with column pivoting: least-squares solve (Eigen colPivHouseholderQr().solve) -------
	// A: rows x cols (rows >= cols), overwritten copy inside.  Rank decided like Eigen: pivots with
	// |R_kk| > eps * cols-ish * |R_00| are kept; free variables are set to zero.
	inline void colpiv_qr_solve(const double *A_in, int rows, int cols, const double *b_in, double *x)
	{
		std::vector<double> A(A_in, A_in + (size_t)rows * cols), b(b_in, b_in + rows);
		std::vector<int> perm(cols);
		std::vector<double> colnorm(cols);
		for (int j = 0; j < cols; ++j)
		{
			perm[j] = j;
			double s = 0;
			for (int i = 0; i < rows; ++i)
				s += A[(size_t)i * cols + j] * A[(size_t)i * cols + j];
			colnorm[j] = s;
		}
		int rank = 0;
		double max_pivot = 0.0;
		const int steps = std::min(rows, cols);
		for (int k = 0; k < steps; ++k)
		{
			// pivot: column with the largest remaining norm (recomputed exactly for robustness)
			int best = k;
			double best_norm = -1;
			for (int j = k; j < cols; ++j)
			{
				double s = 0;
				for (int i = k; i < rows; ++i)
					s += A[(size_t)i * cols + j] * A[(size_t)i * cols + j];
				colnorm[j] = s;
				if (s > best_norm) { best_norm = s; best = j; }
			}
			if (best != k)
			{
				for (int i = 0; i < rows; ++i)
					std::swap(A[(size_t)i * cols + k], A[(size_t)i * cols + best]);
				std::swap(perm[k], perm[best]);
			}
			// Householder vector for column k
			double norm = std::sqrt(best_norm);
			if (norm == 0.0)
				break;
			const double akk = A[(size_t)k * cols + k];
			const double alpha = akk > 0 ? -norm : norm;
			std::vector<double> v(rows - k);
			v[0] = akk - alpha;
			for (int i = k + 1; i < rows; ++i)
				v[i - k] = A[(size_t)i * cols + k];
			double vnorm2 = 0;
			for (double e : v)
				vnorm2 += e * e;
			if (vnorm2 > 0)
			{
				for (int j = k; j < cols; ++j)
				{
					double dot = 0;
					for (int i = k; i < rows; ++i)
						dot += v[i - k] * A[(size_t)i * cols + j];
					const double f = 2.0 * dot / vnorm2;
					for (int i = k; i < rows; ++i)
						A[(size_t)i * cols + j] -= f * v[i - k];
				}
				double dot = 0;
				for (int i = k; i < rows; ++i)
					dot += v[i - k] * b[i];
				const double f = 2.0 * dot / vnorm2;
				for (int i = k; i < rows; ++i)
					b[i] -= f * v[i - k];
			}
			const double piv = std::fabs(A[(size_t)k * cols + k]);
			if (k == 0)
				max_pivot = piv;
			if (piv > max_pivot)
				max_pivot = piv;
			++rank;
		}
		// Eigen's default threshold: eps * min(rows, cols) relative to the biggest pivot
		const double thresh = std::numeric_limits<double>::epsilon() * steps * max_pivot;
		int nonzero = 0;
		for (int k = 0; k < rank; ++k)
			if (std::fabs(A[(size_t)k * cols + k]) > thresh)
				++nonzero;
		std::vector<double> y(cols, 0.0);
		for (int i = nonzero - 1; i >= 0; --i)
		{
			double s = b[i];
			for (int j = i + 1; j < nonzero; ++j)
				s -= A[(size_t)i * cols + j] * y[j];
			y[i] = s / A[(size_t)i * cols + i];
		}
		for (int j = 0; j < cols; ++j)
			x[perm[j]] = y[j];
	}

	// ---- Full-pivot LU (Eigen FullPivLU): rank + kernel basis ------------------------------------------------
	// A: rows x cols.  Returns dimension of the kernel; kernel written as `cols x dim` row-major (each COLUMN is a
	// basis vector) following Eigen's construction: after P A Q = L U, for each non-pivot column k the vector has
	// 1 at that free variable, -U11^-1 U12 at the pivot variables, then is un-permuted by Q (Appendix A.2).
	// Eigen orders things so that kernel column i corresponds to the i-th non-pivot column after the permutation.
	inline int fullpiv_lu_kernel(const double *A_in, int rows, int cols, double *kernel /* cols x (cols-rank) */,
	                             int max_dim)
	{
		std::vector<double> lu(A_in, A_in + (size_t)rows * cols);
		std::vector<int> colperm(cols);
		for (int j = 0; j < cols; ++j)
			colperm[j] = j;
		const int size = std::min(rows, cols);
		std::vector<double> pivots;
		int nonzero_pivots = size;
		double maxpivot = 0.0;
		for (int k = 0; k < size; ++k)
		{
			int pr = k, pc = k;
			double biggest = 0.0;
			for (int j = k; j < cols; ++j) // column-major scan, like Eigen's maxCoeff visitor on a column-major block
				for (int i = k; i < rows; ++i)
				{
					const double a = std::fabs(lu[(size_t)i * cols + j]);
					if (a > biggest) { biggest = a; pr = i; pc = j; }
				}
			if (biggest == 0.0)
			{
				nonzero_pivots = k;
				break;
			}
			if (biggest > maxpivot)
				maxpivot = biggest;
			if (pr != k)
				for (int j = 0; j < cols; ++j)
					std::swap(lu[(size_t)k * cols + j], lu[(size_t)pr * cols + j]);
			if (pc != k)
			{
				for (int i = 0; i < rows; ++i)
					std::swap(lu[(size_t)i * cols + k], lu[(size_t)i * cols + pc]);
				std::swap(colperm[k], colperm[pc]);
			}
			const double piv = lu[(size_t)k * cols + k];
			for (int i = k + 1; i < rows; ++i)
			{
				const double f = lu[(size_t)i * cols + k] / piv;
				lu[(size_t)i * cols + k] = f;
				for (int j = k + 1; j < cols; ++j)
					lu[(size_t)i * cols + j] -= f * lu[(size_t)k * cols + j];
			}
		}
		// rank with Eigen's default threshold eps * diagonalSize
		const double premult = maxpivot * std::numeric_limits<double>::epsilon() * size;
		std::vector<int> pivcols;
		for (int k = 0; k < nonzero_pivots; ++k)
			if (std::fabs(lu[(size_t)k * cols + k]) > premult)
				pivcols.push_back(k);
		const int rank = (int)pivcols.size();
		const int dimker = cols - rank;
		if (dimker == 0 || dimker > max_dim)
			return dimker;
		// Eigen (kernel_retval<FullPivLU>): gather the pivot rows/cols to the top-left (m), solve the upper
		// triangular system m.topLeft(rank,rank) * X = -m.topRight(rank, dimker), set identity below, un-permute.
		// Build m (rank x cols) with columns reordered: pivot columns first (in order), the others after.
		std::vector<int> order; // position -> lu column
		std::vector<char> ispiv(cols, 0);
		for (int p : pivcols) { order.push_back(p); ispiv[p] = 1; }
		for (int j = 0; j < cols; ++j)
			if (!ispiv[j])
				order.push_back(j);
		// Eigen does this with column swaps m.col(i).swap(m.col(pivots(i))) for i < rank, which yields the same
		// set; when every leading pivot is non-negligible (the generic case) order == identity.
		std::vector<double> m((size_t)rank * cols);
		for (int i = 0; i < rank; ++i)
			for (int j = 0; j < cols; ++j)
				m[(size_t)i * cols + j] = (order[j] >= pivcols[i]) ? lu[(size_t)pivcols[i] * cols + order[j]] : 0.0;
		// note: entries left of the diagonal in U are L multipliers; mask them (upper-triangular view)
		for (int i = 0; i < rank; ++i)
			for (int j = 0; j < i; ++j)
				m[(size_t)i * cols + j] = 0.0;
		std::vector<double> X((size_t)cols * dimker, 0.0); // in permuted variable order
		for (int c = 0; c < dimker; ++c)
		{
			// back substitution for pivot variables
			for (int i = rank - 1; i >= 0; --i)
			{
				double s = -m[(size_t)i * cols + rank + c];
				for (int j = i + 1; j < rank; ++j)
					s -= m[(size_t)i * cols + j] * X[(size_t)j * dimker + c];
				X[(size_t)i * dimker + c] = s / m[(size_t)i * cols + i];
			}
			X[(size_t)(rank + c) * dimker + c] = 1.0;
		}
		// un-permute: variable at permuted position p is original column colperm[order[p]]
		for (int p = 0; p < cols; ++p)
			for (int c = 0; c < dimker; ++c)
				kernel[(size_t)colperm[order[p]] * dimker + c] = X[(size_t)p * dimker + c];
		return dimker;
	}

	// ---- Partial-pivot LU solve A X = B (Eigen partialPivLu().solve), A n x n, B n x nrhs --------------------
	inline bool partialpiv_lu_solve(const double *A_in, int n, const double *B_in, int nrhs, double *X)
	{
		std::vector<double> A(A_in, A_in + (size_t)n * n), B(B_in, B_in + (size_t)n * nrhs);
		for (int k = 0; k < n; ++k)
		{
			int p = k;
			double big = std::fabs(A[(size_t)k * n + k]);
			for (int i = k + 1; i < n; ++i)
				if (std::fabs(A[(size_t)i * n + k]) > big) { big = std::fabs(A[(size_t)i * n + k]); p = i; }
			if (p != k)
			{
				for (int j = 0; j < n; ++j) std::swap(A[(size_t)k * n + j], A[(size_t)p * n + j]);
				for (int j = 0; j < nrhs; ++j) std::swap(B[(size_t)k * nrhs + j], B[(size_t)p * nrhs + j]);
			}
			const double piv = A[(size_t)k * n + k];
			for (int i = k + 1; i < n; ++i)
			{
				const double f = A[(size_t)i * n + k] / piv;
				if (f == 0.0) continue;
				for (int j = k + 1; j < n; ++j) A[(size_t)i * n + j] -= f * A[(size_t)k * n + j];
				for (int j = 0; j < nrhs; ++j) B[(size_t)i * nrhs + j] -= f * B[(size_t)k * nrhs + j];
			}
		}
		for (int c = 0; c < nrhs; ++c)
			for (int i = n - 1; i >= 0; --i)
			{
				double s = B[(size_t)i * nrhs + c];
				for (int j = i + 1; j < n; ++j) s -= A[(size_t)i * n + j] * X[(size_t)j * nrhs + c];
				X[(size_t)i * nrhs + c] = s / A[(size_t)i * n + i];
			}
		return true;
	}

	// ---- Full-pivot LU solve A X = B (Eigen FullPivLU::solve), square n x n ---------------------------------
	inline bool fullpiv_lu_solve(const double *A_in, int n, const double *B_in, int nrhs, double *X)
	{
		std::vector<double> A(A_in, A_in + (size_t)n * n), B(B_in, B_in + (size_t)n * nrhs);
		std::vector<int> cp(n);
		for (int j = 0; j < n; ++j) cp[j] = j;
		for (int k = 0; k < n; ++k)
		{
			int pr = k, pc = k;
			double big = 0;
			for (int j = k; j < n; ++j)
				for (int i = k; i < n; ++i)
					if (std::fabs(A[(size_t)i * n + j]) > big) { big = std::fabs(A[(size_t)i * n + j]); pr = i; pc = j; }
			if (big == 0.0) return false;
			if (pr != k)
			{
				for (int j = 0; j < n; ++j) std::swap(A[(size_t)k * n + j], A[(size_t)pr * n + j]);
				for (int j = 0; j < nrhs; ++j) std::swap(B[(size_t)k * nrhs + j], B[(size_t)pr * nrhs + j]);
			}
			if (pc != k)
			{
				for (int i = 0; i < n; ++i) std::swap(A[(size_t)i * n + k], A[(size_t)i * n + pc]);
				std::swap(cp[k], cp[pc]);
			}
			const double piv = A[(size_t)k * n + k];
			for (int i = k + 1; i < n; ++i)
			{
				const double f = A[(size_t)i * n + k] / piv;
				if (f == 0.0) continue;
				for (int j = k + 1; j < n; ++j) A[(size_t)i * n + j] -= f * A[(size_t)k * n + j];
				for (int j = 0; j < nrhs; ++j) B[(size_t)i * nrhs + j] -= f * B[(size_t)k * nrhs + j];
			}
		}
		std::vector<double> Y((size_t)n * nrhs);
		for (int c = 0; c < nrhs; ++c)
			for (int i = n - 1; i >= 0; --i)
			{
				double s = B[(size_t)i * nrhs + c];
				for (int j = i + 1; j < n; ++j) s -= A[(size_t)i * n + j] * Y[(size_t)j * nrhs + c];
				Y[(size_t)i * nrhs + c] = s / A[(size_t)i * n + i];
			}
		for (int j = 0; j < n; ++j)
			for (int c = 0; c < nrhs; ++c)
				X[(size_t)cp[j] * nrhs + c] = Y[(size_t)j * nrhs + c];
		return true;
	}

	// ---- Symmetric eigen-decomposition by cyclic Jacobi (n <= 16).  Eigenvalues ascending, eigenvectors as
	// COLUMNS of V (row-major n x n).  Used for: null vector of A^T A (8-pt, standing in for the last column of
	// FullPivHouseholderQR's Q of the symmetric positive semi-definite A^T A), 4-D null space of the 9x9 (Stewenius),
	// and the SVDs of 3x3 matrices via A^T A.
	inline void jacobi_eigh(const double *A_in, int n, double *evals, double *V)
	{
		std::vector<double> A(A_in, A_in + (size_t)n * n);
		for (int i = 0; i < n; ++i)
			for (int j = 0; j < n; ++j)
				V[(size_t)i * n + j] = (i == j) ? 1.0 : 0.0;
		for (int sweep = 0; sweep < 100; ++sweep)
		{
			double off = 0.0, diag = 0.0;
			for (int i = 0; i < n; ++i)
				for (int j = 0; j < n; ++j)
					(i == j ? diag : off) += A[(size_t)i * n + j] * A[(size_t)i * n + j];
			if (off <= 1e-300 || off <= 1e-34 * diag)
				break;
			for (int p = 0; p < n - 1; ++p)
				for (int q = p + 1; q < n; ++q)
				{
					const double apq = A[(size_t)p * n + q];
					if (apq == 0.0)
						continue;
					const double app = A[(size_t)p * n + p], aqq = A[(size_t)q * n + q];
					const double theta = (aqq - app) / (2.0 * apq);
					const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
					const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
					for (int k = 0; k < n; ++k)
					{
						const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
						A[(size_t)k * n + p] = c * akp - s * akq;
						A[(size_t)k * n + q] = s * akp + c * akq;
					}
					for (int k = 0; k < n; ++k)
					{
						const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
						A[(size_t)p * n + k] = c * apk - s * aqk;
						A[(size_t)q * n + k] = s * apk + c * aqk;
					}
					for (int k = 0; k < n; ++k)
					{
						const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
						V[(size_t)k * n + p] = c * vkp - s * vkq;
						V[(size_t)k * n + q] = s * vkp + c * vkq;
					}
				}
		}
		std::vector<int> order(n);
		for (int i = 0; i < n; ++i) order[i] = i;
		std::sort(order.begin(), order.end(), [&](int a, int b) { return A[(size_t)a * n + a] < A[(size_t)b * n + b]; });
		std::vector<double> Vs((size_t)n * n);
		for (int c = 0; c < n; ++c)
		{
			evals[c] = A[(size_t)order[c] * n + order[c]];
			for (int r = 0; r < n; ++r)
				Vs[(size_t)r * n + c] = V[(size_t)r * n + order[c]];
		}
		std::memcpy(V, Vs.data(), sizeof(double) * n * n);
	}

	// ---- one-sided Jacobi SVD of a small m x n matrix (m >= n): A = U diag(s) V^T, s descending --------------
	// (Eigen JacobiSVD semantics: singular values sorted in decreasing order.)  U: m x n, V: n x n (columns).
	inline void jacobi_svd(const double *A_in, int m, int n, double *U, double *s, double *V)
	{
		std::vector<double> A(A_in, A_in + (size_t)m * n);
		std::vector<double> Vt((size_t)n * n, 0.0);
		for (int i = 0; i < n; ++i) Vt[(size_t)i * n + i] = 1.0;
		for (int sweep = 0; sweep < 100; ++sweep)
		{
			bool rotated = false;
			for (int p = 0; p < n - 1; ++p)
				for (int q = p + 1; q < n; ++q)
				{
					double alpha = 0, beta = 0, gamma = 0;
					for (int i = 0; i < m; ++i)
					{
						const double ap = A[(size_t)i * n + p], aq = A[(size_t)i * n + q];
						alpha += ap * ap; beta += aq * aq; gamma += ap * aq;
					}
					// orthogonal to 4 eps (below that the dot product is its own rounding noise and the sweeps never end)
					if (std::fabs(gamma) <= 8.881784197001252e-16 * std::sqrt(alpha * beta) || gamma == 0.0)
						continue;
					rotated = true;
					const double zeta = (beta - alpha) / (2.0 * gamma);
					const double t = (zeta >= 0 ? 1.0 : -1.0) / (std::fabs(zeta) + std::sqrt(1.0 + zeta * zeta));
					const double c = 1.0 / std::sqrt(1.0 + t * t), sn = c * t;
					for (int i = 0; i < m; ++i)
					{
						const double ap = A[(size_t)i * n + p], aq = A[(size_t)i * n + q];
						A[(size_t)i * n + p] = c * ap - sn * aq;
						A[(size_t)i * n + q] = sn * ap + c * aq;
					}
					for (int i = 0; i < n; ++i)
					{
						const double vp = Vt[(size_t)i * n + p], vq = Vt[(size_t)i * n + q];
						Vt[(size_t)i * n + p] = c * vp - sn * vq;
						Vt[(size_t)i * n + q] = sn * vp + c * vq;
					}
				}
			if (!rotated)
				break;
		}
		std::vector<double> sv(n);
		for (int j = 0; j < n; ++j)
		{
			double nn = 0;
			for (int i = 0; i < m; ++i) nn += A[(size_t)i * n + j] * A[(size_t)i * n + j];
			sv[j] = std::sqrt(nn);
		}
		std::vector<int> order(n);
		for (int i = 0; i < n; ++i) order[i] = i;
		std::sort(order.begin(), order.end(), [&](int a, int b) { return sv[a] > sv[b]; });
		for (int c = 0; c < n; ++c)
		{
			const int o = order[c];
			s[c] = sv[o];
			for (int i = 0; i < m; ++i)
				U[(size_t)i * n + c] = sv[o] > 0 ? A[(size_t)i * n + o] / sv[o] : 0.0;
			for (int i = 0; i < n; ++i)
				V[(size_t)i * n + c] = Vt[(size_t)i * n + o];
		}
	}

	// ---- Eigen ColPivHouseholderQR::matrixQ() of a square n x n matrix (n <= 16) ------------------------------
	// Restates Eigen 3.4 ColPivHouseholderQR::computeInPlace: pivot = first column of largest (down-dated) norm,
	// LAPACK-style norm down-dating with recomputation below sqrt(eps), Householder vectors (1, essential) with
	// beta = -sign(c0) |x| and tau = (beta - c0) / beta (MatrixBase::makeHouseholder).  Q = H_0 H_1 ... H_{n-1},
	// row-major.  Used by the Stewenius non-minimal solver: Q.rightCols<4>() of A^T A
	// (solver_essential_matrix_five_point_stewenius.h:182-187).
	inline void colpiv_householder_q(const double *A_in, int n, double *Q)
	{
		std::vector<double> A(A_in, A_in + (size_t)n * n);
		std::vector<double> updated(n), direct(n), tau(n, 0.0);
		auto at = [&](int i, int j) -> double & { return A[(size_t)i * n + j]; };
		for (int j = 0; j < n; ++j)
		{
			double s = 0;
			for (int i = 0; i < n; ++i) s += at(i, j) * at(i, j);
			updated[j] = direct[j] = std::sqrt(s);
		}
		const double downdate_threshold = std::sqrt(std::numeric_limits<double>::epsilon());
		for (int k = 0; k < n; ++k)
		{
			int big = k;
			for (int j = k + 1; j < n; ++j)
				if (updated[j] > updated[big]) big = j;
			if (big != k)
			{
				for (int i = 0; i < n; ++i) std::swap(at(i, k), at(i, big));
				std::swap(updated[k], updated[big]);
				std::swap(direct[k], direct[big]);
			}
			// makeHouseholderInPlace on A[k:, k]
			double tail = 0;
			for (int i = k + 1; i < n; ++i) tail += at(i, k) * at(i, k);
			const double c0 = at(k, k);
			double beta;
			if (tail <= std::numeric_limits<double>::min())
			{
				tau[k] = 0.0;
				beta = c0;
				for (int i = k + 1; i < n; ++i) at(i, k) = 0.0;
			}
			else
			{
				beta = std::sqrt(c0 * c0 + tail);
				if (c0 >= 0) beta = -beta;
				for (int i = k + 1; i < n; ++i) at(i, k) /= (c0 - beta);
				tau[k] = (beta - c0) / beta;
			}
			at(k, k) = beta;
			// apply H_k = I - tau v v^T (v = [1; essential]) to the trailing columns
			for (int j = k + 1; j < n; ++j)
			{
				double dot = at(k, j);
				for (int i = k + 1; i < n; ++i) dot += at(i, k) * at(i, j);
				dot *= tau[k];
				at(k, j) -= dot;
				for (int i = k + 1; i < n; ++i) at(i, j) -= dot * at(i, k);
			}
			for (int j = k + 1; j < n; ++j)
			{
				if (updated[j] != 0.0)
				{
					double temp = std::fabs(at(k, j)) / updated[j];
					temp = (1.0 + temp) * (1.0 - temp);
					if (temp < 0) temp = 0;
					const double r = updated[j] / direct[j];
					const double temp2 = temp * r * r;
					if (temp2 <= downdate_threshold)
					{
						double s = 0;
						for (int i = k + 1; i < n; ++i) s += at(i, j) * at(i, j);
						direct[j] = std::sqrt(s);
						updated[j] = direct[j];
					}
					else
						updated[j] *= std::sqrt(temp);
				}
			}
		}
		// Q = H_0 ... H_{n-1} applied to the identity (from the last reflector backwards)
		for (int i = 0; i < n; ++i)
			for (int j = 0; j < n; ++j)
				Q[(size_t)i * n + j] = i == j ? 1.0 : 0.0;
		for (int k = n - 1; k >= 0; --k)
			for (int j = 0; j < n; ++j)
			{
				double dot = Q[(size_t)k * n + j];
				for (int i = k + 1; i < n; ++i) dot += at(i, k) * Q[(size_t)i * n + j];
				dot *= tau[k];
				Q[(size_t)k * n + j] -= dot;
				for (int i = k + 1; i < n; ++i) Q[(size_t)i * n + j] -= dot * at(i, k);
			}
	}

	// ---- eigenvalues of a real non-symmetric n x n matrix (n <= 16): Householder reduction to Hessenberg form +
	// Francis double-shift QR (the EISPACK orthes / hqr pair Eigen's RealSchur derives from; no balancing, like
	// Eigen::EigenSolver).  wr / wi = real / imaginary parts.  Returns false if an eigenvalue fails to converge.
	inline bool real_eigenvalues(const double *A_in, int n, double *wr, double *wi)
	{
		std::vector<double> a(A_in, A_in + (size_t)n * n);
		auto A = [&](int i, int j) -> double & { return a[(size_t)i * n + j]; };
		// orthogonal reduction to upper Hessenberg form
		for (int m = 1; m < n - 1; ++m)
		{
			double scale = 0;
			for (int i = m; i < n; ++i) scale += std::fabs(A(i, m - 1));
			if (scale == 0.0) continue;
			std::vector<double> ort(n, 0.0);
			double h = 0;
			for (int i = n - 1; i >= m; --i)
			{
				ort[i] = A(i, m - 1) / scale;
				h += ort[i] * ort[i];
			}
			double g = std::sqrt(h);
			if (ort[m] > 0) g = -g;
			h -= ort[m] * g;
			ort[m] -= g;
			for (int j = m; j < n; ++j)
			{
				double f = 0;
				for (int i = n - 1; i >= m; --i) f += ort[i] * A(i, j);
				f /= h;
				for (int i = m; i < n; ++i) A(i, j) -= f * ort[i];
			}
			for (int i = 0; i < n; ++i)
			{
				double f = 0;
				for (int j = n - 1; j >= m; --j) f += ort[j] * A(i, j);
				f /= h;
				for (int j = m; j < n; ++j) A(i, j) -= f * ort[j];
			}
			A(m, m - 1) = scale * g;
			for (int i = m + 1; i < n; ++i) A(i, m - 1) = 0.0;
		}
		// hqr
		double anorm = 0;
		for (int i = 0; i < n; ++i)
			for (int j = std::max(i - 1, 0); j < n; ++j)
				anorm += std::fabs(A(i, j));
		int nn = n - 1;
		double t = 0.0;
		auto sign = [](double x, double y) { return y >= 0 ? std::fabs(x) : -std::fabs(x); };
		while (nn >= 0)
		{
			int its = 0, l;
			do
			{
				for (l = nn; l >= 1; --l)
				{
					double s = std::fabs(A(l - 1, l - 1)) + std::fabs(A(l, l));
					if (s == 0.0) s = anorm;
					if (std::fabs(A(l, l - 1)) + s == s)
					{
						A(l, l - 1) = 0.0;
						break;
					}
				}
				double x = A(nn, nn);
				if (l == nn)
				{
					wr[nn] = x + t;
					wi[nn--] = 0.0;
				}
				else
				{
					double y = A(nn - 1, nn - 1);
					double w = A(nn, nn - 1) * A(nn - 1, nn);
					if (l == nn - 1)
					{
						double p = 0.5 * (y - x);
						double q = p * p + w;
						double z = std::sqrt(std::fabs(q));
						x += t;
						if (q >= 0.0)
						{
							z = p + sign(z, p);
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
							for (int i = 0; i <= nn; ++i) A(i, i) -= x;
							const double s = std::fabs(A(nn, nn - 1)) + std::fabs(A(nn - 1, nn - 2));
							y = x = 0.75 * s;
							w = -0.4375 * s * s;
						}
						++its;
						int m;
						double p = 0, q = 0, r = 0, z;
						for (m = nn - 2; m >= l; --m)
						{
							z = A(m, m);
							r = x - z;
							double s = y - z;
							p = (r * s - w) / A(m + 1, m) + A(m, m + 1);
							q = A(m + 1, m + 1) - z - r - s;
							r = A(m + 2, m + 1);
							s = std::fabs(p) + std::fabs(q) + std::fabs(r);
							p /= s; q /= s; r /= s;
							if (m == l) break;
							const double u = std::fabs(A(m, m - 1)) * (std::fabs(q) + std::fabs(r));
							const double v = std::fabs(p) * (std::fabs(A(m - 1, m - 1)) + std::fabs(z) + std::fabs(A(m + 1, m + 1)));
							if (u + v == v) break;
						}
						for (int i = m + 2; i <= nn; ++i)
						{
							A(i, i - 2) = 0.0;
							if (i != m + 2) A(i, i - 3) = 0.0;
						}
						for (int k = m; k <= nn - 1; ++k)
						{
							if (k != m)
							{
								p = A(k, k - 1);
								q = A(k + 1, k - 1);
								r = 0.0;
								if (k != nn - 1) r = A(k + 2, k - 1);
								if ((x = std::fabs(p) + std::fabs(q) + std::fabs(r)) != 0.0)
								{
									p /= x; q /= x; r /= x;
								}
							}
							const double s = sign(std::sqrt(p * p + q * q + r * r), p);
							if (s != 0.0)
							{
								if (k == m)
								{
									if (l != m) A(k, k - 1) = -A(k, k - 1);
								}
								else
									A(k, k - 1) = -s * x;
								p += s;
								x = p / s;
								y = q / s;
								z = r / s;
								q /= p;
								r /= p;
								for (int j = k; j <= nn; ++j)
								{
									p = A(k, j) + q * A(k + 1, j);
									if (k != nn - 1)
									{
										p += r * A(k + 2, j);
										A(k + 2, j) -= p * z;
									}
									A(k + 1, j) -= p * y;
									A(k, j) -= p * x;
								}
								const int mmin = nn < k + 3 ? nn : k + 3;
								for (int i = l; i <= mmin; ++i)
								{
									p = x * A(i, k) + y * A(i, k + 1);
									if (k != nn - 1)
									{
										p += z * A(i, k + 2);
										A(i, k + 2) -= p * r;
									}
									A(i, k + 1) -= p * q;
									A(i, k) -= p;
								}
							}
						}
					}
				}
			} while (l < nn - 1);
		}
		return true;
	}

	// ---- eigenvector of the real eigenvalue lambda of A (n <= 16) by inverse iteration on A - lambda I
	// (partial-pivot LU with tiny pivots replaced by eps |A|, two solves), unit 2-norm like EigenSolver's.
	inline void eigenvector_of(const double *A_in, int n, double lambda, double *v)
	{
		std::vector<double> M(A_in, A_in + (size_t)n * n);
		double anorm = 0;
		for (double e : M) anorm = std::max(anorm, std::fabs(e));
		for (int i = 0; i < n; ++i) M[(size_t)i * n + i] -= lambda;
		const double tiny = std::numeric_limits<double>::epsilon() * (anorm > 0 ? anorm : 1.0);
		std::vector<int> piv(n);
		for (int k = 0; k < n; ++k)
		{
			int p = k;
			for (int i = k + 1; i < n; ++i)
				if (std::fabs(M[(size_t)i * n + k]) > std::fabs(M[(size_t)p * n + k])) p = i;
			piv[k] = p;
			if (p != k)
				for (int j = 0; j < n; ++j) std::swap(M[(size_t)k * n + j], M[(size_t)p * n + j]);
			if (std::fabs(M[(size_t)k * n + k]) < tiny) M[(size_t)k * n + k] = tiny;
			for (int i = k + 1; i < n; ++i)
			{
				const double f = M[(size_t)i * n + k] / M[(size_t)k * n + k];
				M[(size_t)i * n + k] = f;
				for (int j = k + 1; j < n; ++j) M[(size_t)i * n + j] -= f * M[(size_t)k * n + j];
			}
		}
		for (int i = 0; i < n; ++i) v[i] = 1.0;
		for (int iter = 0; iter < 3; ++iter)
		{
			if (iter > 0) // the first pass starts from U x = 1 (LAPACK xLAEIN's choice); later ones apply L too
				for (int k = 0; k < n; ++k)
				{
					std::swap(v[k], v[piv[k]]);
					for (int i = k + 1; i < n; ++i) v[i] -= M[(size_t)i * n + k] * v[k];
				}
			for (int i = n - 1; i >= 0; --i)
			{
				double s = v[i];
				for (int j = i + 1; j < n; ++j) s -= M[(size_t)i * n + j] * v[j];
				v[i] = s / M[(size_t)i * n + i];
			}
			double nrm = 0;
			for (int i = 0; i < n; ++i) nrm += v[i] * v[i];
			nrm = std::sqrt(nrm);
			for (int i = 0; i < n; ++i) v[i] /= nrm;
		}
	}
}
