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
					if (std::fabs(gamma) <= 1e-17 * std::sqrt(alpha * beta) || gamma == 0.0)
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
}
