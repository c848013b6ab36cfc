// solve_e5_group.cuh -- 5-point essential solver, TEN LANES PER SAMPLE, working set in shared memory.
//
// Same algorithm, same arithmetic order per output element and therefore the same bits as gcr_e5::solve
// (solve_e5.cuh; reference: estimators/solver_essential_matrix_five_point_nister.h:263-615), re-mapped for the SM:
// the one-thread-per-sample version keeps ~6.5 KB of thread-local arrays (10 x 20 constraint matrix, Sturm chain,
// interval stack), i.e. ~1 GB of local memory for a resident grid, and runs out of HBM at an IPC of 0.08.  Here a
// group of 10 lanes owns one sample: lane r owns row r of the 10 x 20 system, chain polynomial r of the Sturm
// sequence, root r of the degree-10 polynomial; everything the lanes share lives in ~4.1 KB of shared memory per
// sample and is exchanged through it, with a group-wide barrier (`sync`) between a write and the reads of other lanes.
// There are no shuffles or ballots: a group only ever synchronises with itself, so the three groups of a warp may
// run different numbers of bisection steps.
//
// The code is written against a `Sync` functor so that the host can run it with ten threads and a barrier
// (tests / bring-up); on the device Sync is __syncwarp(group mask).
#pragma once

#include "solve_e5.cuh"

namespace gcr_e5g
{
	// shared-memory layout of one sample, in doubles
	constexpr int C_STRIDE = 21; // row stride of the 10 x 20 matrix: 42 words, conflict-free for 10 lanes reading a column
	constexpr int OFF_N = 0;                      // N[4][9] null-space basis
	constexpr int OFF_PT = OFF_N + 36;            // PT[5][4] sample points
	constexpr int OFF_EET = OFF_PT + 20;          // EET[9][10]; later T10[6][11]
	constexpr int OFF_ABS = OFF_EET + 90;         // 11 scratch doubles (pivot candidates / remainder r[11])
	constexpr int OFF_C = OFF_ABS + 11;           // C[10][21]; before it LU5[5][9]
	constexpr int OFF_CHAIN = OFF_C + 10 * C_STRIDE - 121; // the Sturm chain p[11][11] takes the END of the dead matrix ...
	constexpr int STACK_CAP = 62;                 // bisection depth <= 60: one pending sibling per level + the last pair
	constexpr int OFF_STACK = OFF_EET;            // ... so that the interval stack -- 62 x (a, b) + 62 x packed (sa, sb, depth)
	constexpr int STACK_DOUBLES = 2 * STACK_CAP + STACK_CAP / 2; // -- can run over EET / T10, the scratch and its start
	constexpr int OFF_P = OFF_C + 10 * C_STRIDE;  // px[3][4] py[3][4] pc[3][5]
	constexpr int OFF_POLY = OFF_P + 39;          // c[11] -> f[11] in place, df[11]
	constexpr int OFF_DEG = OFF_POLY + 22;        // 12 ints: deg[11], len
	constexpr int OFF_IV = OFF_DEG + 6;           // isolated intervals (a, b) x 10, then the finite / valid flags
	// + 2: 464 doubles = 928 words would put the three samples of a warp on the SAME banks (928 % 32 == 0) -- three wavefronts
	// for every row / chain access of the 30 lanes; shifted by 4 words per sample the three ten-lane patterns (stride 42 or 22
	// words) interleave with at most two lanes per bank: 44 M -> 15 M bank conflicts, k_solve_e5 622 -> 593 us per 65 536 samples
	// (+ 14 pushes the CTA over 45 KB: four CTAs per SM instead of five, slower.)
#ifndef E5_SAMPLE_PAD
#define E5_SAMPLE_PAD 2
#endif
	constexpr int SAMPLE_DOUBLES = OFF_IV + 30 + E5_SAMPLE_PAD;
	static_assert(OFF_STACK + STACK_DOUBLES <= OFF_CHAIN, "the interval stack must end before the Sturm chain");

	__device__ __forceinline__ int sgn_of(double v) { return v > 0.0 ? 1 : (v < 0.0 ? -1 : 0); }

	// Horner from degree 10 down over a zero-padded coefficient row: the leading zeros contribute 0 * x + 0 = 0 and
	// 0 * x + c = c exactly, so the value equals gcr_e5::horner(c, deg, x) bit for bit, without a per-lane trip count
	__device__ __forceinline__ double horner_padded(const double *c, double x)
	{
		double v = c[10];
#pragma unroll
		for (int i = 9; i >= 0; --i)
			v = v * x + c[i];
		return v;
	}

	// sign changes along a chain whose signs are given as bit masks (bit k: polynomial k is > 0 / < 0), zeros skipped
	__device__ __forceinline__ int count_changes(unsigned pos, unsigned neg, int len)
	{
		const unsigned full = (1u << len) - 1u;
		if (((pos | neg) & full) == full)
			return __popc((pos ^ (pos >> 1)) & (full >> 1));
		int changes = 0, last = 0;
		for (int k = 0; k < len; ++k)
		{
			const int s = ((pos >> k) & 1u) ? 1 : (((neg >> k) & 1u) ? -1 : 0);
			if (s != 0)
			{
				if (last != 0 && s != last)
					++changes;
				last = s;
			}
		}
		return changes;
	}

	// sign changes of the Sturm chain at x: lane k evaluates polynomial k (rows are zero-padded to degree 10), the
	// signs meet in two group ballots; the 11th polynomial, when there is one, is the constant c10
	template <class Sync>
	__device__ __forceinline__ int sign_changes(int li, const double *chain, int len, double c10, const Sync &sync, double x)
	{
		const double v = horner_padded(chain + 11 * li, x);
		unsigned pos = sync.ballot(li < len && v > 0.0), neg = sync.ballot(li < len && v < 0.0);
		if (len == 11)
		{
			pos |= c10 > 0.0 ? 1u << 10 : 0u;
			neg |= c10 < 0.0 ? 1u << 10 : 0u;
		}
		return count_changes(pos, neg, len);
	}

	// One sample by one group of ten lanes (li = 0..9).  S: this sample's shared-memory block (SAMPLE_DOUBLES doubles),
	// PT must already hold the five sample points (x0 y0 x1 y1 per row).  Models go to E_out[10][9] (row-major E, in the
	// order of ascending roots, non-finite ones skipped, unused slots zeroed), their orientation-test verdicts to
	// valid_out[10].  Returns the number of valid models (same value in every lane).
	template <class Sync>
	__device__ inline int solve_group(int li, double *S, const Sync &sync, double *E_out, unsigned char *valid_out)
	{
		const double eps = 2.220446049250313e-16;
		double *N = S + OFF_N;
		const double *PT = S + OFF_PT;
		int n_models = 0; // finite models
		bool have = true;

		// ---------------------------------------------------------------- 5 x 9 epipolar rows, complete-pivoting LU ----
		{
			double *lu = S + OFF_C;            // [5][9]
			double *cand = S + OFF_ABS;        // per row: best |value| (5) and its column (5)
			if (li < 5)
			{
				const double x0 = PT[4 * li], y0 = PT[4 * li + 1], x1 = PT[4 * li + 2], y1 = PT[4 * li + 3];
				double *r = lu + 9 * li;
				r[0] = x0 * x1; r[1] = x0 * y1; r[2] = x0; r[3] = y0 * x1; r[4] = y0 * y1; r[5] = y0;
				r[6] = x1; r[7] = y1; r[8] = 1.0;
			}
			sync();
			unsigned long long colperm = 0x876543210ull;
			double maxpivot = 0.0;
			int nonzero = 5;
			for (int k = 0; k < 5; ++k)
			{
				// row-wise candidates: first maximum in ascending column order
				if (li >= k && li < 5)
				{
					double best = 0.0;
					int bj = k;
					for (int j = k; j < 9; ++j)
					{
						const double a = fabs(lu[9 * li + j]);
						if (a > best) { best = a; bj = j; }
					}
					cand[li] = best;
					cand[5 + li] = (double)bj;
				}
				sync();
				// the element Eigen's column-major visitor returns: the maximum, ties to the smallest column, then row
				int pr = k, pc = k;
				double biggest = 0.0;
				for (int i = k; i < 5; ++i)
				{
					const double a = cand[i];
					const int j = (int)cand[5 + i];
					if (a > biggest || (a == biggest && a > 0.0 && j < pc)) { biggest = a; pr = i; pc = j; }
				}
				if (biggest == 0.0) { nonzero = k; break; }
				if (biggest > maxpivot) maxpivot = biggest;
				sync(); // everybody has read the candidates / the old rows
				if (pr != k && li < 9)
				{
					const double t = lu[9 * k + li]; lu[9 * k + li] = lu[9 * pr + li]; lu[9 * pr + li] = t;
				}
				sync();
				if (pc != k)
				{
					if (li < 5)
					{
						const double t = lu[9 * li + k]; lu[9 * li + k] = lu[9 * li + pc]; lu[9 * li + pc] = t;
					}
					const unsigned long long a = (colperm >> (4 * k)) & 15ull, b = (colperm >> (4 * pc)) & 15ull;
					colperm &= ~((15ull << (4 * k)) | (15ull << (4 * pc)));
					colperm |= (b << (4 * k)) | (a << (4 * pc));
				}
				sync();
				if (li > k && li < 5)
				{
					const double piv = lu[9 * k + k];
					const double f = lu[9 * li + k] / piv;
					lu[9 * li + k] = f;
					for (int j = k + 1; j < 9; ++j)
						lu[9 * li + j] -= f * lu[9 * k + j];
				}
				sync();
			}
			const double premult = maxpivot * eps * 5.0;
			int rank = 0;
			for (int k = 0; k < nonzero; ++k)
				if (fabs(lu[9 * k + k]) > premult) ++rank;
			have = rank == 5;
			if (have && li < 4)
			{
				const int c = li;
				double x[5];
				for (int i = 4; i >= 0; --i)
				{
					double s = -lu[9 * i + 5 + c];
					for (int j = i + 1; j < 5; ++j)
						s -= lu[9 * i + j] * x[j];
					x[i] = s / lu[9 * i + i];
				}
				for (int i = 0; i < 5; ++i)
					N[9 * c + (int)((colperm >> (4 * i)) & 15ull)] = x[i];
				for (int q = 0; q < 4; ++q)
					N[9 * c + (int)((colperm >> (4 * (5 + q))) & 15ull)] = q == c ? 1.0 : 0.0;
			}
			sync();
		}

		double *C = S + OFF_C;
		double *P = S + OFF_P;
		double *poly = S + OFF_POLY; // c[11] f[11] df[11]
		const int n = 10;
		double *f = poly, *df = poly + 11; // f overwrites the raw coefficients in place
		int n_roots = 0, top = 0;
		bool go = false;
		if (have)
		{
			// Ee(i, j)[b] = N[b][3 j + i]
#define GCR_EE(i, j, b) N[9 * (b) + 3 * (j) + (i)]
			// ---------------------------------------------------------------- constraints: E E^T, trace, rows of C ----
			double *EET = S + OFF_EET; // [9][10]
			if (li < 9)
			{
				const int i = li / 3, j = li % 3;
				double acc[10];
				for (int m = 0; m < 10; ++m) acc[m] = 0.0;
				for (int k = 0; k < 3; ++k)
					for (int a = 0; a < 4; ++a)
						for (int b = 0; b < 4; ++b)
							acc[gcr_e5::IDX11[a][b]] += 1.0 * GCR_EE(i, k, a) * GCR_EE(j, k, b);
				for (int m = 0; m < 10; ++m) EET[10 * li + m] = acc[m];
			}
			sync();
			{
				const int m = li; // one monomial per lane
				const double t = 0.5 * (EET[0 * 10 + m] + EET[4 * 10 + m] + EET[8 * 10 + m]);
				EET[0 * 10 + m] -= t; EET[4 * 10 + m] -= t; EET[8 * 10 + m] -= t;
			}
			sync();
			{
				double acc[20];
				for (int m = 0; m < 20; ++m) acc[m] = 0.0;
				if (li < 9)
				{
					const int i = li / 3, j = li % 3;
					for (int k = 0; k < 3; ++k)
						for (int a = 0; a < 10; ++a)
							for (int b = 0; b < 4; ++b)
								acc[gcr_e5::IDX21[a][b]] += EET[10 * (3 * i + k) + a] * GCR_EE(k, j, b);
				}
				else
				{
					// determinant row: three 2 x 2 minors of the first two rows times the third row
					const int ia[3][2] = {{1, 2}, {2, 0}, {0, 1}};
					for (int t = 0; t < 3; ++t)
					{
						const int u = ia[t][0], v = ia[t][1];
						double d[10];
						for (int m = 0; m < 10; ++m) d[m] = 0.0;
						for (int a = 0; a < 4; ++a)
							for (int b = 0; b < 4; ++b)
								d[gcr_e5::IDX11[a][b]] += 1.0 * GCR_EE(0, u, a) * GCR_EE(1, v, b);
						for (int a = 0; a < 4; ++a)
							for (int b = 0; b < 4; ++b)
								d[gcr_e5::IDX11[a][b]] += -1.0 * GCR_EE(0, v, a) * GCR_EE(1, u, b);
						for (int a = 0; a < 10; ++a)
							for (int b = 0; b < 4; ++b)
								acc[gcr_e5::IDX21[a][b]] += d[a] * GCR_EE(2, t, b);
					}
				}
				for (int m = 0; m < 20; ++m) C[C_STRIDE * li + m] = acc[m];
			}
			sync();
#undef GCR_EE
			// ---------------------------------------------------------------- 10 x 10 partial-pivot LU, 10 right-hand sides ----
			// lane = physical row; the row swaps of the scalar code become a position table (nibble q of rowat = the row
			// sitting at position q), identical in every lane
			unsigned long long rowat = 0x9876543210ull;
			int mypos = li;
			double *absval = S + OFF_ABS;
			for (int k = 0; k < 10; ++k)
			{
				if (mypos >= k)
					absval[mypos] = fabs(C[C_STRIDE * li + k]);
				sync();
				int p = k;
				double big = absval[k];
				for (int i = k + 1; i < 10; ++i)
					if (absval[i] > big) { big = absval[i]; p = i; }
				if (p != k)
				{
					const unsigned long long a = (rowat >> (4 * k)) & 15ull, b = (rowat >> (4 * p)) & 15ull;
					rowat &= ~((15ull << (4 * k)) | (15ull << (4 * p)));
					rowat |= (b << (4 * k)) | (a << (4 * p));
					if (mypos == k) mypos = p;
					else if (mypos == p) mypos = k;
				}
				const int prow = (int)((rowat >> (4 * k)) & 15ull);
				if (mypos > k)
				{
					const double piv = C[C_STRIDE * prow + k];
					const double f = C[C_STRIDE * li + k] / piv;
					if (f != 0.0)
						for (int j = k + 1; j < 20; ++j)
							C[C_STRIDE * li + j] -= f * C[C_STRIDE * prow + j];
				}
				sync();
			}
			// back substitution of rows 9..4 (rows 0..3 of the solution are never read); lane c owns right-hand side c
			for (int i = 9; i >= 4; --i)
			{
				const int ri = (int)((rowat >> (4 * i)) & 15ull);
				const int c = 10 + li;
				double s = C[C_STRIDE * ri + c];
				for (int j = i + 1; j < 10; ++j)
					s -= C[C_STRIDE * ri + j] * C[C_STRIDE * (int)((rowat >> (4 * j)) & 15ull) + c];
				C[C_STRIDE * ri + c] = s / C[C_STRIDE * ri + i]; // a lane only ever touches its own column: no barrier
			}
			sync();
			// ---------------------------------------------------------------- the three polynomial rows in z ----
			if (li < 3)
			{
				const int i = li;
				const double *ra = C + C_STRIDE * (int)((rowat >> (4 * (4 + 2 * i))) & 15ull) + 10;
				const double *rb = C + C_STRIDE * (int)((rowat >> (4 * (5 + 2 * i))) & 15ull) + 10;
				double *px = P + 4 * i, *py = P + 12 + 4 * i, *pc = P + 24 + 5 * i;
				px[3] = -rb[0]; px[2] = ra[0] - rb[1]; px[1] = ra[1] - rb[2]; px[0] = ra[2];
				py[3] = -rb[3]; py[2] = ra[3] - rb[4]; py[1] = ra[4] - rb[5]; py[0] = ra[5];
				pc[4] = -rb[6]; pc[3] = ra[6] - rb[7]; pc[2] = ra[7] - rb[8]; pc[1] = ra[8] - rb[9]; pc[0] = ra[9];
			}
			sync();
			// ---------------------------------------------------------------- degree-10 determinant ----
			double *T10 = S + OFF_EET; // [6][11]
			if (li < 6)
			{
				const int perm[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {0, 2, 1}, {1, 0, 2}, {2, 1, 0}};
				double t6[7], t10[11];
				gcr_e5::conv(P + 4 * perm[li][0], 3, P + 12 + 4 * perm[li][1], 3, t6);
				gcr_e5::conv(t6, 6, P + 24 + 5 * perm[li][2], 4, t10);
				for (int i = 0; i <= 10; ++i) T10[11 * li + i] = t10[i];
			}
			sync();
			for (int i = li; i <= 10; i += 10)
			{
				double c = 0.0;
				for (int p = 0; p < 6; ++p)
					c += (p < 3 ? 1.0 : -1.0) * T10[11 * p + i];
				poly[i] = c;
			}
			sync();
			// ---------------------------------------------------------------- real roots: Sturm isolation ----
			go = poly[n] != 0.0;
			double R = 0.0;
			if (go)
			{
				const double inv = 1.0 / poly[n];
				sync(); // everybody has read the leading coefficient before it becomes 1
				f[li] = poly[li] * inv;
				if (li == 0) f[n] = 1.0;
				sync();
				double mx = 0.0;
				for (int i = 0; i < n; ++i)
					mx = fmax(mx, fabs(f[i]));
				df[li] = f[li + 1] * (double)(li + 1);
				R = 1.0 + mx;
				// ---- chain ----
				double *chain = S + OFF_CHAIN; // p[11][11]  (the 10 x 20 system is dead)
				int *deg = reinterpret_cast<int *>(S + OFF_DEG);
				double *r = S + OFF_ABS;
				sync(); // df complete, everybody is past the reads of C
				for (int i = li; i <= n; i += 10)
					chain[i] = f[i];
				chain[11 + li] = df[li];
				if (li == 0) { chain[11 + 10] = 0.0; deg[0] = n; deg[1] = n - 1; }
				sync();
				int len = 2;
				while (deg[len - 1] > 0 && len <= 10)
				{
					const double *a = chain + 11 * (len - 2);
					const double *b = chain + 11 * (len - 1);
					const int da = deg[len - 2], db = deg[len - 1];
					for (int i = li; i <= da; i += 10)
						r[i] = a[i];
					sync();
					for (int k = da; k >= db; --k)
					{
						const double q = r[k] / b[db];
						sync(); // q is read before r[k] is cleared
						if (li < db)
							r[k - db + li] -= q * b[li];
						else if (li == db)
							r[k] = 0.0;
						sync();
					}
					int dr = db - 1;
					while (dr >= 0 && r[dr] == 0.0)
						--dr;
					if (dr < 0)
						break;
					double mxr = 0.0;
					for (int i = 0; i <= dr; ++i)
						mxr = fmax(mxr, fabs(r[i]));
					const double s = -1.0 / mxr;
					for (int i = li; i <= n; i += 10) // rows are zero-padded up to degree 10 (horner_padded)
						chain[11 * len + i] = i <= dr ? r[i] * s : 0.0;
					if (li == 0)
						deg[len] = dr;
					++len;
					sync();
				}
				if (li == 0)
					deg[11] = len;
				sync();
				// ---- isolation: depth-first bisection; the interval stack is written by lane 0 ----
				// The loop runs in lockstep for every group of the warp (one vote per iteration) so that the bisection
				// step -- the expensive part -- is issued once for all groups that need one, not once per group.
				double *stk = S + OFF_STACK;
				int *stki = reinterpret_cast<int *>(S + OFF_STACK + 2 * STACK_CAP);
				const double c10 = len == 11 ? chain[110] : 0.0;
				const int s_lo = sign_changes(li, chain, len, c10, sync, -R);
				const int s_hi = sign_changes(li, chain, len, c10, sync, R);
				if (s_lo - s_hi > 0)
				{
					if (li == 0)
					{
						stk[0] = -R; stk[1] = R;
						stki[0] = s_lo | (s_hi << 8);
					}
					top = 1;
				}
				sync();
			}
		}
		// (groups without a polynomial take part in the votes with nothing to do)
		{
			{
				const double *chain = S + OFF_CHAIN;
				const int *deg = reinterpret_cast<const int *>(S + OFF_DEG);
				double *stk = S + OFF_STACK;
				int *stki = reinterpret_cast<int *>(S + OFF_STACK + 2 * STACK_CAP);
				double *iv = S + OFF_IV;
				const int len = go ? deg[11] : 0;
				const double c10 = len == 11 ? chain[110] : 0.0;
				while (sync.any(top > 0 && n_roots < n))
				{
					bool bisect = false;
					double a = 0.0, b = 0.0;
					int sa = 0, sb = 0, depth = 0;
					if (top > 0 && n_roots < n)
					{
						--top;
						a = stk[2 * top]; b = stk[2 * top + 1];
						const int packed = stki[top];
						sa = packed & 255; sb = (packed >> 8) & 255; depth = packed >> 16;
						const int cnt = sa - sb;
						if (cnt > 0)
						{
							bool isolated = false;
							if (cnt == 1)
							{
								const double fa = gcr_e5::horner(f, n, a), fb = gcr_e5::horner(f, n, b);
								if ((fa < 0.0) != (fb < 0.0) || fa == 0.0 || fb == 0.0)
								{
									if (li == 0) { iv[2 * n_roots] = a; iv[2 * n_roots + 1] = b; }
									++n_roots;
									isolated = true;
								}
							}
							bisect = !isolated && !(depth >= 60 || !(b - a > 1e-15 * (fabs(a) + fabs(b))));
						}
					}
					if (bisect)
					{
						// every lane has consumed the popped entry (its values feed the ballots below) before lane 0
						// overwrites the slot
						const double c = 0.5 * (a + b);
						const int sc = sign_changes(li, chain, len, c10, sync, c);
						if (li == 0)
						{
							stk[2 * top] = c; stk[2 * top + 1] = b;
							stki[top] = sc | (sb << 8) | ((depth + 1) << 16);
							stk[2 * top + 2] = a; stk[2 * top + 3] = c;
							stki[top + 1] = sa | (sc << 8) | ((depth + 1) << 16);
						}
						top += 2;
						sync();
					}
				}
				sync();
			}
		}
		if (have)
		{
			// ---------------------------------------------------------------- one root per lane: polish, back-substitute ----
			double *iv = S + OFF_IV;
			double e[9];
			bool finite = false;
			if (li < n_roots)
			{
				const double z = gcr_e5::polish(f, df, n, iv[2 * li], iv[2 * li + 1]);
				const double *px = P, *py = P + 12, *pc = P + 24;
				double B[3][2], b[3];
				for (int i = 0; i < 3; ++i)
				{
					B[i][0] = gcr_e5::horner(px + 4 * i, 3, z);
					B[i][1] = gcr_e5::horner(py + 4 * i, 3, z);
					b[i] = gcr_e5::horner(pc + 5 * i, 4, z);
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
							const double fct = M[i][k] / M[k][k];
							for (int j = k; j < 3; ++j) M[i][j] -= fct * M[k][j];
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
				finite = true;
				for (int j = 0; j < 9; ++j)
				{
					e[j] = (N[j] * x + N[9 + j] * y + N[18 + j] * z + N[27 + j]) * inv;
					if (!isfinite(e[j])) finite = false;
				}
			}
			sync(); // every lane has consumed its interval before the flags overwrite that area
			int *fin = reinterpret_cast<int *>(S + OFF_IV);
			fin[li] = finite ? 1 : 0;
			sync();
			int slot = 0;
			for (int q = 0; q < 10; ++q)
			{
				if (q < li) slot += fin[q];
				n_models += fin[q];
			}
			if (finite)
			{
				double *E = E_out + 9 * slot;
				for (int r = 0; r < 3; ++r)
					for (int cc = 0; cc < 3; ++cc)
						E[3 * r + cc] = e[3 * cc + r];
			}
		}
		// ---------------------------------------------------------------- orientation test per model, unused slots ----
		// (essential_estimator.h:165-171, 338-401, on the five sample points)
		sync();
		int *okf = reinterpret_cast<int *>(S + OFF_IV) + 10;
		{
			bool ok = false;
			double *E = E_out + 9 * li; // now lane = slot
			if (li < n_models)
			{
				struct Pts { const double *p; __device__ __forceinline__ double operator()(int i, int c) const { return p[4 * i + c]; } };
				const Pts pt{PT};
				ok = gcr_f7::orientation_valid(E, pt, 5);
			}
			// (a slot that fails keeps the candidate written above: slots with flag 0 are never read)
			valid_out[li] = ok ? 1 : 0;
			okf[li] = ok ? 1 : 0;
		}
		sync();
		int n_ok = 0;
		for (int q = 0; q < 10; ++q)
			n_ok += okf[q];
		return n_ok;
	}
}
