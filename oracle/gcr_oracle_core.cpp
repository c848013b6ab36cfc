// oracle/gcr_oracle_core.cpp -- TEST INFRASTRUCTURE ONLY (see gcr_oracle.h).
// Residuals, MSAC / MAGSAC++ scoring, counter-based samplers, PROSAC growth function, 4-point homography,
// grid neighbourhood and the termination formula, restated from the reference (citations in gcr_oracle.h and
// beside each function).  Compile with -ffp-contract=off.
#include "gcr_oracle.h"
#include "linalg.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

extern "C" {

// ---------------------------------------------------------------------------------------------- residuals ----
// homography_estimator.h:181-199
double gcro_sqres_h(const double *s, const double *H)
{
	const double x1 = s[0], y1 = s[1], x2 = s[2], y2 = s[3];
	const double t1 = H[0] * x1 + H[1] * y1 + H[2];
	const double t2 = H[3] * x1 + H[4] * y1 + H[5];
	const double t3 = H[6] * x1 + H[7] * y1 + H[8];
	const double d1 = x2 - (t1 / t3);
	const double d2 = y2 - (t2 / t3);
	return d1 * d1 + d2 * d2;
}

// fundamental_estimator.h:195-222
double gcro_sqres_sampson(const double *s, const double *F)
{
	const double x1 = s[0], y1 = s[1], x2 = s[2], y2 = s[3];
	const double e11 = F[0], e12 = F[1], e13 = F[2], e21 = F[3], e22 = F[4], e23 = F[5];
	double rxc = e11 * x2 + e21 * y2 + F[6];
	double ryc = e12 * x2 + e22 * y2 + F[7];
	double rwc = e13 * x2 + e23 * y2 + F[8];
	double r = (x1 * rxc + y1 * ryc + rwc);
	double rx = e11 * x1 + e12 * y1 + e13;
	double ry = e21 * x1 + e22 * y1 + e23;
	return r * r / (rxc * rxc + ryc * ryc + rx * rx + ry * ry);
}

// fundamental_estimator.h:225-253
double gcro_sqres_symepi(const double *s, const double *F)
{
	const double x1 = s[0], y1 = s[1], x2 = s[2], y2 = s[3];
	const double e11 = F[0], e12 = F[1], e13 = F[2], e21 = F[3], e22 = F[4], e23 = F[5];
	const double rxc = e11 * x2 + e21 * y2 + F[6];
	const double ryc = e12 * x2 + e22 * y2 + F[7];
	const double rwc = e13 * x2 + e23 * y2 + F[8];
	const double r = (x1 * rxc + y1 * ryc + rwc);
	const double rx = e11 * x1 + e12 * y1 + e13;
	const double ry = e21 * x1 + e22 * y1 + e23;
	const double a = rxc * rxc + ryc * ryc;
	const double b = rx * rx + ry * ry;
	return r * r * (a + b) / (a * b);
}

// perspective_n_point_estimator.h:148-184
double gcro_sqres_pnp(const double *s, const double *P)
{
	const double u = s[0], v = s[1], x = s[2], y = s[3], z = s[4];
	const double px = P[0] * x + P[1] * y + P[2] * z + P[3];
	const double py = P[4] * x + P[5] * y + P[6] * z + P[7];
	const double pz = P[8] * x + P[9] * y + P[10] * z + P[11];
	const double pu = px / pz, pv = py / pz;
	const double du = pu - u, dv = pv - v;
	return du * du + dv * dv;
}

double gcro_sqres(int problem, const double *pt, const double *model)
{
	switch (problem)
	{
	case GCRO_H: return gcro_sqres_h(pt, model);
	case GCRO_F:
	case GCRO_E: return gcro_sqres_sampson(pt, model);
	default: return gcro_sqres_pnp(pt, model);
	}
}

int gcro_model_len(int problem) { return problem == GCRO_PNP ? 12 : 9; }
int gcro_point_cols(int problem) { return problem == GCRO_PNP ? 5 : 4; }

void gcro_residuals(int problem, const double *pts, int n, const double *model, double *out)
{
	const int cols = gcro_point_cols(problem);
	for (int i = 0; i < n; ++i)
		out[i] = gcro_sqres(problem, pts + (size_t)i * cols, model);
}

// ------------------------------------------------------------------------------------------------ scoring ----
// gamma_values.cpp:2012: stored_gamma_values[k] == Gamma(1.5, k / 1000) (upper incomplete gamma).  Closed form
// Gamma(3/2, x) = sqrt(pi)/2 * erfc(sqrt(x)) + sqrt(x) * exp(-x).  Checked entry-by-entry against the reference
// table in tests (test_oracle_ref.py, 1e-15 abs).
double gcro_gamma_lut(int k)
{
	const double x = k / 1000.0;
	const double sx = std::sqrt(x);
	return 0.88622692545275801365 * std::erfc(sx) + sx * std::exp(-x);
}

// scoring_function.h:137-233.  The early-exit test at :182-183 can never fire, so it is omitted.
void gcro_score_msac(int problem, const double *pts, int n, const double *model, double thr, double *score,
                     uint64_t *n_inl, uint32_t *inl_idx)
{
	const int cols = gcro_point_cols(problem);
	const double truncated_threshold = 3.0 / 2.0 * thr;                                  // GCRANSAC.h:253
	const double squared_truncated_threshold = truncated_threshold * truncated_threshold; // GCRANSAC.h:254
	double value = 0.0;
	uint64_t inliers = 0;
	for (int i = 0; i < n; ++i)
	{
		const double r2 = gcro_sqres(problem, pts + (size_t)i * cols, model);
		if (r2 < squared_truncated_threshold)
		{
			if (inl_idx)
				inl_idx[inliers] = (uint32_t)i;
			++inliers;
			value -= r2;
		}
	}
	if (inliers == 0)
	{
		*score = 0.0;
		*n_inl = 0;
		return;
	}
	*score = (value + inliers * squared_truncated_threshold) / squared_truncated_threshold; // :227-229
	*n_inl = inliers;
}

// scoring_function.h:246-348
void gcro_score_magsac(int problem, const double *pts, int n, const double *model, double thr, double *score,
                       uint64_t *n_inl, uint32_t *inl_idx)
{
	static std::vector<double> lut;
	if (lut.empty())
	{
		lut.resize(38683);
		for (int k = 0; k < 38683; ++k)
			lut[k] = gcro_gamma_lut(k);
	}
	const int cols = gcro_point_cols(problem);
	const unsigned int stored_gamma_number = 36000;      // gamma_values.cpp:3
	const double precision_of_stored_gammas = 1e4;       // gamma_values.cpp:5
	const double dof_minus_one_per_two = (4 - 1.0) / 2.0;
	const double C = 0.25;
	const double two_ad_dof = std::pow(2.0, dof_minus_one_per_two);
	const double C_times_two_ad_dof = C * two_ad_dof;
	const double gamma_value = std::tgamma(dof_minus_one_per_two);
	const double gamma_k = 0.0036572608340910764;
	const double gamma_difference = gamma_value - gamma_k;
	const double increasedThreshold = thr;
	const double squared_sigma_max_2 = increasedThreshold * increasedThreshold * 2.0;
	const double one_over_sigma = C_times_two_ad_dof / increasedThreshold;
	const double weight_zero = one_over_sigma * gamma_difference;

	double value = 0.0;
	uint64_t inliers = 0;
	for (int i = 0; i < n; ++i)
	{
		const double residual = std::sqrt(gcro_sqres(problem, pts + (size_t)i * cols, model));
		if (residual > increasedThreshold)
			continue;
		double weight = 0.0;
		if (residual < std::numeric_limits<double>::epsilon())
			weight = weight_zero;
		else
		{
			const double squared_residual = residual * residual;
			size_t x = (size_t)std::round(precision_of_stored_gammas * squared_residual / squared_sigma_max_2);
			if (stored_gamma_number < x)
				x = stored_gamma_number;
			weight = one_over_sigma * (lut[x] - gamma_k);
		}
		value += weight / weight_zero;
		if (residual > thr)
			continue;
		if (inl_idx)
			inl_idx[inliers] = (uint32_t)i;
		++inliers;
	}
	*score = value;
	*n_inl = inliers;
}

void gcro_score_batch(int problem, int scoring, const double *pts, int n, const double *models, int k, double thr,
                      double *scores, uint64_t *n_inl)
{
	const int ml = gcro_model_len(problem);
	for (int j = 0; j < k; ++j)
	{
		if (scoring == GCRO_MSAC)
			gcro_score_msac(problem, pts, n, models + (size_t)j * ml, thr, scores + j, n_inl + j, nullptr);
		else
			gcro_score_magsac(problem, pts, n, models + (size_t)j * ml, thr, scores + j, n_inl + j, nullptr);
	}
}

// ----------------------------------------------------------------------------------------------- samplers ----
static inline void mulhilo32(uint32_t a, uint32_t b, uint32_t &hi, uint32_t &lo)
{
	const uint64_t p = (uint64_t)a * b;
	hi = (uint32_t)(p >> 32);
	lo = (uint32_t)p;
}

// Philox4x32-10 (Salmon et al., SC'11), the published constants.
void gcro_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4])
{
	uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3];
	uint32_t k0 = key[0], k1 = key[1];
	for (int round = 0; round < 10; ++round)
	{
		uint32_t hi0, lo0, hi1, lo1;
		mulhilo32(0xD2511F53u, c0, hi0, lo0);
		mulhilo32(0xCD9E8D57u, c2, hi1, lo1);
		const uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
		c0 = n0; c1 = n1; c2 = n2; c3 = n3;
		k0 += 0x9E3779B9u;
		k1 += 0xBB67AE85u;
	}
	out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

namespace
{
	// Stream of u32 draws for one sample: Philox counter = (index_lo, index_hi, block, stream), key = seed.
	struct DrawStream
	{
		uint32_t ctr[4], key[2], buf[4];
		int pos;
		DrawStream(uint64_t seed, uint64_t index, uint32_t stream)
		{
			ctr[0] = (uint32_t)index; ctr[1] = (uint32_t)(index >> 32); ctr[2] = 0; ctr[3] = stream;
			key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
			pos = 4;
		}
		uint32_t next()
		{
			if (pos == 4)
			{
				gcro_philox4x32(ctr, key, buf);
				++ctr[2];
				pos = 0;
			}
			return buf[pos++];
		}
		// Unbiased integer in [0, range) (Lemire's multiply-shift with rejection).
		uint32_t bounded(uint32_t range)
		{
			for (;;)
			{
				const uint64_t p = (uint64_t)next() * range;
				const uint32_t lo = (uint32_t)p;
				if (lo < range)
				{
					const uint32_t thresh = (uint32_t)(0u - range) % range;
					if (lo < thresh)
						continue;
				}
				return (uint32_t)(p >> 32);
			}
		}
		// uniform_random_generator.h:90-103: draw-and-reject until `count` distinct values in [0, range).
		void unique_set(uint32_t *out, int count, uint32_t range)
		{
			for (int i = 0; i < count; ++i)
			{
				bool dup;
				do
				{
					out[i] = bounded(range);
					dup = false;
					for (int j = i - 1; j >= 0; --j)
						if (out[i] == out[j]) { dup = true; break; }
				} while (dup);
			}
		}
	};
}

// uniform_sampler.h:81-99 (pool = identity, GCRANSAC.h:258-260)
void gcro_sample_uniform(uint64_t seed, uint64_t sample_index, uint32_t stream, uint32_t n, int m, uint32_t *out)
{
	DrawStream s(seed, sample_index, stream);
	s.unique_set(out, m, n);
}

// prosac_sampler.h:129-160
void gcro_prosac_growth(uint32_t n, int m, uint64_t convergence_iters, uint64_t *growth)
{
	const size_t sample_size = (size_t)m, point_number = n;
	double T_n = (double)convergence_iters;
	for (size_t i = 0; i < sample_size; i++)
		T_n *= static_cast<double>(sample_size - i) / (point_number - i);
	size_t T_n_prime = 1;
	for (size_t i = 0; i < point_number; ++i)
	{
		if (i + 1 <= sample_size)
		{
			growth[i] = T_n_prime;
			continue;
		}
		double Tn_plus1 = static_cast<double>(i + 1) * T_n / (i + 1 - sample_size);
		growth[i] = T_n_prime + (unsigned int)std::ceil(Tn_plus1 - T_n);
		T_n = Tn_plus1;
		T_n_prime = growth[i];
	}
}

// prosac_sampler.h:65-88: subset_size starts at m and grows by one whenever the (incremented) sample number exceeds
// growth[subset_size-1].  growth is strictly increasing past index m-1, so the recurrence has the closed form
// n_k = min{ n >= m : growth[n-1] >= k }, clamped to point_number.
uint32_t gcro_prosac_subset_size(const uint64_t *growth, uint32_t n, int m, uint64_t k)
{
	uint32_t lo = (uint32_t)m, hi = n; // answer in [m, n]
	if (lo >= hi)
		return n < (uint32_t)m ? (uint32_t)m : hi;
	while (lo < hi)
	{
		const uint32_t mid = lo + (hi - lo) / 2;
		if (growth[mid - 1] >= k)
			hi = mid;
		else
			lo = mid + 1;
	}
	return lo;
}

// prosac_sampler.h:207-236.  k = 1 reproduces the reference's first-draw quirk (range [0, m-1] for the m-1 free
// indices while the forced index is also m-1).
void gcro_sample_prosac(uint64_t seed, uint64_t k, uint32_t n, int m, const uint64_t *growth,
                        uint64_t convergence_iters, uint32_t *out)
{
	DrawStream s(seed, k, 1u);
	if (k > convergence_iters)
	{
		s.unique_set(out, m, n);
		return;
	}
	const uint32_t subset = gcro_prosac_subset_size(growth, n, m, k);
	const uint32_t range = (k == 1) ? (uint32_t)m : subset - 1;
	s.unique_set(out, m - 1, range);
	out[m - 1] = subset - 1;
}

// importance_sampler.h:76-78: std::discrete_distribution<int>(begin(probabilities), end(probabilities)); libstdc++'s
// param_type::_M_initialize: sum = accumulate, p_i /= sum, cp = partial_sum(p), cp.back() = 1.
void gcro_importance_cdf(const double *probabilities, uint32_t n, double *cdf)
{
	double sum = 0.0;
	for (uint32_t i = 0; i < n; ++i)
		sum += probabilities[i];
	double value = 0.0;
	for (uint32_t i = 0; i < n; ++i)
	{
		const double p = probabilities[i] / sum;
		value = i == 0 ? p : value + p;
		cdf[i] = value;
	}
	if (n)
		cdf[n - 1] = 1.0;
}

// importance_sampler.h:134-135: subset_[i] = multinomial_distribution(generator), i.e. independent draws.  libstdc++'s
// operator(): p = generate_canonical<double, 53>(urng) -- for a 32-bit generator two calls, (w0 + w1 * 2^32) / 2^64
// in FP64, nextafter(1, 0) if that rounds to 1 -- then lower_bound over the cumulative table.
void gcro_sample_importance(uint64_t seed, uint64_t sample_index, uint32_t n, int m, const double *cdf, uint32_t *out)
{
	DrawStream s(seed, sample_index, 3u);
	for (int j = 0; j < m; ++j)
	{
		const uint32_t w0 = s.next(), w1 = s.next();
		double p = ((double)w0 + (double)w1 * 4294967296.0) / 18446744073709551616.0;
		if (p >= 1.0)
			p = std::nextafter(1.0, 0.0);
		uint32_t lo = 0, hi = n;
		while (lo < hi)
		{
			const uint32_t mid = lo + (hi - lo) / 2;
			if (cdf[mid] < p)
				lo = mid + 1;
			else
				hi = mid;
		}
		out[j] = lo;
	}
}

// --------------------------------------------------------------------------------------------- homography ----
// homography_estimator.h:312-322 ("cross_product" with st_ = 1) and :346-381
static inline void h_cross(double *r, const double *v1, const double *v2)
{
	r[0] = v1[1] - v2[1];
	r[1] = v2[0] - v1[0];
	r[2] = v1[0] * v2[1] - v1[1] * v2[0];
}

int gcro_h_valid_sample(const double *pts, const uint32_t *sample)
{
	const double *a = pts + 4 * (size_t)sample[0], *b = pts + 4 * (size_t)sample[1],
	             *c = pts + 4 * (size_t)sample[2], *d = pts + 4 * (size_t)sample[3];
	double p[3], q[3];
	h_cross(p, a, b);
	h_cross(q, a + 2, b + 2);
	if ((p[0] * c[0] + p[1] * c[1] + p[2]) * (q[0] * c[2] + q[1] * c[3] + q[2]) < 0)
		return 0;
	if ((p[0] * d[0] + p[1] * d[1] + p[2]) * (q[0] * d[2] + q[1] * d[3] + q[2]) < 0)
		return 0;
	h_cross(p, c, d);
	h_cross(q, c + 2, d + 2);
	if ((p[0] * a[0] + p[1] * a[1] + p[2]) * (q[0] * a[2] + q[1] * a[3] + q[2]) < 0)
		return 0;
	if ((p[0] * b[0] + p[1] * b[1] + p[2]) * (q[0] * b[2] + q[1] * b[3] + q[2]) < 0)
		return 0;
	return 1;
}

// math_utils.h:45-87 with _Size = 8 on an 8x9 row-major [A|b]
static void gauss_elimination8(double M[8][9], double *result)
{
	const int S = 8;
	for (int i = 0; i < S; i++)
		for (int k = i + 1; k < S; k++)
			if (std::fabs(M[i][i]) < std::fabs(M[k][i]))
				for (int j = 0; j <= S; j++)
					std::swap(M[i][j], M[k][j]);
	for (int i = 0; i < S - 1; i++)
		for (int k = i + 1; k < S; k++)
		{
			const double temp = M[k][i] / M[i][i];
			for (int j = 0; j <= S; j++)
				M[k][j] = M[k][j] - temp * M[i][j];
		}
	for (int i = S - 1; i >= 0; i--)
	{
		result[i] = M[i][S];
		for (int j = i + 1; j < S; j++)
			result[i] = result[i] - M[i][j] * result[j];
		result[i] = result[i] / M[i][i];
	}
}

// solver_homography_four_point.h:109-190 (weights_ == nullptr on the minimal path)
int gcro_solve_h4(const double *pts, int cols, const uint32_t *sample, double *H9)
{
	double M[8][9];
	int row = 0;
	for (int i = 0; i < 4; ++i)
	{
		const size_t idx = sample ? sample[i] : (size_t)i;
		const double *p = pts + idx * cols;
		const double x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
		const double weight = 1.0;
		const double mwx1 = -weight * x1, mwy1 = -weight * y1, wx2 = weight * x2, wy2 = weight * y2;
		M[row][0] = mwx1; M[row][1] = mwy1; M[row][2] = -weight; M[row][3] = 0; M[row][4] = 0; M[row][5] = 0;
		M[row][6] = wx2 * x1; M[row][7] = wx2 * y1; M[row][8] = -wx2;
		++row;
		M[row][0] = 0; M[row][1] = 0; M[row][2] = 0; M[row][3] = mwx1; M[row][4] = mwy1; M[row][5] = -weight;
		M[row][6] = wy2 * x1; M[row][7] = wy2 * y1; M[row][8] = -wy2;
		++row;
	}
	double h[8];
	gauss_elimination8(M, h);
	for (int i = 0; i < 8; ++i)
		if (std::isnan(h[i]))
			return 0;
	for (int i = 0; i < 8; ++i)
		H9[i] = h[i];
	H9[8] = 1.0;
	return 1;
}

// homography_estimator.h:326-342
int gcro_h_valid_model(const double *H)
{
	const double det = gcro::det3(H);
	return std::fabs(det) < 1e-2 ? 0 : 1;
}

// homography_estimator.h:213-310 (identical arithmetic in fundamental_estimator.h:637-734)
void gcro_normalize_points(const double *pts, int cols, const uint32_t *sample, int n, double *out, double *T1,
                           double *T2)
{
	double ms[2] = {0, 0}, md[2] = {0, 0};
	for (int i = 0; i < n; ++i)
	{
		const double *d = pts + (size_t)cols * sample[i];
		ms[0] += d[0]; ms[1] += d[1]; md[0] += d[2]; md[1] += d[3];
	}
	ms[0] /= n; ms[1] /= n; md[0] /= n; md[1] /= n;
	double ads = 0.0, add = 0.0;
	for (int i = 0; i < n; ++i)
	{
		const double *d = pts + (size_t)cols * sample[i];
		const double dx1 = ms[0] - d[0], dy1 = ms[1] - d[1], dx2 = md[0] - d[2], dy2 = md[1] - d[3];
		ads += std::sqrt(dx1 * dx1 + dy1 * dy1);
		add += std::sqrt(dx2 * dx2 + dy2 * dy2);
	}
	ads /= n;
	add /= n;
	const double rs = M_SQRT2 / ads, rd = M_SQRT2 / add;
	for (int i = 0; i < n; ++i)
	{
		const double *d = pts + (size_t)cols * sample[i];
		*out++ = (d[0] - ms[0]) * rs;
		*out++ = (d[1] - ms[1]) * rs;
		*out++ = (d[2] - md[0]) * rd;
		*out++ = (d[3] - md[1]) * rd;
		for (int c = 4; c < cols; ++c)
			*out++ = d[c];
	}
	const double t1[9] = {rs, 0, -rs * ms[0], 0, rs, -rs * ms[1], 0, 0, 1};
	const double t2[9] = {rd, 0, -rd * md[0], 0, rd, -rd * md[1], 0, 0, 1};
	std::memcpy(T1, t1, sizeof(t1));
	std::memcpy(T2, t2, sizeof(t2));
}

// homography_estimator.h:136-173 + solver_homography_four_point.h:192-264 (n > 4) / :109-190 (n == 4).
// NOTE: the solver is handed the *normalised copy* with sample_ == nullptr, so it reads
// weights_[i] by POSITION i while GCRANSAC.h:657-668 filled the array by original point index.  Kept literally.
int gcro_refit_h(const double *pts, const uint32_t *sample, int n, const double *weights, double *H9)
{
	if (n < 4)
		return 0;
	std::vector<double> np((size_t)n * 4);
	double T1[9], T2[9];
	gcro_normalize_points(pts, 4, sample, n, np.data(), T1, T2);
	double Hn[9];
	if (n == 4)
	{
		// minimal path on the normalised points, honouring weights_[i] like :141-142
		double M[8][9];
		int row = 0;
		for (int i = 0; i < 4; ++i)
		{
			const double *p = np.data() + 4 * i;
			const double x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
			const double weight = weights ? weights[i] : 1.0;
			const double mwx1 = -weight * x1, mwy1 = -weight * y1, wx2 = weight * x2, wy2 = weight * y2;
			M[row][0] = mwx1; M[row][1] = mwy1; M[row][2] = -weight; M[row][3] = 0; M[row][4] = 0; M[row][5] = 0;
			M[row][6] = wx2 * x1; M[row][7] = wx2 * y1; M[row][8] = -wx2;
			++row;
			M[row][0] = 0; M[row][1] = 0; M[row][2] = 0; M[row][3] = mwx1; M[row][4] = mwy1; M[row][5] = -weight;
			M[row][6] = wy2 * x1; M[row][7] = wy2 * y1; M[row][8] = -wy2;
			++row;
		}
		double h[8];
		gauss_elimination8(M, h);
		for (int i = 0; i < 8; ++i)
			if (std::isnan(h[i]))
				return 0;
		std::memcpy(Hn, h, sizeof(h));
		Hn[8] = 1.0;
	}
	else
	{
		const int rows = 2 * n;
		std::vector<double> A((size_t)rows * 8), b(rows);
		int row = 0;
		for (int i = 0; i < n; ++i)
		{
			const double *p = np.data() + 4 * (size_t)i;
			const double x1 = p[0], y1 = p[1], x2 = p[2], y2 = p[3];
			const double weight = weights ? weights[i] : 1.0;
			const double mwx1 = -weight * x1, mwy1 = -weight * y1, wx2 = weight * x2, wy2 = weight * y2;
			double *r = A.data() + (size_t)row * 8;
			r[0] = mwx1; r[1] = mwy1; r[2] = -weight; r[3] = 0; r[4] = 0; r[5] = 0; r[6] = wx2 * x1; r[7] = wx2 * y1;
			b[row] = -wx2;
			++row;
			r = A.data() + (size_t)row * 8;
			r[0] = 0; r[1] = 0; r[2] = 0; r[3] = mwx1; r[4] = mwy1; r[5] = -weight; r[6] = wy2 * x1; r[7] = wy2 * y1;
			b[row] = -wy2;
			++row;
		}
		double h[8];
		gcro::colpiv_qr_solve(A.data(), rows, 8, b.data(), h); // Eigen colPivHouseholderQr().solve (:255-256)
		std::memcpy(Hn, h, sizeof(h));
		Hn[8] = 1.0;
	}
	// :168-171  H = T2^-1 * Hn * T1
	double T2i[9], tmp[9];
	gcro::inv3(T2, T2i);
	gcro::mul3(T2i, Hn, tmp);
	gcro::mul3(tmp, T1, H9);
	return 1;
}

// ------------------------------------------------------------------------------------------ neighbourhood ----
// grid_neighborhood_graph.h:229-292 (+ GridCell key :70-82)
uint64_t gcro_grid_build(const double *pts, int n, int dims, const double *cell_sizes, uint64_t cells_per_axis,
                         uint64_t *keys, uint32_t *sorted_idx, uint32_t *cell_offsets, uint32_t *n_cells)
{
	for (int row = 0; row < n; ++row)
	{
		uint64_t offset = 1, index = 0;
		for (int d = 0; d < dims; ++d)
		{
			const uint64_t idx = (uint64_t)(int64_t)std::floor(pts[(size_t)row * dims + d] / cell_sizes[d]);
			index += offset * idx;
			offset *= cells_per_axis;
		}
		keys[row] = index;
	}
	std::vector<uint32_t> order(n);
	std::iota(order.begin(), order.end(), 0u);
	std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
	uint64_t neighbor_number = 0;
	uint32_t cells = 0;
	int start = 0;
	for (int i = 0; i < n; ++i)
	{
		if (sorted_idx)
			sorted_idx[i] = order[i];
		if (i == 0 || keys[order[i]] != keys[order[i - 1]])
		{
			if (i > 0)
			{
				const uint64_t c = (uint64_t)(i - start);
				neighbor_number += c * (c - 1) / 2;
			}
			if (cell_offsets)
				cell_offsets[cells] = (uint32_t)i;
			++cells;
			start = i;
		}
	}
	if (n > 0)
	{
		const uint64_t c = (uint64_t)(n - start);
		neighbor_number += c * (c - 1) / 2;
	}
	if (cell_offsets)
		cell_offsets[cells] = (uint32_t)n;
	if (n_cells)
		*n_cells = cells;
	return neighbor_number;
}

// ------------------------------------------------------------------------------------------ fixed-K round ----
// GCRANSAC.h:297-471 without the adaptive stop and LO: sample -> isValidSample -> estimateModel -> getScore ->
// strict '<' best-so-far (first wins).  index = sample * max_models + position among the kept models.
void gcro_round(int problem, int scoring, const double *pts, int n, uint64_t seed, uint64_t first, int count,
                double thr, double *best_score, uint64_t *best_index, uint64_t *models_scored, double *best_model)
{
	const int ml = gcro_model_len(problem);
	const int m = problem == GCRO_H ? 4 : problem == GCRO_F ? 7 : problem == GCRO_E ? 5 : 3;
	const int maxsol = problem == GCRO_H ? 1 : problem == GCRO_F ? 3 : problem == GCRO_E ? 10 : 4;
	double best = 0.0;
	uint64_t best_idx = ~0ull, scored = 0;
	uint32_t s[8];
	double models[10 * 12];
	for (int i = 0; i < count; ++i)
	{
		gcro_sample_uniform(seed, first + (uint64_t)i, 0u, (uint32_t)n, m, s);
		int cnt = 0;
		if (problem == GCRO_H)
		{
			if (!gcro_h_valid_sample(pts, s))
				continue;
			cnt = gcro_solve_h4(pts, 4, s, models);
		}
		else if (problem == GCRO_F)
			cnt = gcro_estimate_f7(pts, s, models);
		else if (problem == GCRO_E)
			cnt = gcro_estimate_e5(pts, s, models);
		else
			cnt = gcro_solve_p3p(pts, s, models);
		for (int j = 0; j < cnt; ++j)
		{
			double sc;
			uint64_t ni;
			if (scoring == GCRO_MSAC)
				gcro_score_msac(problem, pts, n, models + j * ml, thr, &sc, &ni, nullptr);
			else
				gcro_score_magsac(problem, pts, n, models + j * ml, thr, &sc, &ni, nullptr);
			++scored;
			if (best < sc)
			{
				best = sc;
				best_idx = (uint64_t)i * maxsol + j;
				if (best_model)
					std::memcpy(best_model, models + j * ml, ml * sizeof(double));
			}
		}
	}
	*best_score = best;
	*best_index = best_idx;
	*models_scored = scored;
}

// ------------------------------------------------------------------------------------------- termination ----
// GCRANSAC.h:158-172
uint64_t gcro_iteration_number(uint64_t inliers, uint64_t n, int m, double log_probability)
{
	const double q = std::pow(static_cast<double>(inliers) / n, m);
	const double log2 = std::log(1 - q);
	if (std::fabs(log2) < std::numeric_limits<double>::epsilon())
		return std::numeric_limits<uint64_t>::max();
	const double iter = log_probability / log2;
	return static_cast<uint64_t>(iter) + 1;
}

} // extern "C"
