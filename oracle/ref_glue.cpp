// oracle/ref_glue.cpp -- TEST INFRASTRUCTURE ONLY.  Built ONLY where /root/reference exists (this container), by
// oracle/build_ref.sh, into oracle/_ref/libgcr_ref.so.  It #includes the reference's own headers UNMODIFIED from
// /root/reference/src/pygcransac/include (GCRANSAC.h, scoring_function.h + gamma_values.cpp, samplers, grid graph,
// BK max-flow) through the two container shims in oracle/shim/, and exposes them over a small C API so that
//   (1) the self-contained restatement in oracle/gcr_oracle*.cpp can be pinned against the real reference code, and
//   (2) golden fixtures (tests/golden/) can be generated from the reference itself (oracle/gen_golden.py).
// The reference's estimator classes need Eigen's decompositions, which are not to be had here, so the Estimator
// template parameter is an adapter over the restated solvers of oracle/gcr_oracle*.cpp (kind = "reference driver +
// restated estimators"; stated wherever results of this library are quoted).  oracle/ref_solvers_glue.cpp compiles those
// estimator classes verbatim over a stand-in for Eigen's API and shows that runs over them and over this adapter are equal
// (tests/test_reference_solver_sources.py).
#include <opencv2/core/core.hpp>

#include <cstdint>
#include <cstddef>
// the reference's random-number class redirected to the counter-based word stream (ProgressiveNapsacSampler and the two
// ProsacSamplers inside it run verbatim on it; the counter-based uniform / PROSAC adapters of this file do not go through it)
#include "ref_rng_redirect.h"

#include "GCRANSAC.h"
#include "neighborhood/grid_neighborhood_graph.h"
#include "samplers/importance_sampler.h"
#include "samplers/prosac_sampler.h"
#include "samplers/progressive_napsac_sampler.h"
#include "samplers/uniform_sampler.h"

#include "gcr_oracle.h"
#include "degensac_front.h"

// The reference's AR-Sampler (samplers/adaptive_reordering_sampler.h) takes its noise from C rand() -- process-global
// state.  The header is compiled UNMODIFIED, with the one call redirected to the counter-based word stream the product
// uses for the same update (stream 4 of (seed, sample index)); everything else -- the heap, the Beta update, the
// clamping -- is the reference's own code.  <cmath> / <cstdlib> / <math.h> first, as in the reference's own build, so
// that its unqualified abs() is the floating-point overload.
#include <cmath>
#include <cstdlib>
#include <math.h>
#include <tuple>
namespace gcr_ref_noise
{
	static uint32_t ctr[4], key[2], buf[4];
	static int pos = 4;
	inline void begin(uint64_t seed, uint64_t index, uint32_t stream)
	{
		ctr[0] = (uint32_t)index; ctr[1] = (uint32_t)(index >> 32); ctr[2] = 0; ctr[3] = stream;
		key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
		pos = 4;
	}
	inline int next_rand() // in [0, RAND_MAX] like rand()
	{
		if (pos == 4)
		{
			gcro_philox4x32(ctr, key, buf);
			++ctr[2];
			pos = 0;
		}
		return (int)(buf[pos++] >> 1);
	}
}
#define rand() gcr_ref_noise::next_rand()
#include "samplers/adaptive_reordering_sampler.h"
#undef rand

#include <cstdint>
#include <cstring>

using namespace gcransac;

namespace
{
	template <int PROBLEM>
	struct ProblemTraits;
	template <> struct ProblemTraits<GCRO_H> { static constexpr size_t m = 4, nonmin = 4, maxsol = 1, rows = 3, cols = 3; static constexpr bool weighting = true; };
	template <> struct ProblemTraits<GCRO_F> { static constexpr size_t m = 7, nonmin = 8, maxsol = 3, rows = 3, cols = 3; static constexpr bool weighting = true; };
	template <> struct ProblemTraits<GCRO_E> { static constexpr size_t m = 5, nonmin = 5, maxsol = 10, rows = 3, cols = 3; static constexpr bool weighting = true; };
	// DEGENSAC's inner estimator: plane-and-parallax minimal solver (2 points + H), 8-point non-minimal solver
	// (fundamental_estimator.h:538-541)
	constexpr int GCRO_FPP = 100;
	template <> struct ProblemTraits<GCRO_FPP> { static constexpr size_t m = 2, nonmin = 8, maxsol = 1, rows = 3, cols = 3; static constexpr bool weighting = true; };
	template <> struct ProblemTraits<GCRO_PNP> { static constexpr size_t m = 3, nonmin = 6, maxsol = 4, rows = 3, cols = 4; static constexpr bool weighting = false; };

	inline void to_flat(const Eigen::MatrixXd &d, double *out)
	{
		for (int i = 0; i < d.rows(); ++i)
			for (int j = 0; j < d.cols(); ++j)
				out[i * d.cols() + j] = d(i, j);
	}
	inline Model from_flat(const double *in, int rows, int cols)
	{
		Model m((Eigen::MatrixXd(rows, cols)));
		for (int i = 0; i < rows; ++i)
			for (int j = 0; j < cols; ++j)
				m.descriptor(i, j) = in[i * cols + j];
		return m;
	}

	template <int PROBLEM> class OracleEstimator;
	// fundamental_estimator.h:341-572 (defined below, after the counter-based sampler it needs)
	bool oracle_apply_degensac(const OracleEstimator<GCRO_F> &est, Model &model_, const cv::Mat &data_,
	                           const std::vector<size_t> &inliers_, const size_t *minimal_sample_, double threshold_,
	                           bool &model_updated_);

	// Estimator concept (estimators/estimator.h:48-108) over the restated solvers.
	template <int PROBLEM>
	class OracleEstimator
	{
	public:
		typedef ProblemTraits<PROBLEM> T;
		double f_min_ratio = PROBLEM == GCRO_FPP ? 0.0 : 0.5; // minimum_inlier_ratio_in_validity_check (:91; inner run: 0.0, :540)
		bool use_degensac = false;
		bool f_nonminimal_ba = false; // F: FundamentalMatrixBundleAdjustmentSolver instead of the bare 8-point solver (types.h:78-80)
		const double *pp_homography = nullptr; // GCRO_FPP: the plane homography (setHomography, :542)
		uint64_t seed = 1;                      // DEGENSAC's inner run draws from streams keyed by (seed, inner run number)
		mutable uint64_t degensac_runs = 0, degensac_updates = 0;

		static constexpr size_t sampleSize() { return T::m; }
		static constexpr size_t nonMinimalSampleSize() { return T::nonmin; }
		static constexpr size_t maximumMinimalSolutions() { return T::maxsol; }
		static constexpr bool isWeightingApplicable() { return T::weighting; }
		size_t inlierLimit() const { return 7 * sampleSize(); }

		double squaredResidual(const cv::Mat &point_, const Eigen::MatrixXd &descriptor_) const
		{
			// the shim's MatrixXd is contiguous row-major, which is the layout the restated residuals take
			const double *flat = descriptor_.rowmajor();
			const double *pt = reinterpret_cast<const double *>(point_.data);
			if constexpr (PROBLEM == GCRO_H)
				return gcro_sqres_h(pt, flat);
			else if constexpr (PROBLEM == GCRO_PNP)
				return gcro_sqres_pnp(pt, flat);
			else
				return gcro_sqres_sampson(pt, flat);
		}
		double squaredResidual(const cv::Mat &point_, const Model &model_) const { return squaredResidual(point_, model_.descriptor); }
		double residual(const cv::Mat &point_, const Eigen::MatrixXd &descriptor_) const { return sqrt(squaredResidual(point_, descriptor_)); }
		double residual(const cv::Mat &point_, const Model &model_) const { return residual(point_, model_.descriptor); }

		bool isValidSample(const cv::Mat &data_, const size_t *sample_) const
		{
			if constexpr (PROBLEM == GCRO_H)
			{
				uint32_t s[4] = {(uint32_t)sample_[0], (uint32_t)sample_[1], (uint32_t)sample_[2], (uint32_t)sample_[3]};
				return gcro_h_valid_sample(reinterpret_cast<const double *>(data_.data), s) != 0;
			}
			return true;
		}

		bool estimateModel(const cv::Mat &data_, const size_t *sample_, std::vector<Model> *models_) const
		{
			uint32_t s[8];
			for (size_t i = 0; i < T::m; ++i)
				s[i] = (uint32_t)sample_[i];
			const double *pts = reinterpret_cast<const double *>(data_.data);
			double out[10 * 12];
			int count = 0;
			if constexpr (PROBLEM == GCRO_H)
				count = gcro_solve_h4(pts, data_.cols, s, out);
			else if constexpr (PROBLEM == GCRO_F)
				count = gcro_estimate_f7(pts, s, out);
			else if constexpr (PROBLEM == GCRO_E)
				count = gcro_estimate_e5(pts, s, out);
			else if constexpr (PROBLEM == GCRO_FPP)
				count = gcro_estimate_f_plane_parallax(pts, s, pp_homography, out);
			else
				count = gcro_solve_p3p(pts, s, out);
			for (int i = 0; i < count; ++i)
				models_->emplace_back(from_flat(out + i * T::rows * T::cols, T::rows, T::cols));
			if constexpr (PROBLEM == GCRO_H)
				return count > 0; // solver_homography_four_point.h returns the solver's bool
			return models_->size() > 0;
		}

		bool estimateModelNonminimal(const cv::Mat &data_, const size_t *sample_, const size_t &sample_number_,
		                             std::vector<Model> *models_, const double *weights_ = nullptr) const
		{
			if (sample_number_ < nonMinimalSampleSize())
				return false;
			std::vector<uint32_t> s(sample_number_);
			for (size_t i = 0; i < sample_number_; ++i)
				s[i] = (uint32_t)sample_[i];
			const double *pts = reinterpret_cast<const double *>(data_.data);
			double out[10 * 12];
			int count = 0;
			if constexpr (PROBLEM == GCRO_H)
				count = gcro_refit_h(pts, s.data(), (int)sample_number_, weights_, out);
			else if constexpr (PROBLEM == GCRO_F)
				count = f_nonminimal_ba ? gcro_refit_f_ba(pts, s.data(), (int)sample_number_, weights_, out)
				                        : gcro_refit_f8(pts, s.data(), (int)sample_number_, weights_, out);
			else if constexpr (PROBLEM == GCRO_FPP)
				count = gcro_refit_f8(pts, s.data(), (int)sample_number_, weights_, out);
			else if constexpr (PROBLEM == GCRO_E)
				count = gcro_refit_e(pts, s.data(), (int)sample_number_, weights_, out);
			else // PnP: the stated non-minimal solver (gcr_oracle_pba.cpp); isWeightingApplicable() is false: never weighted
				count = gcro_refit_pnp(pts, s.data(), (int)sample_number_, nullptr, out);
			for (int i = 0; i < count; ++i)
				models_->emplace_back(from_flat(out + i * T::rows * T::cols, T::rows, T::cols));
			return count > 0;
		}

		bool isValidModel(Model &model, const cv::Mat &data, const std::vector<size_t> &inliers,
		                  const size_t *minimal_sample_, const double threshold_, bool &model_updated_) const
		{
			model_updated_ = false;
			double flat[12];
			to_flat(model.descriptor, flat);
			if constexpr (PROBLEM == GCRO_H)
				return gcro_h_valid_model(flat) != 0;
			else if constexpr (PROBLEM == GCRO_F || PROBLEM == GCRO_E || PROBLEM == GCRO_FPP)
			{
				std::vector<uint32_t> inl(inliers.size());
				for (size_t i = 0; i < inliers.size(); ++i)
					inl[i] = (uint32_t)inliers[i];
				const double ratio = PROBLEM == GCRO_E ? 0.1 : f_min_ratio;
				if (!gcro_f_valid_model_ratio(flat, reinterpret_cast<const double *>(data.data), inl.data(),
				                              (int)inl.size(), (int)T::m, threshold_, ratio))
					return false;
				if constexpr (PROBLEM == GCRO_F) // :325-333 (sampleSize() >= 7)
					if (use_degensac)
						return oracle_apply_degensac(*this, model, data, inliers, minimal_sample_, threshold_, model_updated_);
				return true;
			}
			return true;
		}
	};

	template <int PROBLEM>
	void score_with_reference(int scoring, const cv::Mat &points, const double *model_flat, double thr, double *score,
	                          uint64_t *n_inl, uint32_t *inl_idx)
	{
		typedef OracleEstimator<PROBLEM> Est;
		Est est;
		Model model = from_flat(model_flat, Est::T::rows, Est::T::cols);
		std::vector<size_t> inliers;
		Score s;
		if (scoring == GCRO_MSAC)
		{
			MSACScoringFunction<Est> f;
			const double tt = 3.0 / 2.0 * thr; // GCRANSAC.h:253-255
			f.initialize(tt * tt, points.rows);
			s = f.getScore(points, model, est, thr, inliers);
		}
		else
		{
			MAGSACScoringFunction<Est> f;
			f.initialize(0, points.rows);
			s = f.getScore(points, model, est, thr, inliers);
		}
		*score = s.value;
		*n_inl = s.inlier_number;
		if (inl_idx)
			for (size_t i = 0; i < inliers.size(); ++i)
				inl_idx[i] = (uint32_t)inliers[i];
	}

	class ProsacProbe : public sampler::ProsacSampler
	{
	public:
		ProsacProbe(const cv::Mat *c, size_t m, size_t conv) : sampler::ProsacSampler(c, m, conv) {}
		const std::vector<size_t> &growth() const { return growth_function; }
		size_t subsetSize() const { return subset_size; }
		size_t kth() const { return kth_sample_number; }
	};

	// The reference's ImportanceSampler (samplers/importance_sampler.h) draws `multinomial_distribution(mt19937 seeded
	// from random_device)`.  The probe reaches the distribution object the reference's own constructor built (:76-78)
	// and draws from it with the counter-based word stream instead: the std::discrete_distribution semantics (table
	// construction, generate_canonical, lower bound) are the library's, only the words differ.
	struct PhiloxWords
	{
		typedef uint32_t result_type;
		static constexpr result_type min() { return 0u; }
		static constexpr result_type max() { return 0xFFFFFFFFu; }
		uint32_t ctr[4], key[2], buf[4];
		int pos;
		PhiloxWords(uint64_t seed, uint64_t index, uint32_t stream) : pos(4)
		{
			ctr[0] = (uint32_t)index; ctr[1] = (uint32_t)(index >> 32); ctr[2] = 0; ctr[3] = stream;
			key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
		}
		result_type operator()()
		{
			if (pos == 4)
			{
				gcro_philox4x32(ctr, key, buf);
				++ctr[2];
				pos = 0;
			}
			return buf[pos++];
		}
	};
	class ImportanceProbe : public sampler::ImportanceSampler
	{
	public:
		ImportanceProbe(const cv::Mat *c, const std::vector<double> &p, size_t m) : sampler::ImportanceSampler(c, p, m) {}
		// importance_sampler.h:134-135 with the generator swapped
		void draw(uint64_t seed, uint64_t index, size_t m, size_t *subset)
		{
			PhiloxWords words(seed, index, 3u);
			for (size_t j = 0; j < m; ++j)
				subset[j] = (size_t)multinomial_distribution(words);
		}
	};
	std::vector<double> g_point_probabilities; // ref_set_point_probabilities -> sampler_id 3 / 4 of ref_run
	double g_sampler_variance = 0.1;           // ref_set_sampler_variance -> sampler_id 4

	// The reference's AdaptiveReorderingSampler verbatim behind the Sampler interface; sample() / update() number k take
	// their noise from stream 4 of (seed, k).  run() calls update() once after every sample() (GCRANSAC.h:304-474).
	class CounterARSampler : public sampler::AdaptiveReorderingSampler
	{
		uint64_t seed, next;

	public:
		CounterARSampler(const cv::Mat *c, const std::vector<double> &p, size_t m, double variance, uint64_t seed_)
		    : sampler::AdaptiveReorderingSampler(c, p, m, variance), seed(seed_), next(0)
		{
		}
		uint64_t drawn() const { return next; }
		bool sample(const std::vector<size_t> &pool_, size_t *const subset_, size_t sample_size_)
		{
			gcr_ref_noise::begin(seed, next++, 4u); // the update(s) that follow this sample
			return sampler::AdaptiveReorderingSampler::sample(pool_, subset_, sample_size_);
		}
	};
	double g_image_sizes[4] = {0, 0, 0, 0};    // ref_set_image_sizes (w1, h1, w2, h2) -> sampler_id 2
	const cv::Mat *g_pnapsac_points = nullptr; // E: the pixel coordinates (the run itself is on the K-normalised points)

	// The reference's ProgressiveNapsacSampler<4> verbatim (with the two ProsacSamplers inside it), constructed as the
	// drivers construct it (py/gcransac_python.cpp:4335-4343: layers 16 / 8 / 4 / 2, blending length 0.5); sample()
	// number k (from 1) draws from stream 5 of (seed, k) through the specialised generator members at the top of this file.
	class CounterPNapsacSampler : public sampler::ProgressiveNapsacSampler<4>
	{
		uint64_t seed, next;

	public:
		CounterPNapsacSampler(const cv::Mat *c, size_t m, const double *sizes, uint64_t seed_)
		    : sampler::ProgressiveNapsacSampler<4>(c, {16, 8, 4, 2}, m, {sizes[0], sizes[1], sizes[2], sizes[3]}, 0.5), seed(seed_), next(0)
		{
		}
		uint64_t drawn() const { return next; }
		bool sample(const std::vector<size_t> &pool_, size_t *const subset_, size_t sample_size_)
		{
			gcr_ref_words::begin(seed, ++next, 5u);
			return sampler::ProgressiveNapsacSampler<4>::sample(pool_, subset_, sample_size_);
		}
	};

	// py/gcransac_python.cpp:4350-4354
	std::vector<double> normalized_by_max(std::vector<double> p)
	{
		double max_prob = 0;
		for (const auto &prob : p)
			max_prob = MAX(max_prob, prob);
		for (auto &prob : p)
			prob /= max_prob;
		return p;
	}

	// exposes GCRANSAC<>::getIterationNumber (protected)
	struct IterProbe : public GCRANSAC<OracleEstimator<GCRO_H>, neighborhood::GridNeighborhoodGraph<4>>
	{
		size_t get(size_t a, size_t b, size_t c, double d) { return getIterationNumber(a, b, c, d); }
	};
}

extern "C" {

// scoring_function.h verbatim (MSAC :103-234, MAGSAC :237-349)
void ref_score(int problem, int scoring, const double *pts, int n, const double *model, double thr, double *score,
               uint64_t *n_inl, uint32_t *inl_idx)
{
	const int cols = gcro_point_cols(problem);
	cv::Mat points(n, cols, CV_64F, const_cast<double *>(pts));
	switch (problem)
	{
	case GCRO_H: score_with_reference<GCRO_H>(scoring, points, model, thr, score, n_inl, inl_idx); break;
	case GCRO_F: score_with_reference<GCRO_F>(scoring, points, model, thr, score, n_inl, inl_idx); break;
	case GCRO_E: score_with_reference<GCRO_E>(scoring, points, model, thr, score, n_inl, inl_idx); break;
	default: score_with_reference<GCRO_PNP>(scoring, points, model, thr, score, n_inl, inl_idx); break;
	}
}

// gamma_values.cpp:2012 table entry
double ref_gamma_value(int k) { return stored_gamma_values[k]; }
int ref_gamma_table_size() { return (int)(sizeof(stored_gamma_values) / sizeof(double)); }

// samples [first, first + count) of the reference's ImportanceSampler on the counter-based word stream; out [count][m]
void ref_sample_importance(const double *probabilities, int n, int m, uint64_t seed, uint64_t first, int count, uint32_t *out)
{
	cv::Mat dummy(n, 4, CV_64F);
	std::vector<double> p(probabilities, probabilities + n);
	ImportanceProbe probe(&dummy, p, (size_t)m);
	std::vector<size_t> subset(m);
	for (int i = 0; i < count; ++i)
	{
		probe.draw(seed, first + (uint64_t)i, (size_t)m, subset.data());
		for (int j = 0; j < m; ++j)
			out[(size_t)i * m + j] = (uint32_t)subset[j];
	}
}

// samples [0, count) of the reference's AdaptiveReorderingSampler driven as GCRANSAC::run drives it (sample, update,
// sample, ...); repeat_at >= 0: that sample is followed by a second update (GCRANSAC.h:297-339 falling through to :474)
void ref_sample_adaptive_reordering(const double *probabilities, int n, int m, double variance, uint64_t seed, int count,
                                    long long repeat_at, uint32_t *out)
{
	cv::Mat dummy(n, 4, CV_64F);
	CounterARSampler s(&dummy, normalized_by_max(std::vector<double>(probabilities, probabilities + n)), (size_t)m, variance, seed);
	std::vector<size_t> pool, subset(m);
	for (int i = 0; i < count; ++i)
	{
		s.sample(pool, subset.data(), (size_t)m);
		s.update(subset.data(), (size_t)m, (size_t)i, 0.0);
		if (i == repeat_at)
			s.update(subset.data(), (size_t)m, (size_t)i, 0.0);
		for (int j = 0; j < m; ++j)
			out[(size_t)i * m + j] = (uint32_t)subset[j];
	}
}
void ref_set_sampler_variance(double v) { g_sampler_variance = v; }
void ref_set_image_sizes(double w1, double h1, double w2, double h2)
{
	g_image_sizes[0] = w1; g_image_sizes[1] = h1; g_image_sizes[2] = w2; g_image_sizes[3] = h2;
}
// the first `count` samples of the reference's ProgressiveNapsacSampler<4> over the n x 4 points (pool = all points)
int ref_sample_pnapsac(const double *pts, int n, int m, const double *sizes, uint64_t seed, int count, uint32_t *out)
{
	cv::Mat points(n, 4, CV_64F, const_cast<double *>(pts));
	CounterPNapsacSampler s(&points, (size_t)m, sizes, seed);
	if (!s.isInitialized())
		return -1;
	std::vector<size_t> pool(n), subset(m);
	for (int i = 0; i < n; ++i)
		pool[i] = (size_t)i;
	for (int i = 0; i < count; ++i)
	{
		if (!s.sample(pool, subset.data(), (size_t)m))
			return -2 - i;
		for (int j = 0; j < m; ++j)
			out[(size_t)i * m + j] = (uint32_t)subset[j];
	}
	return 0;
}

// the point probabilities sampler_id 3 of ref_run hands to the reference's ImportanceSampler
void ref_set_point_probabilities(const double *probabilities, int n)
{
	g_point_probabilities.assign(probabilities, probabilities + n);
}

// prosac_sampler.h growth function and the subset size used by each of the first `calls` sample() calls
void ref_prosac(int n, int m, uint64_t conv, uint64_t *growth, int calls, uint32_t *subset_sizes, uint32_t *last_index)
{
	cv::Mat dummy(n, 4, CV_64F);
	ProsacProbe p(&dummy, (size_t)m, (size_t)conv);
	for (int i = 0; i < n; ++i)
		growth[i] = p.growth()[i];
	std::vector<size_t> pool, subset(m);
	for (int c = 0; c < calls; ++c)
	{
		subset_sizes[c] = (uint32_t)p.subsetSize();
		p.sample(pool, subset.data(), (size_t)m);
		last_index[c] = (uint32_t)subset[m - 1];
	}
}

// grid_neighborhood_graph.h verbatim
uint64_t ref_grid(const double *pts, int n, const double *cell_sizes, uint64_t cells_per_axis, uint64_t *keys,
                  uint32_t *sorted_idx, uint32_t *cell_offsets, uint32_t *n_cells)
{
	cv::Mat points(n, 4, CV_64F, const_cast<double *>(pts));
	std::vector<double> cs(cell_sizes, cell_sizes + 4);
	neighborhood::GridNeighborhoodGraph<4> g(&points, cs, (size_t)cells_per_axis);
	// keys through getNeighbors(): every point of a cell reports the same list; recover the key from getCells()
	std::vector<std::pair<size_t, const std::vector<size_t> *>> cells;
	for (const auto &kv : g.getCells())
		cells.emplace_back(kv.first, &kv.second.first);
	std::sort(cells.begin(), cells.end(), [](const auto &a, const auto &b) { return a.first < b.first; });
	uint32_t pos = 0, c = 0;
	for (const auto &cell : cells)
	{
		cell_offsets[c++] = pos;
		for (size_t idx : *cell.second)
		{
			keys[idx] = cell.first;
			sorted_idx[pos++] = (uint32_t)idx;
		}
	}
	cell_offsets[c] = pos;
	*n_cells = (uint32_t)g.filledCellNumber();
	return g.getNeighborNumber();
}

} // extern "C" (templates below need C++ linkage)

// One fixed-K pass of the reference's hypothesise-and-verify inner loop (GCRANSAC.h:297-471 without the adaptive
// stop and LO): counter-based samples -> isValidSample -> estimateModel -> MSAC/MAGSAC getScore (verbatim) -> strict
// '<' best-so-far.  This is what bench.py times as the CPU arm.
template <int PROBLEM, class Scoring>
static void round_with_reference(const cv::Mat &points, uint64_t seed, uint64_t first, int count, double thr,
                                 double *best_score, uint64_t *best_index, uint64_t *models_scored, double *best_model)
{
	typedef OracleEstimator<PROBLEM> Est;
	Est est;
	Scoring scoring;
	const double tt = 3.0 / 2.0 * thr;
	scoring.initialize(tt * tt, points.rows);
	std::vector<Model> models;
	std::vector<size_t> inliers;
	inliers.reserve(points.rows);
	Score best;
	uint64_t best_idx = ~0ull, scored = 0;
	constexpr size_t m = Est::sampleSize();
	uint32_t s32[8];
	size_t s[8];
	for (int i = 0; i < count; ++i)
	{
		gcro_sample_uniform(seed, first + (uint64_t)i, 0u, (uint32_t)points.rows, (int)m, s32);
		for (size_t j = 0; j < m; ++j)
			s[j] = s32[j];
		models.resize(0);
		if (!est.isValidSample(points, s))
			continue;
		if (!est.estimateModel(points, s, &models))
			continue;
		for (size_t j = 0; j < models.size(); ++j)
		{
			Score sc = scoring.getScore(points, models[j], est, thr, inliers, best, true);
			++scored;
			if (best < sc)
			{
				best = sc;
				best_idx = (uint64_t)i * Est::maximumMinimalSolutions() + j;
				if (best_model)
					to_flat(models[j].descriptor, best_model);
			}
		}
	}
	*best_score = best.value;
	*best_index = best_idx;
	*models_scored = scored;
}

extern "C" {

void ref_round(int problem, int scoring, const double *pts, int n, uint64_t seed, uint64_t first, int count, double thr,
               double *best_score, uint64_t *best_index, uint64_t *models_scored, double *best_model)
{
	const int cols = gcro_point_cols(problem);
	cv::Mat points(n, cols, CV_64F, const_cast<double *>(pts));
#define GCR_RR(P)                                                                                                       \
	do                                                                                                                  \
	{                                                                                                                   \
		if (scoring == GCRO_MSAC)                                                                                       \
			round_with_reference<P, MSACScoringFunction<OracleEstimator<P>>>(points, seed, first, count, thr, best_score, \
			                                                                best_index, models_scored, best_model);      \
		else                                                                                                            \
			round_with_reference<P, MAGSACScoringFunction<OracleEstimator<P>>>(points, seed, first, count, thr,          \
			                                                                  best_score, best_index, models_scored,     \
			                                                                  best_model);                               \
	} while (0)
	switch (problem)
	{
	case GCRO_H: GCR_RR(GCRO_H); break;
	case GCRO_F: GCR_RR(GCRO_F); break;
	case GCRO_E: GCR_RR(GCRO_E); break;
	default: GCR_RR(GCRO_PNP); break;
	}
#undef GCR_RR
}

// the exact radius neighbourhood (ExactRadiusGraph below is the class the verbatim driver runs on): CSR
uint64_t ref_radius_graph(const double *pts, int n, double radius, uint32_t *offsets /* n + 1 */, uint32_t *neighbors, uint64_t capacity);

uint64_t ref_iteration_number(uint64_t inliers, uint64_t n, int m, double log_probability)
{
	IterProbe p;
	return p.get(inliers, n, (size_t)m, log_probability);
}

} // extern "C"

// ------------------------------------------------------------------------------------------------------------------
// Whole GCRANSAC<>::run (GCRANSAC.h:202-627, verbatim) on the counter-based sample stream -- config 1's CPU anchor and
// the harness for the drop-in scoring adapter.  The Sampler below implements the reference's virtual Sampler interface
// (samplers/sampler.h:44-86): call k of the main (LO) sampler returns sample k of stream 0 (stream 1) of the Philox
// generator, uniform over the pool exactly like UniformSampler::sample (uniform_sampler.h:81-99).
namespace
{
	class CounterSampler : public sampler::Sampler<cv::Mat, size_t>
	{
		uint64_t seed, next;
		uint32_t stream;
		std::vector<uint32_t> tmp;

	public:
		CounterSampler(const cv::Mat *c, uint64_t seed_, uint32_t stream_) : sampler::Sampler<cv::Mat, size_t>(c), seed(seed_), next(0), stream(stream_)
		{
			initialized = true;
		}
		const std::string getName() const { return "Counter-based uniform sampler"; }
		void reset() {}
		void update(const size_t *const, const size_t &, const size_t &, const double &) {}
		bool initialize(const cv::Mat *const) { return true; }
		uint64_t drawn() const { return next; }
		bool sample(const std::vector<size_t> &pool_, size_t *const subset_, size_t sample_size_)
		{
			if (sample_size_ > pool_.size())
				return false;
			tmp.resize(sample_size_);
			gcro_sample_uniform(seed, next++, stream, (uint32_t)pool_.size(), (int)sample_size_, tmp.data());
			for (size_t i = 0; i < sample_size_; ++i)
				subset_[i] = pool_[tmp[i]];
			return true;
		}
	};

	// PROSAC (samplers/prosac_sampler.h:207-236) on the counter-based stream: call k (from 1) returns PROSAC sample k
	// of gcro_sample_prosac -- the growth function, subset sizes and the k = 1 quirk are pinned against the reference's
	// own ProsacSampler (tests/golden/prosac.npz); only its random_device-seeded draws are replaced.
	class CounterProsacSampler : public sampler::Sampler<cv::Mat, size_t>
	{
		uint64_t seed, next;
		uint32_t n;
		std::vector<uint64_t> growth;
		std::vector<uint32_t> tmp;

	public:
		CounterProsacSampler(const cv::Mat *c, uint64_t seed_) : sampler::Sampler<cv::Mat, size_t>(c), seed(seed_), next(0), n((uint32_t)c->rows)
		{
			initialized = true;
		}
		const std::string getName() const { return "Counter-based PROSAC sampler"; }
		void reset() {}
		void update(const size_t *const, const size_t &, const size_t &, const double &) {}
		bool initialize(const cv::Mat *const) { return true; }
		uint64_t drawn() const { return next; }
		bool sample(const std::vector<size_t> &, size_t *const subset_, size_t sample_size_)
		{
			if (sample_size_ > n)
				return false;
			if (growth.empty())
			{
				growth.resize(n);
				gcro_prosac_growth(n, (int)sample_size_, 100000, growth.data());
			}
			tmp.resize(sample_size_);
			gcro_sample_prosac(seed, ++next, n, (int)sample_size_, growth.data(), 100000, tmp.data());
			for (size_t i = 0; i < sample_size_; ++i)
				subset_[i] = tmp[i];
			return true;
		}
	};

	// ImportanceSampler (samplers/importance_sampler.h) on the counter-based stream: call k (from 0) returns sample k
	// drawn from the reference's own distribution object (ImportanceProbe above).
	class CounterImportanceSampler : public sampler::Sampler<cv::Mat, size_t>
	{
		uint64_t seed, next;
		std::unique_ptr<ImportanceProbe> probe;

	public:
		CounterImportanceSampler(const cv::Mat *c, uint64_t seed_, const std::vector<double> &p, size_t m) : sampler::Sampler<cv::Mat, size_t>(c), seed(seed_), next(0)
		{
			if (p.size() == (size_t)c->rows)
				probe.reset(new ImportanceProbe(c, p, m));
			initialized = (bool)probe;
		}
		const std::string getName() const { return "Counter-based importance sampler"; }
		void reset() {}
		void update(const size_t *const, const size_t &, const size_t &, const double &) {}
		bool initialize(const cv::Mat *const) { return true; }
		uint64_t drawn() const { return next; }
		bool sample(const std::vector<size_t> &, size_t *const subset_, size_t sample_size_)
		{
			if (!probe)
				return false;
			probe->draw(seed, next++, sample_size_, subset_);
			return true;
		}
	};

	// fundamental_estimator.h:341-572.  The H-degenerate sample test is the restated gcro_degensac_sample_homography;
	// the inner GC-RANSAC is the reference's own GCRANSAC<> (verbatim) over the plane-and-parallax estimator with the
	// settings of :547-551.  The reference draws the inner samples from a fresh random_device-seeded UniformSampler
	// shared by the main loop and the local optimisation (:553, not reproducible); here run number r of a call uses
	// the counter-based streams (seed + r * 0xD1B54A32D192ED03, 0 / 1).
	bool oracle_apply_degensac(const OracleEstimator<GCRO_F> &est, Model &model_, const cv::Mat &data_,
	                           const std::vector<size_t> &inliers_, const size_t *minimal_sample_, double threshold_,
	                           bool &model_updated_)
	{
		model_updated_ = false;
		double F[9], Hn[9];
		to_flat(model_.descriptor, F);
		const int front = gcro_degensac_front(F, reinterpret_cast<const double *>(data_.data), inliers_, minimal_sample_, Hn);
		if (front <= 0)
			return front == 0; // 0: not H-degenerate, the model stands (:571); -1: too few homography inliers / no refit (:512, :523)
		typedef OracleEstimator<GCRO_FPP> Inner;
		Inner inner;
		inner.pp_homography = Hn;
		++est.degensac_runs;
		const uint64_t inner_seed = est.seed + est.degensac_runs * 0xD1B54A32D192ED03ull;
		CounterSampler main_sampler(&data_, inner_seed, 0u), lo_sampler(&data_, inner_seed, 1u);
		GCRANSAC<Inner, neighborhood::GridNeighborhoodGraph<4>> gcransac;
		gcransac.settings.threshold = threshold_;
		gcransac.settings.spatial_coherence_weight = 0;
		gcransac.settings.confidence = 0.99;
		gcransac.settings.neighborhood_sphere_radius = 8;
		Model model;
		gcransac.run(data_, inner, &main_sampler, &lo_sampler, nullptr, model);
		const utils::RANSACStatistics &statistics = gcransac.getRansacStatistics();
		if (statistics.inliers.size() > inliers_.size()) // :561-567
		{
			model_updated_ = true;
			model_ = model;
			++est.degensac_updates;
		}
		return true;
	}

	// The neighbourhood of neighborhood=1 (neighborhood/flann_neighborhood_graph.h:99-137) without FLANN: every point's
	// neighbours within `radius` (L2 over the four coordinates converted to FP32 as :115 does, strict <, the point itself
	// dropped as :131-133 drop the first match).  The reference's lists are what OpenCV's approximate kd-tree search
	// (4 trees, 6 checks) happens to return; this is the exact graph, in ascending index order.
	class ExactRadiusGraph : public neighborhood::NeighborhoodGraph<cv::Mat>
	{
		double radius;
		std::vector<std::vector<size_t>> neighbours;
		std::vector<double> no_cell_sizes;
		std::unordered_map<size_t, std::pair<std::vector<size_t>, std::vector<size_t>>> no_cells;

	public:
		ExactRadiusGraph(const cv::Mat *c, double r) : neighborhood::NeighborhoodGraph<cv::Mat>(c), radius(r) { initialized = initialize(c); }
		bool initialize(const cv::Mat *const c)
		{
			const int n = c->rows, cols = c->cols; // 4 (correspondences) or 5 (find6DPose_'s u v X Y Z rows)
			const double *p = reinterpret_cast<const double *>(c->data);
			std::vector<float> f((size_t)n * cols);
			for (size_t i = 0; i < f.size(); ++i) f[i] = (float)p[i];
			const float r = (float)radius, r2 = r * r;
			neighbours.assign(n, std::vector<size_t>());
			neighbor_number = 0;
			for (int i = 0; i < n; ++i)
			{
				for (int j = 0; j < n; ++j)
				{
					const float dx = f[cols * i] - f[cols * j], dy = f[cols * i + 1] - f[cols * j + 1], dz = f[cols * i + 2] - f[cols * j + 2],
					            dw = f[cols * i + 3] - f[cols * j + 3];
					float d2 = ((dx * dx + dy * dy) + dz * dz) + dw * dw;
					if (cols == 5)
					{
						const float dv = f[cols * i + 4] - f[cols * j + 4];
						d2 = d2 + dv * dv;
					}
					if (d2 < r2 && j != i)
						neighbours[i].push_back((size_t)j);
				}
				neighbor_number += neighbours[i].size();
			}
			return neighbor_number > 0;
		}
		const std::vector<size_t> &getNeighbors(size_t i) const { return neighbours[i]; }
		const std::vector<double> &getCellSizes() const { return no_cell_sizes; }
		const std::unordered_map<size_t, std::pair<std::vector<size_t>, std::vector<size_t>>> &getCells() const { return no_cells; }
		size_t getDivisionNumber() const { return 0; }
		size_t filledCellNumber() const { return neighbours.size(); }
	};

	template <int PROBLEM, class Scoring, class Graph>
	int run_reference_on(Graph &graph, const cv::Mat &points, int cells_per_axis, double threshold, double conf,
	                     double spatial_weight, int max_iters, int min_iters, int lo_number, uint64_t seed, double *model_out,
	                     uint8_t *inlier_mask, uint64_t *stats_out, double *score_out, int use_degensac, int f_ba, int sampler_id);

	// radius > 0: the exact radius neighbourhood (neighborhood=1) instead of the grid (neighborhood=0)
	template <int PROBLEM, class Scoring>
	int run_reference(const cv::Mat &points, const cv::Mat &grid_points, const double *cell_sizes, int cells_per_axis, double threshold, double conf,
	                  double spatial_weight, int max_iters, int min_iters, int lo_number, uint64_t seed, double *model_out,
	                  uint8_t *inlier_mask, uint64_t *stats_out, double *score_out, int use_degensac = 0, int f_ba = 0,
	                  int sampler_id = 0, double radius = 0.0)
	{
		if (radius > 0.0)
		{
			ExactRadiusGraph graph(&grid_points, radius);
			return run_reference_on<PROBLEM, Scoring>(graph, points, cells_per_axis, threshold, conf, spatial_weight, max_iters, min_iters,
			                                          lo_number, seed, model_out, inlier_mask, stats_out, score_out, use_degensac, f_ba, sampler_id);
		}
		std::vector<double> cs(cell_sizes, cell_sizes + 4);
		neighborhood::GridNeighborhoodGraph<4> graph(&grid_points, cs, (size_t)cells_per_axis);
		return run_reference_on<PROBLEM, Scoring>(graph, points, cells_per_axis, threshold, conf, spatial_weight, max_iters, min_iters,
		                                          lo_number, seed, model_out, inlier_mask, stats_out, score_out, use_degensac, f_ba, sampler_id);
	}

	template <int PROBLEM, class Scoring, class Grid>
	int run_reference_on(Grid &graph, const cv::Mat &points, int cells_per_axis, double threshold, double conf,
	                     double spatial_weight, int max_iters, int min_iters, int lo_number, uint64_t seed, double *model_out,
	                     uint8_t *inlier_mask, uint64_t *stats_out, double *score_out, int use_degensac, int f_ba, int sampler_id)
	{
		typedef OracleEstimator<PROBLEM> Est;
		Est estimator;
		estimator.use_degensac = use_degensac != 0;
		estimator.f_nonminimal_ba = (f_ba & 1) != 0;
		estimator.seed = seed;
		// sampler_id 1 = PROSAC main sampler (py/gcransac_python.cpp:4320-4323); the LO sampler is uniform either way
		// (stream 2 then: PROSAC draws use stream 1)
		CounterSampler uniform_sampler(&points, seed, 0u), lo_sampler(&points, seed, sampler_id == 1 ? 2u : 1u);
		CounterProsacSampler prosac_sampler(&points, seed);
		// sampler_id 3 = the importance sampler of NG-RANSAC (py/gcransac_python.cpp:4343-4346)
		CounterImportanceSampler importance_sampler(&points, seed, sampler_id == 3 ? g_point_probabilities : std::vector<double>(), Est::sampleSize());
		if (sampler_id == 3 && !importance_sampler.isInitialized())
			return -3;
		// sampler_id 4 = the AR-Sampler over the max-normalised probabilities (py/gcransac_python.cpp:4348-4359)
		std::unique_ptr<CounterARSampler> ar_sampler;
		if (sampler_id == 4)
		{
			if (g_point_probabilities.size() != (size_t)points.rows)
				return -3;
			ar_sampler.reset(new CounterARSampler(&points, normalized_by_max(g_point_probabilities), Est::sampleSize(), g_sampler_variance, seed));
		}
		// sampler_id 2 = Progressive NAPSAC over the points the grids are built on (py/gcransac_python.cpp:4334-4343)
		std::unique_ptr<CounterPNapsacSampler> pnapsac_sampler;
		if (sampler_id == 2)
		{
			pnapsac_sampler.reset(new CounterPNapsacSampler(g_pnapsac_points ? g_pnapsac_points : &points, Est::sampleSize(), g_image_sizes, seed));
			if (!pnapsac_sampler->isInitialized())
				return -3;
		}
		sampler::Sampler<cv::Mat, size_t> *main_sampler_ptr = sampler_id == 1   ? (sampler::Sampler<cv::Mat, size_t> *)&prosac_sampler
		                                                      : sampler_id == 2 ? (sampler::Sampler<cv::Mat, size_t> *)pnapsac_sampler.get()
		                                                      : sampler_id == 4 ? (sampler::Sampler<cv::Mat, size_t> *)ar_sampler.get()
		                                                      : sampler_id == 3 ? (sampler::Sampler<cv::Mat, size_t> *)&importance_sampler
		                                                                        : (sampler::Sampler<cv::Mat, size_t> *)&uniform_sampler;
		Model model;
		GCRANSAC<Est, Grid, Scoring> gcransac;
		// py/gcransac_python.cpp:4498-4504
		gcransac.settings.threshold = threshold;
		gcransac.settings.spatial_coherence_weight = spatial_weight;
		gcransac.settings.confidence = conf;
		gcransac.settings.max_local_optimization_number = lo_number;
		gcransac.settings.max_iteration_number = max_iters;
		gcransac.settings.min_iteration_number = min_iters;
		gcransac.settings.neighborhood_sphere_radius = cells_per_axis;
		gcransac.run(points, estimator, main_sampler_ptr, &lo_sampler, &graph, model);
		utils::RANSACStatistics st = gcransac.getRansacStatistics();
		if constexpr (PROBLEM == GCRO_F)
			if ((f_ba & 2) && st.inliers.size() > 7)
			{
				// the post-run bundle adjustment of findFundamentalMatrix_ (py/gcransac_python.cpp:1832-1887): weights
				// 1 - r^2 / T by inlier position, the BA solver on the pixel coordinates, the reference's own MSAC score
				const double truncated_threshold = 3.0 / 2.0 * threshold;
				const double squared_truncated_threshold = truncated_threshold * truncated_threshold;
				const size_t inlier_number = st.inliers.size();
				std::vector<double> weights(inlier_number, 0.0);
				std::vector<uint32_t> inl(inlier_number);
				for (size_t k = 0; k < inlier_number; ++k)
				{
					inl[k] = (uint32_t)st.inliers[k];
					weights[k] = MAX(0, 1.0 - estimator.squaredResidual(points.row((int)st.inliers[k]), model) / squared_truncated_threshold);
				}
				double polished[9];
				if (gcro_polish_f_ba(reinterpret_cast<const double *>(points.data), inl.data(), (int)inlier_number, weights.data(), polished))
				{
					Model tmp_model = from_flat(polished, 3, 3);
					MSACScoringFunction<Est> scoring;
					scoring.initialize(squared_truncated_threshold, points.rows);
					std::vector<size_t> current_inliers;
					Score score = scoring.getScore(points, tmp_model, estimator, squared_truncated_threshold, current_inliers);
					if (st.score < score.value)
					{
						model.descriptor = tmp_model.descriptor;
						st.score = score.value;
						st.inliers.swap(current_inliers);
					}
				}
			}
		if constexpr (PROBLEM == GCRO_E)
			if (f_ba & 4)
			{
				// the post-run bundle adjustment of findEssentialMatrix_ (py/gcransac_python.cpp:2766-2821): weights 1 - r^2 / T by
				// inlier position, EssentialMatrixBundleAdjustmentSolver on the normalised points (restated:
				// gcr_oracle_eba.cpp), every returned model scored by the reference's own MSAC in order; fewer than six
				// inliers: the identity (:2822-2823)
				if (st.inliers.size() > 5)
				{
					const double truncated_threshold = 3.0 / 2.0 * threshold;
					const double squared_truncated_threshold = truncated_threshold * truncated_threshold;
					const size_t inlier_number = st.inliers.size();
					std::vector<double> weights(inlier_number, 0.0);
					std::vector<uint32_t> inl(inlier_number);
					for (size_t k = 0; k < inlier_number; ++k)
					{
						inl[k] = (uint32_t)st.inliers[k];
						weights[k] = MAX(0, 1.0 - estimator.squaredResidual(points.row((int)st.inliers[k]), model) / squared_truncated_threshold);
					}
					double polished[360];
					const int n_models = gcro_polish_e_ba(reinterpret_cast<const double *>(points.data), points.rows, inl.data(),
					                                      (int)inlier_number, weights.data(), polished);
					MSACScoringFunction<Est> scoring;
					scoring.initialize(squared_truncated_threshold, points.rows);
					for (int q = 0; q < n_models; ++q)
					{
						Model tmp_model = from_flat(polished + 9 * q, 3, 3);
						std::vector<size_t> current_inliers;
						Score score = scoring.getScore(points, tmp_model, estimator, squared_truncated_threshold, current_inliers);
						if (st.score < score.value)
						{
							model.descriptor = tmp_model.descriptor;
							st.score = score.value;
							st.inliers.swap(current_inliers);
						}
					}
				}
				else
				{
					const double identity[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
					model = from_flat(identity, 3, 3);
				}
			}
		if constexpr (PROBLEM == GCRO_PNP)
			if ((f_ba & 8) && st.inliers.size() > 3)
			{
				// find6DPose_ (py/gcransac_python.cpp:1227-1238): the non-minimal solver on the inliers replaces the model
				std::vector<uint32_t> inl(st.inliers.begin(), st.inliers.end());
				double refined[12];
				if (gcro_refit_pnp(reinterpret_cast<const double *>(points.data), inl.data(), (int)inl.size(), nullptr, refined))
					model = from_flat(refined, 3, 4);
			}
		for (int i = 0; i < points.rows; ++i)
			inlier_mask[i] = 0;
		for (size_t idx : st.inliers)
			inlier_mask[idx] = 1;
		if (st.inliers.size() > 0 || (PROBLEM == GCRO_E && (f_ba & 4)))
			to_flat(model.descriptor, model_out);
		if (stats_out)
		{
			stats_out[0] = st.iteration_number;
			stats_out[1] = st.graph_cut_number;
			stats_out[2] = st.local_optimization_number;
			stats_out[3] = st.neighbor_number;
			stats_out[4] = st.accepted_models;
			stats_out[5] = st.rejected_models;
			stats_out[6] = sampler_id == 1 ? prosac_sampler.drawn() : sampler_id == 3 ? importance_sampler.drawn() : sampler_id == 4 ? ar_sampler->drawn() : sampler_id == 2 ? pnapsac_sampler->drawn() : uniform_sampler.drawn();
			stats_out[7] = lo_sampler.drawn();
			stats_out[8] = estimator.degensac_runs;
			stats_out[9] = estimator.degensac_updates;
		}
		if (score_out)
			*score_out = st.score;
		return (int)st.inliers.size();
	}
}

#ifdef GCR_REF_WITH_CUDA
#include "gcr_scoring_function.h" // graph-cut-ransac_b200/host: the drop-in adapter over libgcr_b200.so
#endif

extern "C" {

// scoring_kind: 0 = reference MSACScoringFunction (verbatim CPU), 2 = reference MAGSACScoringFunction (verbatim CPU),
// 1 = gcr::CudaMSACScoringFunction (library built with GCR_REF_WITH_CUDA only).  Returns the inlier count.
int ref_run(int problem, int scoring_kind, const double *pts, int n, const double *cell_sizes, int cells_per_axis,
            double threshold, double conf, double spatial_weight, int max_iters, int min_iters, int lo_number,
            uint64_t seed, double *model_out, uint8_t *inlier_mask, uint64_t *stats_out, double *score_out,
            const double *grid_pts, int use_degensac, int f_ba, int sampler_id, double radius)
{
	const int cols = gcro_point_cols(problem);
	cv::Mat points(n, cols, CV_64F, const_cast<double *>(pts));
	cv::Mat grid_points(n, cols, CV_64F, const_cast<double *>(grid_pts ? grid_pts : pts));
	struct PNapsacPoints // findEssentialMatrix_ hands P-NAPSAC the pixel coordinates (py/gcransac_python.cpp:2650)
	{
		PNapsacPoints(const cv::Mat *p) { g_pnapsac_points = p; }
		~PNapsacPoints() { g_pnapsac_points = nullptr; }
	} pnapsac_points_guard(&grid_points);
#define GCR_RUN(P, S) run_reference<P, S>(points, grid_points, cell_sizes, cells_per_axis, threshold, conf, spatial_weight, max_iters, min_iters, lo_number, seed, model_out, inlier_mask, stats_out, score_out, use_degensac, f_ba, sampler_id, radius)
	if (scoring_kind == 0)
	{
		if (problem == GCRO_H) return GCR_RUN(GCRO_H, MSACScoringFunction<OracleEstimator<GCRO_H>>);
		if (problem == GCRO_F) return GCR_RUN(GCRO_F, MSACScoringFunction<OracleEstimator<GCRO_F>>);
		if (problem == GCRO_E) return GCR_RUN(GCRO_E, MSACScoringFunction<OracleEstimator<GCRO_E>>);
		if (problem == GCRO_PNP) return GCR_RUN(GCRO_PNP, MSACScoringFunction<OracleEstimator<GCRO_PNP>>);
		return -1;
	}
	if (scoring_kind == 2) // GCRANSAC<Estimator, Graph, MAGSACScoringFunction<Estimator>> (scoring_function.h:237-349)
	{
		if (problem == GCRO_H) return GCR_RUN(GCRO_H, MAGSACScoringFunction<OracleEstimator<GCRO_H>>);
		if (problem == GCRO_F) return GCR_RUN(GCRO_F, MAGSACScoringFunction<OracleEstimator<GCRO_F>>);
		if (problem == GCRO_E) return GCR_RUN(GCRO_E, MAGSACScoringFunction<OracleEstimator<GCRO_E>>);
		return -1;
	}
#ifdef GCR_REF_WITH_CUDA
	if (scoring_kind == 1)
	{
		typedef gcr::CudaMSACScoringFunction<OracleEstimator<GCRO_H>, GCR_H> SH;
		typedef gcr::CudaMSACScoringFunction<OracleEstimator<GCRO_F>, GCR_F> SF;
		if (problem == GCRO_H) return GCR_RUN(GCRO_H, SH);
		typedef gcr::CudaMSACScoringFunction<OracleEstimator<GCRO_E>, GCR_E> SE;
		if (problem == GCRO_F) return GCR_RUN(GCRO_F, SF);
		if (problem == GCRO_E) return GCR_RUN(GCRO_E, SE);
		return -1;
	}
#endif
#undef GCR_RUN
	return -2;
}

uint64_t ref_radius_graph(const double *pts, int n, double radius, uint32_t *offsets, uint32_t *neighbors, uint64_t capacity)
{
	cv::Mat points(n, 4, CV_64F, const_cast<double *>(pts));
	ExactRadiusGraph g(&points, radius);
	uint64_t total = 0;
	for (int i = 0; i < n; ++i)
	{
		offsets[i] = (uint32_t)total;
		for (size_t j : g.getNeighbors(i))
		{
			if (total < capacity) neighbors[total] = (uint32_t)j;
			++total;
		}
	}
	offsets[n] = (uint32_t)total;
	return total;
}

} // extern "C"
