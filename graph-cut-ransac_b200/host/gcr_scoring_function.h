// gcr_scoring_function.h -- the reference-side binding of libgcr_b200.so: drop-in replacements for
// gcransac::MSACScoringFunction<Estimator> / MAGSACScoringFunction<Estimator> (scoring_function.h:103-349).
//
//     #include "GCRANSAC.h"                       // the reference's own headers, unmodified
//     #include "gcr_scoring_function.h"
//     gcransac::GCRANSAC<Estimator, Graph, gcr::CudaMSACScoringFunction<Estimator, GCR_H>> gcransac;
//     gcransac.run(points, estimator, &main_sampler, &lo_sampler, &graph, model);
//
// Nothing else changes: GCRANSAC<> owns the scoring object through std::unique_ptr<_ScoringFunction> (GCRANSAC.h:59,
// 112), calls initialize() once (:253-255) and getScore() per model (:428-434, :547-555, :696-700, :875-881).  The
// adapter uploads the point matrix the first time it sees it (keyed by data pointer / shape, like the cv::Mat view the
// drivers build, py/gcransac_python.cpp:4274) and then scores every model with the FP64 kernel K5 (gcr_getscore_host):
// exact inlier count, ascending inlier list (the emplace_back order of scoring_function.h:166-167), score to 1e-12.
// This is the single-hypothesis, HBM-bound regime of the library; the batched regime (K hypotheses per launch, FP32 hot
// kernel) is gcr_round / gcr_score, driven by graph-cut-ransac_b200/host/gcr_driver.h.
//
// There is no CPU fallback: if no CUDA device is present the constructor throws.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "gcr/gcr.h"
#include "scoring_function.h" // the reference's (Score, ScoringFunction<>)

namespace gcr
{
	// one context per device per process, shared by every adapter instance (GCRANSAC<> default-constructs its scorer)
	inline gcr_ctx *shared_context(int device = 0)
	{
		static gcr_ctx *ctx[16] = {nullptr};
		if (device < 0 || device >= 16)
			throw std::invalid_argument("gcr::shared_context: device index out of range");
		if (!ctx[device] && gcr_ctx_create(device, &ctx[device]) != GCR_OK)
			throw std::runtime_error("gcr: no usable CUDA device " + std::to_string(device) + " (libgcr_b200 has no CPU fallback)");
		return ctx[device];
	}

	template <class _ModelEstimator, int PROBLEM, int SCORING>
	class CudaScoringFunction : public gcransac::ScoringFunction<_ModelEstimator>
	{
	protected:
		gcr_ctx *ctx;
		double squared_truncated_threshold = 0.0; // what GCRANSAC::run hands to initialize() (GCRANSAC.h:253-255)
		size_t point_number = 0;
		// which matrix is resident on the device
		mutable const void *resident_data = nullptr;
		mutable int resident_rows = 0, resident_cols = 0;
		mutable std::vector<uint32_t> idx32;

		void ensure_points(const cv::Mat &points_) const
		{
			if (points_.data == resident_data && points_.rows == resident_rows && points_.cols == resident_cols)
				return;
			if (points_.type() != CV_64F)
				throw std::invalid_argument("gcr::CudaScoringFunction: points must be CV_64F");
			if (gcr_points_upload(ctx, reinterpret_cast<const double *>(points_.data), points_.rows, points_.cols, 1) != GCR_OK)
				throw std::runtime_error(std::string("gcr_points_upload: ") + gcr_last_error(ctx));
			resident_data = points_.data;
			resident_rows = points_.rows;
			resident_cols = points_.cols;
		}

	public:
		CudaScoringFunction() : ctx(shared_context(0)) {}

		// re-upload on the next call (the caller changed the matrix contents in place)
		void invalidatePoints() { resident_data = nullptr; }

		void initialize(const double squared_truncated_threshold_, const size_t point_number_) override
		{
			squared_truncated_threshold = squared_truncated_threshold_;
			point_number = point_number_;
		}

		gcransac::Score getScore(const cv::Mat &points_, gcransac::Model &model_, const _ModelEstimator & /*estimator_*/,
		                         const double threshold_, std::vector<size_t> &inliers_,
		                         const gcransac::Score & /*best_score_*/ = gcransac::Score(), const bool store_inliers_ = true,
		                         const std::vector<const std::vector<size_t> *> *index_sets = nullptr) const override
		{
			if (index_sets != nullptr)
				throw std::invalid_argument("gcr::CudaScoringFunction: index-set pre-selection (space partitioning) is not "
				                            "supported: dense GPU scoring makes it moot, pass EmptyInlierSelector");
			ensure_points(points_);
			const int rows = (int)model_.descriptor.rows(), cols = (int)model_.descriptor.cols();
			double flat[12];
			if (rows * cols != gcr_model_len(PROBLEM))
				throw std::invalid_argument("gcr::CudaScoringFunction: model shape does not match the problem");
			for (int r = 0; r < rows; ++r)
				for (int c = 0; c < cols; ++c)
					flat[r * cols + c] = model_.descriptor(r, c);
			double score = 0.0;
			uint32_t count = 0;
			idx32.resize(store_inliers_ ? (size_t)points_.rows : 0);
			// MSAC ignores threshold_ and uses the squared truncated threshold of initialize() (scoring_function.h:163);
			// MAGSAC uses threshold_ as sigma_max (:257, :293)
			if (gcr_getscore_host(ctx, PROBLEM, SCORING, 0, flat, squared_truncated_threshold, threshold_, &score, &count,
			                      store_inliers_ ? idx32.data() : nullptr) != GCR_OK)
				throw std::runtime_error(std::string("gcr_getscore_host: ") + gcr_last_error(ctx));
			inliers_.clear();
			if (store_inliers_)
				inliers_.assign(idx32.begin(), idx32.begin() + count);
			gcransac::Score s;
			s.inlier_number = count;
			s.value = score;
			return s;
		}
	};

	template <class _ModelEstimator, int PROBLEM>
	using CudaMSACScoringFunction = CudaScoringFunction<_ModelEstimator, PROBLEM, GCR_MSAC>;
	template <class _ModelEstimator, int PROBLEM>
	using CudaMAGSACScoringFunction = CudaScoringFunction<_ModelEstimator, PROBLEM, GCR_MAGSAC>;
}
