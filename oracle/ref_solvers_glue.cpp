// oracle/ref_solvers_glue.cpp -- TEST INFRASTRUCTURE ONLY.  Built by oracle/build_ref_solvers.sh into
// oracle/_ref/libgcr_ref_solvers.so: the reference's own SOLVER and ESTIMATOR headers, #included UNMODIFIED from
// /root/reference/src/pygcransac/include, compiled over oracle/shim_solvers (an eager stand-in for the part of Eigen's API
// they use, whose decompositions are oracle/linalg.h -- see the header of shim_solvers/Eigen/Eigen for what that does and
// does not pin).  The C entry points below mirror the oracle's (gcr_oracle.h) one to one, so the tests can hold
// restatement and reference source side by side on the same inputs:
//   solver_homography_four_point.h, homography_estimator.h            <-> gcro_solve_h4, gcro_h_valid_sample, gcro_refit_h
//   solver_fundamental_matrix_seven_point.h / _eight_point.h,
//   fundamental_estimator.h                                           <-> gcro_solve_f7, gcro_estimate_f7, gcro_refit_f8
//   solver_essential_matrix_five_point_nister.h / _stewenius.h,
//   essential_estimator.h                                             <-> gcro_solve_e5_nister, gcro_solve_e5_stewenius,
//                                                                         gcro_estimate_e5, gcro_refit_e
//   solver_p3p.h                                                      <-> gcro_solve_p3p
#include <opencv2/core/core.hpp>
#include <Eigen/Eigen>
#include <cstdint>
#include <vector>

#include "ref_rng_redirect.h" // before any header of the reference: its random-number class on a reproducible word stream

#include "estimators/solver_fundamental_matrix_seven_point.h"
#include "estimators/solver_fundamental_matrix_eight_point.h"
#include "estimators/solver_homography_four_point.h"
#include "estimators/solver_essential_matrix_five_point_nister.h"
#include "estimators/solver_essential_matrix_five_point_stewenius.h"
#include "estimators/solver_p3p.h"
#include "estimators/essential_estimator.h"
#include "estimators/perspective_n_point_estimator.h"
#include "estimators/solver_fundamental_matrix_bundle_adjustment.h"
#include "estimators/solver_essential_matrix_bundle_adjustment.h"

#include "GCRANSAC.h"
#include "neighborhood/grid_neighborhood_graph.h"
#include "samplers/sampler.h"

#include "gcr_oracle.h"
#include "degensac_front.h"

#include "relative_pose/jacobian_impl.h"
#include "relative_pose/robust_loss.h"
#include <type_traits>

// DESIGN.md section 8, checked by the compiler: refine_pnp (relative_pose/bundle.cpp:578-597) constructs its accumulator as
// (correspondences, sample, sample_size, loss_fn, weights), but CameraJacobianAccumulator's fourth parameter is the Camera
// (relative_pose/jacobian_impl.h:12-18) -- the template only compiles because nothing instantiates it.  That is why the
// non-minimal PnP fit has no reference behaviour to match and a stated one is built instead (csrc/k_refit_pnp.cu).
static_assert(!std::is_constructible<pose_lib::CameraJacobianAccumulator<pose_lib::NullCameraModel, pose_lib::TruncatedLoss>,
                                     const cv::Mat &, const size_t *, const size_t &, pose_lib::TruncatedLoss &, const double *>::value,
              "refine_pnp's accumulator construction would compile: the reference's non-minimal PnP fit has a defined behaviour after all");
static_assert(std::is_constructible<pose_lib::CameraJacobianAccumulator<pose_lib::NullCameraModel, pose_lib::TruncatedLoss>,
                                    const cv::Mat &, const size_t *, const size_t &, const pose_lib::Camera &, pose_lib::TruncatedLoss &, const double *>::value,
              "the accumulator's declared constructor");

using namespace gcransac;

// gcr_oracle_e.cpp's test hook (this library carries its own copy of the oracle's sources)
extern int (*gcro_e5_root_finder)(const double *c11, double *roots10);

namespace
{
	cv::Mat view(const double *pts, int n, int cols) { return cv::Mat(n, cols, CV_64F, const_cast<double *>(pts)); }

	std::vector<size_t> indices(const uint32_t *sample, int m)
	{
		std::vector<size_t> s(m);
		for (int i = 0; i < m; ++i) s[i] = sample[i];
		return s;
	}

	// models -> row-major coefficients, `len` each (the descriptor is column-major inside the shim, like Eigen's)
	int copy_out(const std::vector<Model> &models, int rows, int cols, int max_models, double *out)
	{
		int k = 0;
		for (const Model &m : models)
		{
			if (k == max_models) break;
			for (int i = 0; i < rows; ++i)
				for (int j = 0; j < cols; ++j)
					out[(size_t)k * rows * cols + (size_t)i * cols + j] = m.descriptor(i, j);
			++k;
		}
		return k;
	}

	typedef estimator::FundamentalMatrixEstimator<estimator::solver::FundamentalMatrixSevenPointSolver,
	                                              estimator::solver::FundamentalMatrixEightPointSolver> FEstimator;
	typedef estimator::RobustHomographyEstimator<estimator::solver::HomographyFourPointSolver,
	                                             estimator::solver::HomographyFourPointSolver> HEstimator;
	typedef estimator::EssentialMatrixEstimator<estimator::solver::EssentialMatrixFivePointNisterSolver,
	                                            estimator::solver::EssentialMatrixFivePointSteweniusSolver> EEstimator;
}

extern "C"
{
	// ---- the solver classes' estimateModel, as they stand ----
	int gcr_refsrc_solve_h4(const double *pts, int n, int cols, const uint32_t *sample, int m, const double *weights, double *H9)
	{
		estimator::solver::HomographyFourPointSolver s;
		std::vector<Model> models;
		const std::vector<size_t> idx = indices(sample, m);
		if (!s.estimateModel(view(pts, n, cols), idx.data(), m, models, weights)) return 0;
		return copy_out(models, 3, 3, 1, H9);
	}
	int gcr_refsrc_solve_f7(const double *pts, int n, const uint32_t *sample, int m, const double *weights, double *F27)
	{
		estimator::solver::FundamentalMatrixSevenPointSolver s;
		std::vector<Model> models;
		const std::vector<size_t> idx = indices(sample, m);
		if (!s.estimateModel(view(pts, n, 4), idx.data(), m, models, weights)) return 0;
		return copy_out(models, 3, 3, 3, F27);
	}
	int gcr_refsrc_solve_f8(const double *pts, int n, const uint32_t *sample, int m, const double *weights, double *F9)
	{
		estimator::solver::FundamentalMatrixEightPointSolver s;
		std::vector<Model> models;
		const std::vector<size_t> idx = indices(sample, m);
		if (!s.estimateModel(view(pts, n, 4), idx.data(), m, models, weights)) return 0;
		return copy_out(models, 3, 3, 1, F9);
	}
	int gcr_refsrc_solve_e5_nister(const double *pts, int n, const uint32_t *sample, int m, const double *weights, double *E90)
	{
		estimator::solver::EssentialMatrixFivePointNisterSolver s;
		std::vector<Model> models;
		const std::vector<size_t> idx = indices(sample, m);
		if (!s.estimateModel(view(pts, n, 4), idx.data(), m, models, weights)) return 0;
		return copy_out(models, 3, 3, 10, E90);
	}
	int gcr_refsrc_solve_e5_stewenius(const double *pts, int n, const uint32_t *sample, int m, const double *weights, double *E90)
	{
		estimator::solver::EssentialMatrixFivePointSteweniusSolver s;
		std::vector<Model> models;
		const std::vector<size_t> idx = indices(sample, m);
		if (!s.estimateModel(view(pts, n, 4), idx.data(), m, models, weights)) return 0;
		return copy_out(models, 3, 3, 10, E90);
	}
	int gcr_refsrc_solve_p3p(const double *pts5, int n, const uint32_t *sample, double *P48)
	{
		estimator::solver::P3PSolver s;
		std::vector<Model> models;
		const std::vector<size_t> idx = indices(sample, 3);
		if (!s.estimateModel(view(pts5, n, 5), idx.data(), 3, models, nullptr)) return 0;
		return copy_out(models, 3, 4, 4, P48);
	}

	// ---- the estimator classes around them (normalisation, orientation tests, denormalisation, sign) ----
	int gcr_refsrc_h_valid_sample(const double *pts, int n, const uint32_t *sample)
	{
		HEstimator e;
		const std::vector<size_t> idx = indices(sample, 4);
		return e.isValidSample(view(pts, n, 4), idx.data()) ? 1 : 0;
	}
	int gcr_refsrc_refit_h(const double *pts, int n, const uint32_t *sample, int m, const double *weights, double *H9)
	{
		HEstimator e;
		std::vector<Model> models;
		const std::vector<size_t> idx = indices(sample, m);
		if (!e.estimateModelNonminimal(view(pts, n, 4), idx.data(), (size_t)m, &models, weights)) return 0;
		return copy_out(models, 3, 3, 1, H9);
	}
	int gcr_refsrc_estimate_f7(const double *pts, int n, const uint32_t *sample, double *F27)
	{
		FEstimator e(0.0, false);
		std::vector<Model> models;
		const std::vector<size_t> idx = indices(sample, 7);
		if (!e.estimateModel(view(pts, n, 4), idx.data(), &models)) return 0;
		return copy_out(models, 3, 3, 3, F27);
	}
	int gcr_refsrc_refit_f8(const double *pts, int n, const uint32_t *sample, int m, const double *weights, double *F9)
	{
		FEstimator e(0.0, false);
		std::vector<Model> models;
		const std::vector<size_t> idx = indices(sample, m);
		if (!e.estimateModelNonminimal(view(pts, n, 4), idx.data(), (size_t)m, &models, weights)) return 0;
		return copy_out(models, 3, 3, 1, F9);
	}
	int gcr_refsrc_estimate_e5(const double *pts, int n, const uint32_t *sample, double *E90)
	{
		EEstimator e(Eigen::Matrix3d::Identity(), Eigen::Matrix3d::Identity(), 0.0);
		std::vector<Model> models;
		const std::vector<size_t> idx = indices(sample, 5);
		if (!e.estimateModel(view(pts, n, 4), idx.data(), &models)) return 0;
		return copy_out(models, 3, 3, 10, E90);
	}
	int gcr_refsrc_refit_e(const double *pts, int n, const uint32_t *sample, int m, const double *weights, double *E90)
	{
		EEstimator e(Eigen::Matrix3d::Identity(), Eigen::Matrix3d::Identity(), 0.0);
		std::vector<Model> models;
		const std::vector<size_t> idx = indices(sample, m);
		if (!e.estimateModelNonminimal(view(pts, n, 4), idx.data(), (size_t)m, &models, weights)) return 0;
		return copy_out(models, 3, 3, 10, E90);
	}

	// ---- residuals and model tests of the estimator classes ----
	// problem: 0 H (homography_estimator.h:175-199), 1 F Sampson (fundamental_estimator.h:195-222), 2 E Sampson
	// (essential_estimator.h:186-213), 4 F symmetric epipolar (fundamental_estimator.h:225-252)
	void gcr_refsrc_residuals(int problem, const double *pts, int n, const double *model9, double *out)
	{
		Eigen::MatrixXd d(3, 3);
		for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) d(i, j) = model9[3 * i + j];
		const cv::Mat data = view(pts, n, 4);
		HEstimator he;
		FEstimator fe(0.0, false);
		EEstimator ee(Eigen::Matrix3d::Identity(), Eigen::Matrix3d::Identity(), 0.0);
		for (int i = 0; i < n; ++i)
			out[i] = problem == 0 ? he.squaredResidual(data.row(i), d) : problem == 1 ? fe.squaredResidual(data.row(i), d)
			       : problem == 2 ? ee.squaredResidual(data.row(i), d) : fe.squaredSymmetricEpipolarDistance(data.row(i), d);
	}
	// PerspectiveNPointEstimator::squaredResidual (perspective_n_point_estimator.h:148-203): rows x y X Y Z, model 3 x 4
	void gcr_refsrc_residuals_pnp(const double *pts5, int n, const double *model12, double *out)
	{
		Eigen::MatrixXd d(3, 4);
		for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) d(i, j) = model12[4 * i + j];
		const cv::Mat data = view(pts5, n, 5);
		estimator::PerspectiveNPointEstimator<estimator::solver::P3PSolver, estimator::solver::P3PSolver> e;
		for (int i = 0; i < n; ++i) out[i] = e.squaredResidual(data.row(i), d);
	}
	int gcr_refsrc_h_valid_model(const double *H9)
	{
		HEstimator e;
		Model m;
		m.descriptor.resize(3, 3);
		for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m.descriptor(i, j) = H9[3 * i + j];
		bool updated = false;
		return e.isValidModel(m, cv::Mat(), std::vector<size_t>(), nullptr, 0.0, updated) ? 1 : 0;
	}
	// FundamentalMatrixEstimator::isValidModel without DEGENSAC (fundamental_estimator.h:287-338): the symmetric-epipolar ratio test
	int gcr_refsrc_f_valid_model_ratio(const double *F9, const double *pts, int n, const uint32_t *inl, int n_inl, double threshold, double min_ratio)
	{
		FEstimator e(min_ratio, false);
		Model m;
		m.descriptor.resize(3, 3);
		for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m.descriptor(i, j) = F9[3 * i + j];
		const std::vector<size_t> inliers = indices(inl, n_inl);
		bool updated = false;
		return e.isValidModel(m, view(pts, n, 4), inliers, nullptr, threshold, updated) ? 1 : 0;
	}
	// EssentialMatrixEstimator::isValidModel (essential_estimator.h:279-305): the same ratio test over five-point samples
	int gcr_refsrc_e_valid_model_ratio(const double *E9, const double *pts, int n, const uint32_t *inl, int n_inl, double threshold, double min_ratio)
	{
		EEstimator e(Eigen::Matrix3d::Identity(), Eigen::Matrix3d::Identity(), min_ratio);
		Model m = Model();
		m.descriptor.resize(3, 3);
		for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m.descriptor(i, j) = E9[3 * i + j];
		const std::vector<size_t> inliers = indices(inl, n_inl);
		bool updated = false;
		return e.isValidModel(m, view(pts, n, 4), inliers, nullptr, threshold, updated) ? 1 : 0;
	}
	// solver_fundamental_matrix_plane_and_parallax.h:105-160 (two correspondences + the plane's homography); `oriented`: through
	// FundamentalMatrixEstimator<PlaneParallax, EightPoint>::estimateModel, i.e. with the oriented epipolar constraint
	int gcr_refsrc_solve_f_plane_parallax(const double *pts, int n, const uint32_t *sample, const double *H9, int oriented, double *F9)
	{
		Eigen::Matrix3d Hm;
		for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Hm(i, j) = H9[3 * i + j];
		std::vector<Model> models;
		const std::vector<size_t> idx = indices(sample, 2);
		if (oriented)
		{
			estimator::FundamentalMatrixEstimator<estimator::solver::FundamentalMatrixPlaneParallaxSolver,
			                                      estimator::solver::FundamentalMatrixEightPointSolver> e(0.0, false);
			e.getMutableMinimalSolver()->setHomography(&Hm);
			if (!e.estimateModel(view(pts, n, 4), idx.data(), &models)) return 0;
		}
		else
		{
			estimator::solver::FundamentalMatrixPlaneParallaxSolver s;
			s.setHomography(&Hm);
			if (!s.estimateModel(view(pts, n, 4), idx.data(), 2, models, nullptr)) return 0;
		}
		return copy_out(models, 3, 3, 1, F9);
	}

	// ---- bundle adjustment (relative_pose/bundle.cpp, jacobian_impl.h, essential.cpp compiled verbatim beside this file) ----
	// the default F estimator's non-minimal fit (types.h:78-80): estimateModelNonminimal with the BA solver = normalise,
	// eight-point, LM in the Hartley frame, denormalise (fundamental_estimator.h:574-618, solver_fundamental_matrix_bundle_adjustment.h:113-177)
	int gcr_refsrc_refit_f_ba(const double *pts, int n, const uint32_t *sample, int m, const double *weights, double *F9)
	{
		estimator::FundamentalMatrixEstimator<estimator::solver::FundamentalMatrixSevenPointSolver,
		                                      estimator::solver::FundamentalMatrixBundleAdjustmentSolver> e(0.0, false);
		std::vector<Model> models;
		const std::vector<size_t> idx = indices(sample, m);
		if (!e.estimateModelNonminimal(view(pts, n, 4), idx.data(), (size_t)m, &models, weights)) return 0;
		return copy_out(models, 3, 3, 1, F9);
	}
	// the post-run polish of findFundamentalMatrix_ (py/gcransac_python.cpp:1851-1858): the BA solver on the inliers, pixel frame
	int gcr_refsrc_polish_f_ba(const double *pts, int n, const uint32_t *inliers, int m, const double *weights, double *F9)
	{
		estimator::solver::FundamentalMatrixBundleAdjustmentSolver s;
		std::vector<Model> models;
		const std::vector<size_t> idx = indices(inliers, m);
		if (!s.estimateModel(view(pts, n, 4), idx.data(), (size_t)m, models, weights)) return 0;
		return copy_out(models, 3, 3, 1, F9);
	}
	// solver_essential_matrix_bundle_adjustment.h:113-186: Stewenius models -> poses -> LM of the relative pose; every refined
	// model, in the solver's order
	int gcr_refsrc_polish_e_ba(const double *pts, int n, const uint32_t *inliers, int m, const double *weights, int max_models, double *E)
	{
		estimator::solver::EssentialMatrixBundleAdjustmentSolver s;
		std::vector<Model> models;
		const std::vector<size_t> idx = indices(inliers, m);
		if (!s.estimateModel(view(pts, n, 4), idx.data(), (size_t)m, models, weights)) return 0;
		return copy_out(models, 3, 3, max_models, E);
	}

	// ---- the oracle's RESTATED five-point solver over the reference's own root finder (maths/sturm_polynomial_solver.h,
	// compiled verbatim above): what is left between this and gcr_refsrc_solve_e5_nister is the restatement itself ----
	static int reference_sturm(const double *c11, double *roots10) { return poselib::sturm::bisect_sturm<10>(c11, roots10); }
	int gcr_refsrc_oracle_e5_with_reference_sturm(const double *pts, const uint32_t *sample, int estimate, double *E90)
	{
		gcro_e5_root_finder = reference_sturm;
		const int k = estimate ? gcro_estimate_e5(pts, sample, E90) : gcro_solve_e5_nister(pts, sample, E90);
		gcro_e5_root_finder = nullptr;
		return k;
	}
}

// ---- whole runs: the reference's GCRANSAC<> (verbatim) over (a) the reference's OWN estimator classes, (b) an adapter over
// the oracle's restated solvers -- the Estimator the verbatim driver of oracle/_ref/libgcr_ref.so runs with (ref_glue.cpp's
// OracleEstimator, restated here for this library's column-major matrices).  Same counter-based samplers in both (the
// reference's draw from std::mt19937 seeded by std::random_device), same grid graph, MSAC.  Equal runs pin what the adapter
// restates by hand: sampleSize / nonMinimalSampleSize / inlierLimit / isWeightingApplicable, which checks run where. ----
namespace
{
	class CounterSampler : public sampler::Sampler<cv::Mat, size_t>
	{
		uint64_t seed, next;
		uint32_t stream;
		std::vector<uint32_t> tmp;

	public:
		CounterSampler(const cv::Mat *c, uint64_t seed_, uint32_t stream_) : sampler::Sampler<cv::Mat, size_t>(c), seed(seed_), next(0), stream(stream_) { initialized = true; }
		const std::string getName() const { return "counter-based uniform sampler"; }
		void reset() {}
		void update(const size_t *const, const size_t &, const size_t &, const double &) {}
		bool initialize(const cv::Mat *const) { return true; }
		uint64_t drawn() const { return next; }
		bool sample(const std::vector<size_t> &pool_, size_t *const subset_, size_t sample_size_)
		{
			if (sample_size_ > pool_.size()) return false;
			tmp.resize(sample_size_);
			gcro_sample_uniform(seed, next++, stream, (uint32_t)pool_.size(), (int)sample_size_, tmp.data());
			for (size_t i = 0; i < sample_size_; ++i) subset_[i] = pool_[tmp[i]];
			return true;
		}
	};

	void flat_of(const Eigen::MatrixXd &d, double *out)
	{
		for (int i = 0; i < d.rows(); ++i) for (int j = 0; j < d.cols(); ++j) out[i * d.cols() + j] = d(i, j);
	}
	Model model_of(const double *in)
	{
		Model m;
		m.descriptor.resize(3, 3);
		for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) m.descriptor(i, j) = in[3 * i + j];
		return m;
	}

	constexpr int GCRO_FPP = 100; // DEGENSAC's inner estimator: plane-and-parallax minimal solver, eight-point non-minimal solver (fundamental_estimator.h:532-541)
	template <int PROBLEM> class RestatedEstimator
	{
	public:
		bool f_nonminimal_ba = false;
		double min_ratio = 0.5;
		const double *pp_homography = nullptr;
		static constexpr size_t sampleSize() { return PROBLEM == GCRO_H ? 4 : PROBLEM == GCRO_F ? 7 : PROBLEM == GCRO_FPP ? 2 : 5; }
		// FundamentalMatrixBundleAdjustmentSolver::sampleSize() is 7 (it falls back to the 7-point solver), the 8-point solver's 8
		size_t nonMinimalSampleSizeRuntime() const { return PROBLEM == GCRO_H ? 4 : PROBLEM == GCRO_F ? (f_nonminimal_ba ? 7 : 8) : PROBLEM == GCRO_FPP ? 8 : 5; }
		static constexpr size_t nonMinimalSampleSize() { return PROBLEM == GCRO_H ? 4 : (PROBLEM == GCRO_F || PROBLEM == GCRO_FPP) ? 8 : 5; }
		static constexpr size_t maximumMinimalSolutions() { return (PROBLEM == GCRO_H || PROBLEM == GCRO_FPP) ? 1 : PROBLEM == GCRO_F ? 3 : 10; }
		static constexpr bool isWeightingApplicable() { return true; }
		size_t inlierLimit() const { return 7 * sampleSize(); }
		double squaredResidual(const cv::Mat &point_, const Eigen::MatrixXd &d) const
		{
			double flat[9];
			flat_of(d, flat);
			const double *pt = reinterpret_cast<const double *>(point_.data);
			return PROBLEM == GCRO_H ? gcro_sqres_h(pt, flat) : gcro_sqres_sampson(pt, flat);
		}
		double squaredResidual(const cv::Mat &point_, const Model &m) const { return squaredResidual(point_, m.descriptor); }
		double residual(const cv::Mat &point_, const Eigen::MatrixXd &d) const { return sqrt(squaredResidual(point_, d)); }
		double residual(const cv::Mat &point_, const Model &m) const { return residual(point_, m.descriptor); }
		bool isValidSample(const cv::Mat &data_, const size_t *sample_) const
		{
			if (PROBLEM != GCRO_H) return true;
			uint32_t s4[4] = {(uint32_t)sample_[0], (uint32_t)sample_[1], (uint32_t)sample_[2], (uint32_t)sample_[3]};
			return gcro_h_valid_sample(reinterpret_cast<const double *>(data_.data), s4) != 0;
		}
		bool estimateModel(const cv::Mat &data_, const size_t *sample_, std::vector<Model> *models_) const
		{
			uint32_t s8[8];
			for (size_t i = 0; i < sampleSize(); ++i) s8[i] = (uint32_t)sample_[i];
			const double *pts = reinterpret_cast<const double *>(data_.data);
			double out[90];
			const int count = PROBLEM == GCRO_H ? gcro_solve_h4(pts, data_.cols, s8, out) : PROBLEM == GCRO_F ? gcro_estimate_f7(pts, s8, out)
			                  : PROBLEM == GCRO_FPP ? gcro_estimate_f_plane_parallax(pts, s8, pp_homography, out) : gcro_estimate_e5(pts, s8, out);
			for (int i = 0; i < count; ++i) models_->emplace_back(model_of(out + 9 * i));
			return PROBLEM == GCRO_H ? count > 0 : models_->size() > 0;
		}
		bool estimateModelNonminimal(const cv::Mat &data_, const size_t *sample_, const size_t &n_, std::vector<Model> *models_, const double *weights_ = nullptr) const
		{
			if (n_ < nonMinimalSampleSizeRuntime()) return false;
			std::vector<uint32_t> sv(n_);
			for (size_t i = 0; i < n_; ++i) sv[i] = (uint32_t)sample_[i];
			const double *pts = reinterpret_cast<const double *>(data_.data);
			double out[90];
			const int count = PROBLEM == GCRO_H   ? gcro_refit_h(pts, sv.data(), (int)n_, weights_, out)
			                  : PROBLEM == GCRO_F ? (f_nonminimal_ba ? gcro_refit_f_ba(pts, sv.data(), (int)n_, weights_, out) : gcro_refit_f8(pts, sv.data(), (int)n_, weights_, out))
			                  : PROBLEM == GCRO_FPP ? gcro_refit_f8(pts, sv.data(), (int)n_, weights_, out)
			                                      : gcro_refit_e(pts, sv.data(), (int)n_, weights_, out);
			for (int i = 0; i < count; ++i) models_->emplace_back(model_of(out + 9 * i));
			return count > 0;
		}
		bool isValidModel(Model &model, const cv::Mat &data, const std::vector<size_t> &inliers, const size_t *, const double threshold_, bool &model_updated_) const
		{
			model_updated_ = false;
			double flat[9];
			flat_of(model.descriptor, flat);
			if (PROBLEM == GCRO_H) return gcro_h_valid_model(flat) != 0;
			std::vector<uint32_t> inl(inliers.begin(), inliers.end());
			return gcro_f_valid_model_ratio(flat, reinterpret_cast<const double *>(data.data), inl.data(), (int)inl.size(), (int)sampleSize(), threshold_, min_ratio) != 0;
		}
	};

	template <class Est>
	int run_with(Est &estimator, const cv::Mat &points, const double *cell_sizes, int cells, double threshold, double conf, double lambda,
	             int max_iters, int min_iters, int lo_number, uint64_t seed, double *model_out, uint8_t *mask, uint64_t *stats, double *score)
	{
		CounterSampler main_sampler(&points, seed, 0u), lo_sampler(&points, seed, 1u);
		std::vector<double> cs(cell_sizes, cell_sizes + 4);
		neighborhood::GridNeighborhoodGraph<4> graph(&points, cs, (size_t)cells);
		Model model;
		GCRANSAC<Est, neighborhood::GridNeighborhoodGraph<4>> gcransac;
		gcransac.settings.threshold = threshold;
		gcransac.settings.spatial_coherence_weight = lambda;
		gcransac.settings.confidence = conf;
		gcransac.settings.max_local_optimization_number = lo_number;
		gcransac.settings.max_iteration_number = max_iters;
		gcransac.settings.min_iteration_number = min_iters;
		gcransac.settings.neighborhood_sphere_radius = cells;
		gcransac.run(points, estimator, &main_sampler, &lo_sampler, &graph, model);
		const utils::RANSACStatistics &st = gcransac.getRansacStatistics();
		for (int i = 0; i < points.rows; ++i) mask[i] = 0;
		for (size_t idx : st.inliers) mask[idx] = 1;
		if (st.inliers.size() > 0) flat_of(model.descriptor, model_out);
		stats[0] = st.iteration_number; stats[1] = st.graph_cut_number; stats[2] = st.local_optimization_number;
		stats[3] = st.neighbor_number; stats[4] = main_sampler.drawn(); stats[5] = lo_sampler.drawn();
		*score = st.score;
		return (int)st.inliers.size();
	}
}

extern "C" int gcr_refsrc_run(int problem, int own_estimators, const double *pts, int n, const double *cell_sizes, int cells, double threshold,
                              double conf, double lambda, int max_iters, int min_iters, int lo_number, uint64_t seed, int f_ba,
                              double *model_out, uint8_t *mask, uint64_t *stats, double *score)
{
	const cv::Mat points = view(pts, n, 4);
#define GCR_RUN_ARGS points, cell_sizes, cells, threshold, conf, lambda, max_iters, min_iters, lo_number, seed, model_out, mask, stats, score
	if (own_estimators)
	{
		if (problem == GCRO_H) { HEstimator e; return run_with(e, GCR_RUN_ARGS); }
		if (problem == GCRO_F && !(f_ba & 1)) { FEstimator e(0.5, false); return run_with(e, GCR_RUN_ARGS); }
		if (problem == GCRO_F)
		{
			estimator::FundamentalMatrixEstimator<estimator::solver::FundamentalMatrixSevenPointSolver,
			                                      estimator::solver::FundamentalMatrixBundleAdjustmentSolver> e(0.5, false);
			return run_with(e, GCR_RUN_ARGS);
		}
		if (problem == GCRO_E) { EEstimator e(Eigen::Matrix3d::Identity(), Eigen::Matrix3d::Identity(), 0.1); return run_with(e, GCR_RUN_ARGS); }
		return -1;
	}
	if (problem == GCRO_H) { RestatedEstimator<GCRO_H> e; return run_with(e, GCR_RUN_ARGS); }
	if (problem == GCRO_F) { RestatedEstimator<GCRO_F> e; e.f_nonminimal_ba = (f_ba & 1) != 0; return run_with(e, GCR_RUN_ARGS); }
	if (problem == GCRO_E) { RestatedEstimator<GCRO_E> e; e.min_ratio = 0.1; return run_with(e, GCR_RUN_ARGS); }
	return -1;
#undef GCR_RUN_ARGS
}

// ---- DEGENSAC: FundamentalMatrixEstimator::isValidModel with use_degensac (fundamental_estimator.h:287-572), the reference's own
// code (own = 1) against the restatement the verbatim driver's adapter runs (own = 0: gcro_f_valid_model_ratio, degensac_front.h,
// then the same inner run -- the reference's GCRANSAC<> and the reference's UniformSampler, settings :543-553 -- over the
// restated plane-and-parallax estimator).  Both inner runs draw from the same word stream (ref_rng_redirect.h, begun at `key`).
// Returns bit 0: the verdict, bit 1: model_updated_; the (possibly replaced) model in model_out. ----
extern "C" int gcr_refsrc_degensac(int own, const double *F9, const double *pts, int n, const uint32_t *inl, int n_inl, const uint32_t *sample7,
                                   double threshold, uint64_t key, double *model_out)
{
	const cv::Mat data = view(pts, n, 4);
	const std::vector<size_t> inliers = indices(inl, n_inl), sample = indices(sample7, 7);
	Model model = model_of(F9);
	bool updated = false, verdict;
	gcr_ref_words::begin(key, 0, 7u);
	if (own)
	{
		FEstimator e(0.5, true);
		verdict = e.isValidModel(model, data, inliers, sample.data(), threshold, updated);
	}
	else
	{
		verdict = gcro_f_valid_model_ratio(F9, pts, inl, n_inl, 7, threshold, 0.5) != 0;
		if (verdict)
		{
			double Hn[9];
			const int front = gcro_degensac_front(F9, pts, inliers, sample.data(), Hn);
			if (front < 0)
				verdict = false;
			else if (front > 0)
			{
				RestatedEstimator<GCRO_FPP> inner;
				inner.pp_homography = Hn;
				inner.min_ratio = 0.0;
				GCRANSAC<RestatedEstimator<GCRO_FPP>, neighborhood::GridNeighborhoodGraph<4>> gcransac;
				gcransac.settings.threshold = threshold;
				gcransac.settings.spatial_coherence_weight = 0;
				gcransac.settings.confidence = 0.99;
				gcransac.settings.neighborhood_sphere_radius = 8;
				sampler::UniformSampler uniform(&data);
				Model inner_model;
				gcransac.run(data, inner, &uniform, &uniform, nullptr, inner_model);
				if (gcransac.getRansacStatistics().inliers.size() > inliers.size())
				{
					updated = true;
					model = inner_model;
				}
			}
		}
	}
	flat_of(model.descriptor, model_out);
	return (verdict ? 1 : 0) | (updated ? 2 : 0);
}
