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

#include "gcr_oracle.h"

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
