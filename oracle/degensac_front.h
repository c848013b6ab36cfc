// oracle/degensac_front.h -- TEST INFRASTRUCTURE ONLY.  The deterministic first part of DEGENSAC
// (FundamentalMatrixEstimator::applyDegensac, estimators/fundamental_estimator.h:341-528), restated over the oracle's
// functions; shared by the verbatim driver's Estimator adapter (oracle/ref_glue.cpp: what the GPU drivers are compared
// with) and by oracle/ref_solvers_glue.cpp, which holds it against the reference's own applyDegensac.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

#include "gcr_oracle.h"

// F9: the model under test, minimal_sample: its seven points, inliers: its inliers.
// Returns 0 when the sample is not H-degenerate (the model stands, :571), -1 when DEGENSAC rejects the model (fewer than
// four homography inliers :512, or the least-squares homography cannot be fitted :523), 1 with the plane homography refitted
// on its inliers in Hn9 (:501-528) -- the input of the inner plane-and-parallax run.
inline int gcro_degensac_front(const double *F9, const double *pts, const std::vector<size_t> &inliers, const size_t *minimal_sample, double *Hn9)
{
	const double squared_homography_threshold = 2.0 * 2.0; // homography_threshold_ = 2.0 (:93)
	double H[9];
	uint32_t s7[7];
	for (int i = 0; i < 7; ++i)
		s7[i] = (uint32_t)minimal_sample[i];
	if (!gcro_degensac_sample_homography(F9, pts, s7, squared_homography_threshold, H))
		return 0;
	std::vector<uint32_t> h_inliers; // :501-509
	for (const size_t &idx : inliers)
		if (gcro_sqres_h(pts + 4 * idx, H) < squared_homography_threshold)
			h_inliers.push_back((uint32_t)idx);
	if (h_inliers.size() < 4) // homography_estimator.nonMinimalSampleSize() (:512)
		return -1;
	if (gcro_refit_h(pts, h_inliers.data(), (int)h_inliers.size(), nullptr, Hn9) != 1) // :520-528
		return -1;
	return 1;
}
