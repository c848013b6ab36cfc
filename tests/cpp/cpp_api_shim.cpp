// tests/cpp/cpp_api_shim.cpp -- TEST INFRASTRUCTURE.  A C++ translation unit that calls the reference's C++ driver entry
// points exactly as a C++ user of include/gcransac_python.h would (std::vector arguments, the reference's parameter
// order), through graph-cut-ransac_b200/host/gcr_gcransac_python.h; a flat C function hands the results to the Python
// tests.  Built by graph-cut-ransac_b200/host/build_driver.sh into host/_ref/libgcr_cpp_api_shim.so.
#include "gcr_gcransac_python.h"

extern "C" int shim_find(int problem, const double *pts, int n, const double *probabilities, int n_probabilities, const double *K1,
                         const double *K2, int h1, int w1, int h2, int w2, double spatial_coherence_weight, double threshold,
                         double conf, int max_iters, int min_iters, int sampler_id, int neighborhood_id, double neighborhood_size,
                         int use_space_partitioning, double sampler_variance, int lo_number, unsigned long long seed,
                         double *model_out, unsigned char *mask_out, int *model_len, int *mask_len)
{
	const size_t cols = problem == GCR_PNP ? 5 : 4;
	std::vector<double> correspondences(pts, pts + cols * (size_t)n), point_probabilities(probabilities, probabilities + n_probabilities);
	std::vector<bool> inliers;
	std::vector<double> model;
	int count;
	if (problem == GCR_PNP)
		count = find6DPose_(correspondences, point_probabilities, inliers, model, spatial_coherence_weight, threshold, conf, max_iters,
		                    min_iters, true, 0.00001, sampler_id, neighborhood_id, neighborhood_size, sampler_variance, lo_number, seed);
	else if (problem == GCR_H)
		count = findHomography_(correspondences, point_probabilities, inliers, model, h1, w1, h2, w2, spatial_coherence_weight, threshold,
		                        conf, max_iters, min_iters, true, 0.00001, sampler_id, neighborhood_id, neighborhood_size,
		                        use_space_partitioning != 0, sampler_variance, lo_number, seed);
	else if (problem == GCR_F)
		count = findFundamentalMatrix_(correspondences, point_probabilities, inliers, model, h1, w1, h2, w2, spatial_coherence_weight,
		                               threshold, conf, max_iters, min_iters, true, 0.00001, sampler_id, neighborhood_id,
		                               neighborhood_size, sampler_variance, lo_number, seed);
	else
	{
		std::vector<double> k1(K1, K1 + 9), k2(K2, K2 + 9);
		count = findEssentialMatrix_(correspondences, point_probabilities, inliers, model, k1, k2, h1, w1, h2, w2,
		                             spatial_coherence_weight, threshold, conf, max_iters, min_iters, true, 0.01, sampler_id,
		                             neighborhood_id, neighborhood_size, sampler_variance, lo_number, seed);
	}
	*model_len = (int)model.size();
	*mask_len = (int)inliers.size();
	for (size_t i = 0; i < model.size(); ++i)
		model_out[i] = model[i];
	for (size_t i = 0; i < inliers.size(); ++i)
		mask_out[i] = inliers[i] ? 1 : 0;
	return count;
}
