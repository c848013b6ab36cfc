// gcr_gcransac_python.h -- the reference's C++ driver entry points (include/gcransac_python.h:264, 411, 676; definitions
// py/gcransac_python.cpp:1600-1921, 2519-2855, 4221-4550) with their own names, parameter lists and output conventions,
// on top of the batched B200 driver (gcr_driver.h).  A C++ caller of findHomography_ / findFundamentalMatrix_ /
// findEssentialMatrix_ includes this header instead of the reference's and links libgcr_driver.so + libgcr_b200.so.
//
// Kept from the reference: the inlier count as return value (0 = failure or no model, message on stderr); `inliers`
// resized to the number of correspondences, the model resized to 9 row-major doubles (py/gcransac_python.cpp:4520-4540);
// sampler_id 0-4 (uniform, PROSAC, P-NAPSAC, NG-RANSAC importance sampler, AR-Sampler); neighborhood_id 0 = grid with
// `neighborhood_size` cells per axis (H: cell = (size + eps) / cells, :4290-4297; F / E: size / cells, :1672-1677),
// 1 = radius graph (the exact one, where the reference asks FLANN); F = the default estimator (BA non-minimal solver,
// DEGENSAC, post-run BA); E normalises by the intrinsics itself and returns GC-RANSAC's own model (its post-run BA reads
// out of bounds in the reference).  use_sprt / use_space_partitioning / min_inlier_ratio_for_sprt
// are accepted and do not change the scoring: every model is scored on every point; the one side
// effect of use_space_partitioning, forcing the grid neighbourhood (py/gcransac_python.cpp:4282-4286), is kept.
// Added, defaulted, after the reference's parameters: `seed` (the reference seeds from std::random_device) and `device`.
// One gcr_ctx per calling thread and device, created on first use, released at thread exit.
#ifndef GCR_GCRANSAC_PYTHON_H_INCLUDED
#define GCR_GCRANSAC_PYTHON_H_INCLUDED

#include <cstdint>
#include <cstdio>
#include <limits>
#include <map>
#include <vector>

#include "gcr_driver.h"

namespace gcr_api
{
	struct ThreadContexts
	{
		std::map<int, gcr_ctx *> by_device;
		~ThreadContexts()
		{
			for (auto &kv : by_device)
				gcr_ctx_destroy(kv.second);
		}
		gcr_ctx *get(int device)
		{
			auto it = by_device.find(device);
			if (it != by_device.end())
				return it->second;
			gcr_ctx *ctx = nullptr;
			if (gcr_ctx_create(device, &ctx) != GCR_OK)
				return nullptr; // no CUDA device: there is no CPU fallback
			by_device[device] = ctx;
			return ctx;
		}
	};
	inline gcr_ctx *context(int device)
	{
		static thread_local ThreadContexts contexts;
		return contexts.get(device);
	}

	// problem: GCR_H / GCR_F / GCR_E; K1, K2: row-major 3 x 3 intrinsics (E only)
	inline int find(int problem, const char *name, std::vector<double> &correspondences, std::vector<double> &point_probabilities,
	                std::vector<bool> &inliers, std::vector<double> &model, const double *K1, const double *K2, int h1, int w1,
	                int h2, int w2, double spatial_coherence_weight, double threshold, double conf, int max_iters, int min_iters,
	                int sampler_id, int neighborhood_id, double neighborhood_size, double sampler_variance, int lo_number,
	                uint64_t seed, int device)
	{
		const size_t num_tents = correspondences.size() / 4;
		const size_t minimal = problem == GCR_H ? 4 : problem == GCR_F ? 7 : 5;
		if (num_tents < minimal)
		{
			fprintf(stderr, "%s: %zu correspondences, at least %zu are needed\n", name, num_tents, minimal);
			return 0;
		}
		if (sampler_id < 0 || sampler_id > 4)
		{
			fprintf(stderr, "Unknown sampler identifier: %d. The accepted samplers are 0 (uniform sampling), 1 (PROSAC sampling), "
			                "2 (P-NAPSAC sampling), 3 (NG-RANSAC sampling), 4 (AR-Sampler)\n", sampler_id);
			return 0;
		}
		if ((sampler_id == 3 || sampler_id == 4) && point_probabilities.size() != num_tents)
		{
			fprintf(stderr, "The number of correspondences (%zu) and the number of provided probabilities (%zu) do not match.\n",
			        num_tents, point_probabilities.size());
			return 0;
		}
		if (neighborhood_id != 0 && neighborhood_id != 1)
		{
			fprintf(stderr, "Unknown neighborhood-graph identifier: %d. The accepted values are 0 (grid), 1 (radius)\n", neighborhood_id);
			return 0;
		}
		gcr_ctx *ctx = context(device);
		if (!ctx)
		{
			fprintf(stderr, "%s: no usable CUDA device %d (there is no CPU fallback)\n", name, device);
			return 0;
		}
		gcr_run_settings s;
		gcr_run_settings_default(&s);
		s.threshold = threshold;
		s.confidence = conf;
		s.spatial_coherence_weight = spatial_coherence_weight;
		s.max_iteration_number = (uint64_t)max_iters;
		s.min_iteration_number = (uint64_t)min_iters;
		s.max_local_optimization_number = (uint64_t)lo_number;
		s.e_bundle_adjustment = problem == GCR_E ? 1 : 0; // findEssentialMatrix_ ends with its relative-pose BA (py/gcransac_python.cpp:2766-2821)
		const size_t cells = static_cast<size_t>(neighborhood_size);
		s.cells_per_axis = cells > 0 ? cells : 1;
		const double eps = problem == GCR_H ? std::numeric_limits<double>::epsilon() : 0.0;
		const double sizes[4] = {(double)w1, (double)h1, (double)w2, (double)h2};
		for (int d = 0; d < 4; ++d)
		{
			s.cell_sizes[d] = (sizes[d] + eps) / (double)s.cells_per_axis;
			s.image_sizes[d] = sizes[d];
		}
		s.seed = seed;
		s.sampler = sampler_id;
		s.sampler_variance = sampler_variance;
		s.point_probabilities = (sampler_id == 3 || sampler_id == 4) ? point_probabilities.data() : nullptr;
		s.neighborhood = neighborhood_id;
		s.neighborhood_radius = neighborhood_size;
		s.use_degensac = problem == GCR_F ? 1 : 0;
		s.f_bundle_adjustment = problem == GCR_F ? 3 : 0;
		std::vector<uint8_t> mask(num_tents, 0);
		double m[9] = {0};
		const int n_inl = problem == GCR_E ? gcr_find_essential(ctx, correspondences.data(), (int)num_tents, K1, K2, &s, m, mask.data(), nullptr)
		                                   : gcr_find_model(ctx, problem, correspondences.data(), (int)num_tents, &s, m, mask.data(), nullptr);
		if (n_inl < 0)
			return 0; // the driver has printed its message
		inliers.resize(num_tents);
		for (size_t i = 0; i < num_tents; ++i)
			inliers[i] = mask[i] != 0;
		model.resize(9);
		for (int i = 0; i < 9; ++i)
			model[i] = m[i];
		return n_inl;
	}
}

// include/gcransac_python.h:676-728
inline int findHomography_(std::vector<double> &correspondences, std::vector<double> &point_probabilities, std::vector<bool> &inliers,
                           std::vector<double> &homography, int h1, int w1, int h2, int w2, double spatial_coherence_weight,
                           double threshold, double conf, int max_iters, int min_iters, bool use_sprt,
                           double min_inlier_ratio_for_sprt, int sampler_id, int neighborhood_id, double neighborhood_size,
                           bool use_space_partitioning, double sampler_variance, int lo_number, uint64_t seed = 1, int device = 0)
{
	(void)use_sprt; (void)min_inlier_ratio_for_sprt;
	if (use_space_partitioning && neighborhood_id != 0) // py/gcransac_python.cpp:4282-4286
	{
		fprintf(stderr, "Space Partitioning only works with Grid neighorbood yet. Thus, setting neighborhood_id = 0.\n");
		neighborhood_id = 0;
	}
	return gcr_api::find(GCR_H, "findHomography_", correspondences, point_probabilities, inliers, homography, nullptr, nullptr, h1, w1,
	                     h2, w2, spatial_coherence_weight, threshold, conf, max_iters, min_iters, sampler_id, neighborhood_id,
	                     neighborhood_size, sampler_variance, lo_number, seed, device);
}

// include/gcransac_python.h:264-310
inline int findFundamentalMatrix_(std::vector<double> &correspondences, std::vector<double> &point_probabilities,
                                  std::vector<bool> &inliers, std::vector<double> &fundamental_matrix, int h1, int w1, int h2, int w2,
                                  double spatial_coherence_weight, double threshold, double conf, int max_iters, int min_iters,
                                  bool use_sprt, double min_inlier_ratio_for_sprt, int sampler_id, int neighborhood_id,
                                  double neighborhood_size, double sampler_variance, int lo_number, uint64_t seed = 1, int device = 0)
{
	(void)use_sprt; (void)min_inlier_ratio_for_sprt;
	return gcr_api::find(GCR_F, "findFundamentalMatrix_", correspondences, point_probabilities, inliers, fundamental_matrix, nullptr,
	                     nullptr, h1, w1, h2, w2, spatial_coherence_weight, threshold, conf, max_iters, min_iters, sampler_id,
	                     neighborhood_id, neighborhood_size, sampler_variance, lo_number, seed, device);
}

// include/gcransac_python.h:411-461
inline int findEssentialMatrix_(std::vector<double> &correspondences, std::vector<double> &point_probabilities,
                                std::vector<bool> &inliers, std::vector<double> &essential_matrix,
                                std::vector<double> &source_intrinsics, std::vector<double> &destination_intrinsics, int h1, int w1,
                                int h2, int w2, double spatial_coherence_weight, double threshold, double conf, int max_iters,
                                int min_iters, bool use_sprt, double min_inlier_ratio_for_sprt, int sampler_id, int neighborhood_id,
                                double neighborhood_size, double sampler_variance, int lo_number, uint64_t seed = 1, int device = 0)
{
	(void)use_sprt; (void)min_inlier_ratio_for_sprt;
	if (source_intrinsics.size() != 9 || destination_intrinsics.size() != 9)
	{
		fprintf(stderr, "findEssentialMatrix_: the intrinsics should be 3 x 3 row-major matrices\n");
		return 0;
	}
	return gcr_api::find(GCR_E, "findEssentialMatrix_", correspondences, point_probabilities, inliers, essential_matrix,
	                     source_intrinsics.data(), destination_intrinsics.data(), h1, w1, h2, w2, spatial_coherence_weight, threshold,
	                     conf, max_iters, min_iters, sampler_id, neighborhood_id, neighborhood_size, sampler_variance, lo_number, seed,
	                     device);
}
#endif

// include/gcransac_python.h:177-211 (py/gcransac_python.cpp:1026-1272): absolute pose from 2D-3D matches, rows u v X Y Z
// with K-normalised image points; pose = 12 doubles, the row-major 3 x 4 [R | t].  neighborhood_id 0 = the radius
// neighbourhood over the 5-column rows (FLANN in the reference, exact here).  Non-minimal fits: normalised DLT + the LM of
// refine_pnp (the reference's EPnP initialiser is OpenCV's; its accumulator as written has no defined behaviour).
inline int find6DPose_(std::vector<double> &correspondences, std::vector<double> &point_probabilities, std::vector<bool> &inliers,
                       std::vector<double> &pose, double spatial_coherence_weight, double threshold, double conf, int max_iters,
                       int min_iters, bool use_sprt, double min_inlier_ratio_for_sprt, int sampler_id, int neighborhood_id,
                       double neighborhood_size, double sampler_variance, int lo_number, uint64_t seed = 1, int device = 0)
{
	(void)use_sprt; (void)min_inlier_ratio_for_sprt;
	const size_t num_tents = correspondences.size() / 5;
	if (num_tents < 4)
	{
		fprintf(stderr, "find6DPose_: %zu correspondences, at least 4 are needed\n", num_tents);
		return 0;
	}
	if (neighborhood_id != 0) // py/gcransac_python.cpp:1093-1098
	{
		fprintf(stderr, "Unknown neighborhood-graph identifier: %d. The accepted values are 0 (FLANN-based neighborhood)\n", neighborhood_id);
		return 0;
	}
	if (sampler_id == 2) // :1113-1120
	{
		fprintf(stderr, "Progressive NAPSAC is not usable for the absolute pose problem.\n");
		return 0;
	}
	if (sampler_id < 0 || sampler_id > 4)
	{
		fprintf(stderr, "Unknown sampler identifier: %d. The accepted samplers are 0 (uniform sampling), 1 (PROSAC sampling), 2 (P-NAPSAC sampling)\n", sampler_id);
		return 0;
	}
	if ((sampler_id == 3 || sampler_id == 4) && point_probabilities.size() != num_tents)
	{
		fprintf(stderr, "The number of correspondences (%zu) and the number of provided probabilities (%zu) do not match.\n", num_tents,
		        point_probabilities.size());
		return 0;
	}
	gcr_ctx *ctx = gcr_api::context(device);
	if (!ctx)
	{
		fprintf(stderr, "find6DPose_: no usable CUDA device %d (there is no CPU fallback)\n", device);
		return 0;
	}
	gcr_run_settings s;
	gcr_run_settings_default(&s);
	s.threshold = threshold;
	s.confidence = conf;
	s.spatial_coherence_weight = spatial_coherence_weight;
	s.max_iteration_number = (uint64_t)max_iters;
	s.min_iteration_number = (uint64_t)min_iters;
	s.max_local_optimization_number = (uint64_t)lo_number;
	s.seed = seed;
	s.sampler = sampler_id;
	s.sampler_variance = sampler_variance;
	s.point_probabilities = (sampler_id == 3 || sampler_id == 4) ? point_probabilities.data() : nullptr;
	s.neighborhood = 1;
	s.neighborhood_radius = neighborhood_size;
	s.pnp_bundle_adjustment = 1; // :1227-1238
	std::vector<uint8_t> mask(num_tents, 0);
	double m[12] = {0};
	const int n_inl = gcr_find_model(ctx, GCR_PNP, correspondences.data(), (int)num_tents, &s, m, mask.data(), nullptr);
	if (n_inl < 0)
		return 0;
	inliers.resize(num_tents);
	for (size_t i = 0; i < num_tents; ++i)
		inliers[i] = mask[i] != 0;
	pose.assign(m, m + 12);
	return n_inl;
}
