// gcr_backend.h -- the device operations of the batched host driver (gcr_driver.cpp) as ONE interface with two
// implementations:
//   DirectBackend   one image pair on its own gcr_ctx: every operation is a few C-ABI calls (include/gcr/gcr.h) + one
//                   synchronise (what gcr_find_model / gcr_find_essential use);
//   BatchedBackend  (gcr_batch_engine.cpp) many pairs resident on ONE gcr_ctx: a driver thread per pair posts the same
//                   operations as requests, the engine groups the pending requests by kind and executes each group as one
//                   launch (+ one synchronise) for all pairs -- the lock-step batch of SURVEY.md 7.1 step 9.
// The driver's control flow (the sequential replay of GCRANSAC::run, GCRANSAC.h:202-627) is the same code over both.
#ifndef GCR_BACKEND_H_INCLUDED
#define GCR_BACKEND_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "gcr/gcr.h"

namespace gcr_host
{
	// one main-loop batch of one pair: samples [first, first + B) -> model slots
	struct BatchSpec
	{
		int variant = 0;                 // 1: DEGENSAC's inner run (plane-and-parallax solver over pp_H, 2-point samples)
		const double *pp_H = nullptr;
		int sampler = 0;                 // gcr_run_settings.sampler (0 uniform, 1 PROSAC, 3 importance; 2 / 4: explicit)
		const uint32_t *explicit_samples = nullptr; // sampler 2 / 4: [B][m] index sets drawn on the host
		uint64_t first = 0;
		int B = 0, m = 0, mm = 0;
		uint64_t seed = 0;
		int scoring = GCR_MSAC;
		double threshold = 0.0;
		bool want_samples = false;       // DEGENSAC needs the minimal sample behind a candidate
	};
	struct BatchOut
	{
		int B = 0;                       // samples actually solved (the batch is halved when exact counts overflow)
		std::vector<uint8_t> valid;      // [B mm]
		std::vector<float> scores;       // FP32 ranking scores
		std::vector<uint32_t> counts;    // FP64-exact inlier counts
		std::vector<double> models;      // [B mm][ml], entries fetched so far
		std::vector<uint8_t> model_have;
		std::vector<uint32_t> samples;   // [B][m] when wanted
	};
	// local optimisation: `trials` pool samples (or the whole set once) -> non-minimal fits -> FP32 scores
	struct LoSpec
	{
		bool sample_pool = true;         // false: one fit on the whole `pool` (every trial would fit the same model)
		const uint32_t *pool = nullptr;
		size_t pool_size = 0, sample_size = 0;
		uint64_t first_index = 0;
		int trials = 0;
		uint64_t seed = 0;
		uint32_t stream = 1;
		int refit_problem = 0, ba_frame = 0, rm = 1;
		int scoring = GCR_MSAC;
		double threshold = 0.0;
	};

	class Backend
	{
	public:
		int problem = 0, n = 0, ml = 9, cols = 4;
		bool keeps_edges = true; // the edge list handed to set_edges / built by grid_device / radius_graph stays on the device
		std::string error;
		virtual ~Backend() {}
		// may this backend build neighbourhood structures on the device (grids over > 16384 points)?
		virtual bool upload(const double *pts, int n_, int cols_) = 0;
		virtual bool set_point_probabilities(const double *p, int n_) = 0;
		virtual bool grid_device(const double cell_sizes[4], uint64_t cells, std::vector<uint32_t> &sorted_idx,
		                         std::vector<uint32_t> &cell_offsets, uint32_t &n_cells, uint64_t &neighbor_number,
		                         std::vector<uint32_t> *edges) = 0;
		virtual bool radius_graph(double radius, std::vector<uint32_t> &offsets, std::vector<uint32_t> &neighbors,
		                          std::vector<uint32_t> &edges) = 0;
		virtual bool set_edges(const std::vector<uint32_t> &edges) = 0; // host-enumerated edge list (kept on the device when wanted)
		virtual bool run_batch(const BatchSpec &spec, BatchOut &out) = 0;
		virtual bool fetch_models(const std::vector<size_t> &slots, BatchOut &out) = 0;
		// FP64 getScore (scoring_function.h:137-233 / :246-348): T = squared truncated threshold (MSAC)
		virtual bool exact(const double *model, int scoring, double T, double threshold, double &score, uint32_t &count,
		                   uint32_t *inliers /* [n] or null */, double *residuals /* [n] or null */) = 0;
		virtual bool exact_many(const double *models, size_t C, int scoring, double threshold, std::vector<double> &score,
		                        std::vector<uint32_t> &count, std::vector<uint32_t> &idx /* [C][n] */) = 0;
		// labeling()'s energy terms: unary3n = [n][2] capacities then [n] clamped distances; pairwise [E] only when
		// device_pairwise (the edge list lives on the device)
		virtual bool labeling_terms(const double *model, double threshold, double lambda, size_t n_edges, bool device_pairwise,
		                            std::vector<double> &unary3n, std::vector<double> &pairwise) = 0;
		virtual bool refit_sets(int refit_problem, int ba_frame, const uint32_t *flat, const std::vector<uint32_t> &offsets,
		                        const std::vector<uint32_t> &sizes, const double *weights, size_t n_weights, int weight_mode,
		                        std::vector<double> &models, std::vector<uint8_t> &ok) = 0;
		// refit_sets followed by the FP64 getScore of every model it returned, as ONE submission where the backend can (the
		// IRLS round, the final fit and the post-run polish are "fit, then score the fit"): score / count [k], idx [k][n]
		// (the first count entries of a row), residuals [k][n] when wanted; k = sets x models per fit.  Same kernels on the
		// same inputs as the two separate operations.
		virtual bool refit_and_score(int refit_problem, int ba_frame, const uint32_t *flat, const std::vector<uint32_t> &offsets,
		                             const std::vector<uint32_t> &sizes, const double *weights, size_t n_weights, int weight_mode,
		                             int scoring, double T, double threshold, bool want_residuals, std::vector<double> &models,
		                             std::vector<uint8_t> &ok, std::vector<double> &score, std::vector<uint32_t> &count,
		                             std::vector<uint32_t> &idx, std::vector<double> &residuals)
		{
			if (!refit_sets(refit_problem, ba_frame, flat, offsets, sizes, weights, n_weights, weight_mode, models, ok))
				return false;
			const size_t k = ok.size();
			score.assign(k, 0.0);
			count.assign(k, 0u);
			idx.resize(k * (size_t)n);
			residuals.resize(want_residuals ? k * (size_t)n : 0);
			for (size_t q = 0; q < k; ++q)
				if (ok[q] && !exact(&models[q * ml], scoring, T, threshold, score[q], count[q], &idx[q * (size_t)n],
				                    want_residuals ? &residuals[q * (size_t)n] : nullptr))
					return false;
			return true;
		}
		virtual bool lo_batch(const LoSpec &spec, std::vector<double> &models, std::vector<uint8_t> &ok, std::vector<float> &scores32) = 0;
		virtual bool refit_e_ba(const uint32_t *set, size_t count, const double *weights, std::vector<double> &models,
		                        std::vector<uint8_t> &ok) = 0;
	};

	// gcr_malloc'ed block that lives as long as its owner (returned to the context's cache)
	struct DeviceBuffer
	{
		gcr_ctx *ctx = nullptr;
		void *p = nullptr;
		size_t bytes = 0;
		~DeviceBuffer() { if (p) gcr_free(ctx, p); }
		bool ensure(gcr_ctx *c, size_t need)
		{
			ctx = c;
			if (need <= bytes)
				return true;
			if (p) gcr_free(ctx, p);
			p = nullptr;
			bytes = 0;
			if (gcr_malloc(ctx, need, &p) != GCR_OK)
				return false;
			bytes = need;
			return true;
		}
		template <class T> T *as() const { return reinterpret_cast<T *>(p); }
	};

	// ------------------------------------------------------------------------------------------------ direct ----
	class DirectBackend : public Backend
	{
	public:
		gcr_ctx *ctx = nullptr;
		int pair = 0; // which pair of the context this backend works on (0 for a single-pair context)
		explicit DirectBackend(gcr_ctx *c, int pair_ = 0) : ctx(c), pair(pair_) {}

		bool upload(const double *pts, int n_, int cols_) override
		{
			return gcr_points_upload(ctx, pts, n_, cols_, 1) == GCR_OK || fail("gcr_points_upload");
		}
		bool set_point_probabilities(const double *p, int n_) override
		{
			return gcr_set_point_probabilities(ctx, p, n_) == GCR_OK || fail("gcr_set_point_probabilities");
		}
		bool grid_device(const double cell_sizes[4], uint64_t cells, std::vector<uint32_t> &sorted_idx, std::vector<uint32_t> &cell_offsets,
		                 uint32_t &n_cells, uint64_t &neighbor_number, std::vector<uint32_t> *edges) override
		{
			DeviceBuffer dk, ds, dof;
			if (!dk.ensure(ctx, (size_t)n * 8) || !ds.ensure(ctx, (size_t)n * 4) || !dof.ensure(ctx, ((size_t)n + 1) * 4))
				return fail("gcr_malloc");
			if (gcr_grid_build(ctx, pair, cell_sizes, cells, dk.as<uint64_t>(), ds.as<uint32_t>(), dof.as<uint32_t>(), &n_cells,
			                   &neighbor_number) != GCR_OK)
				return fail("gcr_grid_build");
			sorted_idx.resize(n);
			cell_offsets.resize((size_t)n_cells + 1);
			if (gcr_memcpy_d2h(ctx, sorted_idx.data(), ds.p, (size_t)n * 4) != GCR_OK ||
			    gcr_memcpy_d2h(ctx, cell_offsets.data(), dof.p, ((size_t)n_cells + 1) * 4) != GCR_OK)
				return fail("gcr_memcpy_d2h");
			if (!edges)
				return true;
			// labeling()'s edge list from the sorted cells, on the device (neighbor_number = sum over cells of s (s - 1) / 2)
			if (neighbor_number >= (1ull << 31))
			{
				error = "the grid neighbourhood has more than 2^31 point pairs: use more cells per axis";
				return false;
			}
			if (!d_edge_offsets.ensure(ctx, ((size_t)n + 1) * 4) || !d_edges.ensure(ctx, (neighbor_number ? (size_t)neighbor_number : 1) * 8))
				return fail("gcr_malloc");
			if (gcr_grid_edges(ctx, ds.as<uint32_t>(), dof.as<uint32_t>(), n_cells, d_edge_offsets.as<uint32_t>(), d_edges.as<uint32_t>(),
			                   neighbor_number) != GCR_OK)
				return fail("gcr_grid_edges");
			edges->resize(2 * (size_t)neighbor_number);
			if (neighbor_number && gcr_memcpy_d2h(ctx, edges->data(), d_edges.p, edges->size() * sizeof(uint32_t)) != GCR_OK)
				return fail("gcr_memcpy_d2h");
			return true;
		}
		bool radius_graph(double radius, std::vector<uint32_t> &offsets, std::vector<uint32_t> &neighbors, std::vector<uint32_t> &edges) override
		{
			DeviceBuffer dc, doff, dn;
			if (!dc.ensure(ctx, (size_t)n * 4) || !doff.ensure(ctx, ((size_t)n + 1) * 4))
				return fail("gcr_malloc");
			if (gcr_radius_graph_count(ctx, pair, radius, dc.as<uint32_t>()) != GCR_OK)
				return fail("gcr_radius_graph_count");
			std::vector<uint32_t> counts(n);
			if (gcr_memcpy_d2h(ctx, counts.data(), dc.p, (size_t)n * 4) != GCR_OK)
				return fail("gcr_memcpy_d2h");
			offsets.assign((size_t)n + 1, 0u);
			uint64_t total64 = 0;
			for (int i = 0; i < n; ++i)
			{
				total64 += counts[i];
				offsets[i + 1] = (uint32_t)total64;
			}
			if (total64 >= (1ull << 31))
			{
				// (a radius this large against the point spread is not a sparse neighbourhood: the reference's own N x N edge
				// matrix would not fit either)
				error = "neighborhood_radius gives more than 2^31 neighbour entries: choose a smaller radius";
				return false;
			}
			const size_t total = (size_t)total64;
			neighbors.resize(total);
			edges.clear();
			if (!total)
				return true;
			if (!dn.ensure(ctx, total * 4))
				return fail("gcr_malloc");
			if (gcr_memcpy_h2d(ctx, doff.p, offsets.data(), ((size_t)n + 1) * 4) != GCR_OK)
				return fail("gcr_memcpy_h2d");
			if (gcr_radius_graph_fill(ctx, pair, radius, doff.as<uint32_t>(), dn.as<uint32_t>()) != GCR_OK)
				return fail("gcr_radius_graph_fill");
			// labeling()'s edges (i, j > i) straight from the CSR on the device (the distance test is symmetric: half the entries)
			if (!d_edge_offsets.ensure(ctx, ((size_t)n + 1) * 4) || !d_edges.ensure(ctx, (total / 2 ? total / 2 : 1) * 8))
				return fail("gcr_malloc");
			if (gcr_csr_edges(ctx, doff.as<uint32_t>(), dn.as<uint32_t>(), d_edge_offsets.as<uint32_t>(), d_edges.as<uint32_t>(),
			                  total / 2) != GCR_OK)
				return fail("gcr_csr_edges");
			edges.resize(2 * (total / 2));
			if (gcr_memcpy_d2h_async(ctx, neighbors.data(), dn.p, total * 4) != GCR_OK ||
			    (edges.size() && gcr_memcpy_d2h_async(ctx, edges.data(), d_edges.p, edges.size() * sizeof(uint32_t)) != GCR_OK) ||
			    gcr_sync(ctx) != GCR_OK)
				return fail("gcr_memcpy_d2h");
			return true;
		}
		bool set_edges(const std::vector<uint32_t> &edges) override
		{
			if (edges.empty())
				return true;
			if (!d_edges.ensure(ctx, edges.size() * sizeof(uint32_t)))
				return fail("gcr_malloc");
			// (larger than the pinned ring when the neighbourhood is dense: gcr_memcpy_h2d then copies synchronously)
			return gcr_memcpy_h2d(ctx, d_edges.p, edges.data(), edges.size() * sizeof(uint32_t)) == GCR_OK || fail("gcr_memcpy_h2d");
		}

		bool run_batch(const BatchSpec &spec, BatchOut &out) override
		{
			int B = spec.B;
			for (;;)
			{
				const size_t K = (size_t)B * spec.mm;
				if (!d_samples.ensure(ctx, (size_t)B * spec.m * sizeof(uint32_t)) || !d_models.ensure(ctx, K * ml * sizeof(double)) ||
				    !d_valid.ensure(ctx, K) || !d_scores.ensure(ctx, K * sizeof(float)) || !d_counts.ensure(ctx, K * sizeof(uint32_t)))
					return fail("gcr_malloc");
				// PROSAC draws are numbered from k = 1 (prosac_sampler.h:207-236); draw g of the run is PROSAC sample g + 1
				if (spec.explicit_samples)
				{
					// AdaptiveReorderingSampler / ProgressiveNapsacSampler (py/gcransac_python.cpp:4334-4359): the recurrence
					// over the sample number runs on the host, the batch's index sets go up in one copy
					if (gcr_memcpy_h2d(ctx, d_samples.p, spec.explicit_samples, (size_t)B * spec.m * sizeof(uint32_t)) != GCR_OK)
						return fail("gcr_memcpy_h2d");
				}
				// sampler 3 = ImportanceSampler (NG-RANSAC, py/gcransac_python.cpp:4343-4346): draw g = sample g of stream 3
				else if (gcr_sample(ctx, spec.sampler == 1 ? GCR_PROSAC : spec.sampler == 3 ? GCR_IMPORTANCE : GCR_UNIFORM, spec.m,
				                    spec.sampler == 1 ? spec.first + 1 : spec.first, B, spec.seed, d_samples.as<uint32_t>()) != GCR_OK)
					return fail("gcr_sample");
				if (spec.variant == 1)
				{
					if (gcr_solve_plane_parallax(ctx, d_samples.as<uint32_t>(), B, spec.pp_H, d_models.as<double>(), d_valid.as<uint8_t>()) != GCR_OK)
						return fail("gcr_solve_plane_parallax");
				}
				else if (gcr_solve_minimal(ctx, problem, d_samples.as<uint32_t>(), B, d_models.as<double>(), d_valid.as<uint8_t>()) != GCR_OK)
					return fail("gcr_solve_minimal");
				if (gcr_score(ctx, problem, spec.scoring, d_models.as<double>(), d_valid.as<uint8_t>(), (int)K, spec.threshold,
				              GCR_SCORE_EXACT_COUNTS, d_scores.as<float>(), d_counts.as<uint32_t>()) != GCR_OK)
					return fail("gcr_score");
				out.valid.resize(K);
				out.scores.resize(K);
				out.counts.resize(K);
				out.models.resize(K * ml);
				// small batches (H, F: up to a few hundred KB) come back whole with the flags and scores -- one more queued copy,
				// no later round trip; large ones (five-point batches: 0.7-2.9 MB) stay on the device and are fetched on demand
				const bool models_with_batch = K * ml * sizeof(double) <= (size_t)(256 << 10);
				out.model_have.assign(K, models_with_batch ? 1 : 0);
				if (models_with_batch && gcr_memcpy_d2h_async(ctx, out.models.data(), d_models.p, K * ml * sizeof(double)) != GCR_OK)
					return fail("gcr_memcpy_d2h");
				if (spec.want_samples)
				{
					// applyDegensac needs the minimal sample behind a candidate model (fundamental_estimator.h:344)
					out.samples.resize((size_t)B * spec.m);
					if (gcr_memcpy_d2h_async(ctx, out.samples.data(), d_samples.p, (size_t)B * spec.m * sizeof(uint32_t)) != GCR_OK)
						return fail("gcr_memcpy_d2h");
				}
				uint32_t dropped = 0;
				if (gcr_memcpy_d2h_async(ctx, out.valid.data(), d_valid.p, K) != GCR_OK ||
				    gcr_memcpy_d2h_async(ctx, out.scores.data(), d_scores.p, K * sizeof(float)) != GCR_OK ||
				    gcr_memcpy_d2h_async(ctx, out.counts.data(), d_counts.p, K * sizeof(uint32_t)) != GCR_OK ||
				    gcr_score_dropped_async(ctx, &dropped) != GCR_OK || gcr_sync(ctx) != GCR_OK)
					return fail("gcr_memcpy_d2h");
				if (!dropped)
				{
					out.B = B;
					return true;
				}
				// the FP64 re-check list overflowed (N x batch very large): these counts feed the inlier-container bookkeeping
				// (lazy_count, the size check before IRLS), so they must be exact -- the same samples again, half at a time
				// (a sample is a function of its index: batch boundaries change no result)
				if (B <= 1)
				{
					error = "gcr_score: the FP64 re-check list overflows for a single sample";
					return false;
				}
				B = (B + 1) / 2;
			}
		}
		bool fetch_models(const std::vector<size_t> &slots, BatchOut &out) override
		{
			bool queued = false;
			for (size_t slot : slots)
				if (!out.model_have[slot])
				{
					if (gcr_memcpy_d2h_async(ctx, &out.models[slot * ml], d_models.as<double>() + slot * ml, ml * sizeof(double)) != GCR_OK)
						return fail("gcr_memcpy_d2h");
					out.model_have[slot] = 1;
					queued = true;
				}
			return !queued || gcr_sync(ctx) == GCR_OK || fail("gcr_sync");
		}
		bool exact(const double *model, int scoring, double T, double threshold, double &score, uint32_t &count, uint32_t *inliers,
		           double *residuals) override
		{
			if (!residuals)
				return gcr_getscore_host(ctx, problem, scoring, pair, model, T, threshold, &score, &count, inliers) == GCR_OK ||
				       fail("gcr_getscore_host");
			if (!d_one_model.ensure(ctx, ml * sizeof(double)) || !d_one_score.ensure(ctx, sizeof(double)) ||
			    !d_one_count.ensure(ctx, sizeof(uint32_t)) || !d_idx.ensure(ctx, (size_t)n * sizeof(uint32_t)) ||
			    !d_res.ensure(ctx, (size_t)n * sizeof(double)))
				return fail("gcr_malloc");
			if (gcr_memcpy_h2d(ctx, d_one_model.p, model, ml * sizeof(double)) != GCR_OK)
				return fail("gcr_memcpy_h2d");
			if (gcr_score_exact(ctx, problem, scoring, pair, d_one_model.as<double>(), 1, threshold, d_one_score.as<double>(),
			                    d_one_count.as<uint32_t>(), d_idx.as<uint32_t>(), d_res.as<double>()) != GCR_OK)
				return fail("gcr_score_exact");
			// every read-back of the call is queued, one synchronise delivers them (the list is copied whole: 4 n bytes)
			if (gcr_memcpy_d2h_async(ctx, &score, d_one_score.p, sizeof(double)) != GCR_OK ||
			    gcr_memcpy_d2h_async(ctx, &count, d_one_count.p, sizeof(uint32_t)) != GCR_OK ||
			    gcr_memcpy_d2h_async(ctx, residuals, d_res.p, (size_t)n * sizeof(double)) != GCR_OK ||
			    (inliers && gcr_memcpy_d2h_async(ctx, inliers, d_idx.p, (size_t)n * sizeof(uint32_t)) != GCR_OK) || gcr_sync(ctx) != GCR_OK)
				return fail("gcr_memcpy_d2h");
			return true;
		}
		// FP64 getScore of C models (host, ml doubles each) in one launch and one synchronise: scores, inlier counts and
		// the inlier lists ([C][n], the first count entries of each row)
		bool exact_many(const double *models, size_t C, int scoring, double threshold, std::vector<double> &score,
		                std::vector<uint32_t> &count, std::vector<uint32_t> &idx) override
		{
			if (!d_cand_models.ensure(ctx, C * ml * sizeof(double)) || !d_cand_scores.ensure(ctx, C * sizeof(double)) ||
			    !d_cand_counts.ensure(ctx, C * sizeof(uint32_t)) || !d_cand_idx.ensure(ctx, C * (size_t)n * sizeof(uint32_t)))
				return fail("gcr_malloc");
			if (gcr_memcpy_h2d(ctx, d_cand_models.p, models, C * ml * sizeof(double)) != GCR_OK)
				return fail("gcr_memcpy_h2d");
			if (gcr_score_exact(ctx, problem, scoring, pair, d_cand_models.as<double>(), (int)C, threshold, d_cand_scores.as<double>(),
			                    d_cand_counts.as<uint32_t>(), d_cand_idx.as<uint32_t>(), nullptr) != GCR_OK)
				return fail("gcr_score_exact");
			score.resize(C);
			count.resize(C);
			idx.resize(C * (size_t)n);
			if (gcr_memcpy_d2h_async(ctx, score.data(), d_cand_scores.p, C * sizeof(double)) != GCR_OK ||
			    gcr_memcpy_d2h_async(ctx, count.data(), d_cand_counts.p, C * sizeof(uint32_t)) != GCR_OK ||
			    gcr_memcpy_d2h_async(ctx, idx.data(), d_cand_idx.p, C * (size_t)n * sizeof(uint32_t)) != GCR_OK || gcr_sync(ctx) != GCR_OK)
				return fail("gcr_memcpy_d2h");
			return true;
		}
		bool labeling_terms(const double *model, double threshold, double lambda, size_t n_edges, bool device_pairwise,
		                    std::vector<double> &unary3n, std::vector<double> &pairwise) override
		{
			const size_t E = n_edges;
			if (!d_lab_model.ensure(ctx, ml * sizeof(double)) || !d_unary.ensure(ctx, (size_t)n * 3 * sizeof(double)) ||
			    !d_pairwise.ensure(ctx, (device_pairwise && E ? E : 1) * sizeof(double)))
				return fail("gcr_malloc");
			unary3n.resize((size_t)n * 3); // [n][2] capacities, then [n] clamped distances
			pairwise.resize(device_pairwise ? E : 0);
			double *d_dist = d_unary.as<double>() + 2 * (size_t)n;
			if (gcr_memcpy_h2d(ctx, d_lab_model.p, model, ml * sizeof(double)) != GCR_OK)
				return fail("gcr_memcpy_h2d");
			if (gcr_labeling_terms(ctx, problem, pair, d_lab_model.as<double>(), threshold, lambda, d_edges.as<uint32_t>(), E,
			                       d_unary.as<double>(), device_pairwise ? d_pairwise.as<double>() : nullptr, d_dist) != GCR_OK)
				return fail("gcr_labeling_terms");
			if (gcr_memcpy_d2h_async(ctx, unary3n.data(), d_unary.p, (size_t)n * 3 * sizeof(double)) != GCR_OK ||
			    (device_pairwise && E && gcr_memcpy_d2h_async(ctx, pairwise.data(), d_pairwise.p, E * sizeof(double)) != GCR_OK) ||
			    gcr_sync(ctx) != GCR_OK)
				return fail("gcr_memcpy_d2h");
			return true;
		}
		// non-minimal refit of `count` index sets (one GPU launch)
		bool refit_sets(int rp, int ba_frame, const uint32_t *flat, const std::vector<uint32_t> &offsets, const std::vector<uint32_t> &sizes,
		                const double *weights, size_t n_weights, int weight_mode, std::vector<double> &models, std::vector<uint8_t> &ok) override
		{
			if (!launch_refit(rp, ba_frame, flat, offsets, sizes, weights, n_weights, weight_mode, models, ok))
				return false;
			return gcr_sync(ctx) == GCR_OK || fail("gcr_sync");
		}
		bool refit_and_score(int rp, int ba_frame, const uint32_t *flat, const std::vector<uint32_t> &offsets, const std::vector<uint32_t> &sizes,
		                     const double *weights, size_t n_weights, int weight_mode, int scoring, double, double threshold,
		                     bool want_residuals, std::vector<double> &models, std::vector<uint8_t> &ok, std::vector<double> &score,
		                     std::vector<uint32_t> &count, std::vector<uint32_t> &idx, std::vector<double> &residuals) override
		{
			if (gcr_model_len(rp) != ml) // (a fit of another model family: scored by its own call)
				return Backend::refit_and_score(rp, ba_frame, flat, offsets, sizes, weights, n_weights, weight_mode, scoring, 0.0, threshold,
				                                want_residuals, models, ok, score, count, idx, residuals);
			if (!launch_refit(rp, ba_frame, flat, offsets, sizes, weights, n_weights, weight_mode, models, ok))
				return false; // (its read-backs are queued; the scores below travel with them)
			const size_t k = ok.size();
			if (!d_cand_scores.ensure(ctx, k * sizeof(double)) || !d_cand_counts.ensure(ctx, k * sizeof(uint32_t)) ||
			    !d_cand_idx.ensure(ctx, k * (size_t)n * sizeof(uint32_t)) || (want_residuals && !d_res.ensure(ctx, k * (size_t)n * sizeof(double))))
				return fail("gcr_malloc");
			// every slot of the fit is scored, a failed fit's zero model included (its numbers are never read)
			if (gcr_score_exact(ctx, problem, scoring, pair, d_lmodels.as<double>(), (int)k, threshold, d_cand_scores.as<double>(),
			                    d_cand_counts.as<uint32_t>(), d_cand_idx.as<uint32_t>(), want_residuals ? d_res.as<double>() : nullptr) != GCR_OK)
				return fail("gcr_score_exact");
			score.resize(k);
			count.resize(k);
			idx.resize(k * (size_t)n);
			residuals.resize(want_residuals ? k * (size_t)n : 0);
			if (gcr_memcpy_d2h_async(ctx, score.data(), d_cand_scores.p, k * sizeof(double)) != GCR_OK ||
			    gcr_memcpy_d2h_async(ctx, count.data(), d_cand_counts.p, k * sizeof(uint32_t)) != GCR_OK ||
			    gcr_memcpy_d2h_async(ctx, idx.data(), d_cand_idx.p, k * (size_t)n * sizeof(uint32_t)) != GCR_OK ||
			    (want_residuals && gcr_memcpy_d2h_async(ctx, residuals.data(), d_res.p, k * (size_t)n * sizeof(double)) != GCR_OK) ||
			    gcr_sync(ctx) != GCR_OK)
				return fail("gcr_memcpy_d2h");
			return true;
		}
		bool lo_batch(const LoSpec &spec, std::vector<double> &models, std::vector<uint8_t> &ok, std::vector<float> &scores32) override
		{
			const int rp = spec.refit_problem, rml = gcr_model_len(rp);
			int n_models;
			if (spec.sample_pool)
			{
				// inner RANSAC as one batch: `trials` pool samples -> refits -> FP32 scores
				const int trials = spec.trials;
				if (!d_pool.ensure(ctx, spec.pool_size * sizeof(uint32_t)) || !d_sets.ensure(ctx, (size_t)trials * spec.sample_size * sizeof(uint32_t)))
					return fail("gcr_malloc");
				if (gcr_memcpy_h2d(ctx, d_pool.p, spec.pool, spec.pool_size * sizeof(uint32_t)) != GCR_OK)
					return fail("gcr_memcpy_h2d");
				if (gcr_sample_from_pool(ctx, d_pool.as<uint32_t>(), (uint32_t)spec.pool_size, (int)spec.sample_size, spec.first_index, trials,
				                         spec.seed, spec.stream, d_sets.as<uint32_t>()) != GCR_OK)
					return fail("gcr_sample_from_pool");
				std::vector<uint32_t> offsets(trials), sizes(trials, (uint32_t)spec.sample_size);
				for (int t = 0; t < trials; ++t) offsets[t] = (uint32_t)(t * spec.sample_size);
				if (!launch_refit(rp, spec.ba_frame, nullptr, offsets, sizes, nullptr, 0, GCR_WEIGHT_NONE, models, ok))
					return false; // (read-backs queued: delivered together with the scores below)
				n_models = trials * spec.rm; // trial-major, the solver's model order inside a trial (GCRANSAC.h:866)
			}
			else
			{
				// all inliers, every trial fits the same model: once is enough (later trials cannot beat it, strict '<')
				std::vector<uint32_t> offsets(1, 0u), sizes(1, (uint32_t)spec.pool_size);
				if (!launch_refit(rp, spec.ba_frame, spec.pool, offsets, sizes, nullptr, 0, GCR_WEIGHT_NONE, models, ok))
					return false;
				n_models = spec.rm;
			}
			if (!d_lscores.ensure(ctx, n_models * sizeof(float)))
				return fail("gcr_malloc");
			if (gcr_score(ctx, problem, spec.scoring, d_lmodels.as<double>(), d_lok.as<uint8_t>(), n_models, spec.threshold, 0,
			              d_lscores.as<float>(), nullptr) != GCR_OK)
				return fail("gcr_score");
			scores32.resize(n_models);
			(void)rml;
			return gcr_memcpy_d2h(ctx, scores32.data(), d_lscores.p, n_models * sizeof(float)) == GCR_OK || fail("gcr_memcpy_d2h");
		}
		bool refit_e_ba(const uint32_t *set, size_t count, const double *weights, std::vector<double> &models, std::vector<uint8_t> &ok) override
		{
			const int R = gcr_max_e_ba_models();
			if (!d_sets.ensure(ctx, count * sizeof(uint32_t)) || !d_weights.ensure(ctx, (count > (size_t)n ? count : (size_t)n) * sizeof(double)) ||
			    !d_lmodels.ensure(ctx, (size_t)R * 9 * sizeof(double)) || !d_lok.ensure(ctx, (size_t)R))
				return fail("gcr_malloc");
			models.resize((size_t)R * 9);
			ok.resize(R);
			if (gcr_memcpy_h2d(ctx, d_sets.p, set, count * sizeof(uint32_t)) != GCR_OK ||
			    gcr_memcpy_h2d(ctx, d_weights.p, weights, count * sizeof(double)) != GCR_OK ||
			    gcr_refit_e_ba(ctx, pair, d_sets.as<uint32_t>(), (uint32_t)count, d_weights.as<double>(), d_lmodels.as<double>(),
			                   d_lok.as<uint8_t>()) != GCR_OK ||
			    gcr_memcpy_d2h_async(ctx, models.data(), d_lmodels.p, models.size() * sizeof(double)) != GCR_OK ||
			    gcr_memcpy_d2h_async(ctx, ok.data(), d_lok.p, ok.size()) != GCR_OK || gcr_sync(ctx) != GCR_OK)
				return fail("gcr_refit_e_ba");
			return true;
		}

	private:
		DeviceBuffer d_samples, d_models, d_valid, d_scores, d_counts;
		DeviceBuffer d_one_model, d_one_score, d_one_count, d_idx, d_res;
		DeviceBuffer d_cand_models, d_cand_scores, d_cand_counts, d_cand_idx;
		DeviceBuffer d_pool, d_sets, d_offsets, d_lmodels, d_lok, d_lscores, d_weights;
		DeviceBuffer d_edges, d_edge_offsets, d_unary, d_pairwise, d_lab_model;

		bool fail(const char *what)
		{
			error = std::string(what) + ": " + gcr_last_error(ctx);
			return false;
		}
		// uploads the sets (flat == nullptr: they are in d_sets already), launches the fit, queues the read-backs (no synchronise)
		bool launch_refit(int rp, int ba_frame, const uint32_t *flat, const std::vector<uint32_t> &offsets, const std::vector<uint32_t> &sizes,
		                  const double *weights, size_t n_weights, int weight_mode, std::vector<double> &models, std::vector<uint8_t> &ok)
		{
			const int rm = gcr_max_refit_models(rp), rml = gcr_model_len(rp);
			const int count = (int)sizes.size();
			size_t total = 0;
			for (uint32_t v : sizes) total += v;
			if (flat)
			{
				if (!d_sets.ensure(ctx, (total ? total : 1) * sizeof(uint32_t)))
					return fail("gcr_malloc");
				if (gcr_memcpy_h2d(ctx, d_sets.p, flat, total * sizeof(uint32_t)) != GCR_OK)
					return fail("gcr_memcpy_h2d");
			}
			// offsets and sizes travel as one array (one copy instead of two): [offsets | sizes]
			if (!d_offsets.ensure(ctx, 2 * count * sizeof(uint32_t)) || !d_lmodels.ensure(ctx, (size_t)count * rm * rml * sizeof(double)) ||
			    !d_lok.ensure(ctx, (size_t)count * rm))
				return fail("gcr_malloc");
			std::vector<uint32_t> offsets_sizes(offsets.begin(), offsets.begin() + count);
			offsets_sizes.insert(offsets_sizes.end(), sizes.begin(), sizes.begin() + count);
			if (gcr_memcpy_h2d(ctx, d_offsets.p, offsets_sizes.data(), 2 * count * sizeof(uint32_t)) != GCR_OK)
				return fail("gcr_memcpy_h2d");
			const uint32_t *d_set_offsets = d_offsets.as<uint32_t>(), *d_set_sizes = d_offsets.as<uint32_t>() + count;
			const double *d_w = nullptr;
			if (weights)
			{
				if (!d_weights.ensure(ctx, (n_weights > (size_t)n ? n_weights : (size_t)n) * sizeof(double)))
					return fail("gcr_malloc");
				if (gcr_memcpy_h2d(ctx, d_weights.p, weights, n_weights * sizeof(double)) != GCR_OK)
					return fail("gcr_memcpy_h2d");
				d_w = d_weights.as<double>();
			}
			if (ba_frame)
			{
				if (gcr_refit_f_ba(ctx, pair, d_sets.as<uint32_t>(), d_set_offsets, d_set_sizes, count, d_w, weight_mode, ba_frame,
				                   d_lmodels.as<double>(), d_lok.as<uint8_t>()) != GCR_OK)
					return fail("gcr_refit_f_ba");
			}
			else if (gcr_refit(ctx, rp, pair, d_sets.as<uint32_t>(), d_set_offsets, d_set_sizes, count, d_w, weight_mode,
			                   d_lmodels.as<double>(), d_lok.as<uint8_t>()) != GCR_OK)
				return fail("gcr_refit");
			models.resize((size_t)count * rm * rml);
			ok.resize((size_t)count * rm);
			return (gcr_memcpy_d2h_async(ctx, models.data(), d_lmodels.p, models.size() * sizeof(double)) == GCR_OK &&
			        gcr_memcpy_d2h_async(ctx, ok.data(), d_lok.p, ok.size()) == GCR_OK) ||
			       fail("gcr_memcpy_d2h");
		}
	};
}
#endif
