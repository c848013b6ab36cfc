// gcr_driver.cpp -- see gcr_driver.h.  Host C++; the only compile-time dependency on the reference tree is its
// vendored Boykov-Kolmogorov max-flow (energy.h / graph.h + graph.cpp / maxflow.cpp), which stays the reference's code.
#include "gcr_driver.h"
#include "gcr_backend.h"
#include "gcr_host_samplers.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <memory>
#include <string>
#include <vector>

// The reference's own Boykov-Kolmogorov max-flow, included from the reference tree exactly as its GCoptimization.h:110-111
// does (graph.cpp / maxflow.cpp hold the template definitions): the min-cut stays the reference's host code.
#include "energy.h"
#include "graph.cpp"
#include "maxflow.cpp"

namespace
{
	struct Score
	{
		double value = 0.0;
		uint64_t inliers = 0;
	};

	// GCRANSAC.h:158-172
	uint64_t iteration_number(uint64_t inliers, uint64_t n, int m, double log_probability)
	{
		const double q = std::pow(static_cast<double>(inliers) / n, m);
		const double log2 = std::log(1 - q);
		if (std::fabs(log2) < std::numeric_limits<double>::epsilon())
			return std::numeric_limits<uint64_t>::max();
		const double iter = log_probability / log2;
		return static_cast<uint64_t>(iter) + 1;
	}

	// fundamental_estimator.h:225-253
	// homography_estimator.h:181-199
	double sq_homography_transfer(const double *s, const double *H)
	{
		const double x1 = s[0], y1 = s[1], x2 = s[2], y2 = s[3];
		const double t1 = H[0] * x1 + H[1] * y1 + H[2], t2 = H[3] * x1 + H[4] * y1 + H[5], t3 = H[6] * x1 + H[7] * y1 + H[8];
		const double d1 = x2 - (t1 / t3), d2 = y2 - (t2 / t3);
		return d1 * d1 + d2 * d2;
	}

	double sq_symmetric_epipolar(const double *s, const double *F)
	{
		const double x1 = s[0], y1 = s[1], x2 = s[2], y2 = s[3];
		const double rxc = F[0] * x2 + F[3] * y2 + F[6];
		const double ryc = F[1] * x2 + F[4] * y2 + F[7];
		const double rwc = F[2] * x2 + F[5] * y2 + F[8];
		const double r = (x1 * rxc + y1 * ryc + rwc);
		const double rx = F[0] * x1 + F[1] * y1 + F[2];
		const double ry = F[3] * x1 + F[4] * y1 + F[5];
		const double a = rxc * rxc + ryc * ryc;
		const double b = rx * rx + ry * ry;
		return r * r * (a + b) / (a * b);
	}

	using gcr_host::Backend;
	using gcr_host::BatchOut;
	using gcr_host::BatchSpec;
	using gcr_host::DirectBackend;
	using gcr_host::LoSpec;

	// One of the two inlier containers of GCRANSAC::run (temp_inner_inliers, GCRANSAC.h:245-247).  Every getScore of the
	// reference overwrites the current container; here the contents are only materialised when somebody reads them.
	struct InlierBuffer
	{
		bool lazy = false;      // holds "the inliers of `model`", not fetched yet
		long long slot = -1;    // lazy: `model` itself is still slot `slot` of the current batch on the device (>= 0), not in `model`
		double model[12] = {0}; // 9 (H / F / E) or 12 (PnP) entries
		uint64_t lazy_count = 0;
		std::vector<uint32_t> idx;
		size_t size() const { return lazy ? (size_t)lazy_count : idx.size(); }
	};

	struct PhaseTimer // adds the scope's wall clock to one slot of gcr_run_stats::phase_seconds
	{
		double &slot;
		std::chrono::steady_clock::time_point t0;
		explicit PhaseTimer(double &s) : slot(s), t0(std::chrono::steady_clock::now()) {}
		~PhaseTimer() { slot += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
	};

	class Driver
	{
	public:
		Backend *be = nullptr; // the device operations: one pair on its own context, or a seat in the batch engine
		int problem, n, m, mm, rm; // rm = models per non-minimal refit (1: H, F, PnP; up to 10: E)
		int ml = 9, cols = 4;      // model length (12: PnP's 3 x 4 pose) and point columns (5: PnP's u v X Y Z)
		const double *pts;
		const double *grid_pts = nullptr; // points the neighbourhood grid is built on (E: pixel coordinates; else = pts)
		gcr_run_settings S;
		gcr_run_stats st;
		// the _ScoringFunction template argument of GCRANSAC<>: GCR_MSAC (every Python entry point) or GCR_MAGSAC
		// (scoring_function.h:237-349; BASELINE configs[2]).  FP32 ranking error: 2e-5 (MSAC), up to 2e-4 (MAGSAC: the
		// step-function LUT index flips under FP32 residual error) -- rank_slack is what "may beat the best so far" allows
		int scoring = GCR_MSAC;
		double rank_slack = 2e-4;
		double T; // squared truncated threshold (GCRANSAC.h:253-254)
		double truncated_threshold;
		std::string error;
		// DEGENSAC (fundamental_estimator.h:341-572).  The inner run is this same driver with the plane-and-parallax
		// minimal solver (variant 1: 2-point samples + pp_H), on the points the outer run already uploaded.
		int variant = 0;
		bool nested = false;
		double pp_H[9] = {0};
		double f_min_ratio = 0.5; // minimum_inlier_ratio_in_validity_check (:91); the inner estimator is built with 0.0 (:540)
		std::vector<uint32_t> &h_samples = batch.samples;
		std::vector<uint32_t> explicit_samples; // samplers 2 / 4: the batch's index sets, drawn on the host
		// FP64 scores / inlier lists of the batch's record candidates, evaluated in ONE launch (prefetch_candidates)
		std::vector<size_t> cand_slot;
		std::vector<double> cand_score;
		std::vector<uint32_t> cand_count;
		std::vector<uint32_t> cand_idx; // [candidate][n]

		// sampler 4: the AR-Sampler's sequence, drawn ahead on the host (gcr_host_samplers.h)
		gcr_host::AdaptiveReordering ar_sampler;
		// sampler 2: Progressive NAPSAC's sequence (its four grid layers are built on the GPU, gcr_grid_build)
		gcr_host::ProgressiveNapsac pnapsac_sampler;
		// main-loop batch
		int B = 0;
		uint64_t batch_first = 0;
		int batch_count = 0;
		// the current batch: flags, FP32 scores, exact counts of every slot; the models of a large batch stay on the device
		// (72 B per slot, 2.9 MB per 4096 five-point samples) and the host fetches the few it looks at -- record candidates,
		// the model behind a lazy inlier container -- into batch.models
		BatchOut batch;
		std::vector<uint8_t> &h_valid = batch.valid;
		std::vector<float> &h_scores = batch.scores;
		std::vector<uint32_t> &h_counts = batch.counts;
		std::vector<double> &h_models = batch.models;
		std::vector<uint8_t> &h_model_have = batch.model_have;
		InlierBuffer *live_buffers = nullptr; // the two containers of the run in progress (their lazy models are resolved
		                                      // before a new batch overwrites d_models)
		// grid
		std::vector<uint32_t> cell_of_point, sorted_idx, cell_offsets;
		bool have_grid = false;
		// radius neighbourhood (S.neighborhood == 1): CSR, neighbours ascending
		std::vector<uint32_t> nbr_offsets, nbr_idx;
		bool have_radius_graph = false;
		// labeling()'s pairwise structure: every unordered pair of neighbours once, (i, j > i), i ascending then j ascending
		// -- enumerated once per run (on the GPU from the grid cells / radius CSR when those were built there), kept on both
		// sides: the device copy feeds gcr_labeling_terms, the host copy the max-flow graph
		std::vector<uint32_t> edges; // [n_edges][2]
		size_t n_edges = 0;
		std::vector<double> h_unary, h_pairwise;

		// DEGENSAC's inner run needs the same points on a backend of its own when the outer one is a seat of the batch engine
		// (its minimal solver is not one of the engine's batched operations): null = the inner run shares `be`
		Backend *(*nested_backend)(Driver *) = nullptr;
		void *owner = nullptr; // the batch engine's per-worker state, for nested_backend
		bool fail(const char *)
		{
			error = be->error;
			return false;
		}

		// ---- FP64 getScore of one model (scoring_function.h:137-233): score, count, ascending inlier list ----
		bool exact(const double *model, Score &s, std::vector<uint32_t> *inliers, std::vector<double> *residuals = nullptr)
		{
			++st.exact_rescores;
			uint32_t count = 0;
			if (inliers) inliers->resize(n);
			if (residuals) residuals->resize(n);
			if (!be->exact(model, scoring, T, S.threshold, s.value, count, inliers ? inliers->data() : nullptr,
			               residuals ? residuals->data() : nullptr))
				return fail("exact");
			s.inliers = count;
			if (inliers) inliers->resize(count);
			return true;
		}

		// model of slot `slot` of the current batch, fetched on first use (one 72-byte read-back)
		const double *model_of(size_t slot)
		{
			if (!h_model_have[slot] && !fetch_models(std::vector<size_t>(1, slot)))
				return nullptr;
			return &h_models[slot * ml];
		}
		// several at once: one round trip
		bool fetch_models(const std::vector<size_t> &slots)
		{
			return be->fetch_models(slots, batch) || fail("fetch_models");
		}
		bool resolve_model(InlierBuffer &b)
		{
			if (b.slot < 0)
				return true;
			const double *model = model_of((size_t)b.slot);
			if (!model)
				return false;
			memcpy(b.model, model, ml * sizeof(double));
			b.slot = -1;
			return true;
		}

		bool materialise(InlierBuffer &b)
		{
			if (!b.lazy)
				return true;
			if (!resolve_model(b))
				return false;
			Score s;
			if (!exact(b.model, s, &b.idx))
				return false;
			b.lazy = false;
			return true;
		}

		// ---- isValidModel ----
		// `model` may be replaced (DEGENSAC, F only): then `updated` is set and the caller re-scores it (GCRANSAC.h:446-454)
		bool is_valid_model(double *model, const std::vector<uint32_t> &inliers, const uint32_t *minimal_sample, bool &updated,
		                    bool &failed)
		{
			updated = false;
			failed = false;
			if (problem == GCR_PNP) // perspective_n_point_estimator.h:219-227: every pose is a valid model
				return true;
			if (problem == GCR_H)
			{
				// homography_estimator.h:326-342: |det H| >= 1e-2
				const double *H = model;
				const double det = H[0] * (H[4] * H[8] - H[5] * H[7]) - H[1] * (H[3] * H[8] - H[5] * H[6]) +
				                   H[2] * (H[3] * H[7] - H[4] * H[6]);
				return !(std::fabs(det) < 1e-2);
			}
			// fundamental_estimator.h:287-318 (symmetric-epipolar ratio test), essential_estimator.h:279-305 (ratio 0.1).
			// `const size_t x = MAX(size_t, double)`: truncated.
			const double a = (double)m, b = inliers.size() * (problem == GCR_E ? 0.1 : f_min_ratio);
			const size_t minimum_inlier_number = (size_t)(a > b ? a : b);
			const double squared_threshold = truncated_threshold * truncated_threshold;
			size_t inlier_number = 0;
			bool passed = false;
			for (uint32_t i : inliers)
				if (sq_symmetric_epipolar(pts + 4 * (size_t)i, model) < squared_threshold)
					if (++inlier_number >= minimum_inlier_number)
					{
						passed = true;
						break;
					}
			if (!passed)
				return false;
			if (problem == GCR_F && variant == 0 && S.use_degensac && minimal_sample) // :325-333
				return apply_degensac(model, inliers, minimal_sample, updated, failed);
			return true;
		}

		// ---- applyDegensac (fundamental_estimator.h:341-572) ----
		// H-degenerate sample test on the host (seven points, five triplets); when it fires: the plane homography is
		// refitted on its inliers among the model's inliers (GPU, K6) and an inner GC-RANSAC with the plane-and-parallax
		// solver runs as a nested driver on this context.  The reference seeds the inner sampler from random_device
		// (:553); here inner run r draws from the counter-based streams of seed + r * 0xD1B54A32D192ED03.
		bool apply_degensac(double *model, const std::vector<uint32_t> &inliers, const uint32_t *sample, bool &updated,
		                    bool &failed)
		{
			static const int triplets[15] = {0, 1, 2, 3, 4, 5, 0, 1, 6, 3, 4, 6, 2, 5, 6};
			const double squared_homography_threshold = 2.0 * 2.0; // homography_threshold_ = 2.0 (:93)
			const double *F = model;
			// epipole in the second image: e^T F = 0 (the reference takes JacobiSVD's U.col(2), :365-372).  F comes from
			// the 7-point solver, rank 2 up to the cubic's polish: e = the largest cross product of two columns of F.
			double e[3] = {0, 0, 0}, best = -1.0;
			for (int c0 = 0; c0 < 3; ++c0)
				for (int c1 = c0 + 1; c1 < 3; ++c1)
				{
					const double u[3] = {F[c0], F[3 + c0], F[6 + c0]}, v[3] = {F[c1], F[3 + c1], F[6 + c1]};
					const double w[3] = {u[1] * v[2] - u[2] * v[1], u[2] * v[0] - u[0] * v[2], u[0] * v[1] - u[1] * v[0]};
					const double nn = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
					if (nn > best) { best = nn; e[0] = w[0]; e[1] = w[1]; e[2] = w[2]; }
				}
			e[0] /= e[2]; e[1] /= e[2]; e[2] = 1.0;
			auto cross = [](const double *x, const double *y, double *z) {
				z[0] = x[1] * y[2] - x[2] * y[1]; z[1] = x[2] * y[0] - x[0] * y[2]; z[2] = x[0] * y[1] - x[1] * y[0];
			};
			const double ex[9] = {0, -e[2], e[1], e[2], 0, -e[0], -e[1], e[0], 0};
			double A[9];
			for (int r = 0; r < 3; ++r)
				for (int c = 0; c < 3; ++c)
					A[3 * r + c] = ex[3 * r] * F[c] + ex[3 * r + 1] * F[3 + c] + ex[3 * r + 2] * F[6 + c];
			bool h_degenerate = false;
			double H[9];
			for (int t = 0; t < 5 && !h_degenerate; ++t)
			{
				uint32_t idx[3];
				double p1[3][3], b[3];
				for (int q = 0; q < 3; ++q)
				{
					idx[q] = sample[triplets[3 * t + q]];
					const double *p = pts + 4 * (size_t)idx[q];
					p1[q][0] = p[0]; p1[q][1] = p[1]; p1[q][2] = 1;
					const double p2[3] = {p[2], p[3], 1};
					double pxe[3], Ap[3], pxAp[3];
					cross(p2, e, pxe);
					for (int r = 0; r < 3; ++r)
						Ap[r] = A[3 * r] * p1[q][0] + A[3 * r + 1] * p1[q][1] + A[3 * r + 2] * p1[q][2];
					cross(p2, Ap, pxAp);
					b[q] = (pxAp[0] * pxe[0] + pxAp[1] * pxe[1] + pxAp[2] * pxe[2]) / (pxe[0] * pxe[0] + pxe[1] * pxe[1] + pxe[2] * pxe[2]);
				}
				// (M^-1 b) by cofactors, M = rows (x, y, 1) of the triplet in the first image
				const double *M0 = p1[0], *M1 = p1[1], *M2 = p1[2];
				const double c00 = M1[1] * M2[2] - M1[2] * M2[1], c01 = M1[2] * M2[0] - M1[0] * M2[2], c02 = M1[0] * M2[1] - M1[1] * M2[0];
				const double det = M0[0] * c00 + M0[1] * c01 + M0[2] * c02;
				const double id = 1.0 / det;
				const double Mi[9] = {c00 * id, (M0[2] * M2[1] - M0[1] * M2[2]) * id, (M0[1] * M1[2] - M0[2] * M1[1]) * id,
				                      c01 * id, (M0[0] * M2[2] - M0[2] * M2[0]) * id, (M0[2] * M1[0] - M0[0] * M1[2]) * id,
				                      c02 * id, (M0[1] * M2[0] - M0[0] * M2[1]) * id, (M0[0] * M1[1] - M0[1] * M1[0]) * id};
				double Mb[3];
				for (int r = 0; r < 3; ++r)
					Mb[r] = Mi[3 * r] * b[0] + Mi[3 * r + 1] * b[1] + Mi[3 * r + 2] * b[2];
				for (int r = 0; r < 3; ++r)
					for (int c = 0; c < 3; ++c)
						H[3 * r + c] = A[3 * r + c] - e[r] * Mb[c];
				int inlier_number = 3;
				for (int i = 0; i < 7; ++i)
				{
					const uint32_t id7 = sample[i];
					if (id7 == idx[0] || id7 == idx[1] || id7 == idx[2])
						continue;
					if (sq_homography_transfer(pts + 4 * (size_t)id7, H) < squared_homography_threshold)
						++inlier_number;
				}
				if (inlier_number >= 5)
					h_degenerate = true;
			}
			if (!h_degenerate)
				return true; // :571
			std::vector<uint32_t> h_inliers; // :501-509
			for (uint32_t i : inliers)
				if (sq_homography_transfer(pts + 4 * (size_t)i, H) < squared_homography_threshold)
					h_inliers.push_back(i);
			if (h_inliers.size() < 4) // homography_estimator.nonMinimalSampleSize()
				return false;
			std::vector<double> models;
			std::vector<uint8_t> ok;
			{
				std::vector<uint32_t> offsets(1, 0u), sizes(1, (uint32_t)h_inliers.size());
				if (!refit(h_inliers.data(), offsets, sizes, nullptr, 0, GCR_WEIGHT_NONE, models, ok, GCR_H))
				{
					failed = true;
					return false;
				}
				if (!ok[0]) // :530-531
					return false;
			}
			++st.degensac_runs;
			Driver inner;
			Backend *inner_be = nested_backend ? nested_backend(this) : be; // the inner run works on the points the outer run uploaded
			if (!inner_be)
			{
				failed = true;
				return false;
			}
			inner.be = inner_be;
			inner.problem = GCR_F;
			inner.variant = 1;
			inner.nested = true;
			inner.f_min_ratio = 0.0;
			memcpy(inner.pp_H, models.data(), sizeof(inner.pp_H));
			inner.n = n;
			inner.pts = pts;
			gcr_run_settings_default(&inner.S); // a default-constructed GCRANSAC<> (settings.h:63-84) + :547-551
			inner.S.threshold = truncated_threshold;
			inner.S.spatial_coherence_weight = 0;
			inner.S.confidence = 0.99;
			inner.S.use_degensac = 0;
			inner.S.sampler = 0; // sampler::UniformSampler (:546)
			inner.S.seed = S.seed + st.degensac_runs * 0xD1B54A32D192ED03ull;
			inner.S.batch = S.batch;
			double inner_model[9];
			std::vector<uint8_t> inner_mask(n);
			const int r = inner.run(inner_model, inner_mask.data());
			st.gpu_batches += inner.st.gpu_batches;
			st.exact_rescores += inner.st.exact_rescores;
			st.models_scored += inner.st.models_scored;
			st.graph_cut_seconds += inner.st.graph_cut_seconds;
			if (r < 0)
			{
				error = inner.error;
				failed = true;
				return false;
			}
			if ((size_t)r > inliers.size()) // :561-567
			{
				updated = true;
				memcpy(model, inner_model, sizeof(inner_model));
				++st.degensac_updates;
			}
			return true;
		}

		// ---- one GPU batch of the main loop: samples [first, first + B) -> models, valid flags, FP32 scores, exact counts ----
		bool run_batch(uint64_t first)
		{
			if (live_buffers && !h_model_have.empty()) // a lazy container may still point into the batch about to be overwritten
				for (int b = 0; b < 2; ++b)
					if (live_buffers[b].lazy && !resolve_model(live_buffers[b]))
						return false;
			BatchSpec spec;
			spec.variant = variant;
			spec.pp_H = pp_H;
			spec.sampler = S.sampler;
			spec.first = first;
			spec.B = B;
			spec.m = m;
			spec.mm = mm;
			spec.seed = S.seed;
			spec.scoring = scoring;
			spec.threshold = S.threshold;
			if (S.sampler == 4 || S.sampler == 2)
			{
				// AdaptiveReorderingSampler / ProgressiveNapsacSampler (py/gcransac_python.cpp:4334-4359): the recurrence
				// over the sample number runs on the host, the batch's index sets go up in one copy
				explicit_samples.resize((size_t)B * m);
				if (S.sampler == 4)
					ar_sampler.samples(first, (size_t)B, explicit_samples.data());
				else
					pnapsac_sampler.samples(first, (size_t)B, explicit_samples.data());
				spec.explicit_samples = explicit_samples.data();
			}
			// applyDegensac needs the minimal sample behind a candidate model (fundamental_estimator.h:344)
			spec.want_samples = S.use_degensac && problem == GCR_F && variant == 0;
			batch.samples.clear();
			if (!be->run_batch(spec, batch))
				return fail("run_batch");
			if (spec.want_samples && spec.explicit_samples)
				batch.samples.assign(explicit_samples.begin(), explicit_samples.begin() + (size_t)batch.B * m);
			B = batch.B;
			batch_first = first;
			batch_count = B;
			++st.gpu_batches;
			return true;
		}

		// The sequential loop re-scores in FP64 every model whose FP32 score comes within 2e-4 of the best so far.  Only
		// the running records of the batch can do that, O(log B) of them: they are found here from the FP32 scores
		// (running maximum started at the current best) and evaluated in one launch instead of one launch + two
		// synchronisations each.  A model the loop wants and this prediction missed (the best so far can lag behind the
		// FP32 record when isValidModel rejects it) is evaluated on demand as before.
		bool prefetch_candidates(double best_value, uint64_t from)
		{
			cand_slot.clear();
			double record = best_value;
			for (size_t j = (size_t)(from - batch_first); j < (size_t)batch_count; ++j)
				for (int q = 0; q < mm; ++q)
				{
					const size_t slot = j * mm + q;
					if (!h_valid[slot] || !((double)h_scores[slot] > record * (1.0 - rank_slack)))
						continue;
					cand_slot.push_back(slot);
					record = std::max(record, (double)h_scores[slot]);
				}
			const size_t C = cand_slot.size();
			if (C < 2 || C > 64)
			{
				cand_slot.clear(); // nothing to gain / an implausible batch (pure noise): on demand
				return true;
			}
			if (!fetch_models(cand_slot))
				return false;
			std::vector<double> models;
			for (size_t slot : cand_slot)
				models.insert(models.end(), &h_models[slot * ml], &h_models[slot * ml] + ml);
			return exact_many(models.data(), C, cand_score, cand_count, cand_idx);
		}

		// FP64 getScore of C models (host, 9 doubles each) in one launch and one synchronise: scores, inlier counts and
		// the inlier lists ([C][n], the first count entries of each row)
		bool exact_many(const double *models, size_t C, std::vector<double> &score, std::vector<uint32_t> &count,
		                std::vector<uint32_t> &idx)
		{
			if (!be->exact_many(models, C, scoring, S.threshold, score, count, idx))
				return fail("exact_many");
			st.exact_rescores += C;
			return true;
		}

		// exact() of slot `slot` of the current batch, from the prefetched candidates when it is one of them
		bool exact_slot(size_t slot, const double *model, Score &s, std::vector<uint32_t> &inliers)
		{
			for (size_t c = 0; c < cand_slot.size(); ++c)
				if (cand_slot[c] == slot)
				{
					s.value = cand_score[c];
					s.inliers = cand_count[c];
					inliers.assign(cand_idx.begin() + c * (size_t)n, cand_idx.begin() + c * (size_t)n + cand_count[c]);
					return true;
				}
			return exact(model, s, &inliers);
		}

		// ---- GCRANSAC::labeling (GCRANSAC.h:913-1021): the energy terms on the GPU (unary from the model's FP64
		// residuals, pairwise over the run's edge list: gcr_labeling_terms), the BK min-cut on the host ----
		bool labeling(const double *model, std::vector<uint32_t> &inliers)
		{
			const double lambda = S.spatial_coherence_weight;
			const size_t E = lambda > 0 ? n_edges : 0;
			// pairwise weights: 8 bytes per edge read back from the device kernel when the graph came from the device anyway
			// (n > 16384), else formed here from the per-point clamped distances (24 n bytes come back instead of 16 n + 8 E)
			static const bool force_device_pairwise = getenv("GCR_DEVICE_PAIRWISE") != nullptr;
			const bool device_pairwise = E > 0 && be->keeps_edges && (n > 16384 || force_device_pairwise);
			if (!be->labeling_terms(model, S.threshold, lambda, E, device_pairwise, h_unary, h_pairwise))
				return fail("labeling_terms");
			++st.exact_rescores; // (one FP64 pass over the points, as the residual read-back it replaces)
			const auto t0 = std::chrono::steady_clock::now();
			const double *dist = h_unary.data() + 2 * (size_t)n;
			// (the reference passes edge_num_max = 0 and lets the arc array grow by reallocation, ~20 times with a pointer fix-up
			// pass each; the edge count is known here -- same graph, same arc order, same cut)
			Energy<double, double, double> *problem_graph = new Energy<double, double, double>(n, (int)E, NULL);
			for (int i = 0; i < n; ++i)
				problem_graph->add_node();
			for (int i = 0; i < n; ++i)
				problem_graph->add_term1(i, h_unary[2 * (size_t)i], h_unary[2 * (size_t)i + 1]);
			for (size_t e = 0; e < E; ++e)
			{
				const uint32_t i = edges[2 * e], j = edges[2 * e + 1];
				const double w = device_pairwise ? h_pairwise[e] : 0.5 * (dist[i] + dist[j]) * lambda; // e00 * lambda (:996-1003)
				problem_graph->add_term2((int)i, (int)j, w, lambda, lambda, 0 * lambda);
			}
			problem_graph->minimize();
			inliers.clear();
			for (int i = 0; i < n; ++i)
				if (problem_graph->what_segment(i) == Graph<double, double, double>::SINK)
					inliers.push_back((uint32_t)i);
			delete problem_graph;
			st.graph_cut_seconds += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			return true;
		}

		// the run's edge list from a host-side structure (grids of up to 16384 points are built on the host)
		bool upload_edges()
		{
			n_edges = edges.size() / 2;
			return be->set_edges(edges) || fail("set_edges");
		}

		// ---- non-minimal refit of `count` index sets (one GPU launch) ----
		// the reference's default F estimator refines every non-minimal fit by bundle adjustment (types.h:78-80); DEGENSAC's
		// inner estimator (variant 1) keeps the bare 8-point solver (fundamental_estimator.h:532-533)
		int ba_frame_of(int rp, int ba_frame_override) const
		{
			return ba_frame_override ? ba_frame_override : (rp == GCR_F && variant == 0 && (S.f_bundle_adjustment & 1)) ? GCR_BA_HARTLEY : 0;
		}
		bool refit(const uint32_t *flat, const std::vector<uint32_t> &offsets, const std::vector<uint32_t> &sizes,
		           const double *weights, size_t n_weights, int weight_mode, std::vector<double> &models, std::vector<uint8_t> &ok,
		           int refit_problem = -1, int ba_frame_override = 0)
		{
			const int rp = refit_problem < 0 ? problem : refit_problem;
			return be->refit_sets(rp, ba_frame_of(rp, ba_frame_override), flat, offsets, sizes, weights, n_weights, weight_mode, models, ok) ||
			       fail("refit_sets");
		}

		// ---- graphCutLocalOptimization (GCRANSAC.h:780-910) ----
		bool local_optimization(InlierBuffer &best_inliers_out, double *best_model, Score &best_score)
		{
			const size_t inlier_limit = 7 * (size_t)m; // estimator.inlierLimit()
			const int trials = (int)S.max_local_optimization_number;
			Score max_score = best_score;
			double lbest_model[12];
			memcpy(lbest_model, best_model, ml * sizeof(double));
			std::vector<uint32_t> lbest_inliers, inliers, tmp;
			++st.local_optimization_number;
			while (++st.graph_cut_number < S.max_graph_cut_number)
			{
				bool updated = false;
				if (!labeling(lbest_model, inliers))
					return false;
				const size_t sample_size = std::min(inlier_limit, inliers.size());
				std::vector<double> models;
				std::vector<uint8_t> ok;
				std::vector<float> scores32;
				int n_models = 0;
				LoSpec lo;
				lo.pool = inliers.data();
				lo.pool_size = inliers.size();
				lo.sample_size = sample_size;
				lo.first_index = st.lo_samples;
				lo.trials = trials;
				lo.seed = S.seed;
				lo.stream = S.sampler == 1 ? 2u : 1u;
				lo.refit_problem = problem;
				lo.ba_frame = ba_frame_of(problem, 0);
				lo.rm = rm;
				lo.scoring = scoring;
				lo.threshold = S.threshold;
				if (sample_size < inliers.size() && trials > 0)
				{
					// inner RANSAC as one batch: `trials` pool samples -> refits -> FP32 scores
					lo.sample_pool = true;
					if (!be->lo_batch(lo, models, ok, scores32))
						return fail("lo_batch");
					st.lo_samples += trials;
					n_models = trials * rm; // trial-major, the solver's model order inside a trial (:866)
				}
				else if ((size_t)m < inliers.size() && trials > 0)
				{
					// all inliers, every trial fits the same model: once is enough (later trials cannot beat it, strict '<')
					lo.sample_pool = false;
					if (!be->lo_batch(lo, models, ok, scores32))
						return fail("lo_batch");
					n_models = 0;
					for (int q = 0; q < rm; ++q)
						if (ok[q]) n_models = rm;
				}
				if (n_models > 0)
				{
					st.models_scored += n_models;
					// the trials that can become the round's best: running records of the FP32 scores, evaluated in FP64
					// in one launch (the sequential loop below falls back to a single evaluation for anything else)
					std::vector<int> rec;
					std::vector<double> rec_models, rec_score;
					std::vector<uint32_t> rec_count, rec_idx;
					{
						double record = max_score.value;
						for (int t = 0; t < n_models && rec.size() < 64; ++t)
							if (ok[t] && (double)scores32[t] > record * (1.0 - rank_slack))
							{
								rec.push_back(t);
								rec_models.insert(rec_models.end(), &models[(size_t)t * ml], &models[(size_t)t * ml] + ml);
								record = std::max(record, (double)scores32[t]);
							}
						if (rec.size() >= 2)
						{
							if (!exact_many(rec_models.data(), rec.size(), rec_score, rec_count, rec_idx))
								return false;
						}
						else
							rec.clear();
					}
					for (int t = 0; t < n_models; ++t)
					{
						if (!ok[t])
							continue;
						if (!((double)scores32[t] > max_score.value * (1.0 - rank_slack)))
							continue; // FP32 ranking says it cannot beat the best so far
						Score s;
						size_t c = 0;
						while (c < rec.size() && rec[c] != t) ++c;
						if (c < rec.size())
						{
							s.value = rec_score[c];
							s.inliers = rec_count[c];
							tmp.assign(rec_idx.begin() + c * (size_t)n, rec_idx.begin() + c * (size_t)n + rec_count[c]);
						}
						else if (!exact(&models[(size_t)t * ml], s, &tmp))
							return false;
						if (max_score.value < s.value) // :880
						{
							updated = true;
							max_score = s;
							memcpy(lbest_model, &models[(size_t)t * ml], ml * sizeof(double));
							lbest_inliers.swap(tmp);
						}
					}
				}
				if (!updated)
					break;
			}
			if (best_score.value < max_score.value) // :900
			{
				best_score = max_score;
				memcpy(best_model, lbest_model, ml * sizeof(double));
				best_inliers_out.lazy = false;
				best_inliers_out.idx.swap(lbest_inliers);
			}
			return true;
		}

		// refit + FP64 getScore of every model the fit returned, one submission (Backend::refit_and_score)
		struct Fitted
		{
			std::vector<double> models, score, residuals;
			std::vector<uint8_t> ok;
			std::vector<uint32_t> count, idx;
		};
		bool refit_scored(const uint32_t *flat, uint32_t size, const double *weights, size_t n_weights, int weight_mode, bool want_residuals,
		                  Fitted &f, int refit_problem = -1, int ba_frame_override = 0)
		{
			const int rp = refit_problem < 0 ? problem : refit_problem;
			std::vector<uint32_t> offsets(1, 0u), sizes(1, size);
			if (!be->refit_and_score(rp, ba_frame_of(rp, ba_frame_override), flat, offsets, sizes, weights, n_weights, weight_mode, scoring, T,
			                         S.threshold, want_residuals, f.models, f.ok, f.score, f.count, f.idx, f.residuals))
				return fail("refit_and_score");
			for (uint8_t v : f.ok)
				st.exact_rescores += v ? 1 : 0;
			return true;
		}
		// model q of a scored fit as the reference's getScore would have left it: score, inlier list
		void take_scored(const Fitted &f, int q, Score &s, std::vector<uint32_t> &inliers) const
		{
			s.value = f.score[q];
			s.inliers = f.count[q];
			inliers.assign(f.idx.begin() + (size_t)q * n, f.idx.begin() + (size_t)q * n + f.count[q]);
		}

		// ---- iteratedLeastSquaresFitting (GCRANSAC.h:630-758), single-model estimators ----
		// One round of the reference = getScore (residuals for the weights) -> weighted fit -> getScore of the fit.  The
		// fit and the score of its result travel as one submission, and an accepted fit's residuals come back with it: they
		// are the next round's "getScore of the current model".
		bool iterated_least_squares(std::vector<uint32_t> &inliers, double *model, bool &success)
		{
			success = false;
			if (inliers.size() <= (size_t)m)
				return true;
			size_t iterations = 0;
			std::vector<double> weights(n, 0.0), r2;
			bool have_r2 = false;
			std::vector<uint32_t> tmp;
			Fitted fit;
			Score irls_best; // :648
			const bool weighting = problem != GCR_PNP; // _ModelEstimator::isWeightingApplicable() (:652): false for PnP
			while (++iterations < S.max_least_squares_iterations)
			{
				Score s;
				if (weighting)
				{
					if (!have_r2 && !exact(model, s, nullptr, &r2))
						return false;
					for (uint32_t idx : inliers) // Tukey bisquare weights, written by POINT index (:657-668)
					{
						const double w = std::max(0.0, 1.0 - r2[idx] / T);
						weights[idx] = w * w;
					}
					// ... and read by POSITION by the H / 8-point solvers, by POINT by the Stewenius solver
					// (solver_essential_matrix_five_point_stewenius.h:143): reproduced literally
					if (!refit_scored(inliers.data(), (uint32_t)inliers.size(), weights.data(), weights.size(),
					                  problem == GCR_E ? GCR_WEIGHT_BY_POINT : GCR_WEIGHT_POSITIONAL, true, fit))
						return false;
					for (uint32_t idx : inliers)
						weights[idx] = 0;
				}
				else if (!refit_scored(inliers.data(), (uint32_t)inliers.size(), nullptr, 0, GCR_WEIGHT_NONE, false, fit)) // :682-688
					return false;
				std::vector<int> found;
				for (int q = 0; q < rm; ++q)
					if (fit.ok[q]) found.push_back(q);
				if (found.empty()) // :690
					break;
				auto accept = [&](int q) {
					memcpy(model, &fit.models[(size_t)q * ml], ml * sizeof(double));
					inliers.swap(tmp);
					if (weighting)
					{
						r2.assign(fit.residuals.begin() + (size_t)q * n, fit.residuals.begin() + (size_t)(q + 1) * n);
						have_r2 = true;
					}
				};
				if (found.size() == 1) // :692-714
				{
					take_scored(fit, found[0], s, tmp);
					if (tmp.size() < (size_t)m)
						break;
					if (s.inliers <= inliers.size())
						break;
					accept(found[0]);
				}
				else // :715-748 several models: the best by inlier count, `best_score` lives across the rounds
				{
					bool updated = false;
					for (int q : found)
					{
						take_scored(fit, q, s, tmp);
						if (tmp.size() < (size_t)m)
							continue;
						if (s.inliers <= inliers.size())
							continue;
						if (s.inliers >= irls_best.inliers)
						{
							updated = true;
							irls_best = s;
							accept(q);
						}
					}
					if (!updated)
						break;
				}
			}
			success = iterations > 1;
			return true;
		}

		bool build_radius_graph()
		{
			have_radius_graph = false;
			// the CSR and labeling()'s edges (i, j > i), both built on the device
			if (!be->radius_graph(S.neighborhood_radius, nbr_offsets, nbr_idx, edges))
				return fail("radius_graph");
			n_edges = edges.size() / 2;
			have_radius_graph = true;
			return true;
		}

		// ProgressiveNapsacSampler<4>(&points, {16, 8, 4, 2}, m, {w1, h1, w2, h2}, 0.5) (py/gcransac_python.cpp:4335-4343):
		// one GridNeighborhoodGraph<4> per layer over the points currently uploaded, cell size = image size / cells
		bool build_pnapsac()
		{
			static const uint64_t layer_cells[4] = {16, 8, 4, 2};
			std::vector<gcr_host::GridLayer> layers(4);
			for (int d = 0; d < 4; ++d)
				if (!(S.image_sizes[d] > 0.0) || !std::isfinite(S.image_sizes[d]))
				{
					error = "sampler 2 (Progressive NAPSAC) needs gcr_run_settings.image_sizes = (w1, h1, w2, h2), all positive";
					return false;
				}
			if (n <= 16384)
			{
				// a few thousand points: four sorts on the host take less than the four round trips of the GPU build
				const double *p = (grid_pts && grid_pts != pts) ? grid_pts : pts;
				for (int l = 0; l < 4; ++l)
					gcr_host::build_grid_layer(p, (size_t)n, S.image_sizes, (size_t)layer_cells[l], layers[l]);
				if (!pnapsac_sampler.initialize(std::move(layers), (size_t)n, m, 0.5, S.seed))
				{
					error = "sampler 2 (Progressive NAPSAC): bad sample size / point number";
					return false;
				}
				return true;
			}
			for (int l = 0; l < 4; ++l)
			{
				double cell_sizes[4];
				for (int d = 0; d < 4; ++d)
					cell_sizes[d] = S.image_sizes[d] / layer_cells[l];
				uint32_t n_cells = 0;
				uint64_t neighbor_number = 0;
				gcr_host::GridLayer &g = layers[l];
				if (!be->grid_device(cell_sizes, layer_cells[l], g.sorted_idx, g.cell_offsets, n_cells, neighbor_number, nullptr))
					return fail("grid_device");
				g.cell_of_point.resize(n);
				for (uint32_t c = 0; c < n_cells; ++c)
					for (uint32_t q = g.cell_offsets[c]; q < g.cell_offsets[c + 1]; ++q)
						g.cell_of_point[g.sorted_idx[q]] = c;
			}
			if (!pnapsac_sampler.initialize(std::move(layers), (size_t)n, m, 0.5, S.seed))
			{
				error = "sampler 2 (Progressive NAPSAC): bad sample size / point number";
				return false;
			}
			return true;
		}

		bool build_grid()
		{
			if (problem == GCR_PNP && S.neighborhood != 1 && S.spatial_coherence_weight > 0.0 && S.do_local_optimization)
			{
				// find6DPose_ knows one neighbourhood: FlannNeighborhoodGraph over the 5-column rows (py/gcransac_python.cpp:1087-1098)
				error = "the absolute-pose problem uses the radius neighbourhood (gcr_run_settings.neighborhood = 1)";
				return false;
			}
			if (S.neighborhood == 1)
			{
				have_grid = false;
				if (!(S.spatial_coherence_weight > 0.0) || !S.do_local_optimization)
					return true;
				return build_radius_graph();
			}
			have_grid = false;
			if (!(S.spatial_coherence_weight > 0.0) || !S.do_local_optimization)
				return true;
			if (n <= 16384)
			{
				// a few thousand points: binning + one sort on the host costs less than the ~10 launches and two
				// synchronising read-backs of the GPU build (which every host thread's pair would queue on the context)
				gcr_host::GridLayer g;
				gcr_host::build_grid_cells((grid_pts && grid_pts != pts) ? grid_pts : pts, (size_t)n, S.cell_sizes, (size_t)S.cells_per_axis, g);
				cell_of_point.swap(g.cell_of_point);
				sorted_idx.swap(g.sorted_idx);
				cell_offsets.swap(g.cell_offsets);
				have_grid = true;
				// the pairs (i, j > i) of every cell, i ascending, the cell's list ascending: once per run
				edges.clear();
				for (int i = 0; i < n; ++i)
				{
					const uint32_t c = cell_of_point[i];
					for (uint32_t q = cell_offsets[c]; q < cell_offsets[c + 1]; ++q)
						if ((int)sorted_idx[q] > i)
						{
							edges.push_back((uint32_t)i);
							edges.push_back(sorted_idx[q]);
						}
				}
				return upload_edges();
			}
			uint32_t n_cells = 0;
			uint64_t neighbor_number = 0;
			// cells and labeling()'s edge list from the sorted cells, both on the device
			if (!be->grid_device(S.cell_sizes, S.cells_per_axis, sorted_idx, cell_offsets, n_cells, neighbor_number, &edges))
				return fail("grid_device");
			cell_of_point.resize(n);
			for (uint32_t c = 0; c < n_cells; ++c)
				for (uint32_t q = cell_offsets[c]; q < cell_offsets[c + 1]; ++q)
					cell_of_point[sorted_idx[q]] = c;
			have_grid = true;
			n_edges = edges.size() / 2;
			return true;
		}

		// ---- GCRANSAC::run ----
		int run(double *model_out, uint8_t *inlier_mask)
		{
			memset(&st, 0, sizeof(st));
			ml = gcr_model_len(problem);
			cols = problem == GCR_PNP ? 5 : 4;
			be->problem = problem;
			be->n = n;
			be->ml = ml;
			be->cols = cols;
			m = variant == 1 ? 2 : gcr_sample_size(problem);
			mm = variant == 1 ? 1 : gcr_max_models(problem);
			B = S.batch > 0 ? S.batch : 1024;
			truncated_threshold = 3.0 / 2.0 * S.threshold;
			T = truncated_threshold * truncated_threshold;
			scoring = S.scoring == GCR_MAGSAC ? GCR_MAGSAC : GCR_MSAC;
			rank_slack = scoring == GCR_MAGSAC ? 1e-3 : 2e-4;
			rm = gcr_max_refit_models(problem);
			std::unique_ptr<PhaseTimer> setup_timer(new PhaseTimer(st.phase_seconds[0]));
			if (nested)
				have_grid = false; // spatial_coherence_weight = 0: no neighbourhood (the reference passes nullptr, :558)
			else if (grid_pts && grid_pts != pts)
			{
				// findEssentialMatrix_ builds the grid on the pixel coordinates and runs on the K-normalised ones
				// (py/gcransac_python.cpp:2592-2600, 2641-2645).  The pixel coordinates only go to the GPU when a GPU build
				// will read them: the radius graph, or grids over more than 16384 points.
				const bool graph_wanted = S.spatial_coherence_weight > 0.0 && S.do_local_optimization;
				if (((graph_wanted && S.neighborhood == 1) || ((graph_wanted || S.sampler == 2) && n > 16384)) &&
				    !be->upload(grid_pts, n, cols))
				{
					fail("upload");
					return -1;
				}
				if (!build_grid())
					return -1;
				if (S.sampler == 2 && !build_pnapsac()) // on the pixel coordinates (py/gcransac_python.cpp:2650)
					return -1;
			}
			if (!nested)
			{
				if (!be->upload(pts, n, cols))
				{
					fail("upload");
					return -1;
				}
				if (!(grid_pts && grid_pts != pts) && !build_grid())
					return -1;
				if (S.sampler == 2 && problem == GCR_PNP)
				{
					error = "Progressive NAPSAC is not usable for the absolute pose problem."; // py/gcransac_python.cpp:1113-1120
					return -1;
				}
				if (S.sampler == 2 && !(grid_pts && grid_pts != pts) && !build_pnapsac())
					return -1;
				if (S.sampler == 4)
				{
					if (!S.point_probabilities ||
					    !ar_sampler.initialize(S.point_probabilities, (size_t)n, m, S.sampler_variance, S.seed))
					{
						error = "sampler 4 (AR-Sampler) needs gcr_run_settings.point_probabilities: one finite, non-negative value per "
						        "correspondence, not all zero";
						return -1;
					}
				}
				if (S.sampler == 3)
				{
					// ImportanceSampler's constructor (samplers/importance_sampler.h:67-82): one probability per point
					if (!S.point_probabilities)
					{
						error = "sampler 3 (importance sampler) needs gcr_run_settings.point_probabilities";
						return -1;
					}
					if (!be->set_point_probabilities(S.point_probabilities, n))
					{
						fail("set_point_probabilities");
						return -1;
					}
				}
			}
			setup_timer.reset();
			const double log_probability = std::log(1.0 - S.confidence);
			uint64_t max_iteration = iteration_number(1, n, m, log_probability);
			const auto loop_t0 = std::chrono::steady_clock::now();
			const double batch_before = st.phase_seconds[1], lo_before = st.phase_seconds[3];
			uint64_t &it = st.iteration_number;
			uint64_t cursor = 0; // next sample of the main stream
			Score best;
			double best_model[12] = {0};
			InlierBuffer buf[2];
			live_buffers = buf;
			int offset = 0;
			batch_count = 0;

			while (S.min_iteration_number > it || it < std::min<uint64_t>(max_iteration, S.max_iteration_number))
			{
				bool do_lo = false;
				++it;
				uint64_t unsuccessful = 0;
				long long success_at = -1;
				while (unsuccessful < S.max_unsuccessful_model_generations)
				{
					const uint64_t g = cursor++;
					if (batch_count == 0 || g < batch_first || g >= batch_first + (uint64_t)batch_count)
					{
						PhaseTimer t(st.phase_seconds[1]);
						// batch size: the configured one, 4 x / 16 x that while the loop still has far to go -- a 5-point run at
						// 85 % outliers needs ~60 000 samples, and 4096 of them fill a third of the GPU (the samples are a
						// function of their index, so batch boundaries do not change any result)
						const uint64_t horizon = std::min<uint64_t>(max_iteration, S.max_iteration_number);
						const int base = S.batch > 0 ? S.batch : 1024;
						const uint64_t far = (st.gpu_batches > 0 && horizon > it) ? horizon - it : 0;
						B = far > 64ull * (uint64_t)base ? 16 * base : far > 8ull * (uint64_t)base ? 4 * base : base; // (the first
						// batch usually finds a model that shortens the horizon from "unbounded" to a few hundred samples)
						if (!run_batch(g) || !prefetch_candidates(best.value, g))
							return -1;
					}
					const size_t j = (size_t)(g - batch_first);
					bool any = false;
					for (int q = 0; q < mm; ++q)
						any |= h_valid[j * mm + q] != 0;
					if (any) { success_at = (long long)j; break; }
					++unsuccessful;
				}
				st.main_samples = cursor;
				it += unsuccessful;
				if (success_at < 0 && S.sampler == 4)
				{
					// GCRANSAC.h:297-339 gave up: the end-of-iteration update (:474) runs on the last sample once more
					ar_sampler.repeat_update(cursor - 1);
					batch_count = 0;
				}
				if (success_at >= 0)
				{
					for (int q = 0; q < mm; ++q)
					{
						const size_t slot = (size_t)success_at * mm + q;
						if (!h_valid[slot])
							continue;
						++st.models_scored;
						// the reference's getScore writes this model's inliers into the current container
						InlierBuffer &cur = buf[offset];
						cur.lazy = true;
						cur.slot = (long long)slot; // the model stays on the device unless somebody asks for it
						cur.lazy_count = h_counts[slot];
						cur.idx.clear();
						if (!((double)h_scores[slot] > best.value * (1.0 - rank_slack)))
							continue; // FP32 ranking: cannot beat the best so far (:440)
						const double *model = model_of(slot);
						if (!model)
							return -1;
						memcpy(cur.model, model, ml * sizeof(double));
						cur.slot = -1;
						Score s;
						if (!exact_slot(slot, model, s, cur.idx))
							return -1;
						cur.lazy = false;
						if (!(best.value < s.value)) // :440
							continue;
						double candidate[12];
						memcpy(candidate, model, ml * sizeof(double));
						bool updated = false, failed = false;
						const bool valid_model = is_valid_model(candidate, cur.idx, h_samples.empty() ? nullptr : &h_samples[(size_t)success_at * m],
						                                        updated, failed);
						if (failed)
							return -1;
						if (valid_model) // :441-446
						{
							if (updated) // :449-456: the replaced model is scored again, into the same container
							{
								memcpy(cur.model, candidate, ml * sizeof(double));
								if (!exact(candidate, s, &cur.idx))
									return -1;
							}
							offset = (offset + 1) % 2;
							memcpy(best_model, candidate, ml * sizeof(double));
							best = s;
							do_lo = it > S.min_iteration_number_before_lo && best.inliers > (uint64_t)m; // :463-464
							max_iteration = iteration_number(best.inliers, n, m, log_probability);
						}
					}
				}
				if (S.do_local_optimization && do_lo)
				{
					++st.local_optimization_number; // :485 (and once more inside, :805 -- the reference counts twice)
					PhaseTimer t(st.phase_seconds[3]);
					if (!local_optimization(buf[offset], best_model, best))
						return -1;
					max_iteration = iteration_number(best.inliers, n, m, log_probability);
				}
			}
			st.phase_seconds[2] += std::chrono::duration<double>(std::chrono::steady_clock::now() - loop_t0).count() -
			                       (st.phase_seconds[1] - batch_before) - (st.phase_seconds[3] - lo_before);

			for (int i = 0; i < n; ++i)
				inlier_mask[i] = 0;
			if (best.inliers <= (uint64_t)m) // :521-527
				return 0;
			if (S.do_local_optimization && st.local_optimization_number == 0) // :530-543
			{
				++st.local_optimization_number;
				PhaseTimer t(st.phase_seconds[3]);
				if (!local_optimization(buf[offset], best_model, best))
					return -1;
			}
			// :547-555 consistency of the stored inlier set (by SIZE, as the reference does)
			if (buf[offset].size() != best.inliers)
				offset = (offset + 1) % 2;
			if (buf[offset].size() != best.inliers)
			{
				buf[offset].lazy = false;
				if (!exact(best_model, best, &buf[offset].idx))
					return -1;
			}
			if (!materialise(buf[offset]))
				return -1;

			bool iterative_refitting_applied = false;
			if (S.do_final_iterated_least_squares) // :559-586
			{
				PhaseTimer t(st.phase_seconds[4]);
				double model[12];
				memcpy(model, best_model, ml * sizeof(double));
				bool success = false;
				if (!iterated_least_squares(buf[offset].idx, model, success))
					return -1;
				if (success)
				{
					const int idx2 = (offset + 1) % 2;
					Score cur;
					buf[idx2].lazy = false;
					if (!exact(model, cur, &buf[idx2].idx))
						return -1;
					if (best.value < cur.value)
					{
						iterative_refitting_applied = true;
						memcpy(best_model, model, ml * sizeof(double));
						offset = idx2;
					}
				}
			}
			PhaseTimer final_timer(st.phase_seconds[5]);
			if (!iterative_refitting_applied) // :588-617: one least-squares fit on the first best.inliers entries
			{
				const uint32_t cnt = (uint32_t)std::min<size_t>(best.inliers, buf[offset].idx.size());
				Fitted fit;
				if (!refit_scored(buf[offset].idx.data(), cnt, nullptr, 0, GCR_WEIGHT_NONE, false, fit))
					return -1;
				for (int q = 0; q < rm; ++q) // :597-613 every returned model, in order; the best SCORE is not updated there
				{
					if (!fit.ok[q])
						continue;
					const int idx2 = (offset + 1) % 2;
					Score cur;
					buf[idx2].lazy = false;
					take_scored(fit, q, cur, buf[idx2].idx);
					if (best.value < cur.value)
					{
						memcpy(best_model, &fit.models[(size_t)q * ml], ml * sizeof(double));
						offset = idx2;
					}
				}
			}
			st.score = best.value;
			if (problem == GCR_F && !nested && (S.f_bundle_adjustment & 2) && buf[offset].idx.size() > 7)
			{
				// the post-run bundle adjustment of findFundamentalMatrix_ (py/gcransac_python.cpp:1832-1887)
				std::vector<uint32_t> &inl = buf[offset].idx;
				std::vector<double> r2, weights(inl.size(), 0.0), models;
				std::vector<uint8_t> ok;
				Score s;
				if (!exact(best_model, s, nullptr, &r2))
					return -1;
				for (size_t k = 0; k < inl.size(); ++k)
					weights[k] = std::max(0.0, 1.0 - r2[inl[k]] / T);
				// the driver code scores the polished model with an MSACScoringFunction of its own (:1866-1875), whatever
				// scoring function the run was instantiated with
				const int run_scoring = scoring;
				scoring = GCR_MSAC;
				Fitted fit;
				const bool fitted = refit_scored(inl.data(), (uint32_t)inl.size(), weights.data(), weights.size(), GCR_WEIGHT_POSITIONAL, false, fit,
				                                 GCR_F, GCR_BA_PIXEL);
				scoring = run_scoring;
				if (!fitted)
					return -1;
				if (fit.ok[0])
				{
					const int idx2 = (offset + 1) % 2;
					Score cur;
					buf[idx2].lazy = false;
					take_scored(fit, 0, cur, buf[idx2].idx);
					if (st.score < cur.value) // :1877
					{
						memcpy(best_model, fit.models.data(), 9 * sizeof(double));
						st.score = cur.value;
						offset = idx2;
					}
				}
			}
			if (problem == GCR_PNP && S.pnp_bundle_adjustment && buf[offset].idx.size() > 3)
			{
				// find6DPose_ (py/gcransac_python.cpp:1227-1238): PnPBundleAdjustment on the inliers, its result REPLACES the model
				// (no score comparison, the inlier set stays)
				std::vector<double> models;
				std::vector<uint8_t> ok;
				std::vector<uint32_t> offsets(1, 0u), sizes(1, (uint32_t)buf[offset].idx.size());
				if (!refit(buf[offset].idx.data(), offsets, sizes, nullptr, 0, GCR_WEIGHT_NONE, models, ok))
					return -1;
				if (ok[0])
					memcpy(best_model, models.data(), ml * sizeof(double));
			}
			if (problem == GCR_E && !nested && S.e_bundle_adjustment && buf[offset].idx.size() > 5)
			{
				// the post-run bundle adjustment of findEssentialMatrix_ (py/gcransac_python.cpp:2766-2821): weights 1 - r^2 / T by
				// inlier POSITION, EssentialMatrixBundleAdjustmentSolver on the inliers (gcr_refit_e_ba: Stewenius -> poses -> LM),
				// every model it returns scored by an MSAC scoring function of the driver's own, in order, strict '<'
				std::vector<uint32_t> &inl = buf[offset].idx;
				std::vector<double> r2, weights(inl.size(), 0.0);
				Score s;
				if (!exact(best_model, s, nullptr, &r2))
					return -1;
				for (size_t k = 0; k < inl.size(); ++k)
					weights[k] = std::max(0.0, 1.0 - r2[inl[k]] / T);
				const int R = gcr_max_e_ba_models();
				std::vector<double> models;
				std::vector<uint8_t> ok;
				if (!be->refit_e_ba(inl.data(), inl.size(), weights.data(), models, ok))
				{
					fail("refit_e_ba");
					return -1;
				}
				const int run_scoring = scoring;
				scoring = GCR_MSAC; // (:2797-2798, whatever scoring function the run was instantiated with)
				for (int q = 0; q < R; ++q)
				{
					if (!ok[q])
						continue;
					const int idx2 = (offset + 1) % 2;
					Score cur;
					buf[idx2].lazy = false;
					if (!exact(&models[(size_t)q * 9], cur, &buf[idx2].idx))
					{
						scoring = run_scoring;
						return -1;
					}
					if (st.score < cur.value) // :2813
					{
						memcpy(best_model, &models[(size_t)q * 9], 9 * sizeof(double));
						st.score = cur.value;
						offset = idx2;
					}
				}
				scoring = run_scoring;
			}
			for (uint32_t i : buf[offset].idx)
				inlier_mask[i] = 1;
			memcpy(model_out, best_model, ml * sizeof(double));
			return (int)buf[offset].idx.size();
		}
	};

	thread_local std::string g_driver_error;
}

#include "gcr_batch_engine_impl.h" // the lock-step multi-pair driver over the same Driver class

extern "C" void gcr_run_settings_default(gcr_run_settings *s)
{
	if (!s)
		return;
	memset(s, 0, sizeof(*s));
	s->threshold = 2.0;              // settings.h:82
	s->confidence = 0.95;            // :83
	s->spatial_coherence_weight = 0.14;
	s->max_iteration_number = std::numeric_limits<uint64_t>::max();
	s->min_iteration_number = 20;
	s->max_local_optimization_number = 10;
	s->min_iteration_number_before_lo = 20;
	s->max_graph_cut_number = 10;
	s->max_least_squares_iterations = 10;
	s->max_unsuccessful_model_generations = 100;
	s->do_local_optimization = 1;
	s->do_final_iterated_least_squares = 1;
	s->cells_per_axis = 4;
	s->seed = 1;
	s->batch = 1024;
	s->use_degensac = 0;
	s->f_bundle_adjustment = 0;
	s->sampler = 0;
	s->neighborhood = 0;
	s->neighborhood_radius = 20.0;
	s->point_probabilities = nullptr;
	s->sampler_variance = 0.1;
	s->scoring = GCR_MSAC;
	s->e_bundle_adjustment = 0;
	s->pnp_bundle_adjustment = 0;
	for (int d = 0; d < 4; ++d)
		s->image_sizes[d] = 0.0;
}

// samples [first, first + count) of the host-side Progressive NAPSAC sequence over n x 4 host points; the grid layers are
// built here on the host (the driver builds them with gcr_grid_build): no GPU involved, for the CPU tests.
extern "C" int gcr_host_sample_progressive_napsac(const double *points, int n, int m, const double *image_sizes, uint64_t seed,
                                                  uint64_t first, int count, uint32_t *out)
{
	if (!points || !image_sizes || !out || n < 1 || count < 0)
		return -1;
	static const size_t layer_cells[4] = {16, 8, 4, 2};
	std::vector<gcr_host::GridLayer> layers(4);
	for (int l = 0; l < 4; ++l)
		gcr_host::build_grid_layer(points, (size_t)n, image_sizes, layer_cells[l], layers[l]);
	gcr_host::ProgressiveNapsac s;
	if (!s.initialize(std::move(layers), (size_t)n, m, 0.5, seed))
		return -2;
	s.samples(first, (size_t)count, out);
	return 0;
}

// samples [first, first + count) of the host-side AR-Sampler sequence (gcr_host_samplers.h) -- what the driver uploads
// for sampler 4; repeat_at >= 0: that sample's update ran twice.  No GPU involved: the CPU tests check it against the
// reference's own class.
extern "C" int gcr_host_sample_adaptive_reordering(const double *probabilities, int n, int m, double variance, uint64_t seed,
                                                   uint64_t first, int count, long long repeat_at, uint32_t *out)
{
	gcr_host::AdaptiveReordering s;
	if (!out || count < 0 || !s.initialize(probabilities, (size_t)n, m, variance, seed))
		return -1;
	if (repeat_at >= 0)
		s.repeat_update((uint64_t)repeat_at);
	s.samples(first, (size_t)count, out);
	return 0;
}

extern "C" int gcr_find_model(gcr_ctx *ctx, int problem, const double *h_points, int n, const gcr_run_settings *settings,
                              double *model_out, uint8_t *inlier_mask, gcr_run_stats *stats)
{
	if (!ctx || !h_points || !settings || !model_out || !inlier_mask || n <= 0)
		return -2;
	if (problem != GCR_H && problem != GCR_F && problem != GCR_E && problem != GCR_PNP)
		return -3;
	DirectBackend direct(ctx);
	Driver d;
	d.be = &direct;
	d.problem = problem;
	d.n = n;
	d.pts = h_points;
	d.S = *settings;
	const int r = d.run(model_out, inlier_mask);
	if (stats)
		*stats = d.st;
	if (r < 0)
	{
		g_driver_error = d.error;
		fprintf(stderr, "gcr_find_model: %s\n", d.error.c_str());
	}
	return r;
}

// findEssentialMatrix_ (py/gcransac_python.cpp:2519-2855) up to, not including, its LM bundle-adjustment polish:
// grid on the pixel coordinates, K^-1 normalisation (utils.h:65-110), threshold / mean focal length (:2624-2626, 2737).
extern "C" int gcr_find_essential(gcr_ctx *ctx, const double *h_points, int n, const double *K1, const double *K2,
                                  const gcr_run_settings *settings, double *model_out, uint8_t *inlier_mask,
                                  gcr_run_stats *stats)
{
	if (!ctx || !h_points || !K1 || !K2 || !settings || !model_out || !inlier_mask || n <= 0)
		return -2;
	auto inverse3 = [](const double *M, double *R) {
		// Eigen's 3 x 3 inverse: cofactors / determinant
		const double c00 = M[4] * M[8] - M[5] * M[7], c01 = M[5] * M[6] - M[3] * M[8], c02 = M[3] * M[7] - M[4] * M[6];
		const double det = M[0] * c00 + M[1] * c01 + M[2] * c02;
		const double inv = 1.0 / det;
		R[0] = c00 * inv; R[1] = (M[2] * M[7] - M[1] * M[8]) * inv; R[2] = (M[1] * M[5] - M[2] * M[4]) * inv;
		R[3] = c01 * inv; R[4] = (M[0] * M[8] - M[2] * M[6]) * inv; R[5] = (M[2] * M[3] - M[0] * M[5]) * inv;
		R[6] = c02 * inv; R[7] = (M[1] * M[6] - M[0] * M[7]) * inv; R[8] = (M[0] * M[4] - M[1] * M[3]) * inv;
	};
	double K1i[9], K2i[9];
	inverse3(K1, K1i);
	inverse3(K2, K2i);
	std::vector<double> normalized((size_t)n * 4);
	for (int r = 0; r < n; ++r)
	{
		const double x0 = h_points[4 * (size_t)r], y0 = h_points[4 * (size_t)r + 1], x1 = h_points[4 * (size_t)r + 2],
		             y1 = h_points[4 * (size_t)r + 3];
		normalized[4 * (size_t)r + 0] = K1i[0] * x0 + K1i[1] * y0 + K1i[2] * 1.0;
		normalized[4 * (size_t)r + 1] = K1i[3] * x0 + K1i[4] * y0 + K1i[5] * 1.0;
		normalized[4 * (size_t)r + 2] = K2i[0] * x1 + K2i[1] * y1 + K2i[2] * 1.0;
		normalized[4 * (size_t)r + 3] = K2i[3] * x1 + K2i[4] * y1 + K2i[5] * 1.0;
	}
	const double threshold_normalizer = 0.25 * (K1[0] + K1[4] + K2[0] + K2[4]);
	DirectBackend direct(ctx);
	Driver d;
	d.be = &direct;
	d.problem = GCR_E;
	d.n = n;
	d.pts = normalized.data();
	d.grid_pts = h_points;
	d.S = *settings;
	d.S.threshold = settings->threshold / threshold_normalizer;
	const int r = d.run(model_out, inlier_mask);
	if (stats)
		*stats = d.st;
	if (r < 0)
	{
		g_driver_error = d.error;
		fprintf(stderr, "gcr_find_essential: %s\n", d.error.c_str());
	}
	return r;
}
