// gcr_host_samplers.h -- the reference's SEQUENTIAL main samplers, drawn ahead on the host for the batched driver:
// AdaptiveReordering (sampler 4) and ProgressiveNapsac (sampler 2, see the class).
//
// AdaptiveReorderingSampler (samplers/adaptive_reordering_sampler.h, `sampler=4`, the AR-Sampler the reference's F / E
// tests use) keeps a priority queue of (inlier probability, point): sample() pops the m best points (:205-215),
// update() lowers their probability by one more appearance, adds a little noise and pushes them back (:135-160).
// GCRANSAC::run calls update() after EVERY sample() -- rejected ones included (GCRANSAC.h:304, 317, 331, 474) -- and
// update() reads neither the inlier ratio nor the iteration number it is handed, so sample k is a function of the
// probabilities and the noise alone, not of any verdict on earlier samples.  The recurrence over k is sequential
// (a heap), O(m log n) per sample: it runs here, on the host, any number of samples ahead of the GPU batch that
// solves and scores them.  The reference's noise is C rand() (process-global, shared with SPRT); here update k takes
// its m values from stream 4 of the counter-based generator, word >> 1 in [0, RAND_MAX] like rand().
//
// One thing does feed back: when max_unsuccessful_model_generations samples in a row give no model, run() falls through
// to the end-of-iteration update() with the last sample once more (GCRANSAC.h:297-339 -> :474): that sample's points
// are pushed twice.  The driver reports it (repeat_update) and the sequence is rebuilt from there.
#ifndef GCR_HOST_SAMPLERS_H_INCLUDED
#define GCR_HOST_SAMPLERS_H_INCLUDED

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <queue>
#include <set>
#include <utility>
#include <vector>

namespace gcr_host
{
	// Philox4x32-10 (Salmon et al., SC'11): the same block function as the device generator (csrc/gcr_internal.cuh)
	inline void philox4x32(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
	{
		for (int round = 0; round < 10; ++round)
		{
			const uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
			const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
			c0 = n0; c1 = n1; c2 = n2; c3 = n3;
			k0 += 0x9E3779B9u;
			k1 += 0xBB67AE85u;
		}
		out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
	}

	// the words of (seed, index, stream), in order: counter = (index_lo, index_hi, block, stream), key = seed
	struct WordStream
	{
		uint32_t c0, c1, block, stream, k0, k1, buf[4];
		int pos;
		WordStream(uint64_t seed, uint64_t index, uint32_t stream_)
		    : c0((uint32_t)index), c1((uint32_t)(index >> 32)), block(0), stream(stream_), k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), pos(4)
		{
		}
		uint32_t next()
		{
			if (pos == 4)
			{
				philox4x32(c0, c1, block, stream, k0, k1, buf);
				++block;
				pos = 0;
			}
			return buf[pos++];
		}
	};

	// unbiased integer in [lo, hi]: Lemire's multiply-shift with rejection, as GcrDrawStream::bounded on the device
	inline size_t in_range(WordStream &w, size_t lo, size_t hi)
	{
		const uint32_t range = (uint32_t)(hi - lo + 1);
		for (;;)
		{
			const uint64_t p = (uint64_t)w.next() * range;
			const uint32_t l = (uint32_t)p;
			if (l < range)
			{
				const uint32_t thresh = (uint32_t)(0u - range) % range;
				if (l < thresh)
					continue;
			}
			return lo + (size_t)(p >> 32);
		}
	}

	// ProsacSampler (samplers/prosac_sampler.h) as a state machine: the subset size, the sample counter and the range of
	// its generator (which the reference only resets when the subset grows, :65-88, 178-203) -- the two PROSAC samplers
	// inside Progressive NAPSAC are built with other arguments than the stand-alone one (sample size 1; convergence
	// after n samples), and set to arbitrary sample numbers.
	struct ProsacState
	{
		size_t m = 0, n = 0, convergence = 0, kth = 1, subset_size = 0, range_lo = 0, range_hi = 0;
		std::vector<size_t> growth;
		void initialize(size_t sample_size, size_t point_number, size_t convergence_iterations)
		{
			m = sample_size; n = point_number; convergence = convergence_iterations; kth = 1;
			growth.assign(n, 0); // :129-160
			double T_n = (double)convergence;
			for (size_t i = 0; i < m; i++)
				T_n *= static_cast<double>(m - i) / (n - i);
			size_t T_n_prime = 1;
			for (size_t i = 0; i < n; ++i)
			{
				if (i + 1 <= m) { growth[i] = T_n_prime; continue; }
				const double Tn_plus1 = static_cast<double>(i + 1) * T_n / (i + 1 - m);
				growth[i] = T_n_prime + (unsigned int)std::ceil(Tn_plus1 - T_n);
				T_n = Tn_plus1;
				T_n_prime = growth[i];
			}
			subset_size = m;
			range_lo = 0; range_hi = m - 1; // :171-172
		}
		void grow()
		{
			++subset_size;
			if (subset_size > n)
				subset_size = n;
			range_lo = 0; range_hi = subset_size - 2;
		}
		void increment() // :65-88
		{
			++kth;
			if (kth > convergence) { range_lo = 0; range_hi = n - 1; }
			else if (kth > growth[subset_size - 1])
				grow();
		}
		void set_sample_number(size_t k) // :178-203
		{
			kth = k;
			if (kth > convergence) { range_lo = 0; range_hi = n - 1; }
			else
				while (kth > growth[subset_size - 1] && subset_size != n)
					grow();
		}
		void unique_set(WordStream &w, size_t *out, size_t count) const // uniform_random_generator.h:75-88
		{
			for (size_t i = 0; i < count; i++)
			{
				out[i] = in_range(w, range_lo, range_hi);
				for (int j = (int)i - 1; j >= 0; j--)
					if (out[i] == out[j]) { i--; break; }
			}
		}
		void sample(WordStream &w, size_t *out) // :207-236
		{
			if (kth > convergence) { unique_set(w, out, m); return; }
			unique_set(w, out, m - 1);
			out[m - 1] = subset_size - 1;
			increment();
		}
	};

	// one layer of Progressive NAPSAC's grids: GridNeighborhoodGraph<4> as gcr_grid_build returns it -- the points sorted
	// by cell (ascending index inside a cell = the reference's insertion order) and each point's cell
	struct GridLayer
	{
		std::vector<uint32_t> cell_of_point, sorted_idx, cell_offsets;
		size_t cell_size(uint32_t point) const { const uint32_t c = cell_of_point[point]; return cell_offsets[c + 1] - cell_offsets[c]; }
		const uint32_t *cell_points(uint32_t point) const { return &sorted_idx[cell_offsets[cell_of_point[point]]]; }
	};

	// GridNeighborhoodGraph<4>::initialize (neighborhood/grid_neighborhood_graph.h:229-292) on the host, for point sets
	// too small to be worth a GPU round trip: key = sum_d floor(x_d / cell_d) * cells^d, cells in ascending key order,
	// points in ascending index order inside a cell (what gcr_grid_build returns for larger sets)
	inline void build_grid_cells(const double *points, size_t n, const double *cell_sizes, size_t cells, GridLayer &g);
	inline void build_grid_layer(const double *points, size_t n, const double *image_sizes, size_t cells, GridLayer &g)
	{
		double cell_sizes[4];
		for (int d = 0; d < 4; ++d)
			cell_sizes[d] = image_sizes[d] / cells; // progressive_napsac_sampler.h:171-173
		build_grid_cells(points, n, cell_sizes, cells, g);
	}
	inline void build_grid_cells(const double *points, size_t n, const double *cell_sizes, size_t cells, GridLayer &g)
	{
		std::vector<std::pair<uint64_t, uint32_t>> keyed(n);
		for (size_t i = 0; i < n; ++i)
		{
			uint64_t key = 0, offset = 1;
			for (int d = 0; d < 4; ++d)
			{
				const size_t idx = (size_t)std::floor(points[4 * i + d] / cell_sizes[d]);
				key += offset * idx;
				offset *= cells;
			}
			keyed[i] = std::make_pair(key, (uint32_t)i);
		}
		std::sort(keyed.begin(), keyed.end());
		g.cell_of_point.assign(n, 0);
		g.sorted_idx.assign(n, 0);
		g.cell_offsets.clear();
		for (size_t q = 0; q < n; ++q)
		{
			if (q == 0 || keyed[q].first != keyed[q - 1].first)
				g.cell_offsets.push_back((uint32_t)q);
			g.sorted_idx[q] = keyed[q].second;
			g.cell_of_point[keyed[q].second] = (uint32_t)g.cell_offsets.size() - 1;
		}
		g.cell_offsets.push_back((uint32_t)n);
	}

	// ProgressiveNapsacSampler<4> (samplers/progressive_napsac_sampler.h, `sampler=2`, the sampler of the reference's H
	// test).  Sample k takes its centre from a one-point PROSAC sampler (the k-th best point), the rest from the centre's
	// cell in the finest grid layer that holds enough points, the neighbourhood growing with the number of times each
	// point has been hit (:246-310); after 0.5 n samples it is PROSAC over all points (:236-242).  Its update() is empty
	// (:111-118): the sequence depends on the points, the layers and the random words only -- a sequential recurrence over
	// k (the hit counters), drawn ahead on the host like the AR-Sampler's.  Sample k (from 1) takes stream 5 of (seed, k).
	class ProgressiveNapsac
	{
		std::vector<GridLayer> layers;
		std::vector<size_t> layer_of_point, hits, subset_size_of_point, growth;
		ProsacState one_point, prosac;
		size_t n = 0, m = 0, max_iterations = 0, kth = 0;
		uint64_t seed = 0;
		std::vector<uint32_t> drawn;

		void generate_one()
		{
			++kth;
			WordStream words(seed, kth, 5u);
			size_t subset[16];
			auto emit = [&]() { for (size_t i = 0; i < m; ++i) drawn.push_back((uint32_t)subset[i]); };
			if (kth > max_iterations) // :236-242
			{
				prosac.set_sample_number(kth);
				prosac.sample(words, subset);
				emit();
				return;
			}
			size_t centre[1];
			one_point.sample(words, centre);
			const size_t initial_point = centre[0];
			const size_t hit_count = ++hits[initial_point];
			size_t &subset_size = subset_size_of_point[initial_point]; // :259-262
			while (hit_count > growth[subset_size - 1] && subset_size < n)
				subset_size = std::min(subset_size + 1, n);
			size_t &layer = layer_of_point[initial_point]; // :265-290
			bool is_last_layer = false;
			for (;;)
			{
				if (layer >= layers.size()) { is_last_layer = true; break; }
				if (layers[layer].cell_size((uint32_t)initial_point) < subset_size) { ++layer; continue; }
				break;
			}
			if (!is_last_layer) // :293-318
			{
				const uint32_t *neighbors = layers[layer].cell_points((uint32_t)initial_point);
				subset[m - 1] = initial_point;
				subset[m - 2] = neighbors[subset_size - 1];
				// uniform_random_generator.h:105-124: m - 2 distinct POSITIONS in [0, subset_size - 2], never the value
				// `initial_point` (the reference skips the centre's INDEX among neighbour positions; kept)
				for (size_t i = 0; i < m - 2; i++)
				{
					subset[i] = in_range(words, 0, subset_size - 2);
					if (subset[i] == initial_point) { i--; continue; }
					for (int j = (int)i - 1; j >= 0; j--)
						if (subset[i] == subset[j]) { i--; break; }
				}
				for (size_t i = 0; i < m - 2; ++i)
				{
					subset[i] = neighbors[subset[i]];
					++hits[subset[i]];
				}
				++hits[subset[m - 2]];
			}
			else // :321-329
			{
				prosac.set_sample_number(kth);
				prosac.sample(words, subset);
				subset[m - 1] = initial_point;
			}
			emit();
		}

	public:
		// grid_layers: finest first (the drivers use 16, 8, 4, 2 cells per axis); sampler_length 0.5 (py/gcransac_python.cpp:4343)
		bool initialize(std::vector<GridLayer> grid_layers, size_t point_number, int sample_size, double sampler_length, uint64_t seed_)
		{
			if (sample_size < 3 || sample_size > 16 || point_number < (size_t)sample_size)
				return false;
			layers = std::move(grid_layers);
			n = point_number; m = (size_t)sample_size; seed = seed_; kth = 0;
			drawn.clear();
			layer_of_point.assign(n, 0);
			hits.assign(n, 0);
			subset_size_of_point.assign(n, m);
			one_point.initialize(1, n, n);
			prosac.initialize(m, n, n);
			max_iterations = static_cast<size_t>(sampler_length * n);
			growth.assign(n, 0); // :179-199
			const size_t local_sample_size = m - 1;
			double T_n = (double)max_iterations;
			for (size_t i = 0; i < local_sample_size; ++i)
				T_n *= static_cast<double>(local_sample_size - i) / (n - i);
			unsigned int T_n_prime = 1;
			for (size_t i = 0; i < n; ++i)
			{
				if (i + 1 <= local_sample_size) { growth[i] = T_n_prime; continue; }
				const double Tn_plus1 = static_cast<double>(i + 1) * T_n / (i + 1 - local_sample_size);
				growth[i] = T_n_prime + static_cast<size_t>(std::ceil(Tn_plus1 - T_n));
				T_n = Tn_plus1;
				T_n_prime = (unsigned int)growth[i];
			}
			return true;
		}
		// samples [first, first + count) (0-based: sample g is the reference's call g + 1) -> out [count][m]
		void samples(uint64_t first, size_t count, uint32_t *out)
		{
			while (drawn.size() / m < first + count)
				generate_one();
			std::copy(drawn.begin() + first * m, drawn.begin() + (first + count) * m, out);
		}
	};

	class AdaptiveReordering
	{
		struct Point
		{
			double probability, a, b;
			uint64_t appearances;
		};
		std::vector<double> initial;
		std::vector<Point> points;
		std::priority_queue<std::pair<double, size_t>> queue; // max first; ties: the larger index (std::pair's order)
		std::vector<uint32_t> drawn;                          // samples generated so far, [k][m]
		std::set<uint64_t> repeated;                          // samples whose update ran twice (see above)
		int m = 0;
		uint64_t seed = 0;
		double variance = 0.9765, randomness = 0.01;

		void restart()
		{
			points.clear();
			queue = decltype(queue)();
			drawn.clear();
			for (size_t i = 0; i < initial.size(); ++i)
			{
				// adaptive_reordering_sampler.h:84-94
				double probability = initial[i];
				if (probability == 1.0)
					probability -= 1e-6;
				const double a = probability * probability * (1.0 - probability) / variance - probability;
				const double b = a * (1.0 - probability) / probability;
				points.push_back(Point{probability, a, b, 0});
				queue.emplace(probability, i);
			}
		}
		void update(uint64_t k, const uint32_t *subset, WordStream &noise)
		{
			// :135-160; randomness_rand_max * rand() - randomness / 2 with rand() = the next word >> 1
			const double randomness_2 = randomness / 2.0, randomness_rand_max = randomness / 2147483647.0;
			for (int i = 0; i < m; ++i)
			{
				Point &p = points[subset[i]];
				++p.appearances;
				double r = std::fabs(p.a / (p.a + p.b + p.appearances)) + randomness_rand_max * static_cast<double>(noise.next() >> 1) - randomness_2;
				r = std::max(0.0, std::min(0.999, r));
				p.probability = r;
				queue.emplace(r, subset[i]);
			}
			(void)k;
		}
		void generate_one()
		{
			const uint64_t k = drawn.size() / m;
			uint32_t subset[16];
			for (int i = 0; i < m; ++i) // :205-215
			{
				subset[i] = (uint32_t)queue.top().second;
				queue.pop();
			}
			drawn.insert(drawn.end(), subset, subset + m);
			WordStream noise(seed, k, 4u);
			update(k, subset, noise);
			if (repeated.count(k))
				update(k, subset, noise);
		}

	public:
		// probabilities: as the caller gave them; the drivers divide by the largest first (py/gcransac_python.cpp:4350-4354)
		bool initialize(const double *probabilities, size_t n, int sample_size, double estimator_variance, uint64_t seed_)
		{
			if (!probabilities || n == 0 || sample_size < 1 || sample_size > 16 || (size_t)sample_size > n)
				return false;
			double max_prob = 0;
			for (size_t i = 0; i < n; ++i)
			{
				if (!std::isfinite(probabilities[i]) || probabilities[i] < 0.0)
					return false; // (the reference does not check: NaN keys would break its heap's ordering)
				max_prob = std::max(max_prob, probabilities[i]);
			}
			if (!(max_prob > 0.0))
				return false;
			initial.resize(n);
			for (size_t i = 0; i < n; ++i)
				initial[i] = probabilities[i] / max_prob;
			m = sample_size;
			seed = seed_;
			variance = estimator_variance;
			repeated.clear();
			restart();
			return true;
		}
		// samples [first, first + count) -> out [count][m]
		void samples(uint64_t first, size_t count, uint32_t *out)
		{
			while (drawn.size() / m < first + count)
				generate_one();
			std::copy(drawn.begin() + first * m, drawn.begin() + (first + count) * m, out);
		}
		// sample k's update ran twice: everything after k changes
		void repeat_update(uint64_t k)
		{
			if (!repeated.insert(k).second)
				return;
			const uint64_t keep = std::min<uint64_t>(k, drawn.size() / m);
			restart();
			while (drawn.size() / m < keep)
				generate_one();
		}
	};
}
#endif
