// gcr_host_samplers.h -- the reference's SEQUENTIAL main samplers, drawn ahead on the host for the batched driver.
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
				max_prob = std::max(max_prob, probabilities[i]);
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
