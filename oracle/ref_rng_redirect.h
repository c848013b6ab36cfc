// oracle/ref_rng_redirect.h -- TEST INFRASTRUCTURE ONLY.  Include BEFORE any header of the reference.
//
// The reference's samplers draw through utils::UniformRandomGenerator<size_t> (uniform_random_generator.h): a std::mt19937
// seeded from std::random_device.  Where the reference's own sampler code has to run verbatim on a reproducible stream, the
// member functions that consume random numbers are explicitly specialised HERE, before anything instantiates them: same
// contracts (values uniform in the range of the last resetGenerator, draw-and-reject for distinct sets, the `to_skip_` value
// never returned), the words taken from the counter-based stream gcr_ref_words, which the caller points at (seed, index,
// stream) with begin().  Everything else in uniform_random_generator.h and every sampler header is the reference's text.
#pragma once
#include <cstddef>
#include <cstdint>
extern "C" void gcro_philox4x32(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
namespace gcr_ref_words
{
	static uint32_t ctr[4], key[2], buf[4];
	static int pos = 4;
	inline void begin(uint64_t seed, uint64_t index, uint32_t stream)
	{
		ctr[0] = (uint32_t)index; ctr[1] = (uint32_t)(index >> 32); ctr[2] = 0; ctr[3] = stream;
		key[0] = (uint32_t)seed; key[1] = (uint32_t)(seed >> 32);
		pos = 4;
	}
	inline uint32_t next()
	{
		if (pos == 4)
		{
			gcro_philox4x32(ctr, key, buf);
			++ctr[2];
			pos = 0;
		}
		return buf[pos++];
	}
	// unbiased integer in [lo, hi] (Lemire's multiply-shift with rejection; the product's GcrDrawStream::bounded)
	inline size_t in_range(size_t lo, size_t hi)
	{
		const uint32_t range = (uint32_t)(hi - lo + 1);
		for (;;)
		{
			const uint64_t p = (uint64_t)next() * range;
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
}
#ifndef OLGA_INLINE
#define OLGA_INLINE inline
#define GCR_REF_DEFINED_OLGA_INLINE
#endif
#include "uniform_random_generator.h"
namespace gcransac
{
	namespace utils
	{
		template <> inline int UniformRandomGenerator<size_t>::getRandomNumber() { return (int)gcr_ref_words::in_range(generate.a(), generate.b()); }
		// uniform_random_generator.h:75-88
		template <> inline void UniformRandomGenerator<size_t>::generateUniqueRandomSet(size_t *sample_, size_t sample_size_)
		{
			for (size_t i = 0; i < sample_size_; i++)
			{
				sample_[i] = gcr_ref_words::in_range(generate.a(), generate.b());
				for (int j = (int)i - 1; j >= 0; j--)
					if (sample_[i] == sample_[j]) { i--; break; }
			}
		}
		// :90-103
		template <> inline void UniformRandomGenerator<size_t>::generateUniqueRandomSet(size_t *sample_, size_t sample_size_, size_t max_)
		{
			resetGenerator(0, max_);
			for (size_t i = 0; i < sample_size_; i++)
			{
				sample_[i] = gcr_ref_words::in_range(generate.a(), generate.b());
				for (int j = (int)i - 1; j >= 0; j--)
					if (sample_[i] == sample_[j]) { i--; break; }
			}
		}
		// :105-124
		template <> inline void UniformRandomGenerator<size_t>::generateUniqueRandomSet(size_t *sample_, size_t sample_size_, size_t max_, size_t to_skip_)
		{
			resetGenerator(0, max_);
			for (size_t i = 0; i < sample_size_; i++)
			{
				sample_[i] = gcr_ref_words::in_range(generate.a(), generate.b());
				if (sample_[i] == to_skip_) { i--; continue; }
				for (int j = (int)i - 1; j >= 0; j--)
					if (sample_[i] == sample_[j]) { i--; break; }
			}
		}
	}
}
#ifdef GCR_REF_DEFINED_OLGA_INLINE
#undef OLGA_INLINE
#undef GCR_REF_DEFINED_OLGA_INLINE
#endif
