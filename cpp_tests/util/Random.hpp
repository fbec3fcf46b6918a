// util/Random.hpp — input generator of the reference's test-suite, restated (test/util/Random.hpp:12-38).
//
// The seeded test inputs ARE the golden inputs, so the observable sequence must match the reference exactly:
//   * engine: std::minstd_rand (fully specified by the C++ standard), default-constructed for seed 0;
//   * a sample in [lo, hi) is  lo + engine() % (hi - lo)   — modulo bias and all.
// Since minstd_rand yields at most 2^31 - 2, "full range" uint32 keys never have bit 31 set;
// tests/golden/reference_golden.json pins the first outputs.
#pragma once

#include <cstddef>
#include <cstdint>
#include <random>
#include <vector>

#include "glu/errors.hpp"

namespace glu
{
    class Random
    {
    public:
        explicit Random(uint64_t seed = 0)
        {
            if (seed != 0)
                m_engine.seed(static_cast<std::minstd_rand::result_type>(seed));
        }

        /// One draw from the half-open range [lo, hi).
        template<typename IntegerT> IntegerT sample_int(IntegerT lo, IntegerT hi)
        {
            GLU_CHECK_ARGUMENT(lo < hi, "Min must be strictly lower than Max");
            const auto span = hi - lo;
            return static_cast<IntegerT>(lo + m_engine() % span);
        }

        /// `count` consecutive draws.
        template<typename IntegerT> std::vector<IntegerT> sample_int_vector(size_t count, IntegerT lo, IntegerT hi)
        {
            std::vector<IntegerT> draws(count);
            for (IntegerT& d : draws)
                d = sample_int(lo, hi);
            return draws;
        }

    private:
        std::minstd_rand m_engine;
    };
} // namespace glu
