// util/Random.hpp — the input generator of the reference's test-suite (test/util/Random.hpp:12-38).
// It defines the golden inputs, so its observable behaviour is reproduced exactly:
//   engine  = std::minstd_rand seeded with `seed` (default-seeded when seed == 0);
//   sample  = engine() % (max - min) + min           -> half-open [min, max), modulo bias included;
// "full range" keys sample_int_vector<uint32_t>(n, 0, UINT32_MAX) are therefore 31-bit.
#pragma once

#include <cstdint>
#include <random>
#include <vector>

#include "glu/errors.hpp"

namespace glu
{
    class Random
    {
    private:
        std::minstd_rand m_engine;

    public:
        explicit Random(uint64_t seed = 0) : m_engine(seed != 0 ? std::minstd_rand(seed) : std::minstd_rand()) {}

        template<typename IntegerT> IntegerT sample_int(IntegerT min, IntegerT max)
        {
            GLU_CHECK_ARGUMENT(min < max, "Min must be strictly lower than Max");
            return IntegerT(m_engine() % (max - min) + min);
        }

        template<typename IntegerT> std::vector<IntegerT> sample_int_vector(size_t num_elements, IntegerT min, IntegerT max)
        {
            std::vector<IntegerT> out;
            out.reserve(num_elements);
            while (out.size() < num_elements)
                out.push_back(sample_int(min, max));
            return out;
        }
    };
} // namespace glu
