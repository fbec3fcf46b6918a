// make_golden.cpp — generates tests/golden/reference_golden.json.
//
// Compiled AGAINST THE REFERENCE'S OWN HEADERS where they lie (-I/root/reference -I/root/reference/test):
// the inputs come from the reference's generator glu::Random (test/util/Random.hpp) and the expected
// values from the very std:: algorithms the reference's tests use as their oracle
// (test/reduce_tests.cpp:155,174; test/blelloch_scan_tests.cpp:44,75; test/radix_sort_tests.cpp:20-51).
// The GL part of the reference cannot run here, so these are the strongest vectors available.
// Run ./make_golden.sh in this container (needs /root/reference); the JSON is committed.
#include <algorithm>
#include <cinttypes>
#include <cstdint>
#include <cstdio>
#include <numeric>
#include <vector>

#include "glu/errors.hpp"
#include "util/Random.hpp"

using glu::Random;
typedef uint32_t GLuint;

static uint64_t fnv1a(const std::vector<uint32_t>& v)
{
    uint64_t h = 1469598103934665603ull;
    for (uint32_t x : v)
        for (int b = 0; b < 4; b++)
        {
            h ^= (x >> (8 * b)) & 0xff;
            h *= 1099511628211ull;
        }
    return h;
}

int main()
{
    printf("{\n");
    { // generator
        Random r(1);
        auto v = r.sample_int_vector<GLuint>(16, 0, UINT32_MAX);
        printf(" \"random_seed1_full_first16\": [");
        for (size_t i = 0; i < v.size(); i++) printf("%s%u", i ? ", " : "", v[i]);
        printf("],\n");
        Random r0(0);
        auto v0 = r0.sample_int_vector<GLuint>(4, 0, UINT32_MAX);
        printf(" \"random_seed0_full_first4\": [%u, %u, %u, %u],\n", v0[0], v0[1], v0[2], v0[3]);
        Random r2(123);
        auto v2 = r2.sample_int_vector<GLuint>(8, 0, 100);
        printf(" \"random_seed123_0_100_first8\": [");
        for (size_t i = 0; i < v2.size(); i++) printf("%s%u", i ? ", " : "", v2[i]);
        printf("],\n");
    }
    { // Reduce-subgroup-fitting-size / non-fitting-size (test/reduce_tests.cpp:147-183)
        const size_t sizes[] = {32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072,
                                1, 31, 93, 201, 693, 2087, 7358, 88289, 345897, 6094798, 5238082, 10043898};
        printf(" \"reduce_seed1_0_100\": {");
        bool first = true;
        for (size_t n : sizes)
        {
            Random r(1);
            auto data = r.sample_int_vector<GLuint>(n, 0, 100);
            GLuint sum = std::accumulate(data.begin(), data.end(), GLuint(0));
            GLuint mn = *std::min_element(data.begin(), data.end());
            GLuint mx = *std::max_element(data.begin(), data.end());
            printf("%s\"%zu\": {\"sum\": %u, \"min\": %u, \"max\": %u}", first ? "" : ", ", n, sum, mn, mx);
            first = false;
        }
        printf("},\n");
    }
    { // BlellochScan-multiple-sizes (test/blelloch_scan_tests.cpp:28-46)
        printf(" \"scan_seed123_0_100\": {");
        bool first = true;
        for (size_t n = 1024; n <= 1048576; n <<= 1)
        {
            Random r(123);
            auto data = r.sample_int_vector<GLuint>(n, 0, 100);
            std::vector<GLuint> expected(n);
            std::exclusive_scan(data.begin(), data.end(), expected.begin(), 0);
            printf("%s\"%zu\": {\"last\": %u, \"fnv1a\": \"%016" PRIx64 "\"}", first ? "" : ", ", n, expected[n - 1],
                   fnv1a(expected));
            first = false;
        }
        printf("},\n");
    }
    { // BlellochScan-multiple-partitions (test/blelloch_scan_tests.cpp:48-82)
        printf(" \"scan_partitions_seed123_1024\": {");
        bool first = true;
        for (size_t parts : {1, 32, 100, 1000})
        {
            Random r(123);
            auto data = r.sample_int_vector<GLuint>(1024 * parts, 0, 100);
            std::vector<GLuint> expected(data.size());
            for (size_t p = 0; p < parts; p++)
                std::exclusive_scan(data.begin() + p * 1024, data.begin() + (p + 1) * 1024, expected.begin() + p * 1024, 0);
            printf("%s\"%zu\": {\"fnv1a\": \"%016" PRIx64 "\"}", first ? "" : ", ", parts, fnv1a(expected));
            first = false;
        }
        printf("},\n");
    }
    { // RadixSort-* (test/radix_sort_tests.cpp:54-158): keys sorted; vals = index, stable
        printf(" \"sort_seed1\": {");
        bool first = true;
        struct C { size_t n; uint32_t hi; };
        const C cases[] = {{10, UINT32_MAX}, {128, UINT32_MAX}, {256, UINT32_MAX}, {512, UINT32_MAX}, {1024, UINT32_MAX},
                           {2048, 10}, {10993, UINT32_MAX}, {14978, UINT32_MAX}, {16243, UINT32_MAX}, {18985, UINT32_MAX},
                           {23857, UINT32_MAX}, {27865, UINT32_MAX}, {33363, UINT32_MAX}, {41298, UINT32_MAX},
                           {45821, UINT32_MAX}, {47487, UINT32_MAX}, {1048576, UINT32_MAX}};
        for (const C& c : cases)
        {
            Random r(1);
            auto keys = r.sample_int_vector<GLuint>(c.n, 0, c.hi);
            std::vector<std::pair<GLuint, GLuint>> pairs(c.n);
            for (size_t i = 0; i < c.n; i++) pairs[i] = {keys[i], GLuint(i)};
            std::stable_sort(pairs.begin(), pairs.end(), [](auto& a, auto& b) { return a.first < b.first; });
            std::vector<uint32_t> sk(c.n), sv(c.n);
            for (size_t i = 0; i < c.n; i++) { sk[i] = pairs[i].first; sv[i] = pairs[i].second; }
            printf("%s\"%zu\": {\"hi\": %u, \"min\": %u, \"max\": %u, \"keys_fnv1a\": \"%016" PRIx64 "\", \"vals_fnv1a\": \"%016" PRIx64 "\"}",
                   first ? "" : ", ", c.n, c.hi, sk.front(), sk.back(), fnv1a(sk), fnv1a(sv));
            first = false;
        }
        printf("},\n");
        Random r(1);
        auto keys = r.sample_int_vector<GLuint>(2048, 0, 10);
        size_t counts[10] = {0};
        for (auto k : keys) counts[k]++;
        printf(" \"sort_2048_digit_counts\": [");
        for (int i = 0; i < 10; i++) printf("%s%zu", i ? ", " : "", counts[i]);
        printf("]\n");
    }
    printf("}\n");
    return 0;
}
