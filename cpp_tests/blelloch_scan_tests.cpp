// blelloch_scan_tests.cpp — the reference's BlellochScan cases (test/blelloch_scan_tests.cpp:12-108) against
// device pointers, plus sizes that are not powers of two (which the reference rejects).
#include <cstdint>
#include <cstring>
#include <numeric>
#include <vector>

#include "glu/BlellochScan.hpp"
#include "harness.hpp"
#include "util/Random.hpp"
#include "util/StopWatch.hpp"

using namespace glu;

TEST_CASE("BlellochScan-simple", "[.]")
{
    const std::vector<uint32_t> data{1, 2, 3, 4, 5, 6, 7, 8};
    DeviceBuffer buffer(data);
    BlellochScan blelloch_scan(DataType_Uint);
    blelloch_scan(buffer.handle(), data.size());
    print_buffer<uint32_t>(buffer);
    CHECK(buffer.get_data<uint32_t>() == std::vector<uint32_t>({0, 1, 3, 6, 10, 15, 21, 28}));
}

TEST_CASE("BlellochScan-multiple-sizes", "")
{
    for (size_t k_num_elements : {1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072, 262144, 524288, 1048576})
    {
        Random random(123);
        const std::vector<uint32_t> data = random.sample_int_vector<uint32_t>(k_num_elements, 0, 100);
        DeviceBuffer buffer(data);
        BlellochScan blelloch_scan(DataType_Uint);
        blelloch_scan(buffer.handle(), data.size());
        std::vector<uint32_t> expected(k_num_elements);
        std::exclusive_scan(data.begin(), data.end(), expected.begin(), uint32_t(0));
        REQUIRE(buffer.get_data<uint32_t>() == expected);
    }
}

TEST_CASE("BlellochScan-multiple-partitions", "")
{
    const size_t k_num_elements = 1024;
    for (size_t k_num_partitions : {1, 32, 100, 1000})
    {
        Random random(123);
        const std::vector<uint32_t> data = random.sample_int_vector<uint32_t>(k_num_elements * k_num_partitions, 0, 100);
        DeviceBuffer buffer(data);
        BlellochScan blelloch_scan(DataType_Uint);
        blelloch_scan(buffer.handle(), k_num_elements, k_num_partitions);
        const std::vector<uint32_t> result = buffer.get_data<uint32_t>();
        std::vector<uint32_t> expected(k_num_elements);
        for (size_t partition = 0; partition < k_num_partitions; partition++)
        {
            const uint32_t* first = data.data() + partition * k_num_elements;
            std::exclusive_scan(first, first + k_num_elements, expected.begin(), uint32_t(0));
            REQUIRE(std::memcmp(expected.data(), result.data() + partition * k_num_elements,
                                k_num_elements * sizeof(uint32_t)) == 0);
        }
    }
}

// Not in the reference (it aborts on these, glu/BlellochScan.hpp:134): arbitrary counts, full-range values.
TEST_CASE("BlellochScan-non-power-of-2", "")
{
    std::mt19937 engine(99);
    for (size_t n : {1, 2, 3, 31, 33, 1000, 4097, 70001, 1000003, 5000011})
    {
        std::vector<uint32_t> data(n);
        for (uint32_t& x : data)
            x = engine();
        DeviceBuffer buffer(data);
        BlellochScan blelloch_scan(DataType_Uint);
        blelloch_scan(buffer.handle(), n);
        std::vector<uint32_t> expected(n);
        std::exclusive_scan(data.begin(), data.end(), expected.begin(), uint32_t(0));
        CHECK(buffer.get_data<uint32_t>() == expected);
    }
}

TEST_CASE("BlellochScan-benchmark", "[.][benchmark]")
{
    for (size_t k_num_elements : {1024, 16384, 65536, 131072, 524288, 1048576, 16777216, 67108864, 134217728, 268435456})
    {
        std::vector<uint32_t> data(k_num_elements);
        DeviceBuffer buffer(data);
        BlellochScan blelloch_scan(DataType_Uint);
        blelloch_scan(buffer.handle(), k_num_elements); // warm-up
        const uint64_t ns = measure_elapsed_time([&]() { blelloch_scan(buffer.handle(), k_num_elements); });
        std::printf("BlellochScan; Num elements: %zu, Elapsed: %s\n", k_num_elements, ns_to_human_string(ns).c_str());
    }
}
