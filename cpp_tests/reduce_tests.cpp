// reduce_tests.cpp — the reference's Reduce cases (test/reduce_tests.cpp:14-209) against device pointers:
// same case names, inputs, known answers and std::accumulate oracle; plus the operator/type combinations the
// reference never exercises (SURVEY.md §4 "gaps").
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

#include "glu/Reduce.hpp"
#include "harness.hpp"
#include "reference_vectors.hpp"
#include "util/Random.hpp"
#include "util/StopWatch.hpp"

using namespace glu;
namespace rv = reference_vectors;

namespace
{
    void check_sum_against_accumulate(size_t num_elements)
    {
        Random random(1);
        const std::vector<uint32_t> data = random.sample_int_vector<uint32_t>(num_elements, 0, 100);
        const uint32_t sum = std::accumulate(data.begin(), data.end(), uint32_t(0));
        DeviceBuffer buffer(data);
        Reduce reduce(DataType_Uint, ReduceOperator_Sum);
        reduce(buffer.handle(), num_elements);
        const uint32_t calc_sum = buffer.get_data<uint32_t>()[0];
        if (calc_sum != sum)
            std::printf("  N=%zu: got %u expected %u\n", num_elements, calc_sum, sum);
        CHECK(calc_sum == sum);
    }
} // namespace

TEST_CASE("Reduce-simple-uint", "")
{
    const std::vector<uint32_t>& data = rv::k_simple_uint;
    struct Section
    {
        ReduceOperator op;
        size_t count;
        uint32_t expected;
    };
    for (const Section& s : {Section{ReduceOperator_Sum, data.size(), 4951u}, Section{ReduceOperator_Mul, 5, 319200u},
                             Section{ReduceOperator_Min, data.size(), 1u}, Section{ReduceOperator_Max, data.size(), 99u}})
    {
        DeviceBuffer buffer(data); // a fresh buffer per section, as Catch2 SECTIONs re-run the set-up
        Reduce reduce(DataType_Uint, s.op);
        reduce(buffer.handle(), s.count);
        CHECK(buffer.get_data<uint32_t>()[0] == s.expected);
    }
}

namespace
{
    /// Sum-reduces `data` on the device as `data_type` and returns element 0.
    template<typename T> T device_sum(DataType data_type, const std::vector<T>& data)
    {
        Reduce reduce(data_type, ReduceOperator_Sum);
        DeviceBuffer buffer(data);
        reduce(buffer.handle(), data.size());
        return buffer.get_data<T>()[0];
    }
} // namespace

TEST_CASE("Reduce-all", "")
{
    CHECK(device_sum(DataType_Uint, rv::k_all_uint) == 505);
    CHECK_WITHIN_ABS(device_sum(DataType_Float, rv::k_all_float), 167.9f, 0.1f);
    CHECK_WITHIN_ABS(device_sum(DataType_Double, rv::k_all_double), 155.6, 0.1);
    const rv::Vec2 v2 = device_sum(DataType_Vec2, rv::k_all_vec2);
    CHECK_WITHIN_ABS(v2.x, 66.29f, 0.1f);
    CHECK_WITHIN_ABS(v2.y, -23.75f, 0.1f);
    const rv::Vec4 v4 = device_sum(DataType_Vec4, rv::k_all_vec4);
    CHECK_WITHIN_ABS(v4.x, -135.24f, 0.1f);
    CHECK_WITHIN_ABS(v4.y, 192.97f, 0.1f);
    CHECK_WITHIN_ABS(v4.z, 69.49f, 0.1f);
    CHECK_WITHIN_ABS(v4.w, 208.59f, 0.1f);
    const rv::IVec2 i2 = device_sum(DataType_IVec2, rv::k_all_ivec2);
    CHECK(i2.x == -226);
    CHECK(i2.y == -53);
    const rv::IVec4 i4 = device_sum(DataType_IVec4, rv::k_all_ivec4);
    CHECK(i4.x == -90);
    CHECK(i4.y == -2);
    CHECK(i4.z == -49);
    CHECK(i4.w == 58);
}

TEST_CASE("Reduce-subgroup-fitting-size", "")
{
    for (size_t n : {32, 64, 128, 256, 512, 1024, 2048, 4096, 8192, 16384, 32768, 65536, 131072})
        check_sum_against_accumulate(n);
}

TEST_CASE("Reduce-subgroup-non-fitting-size", "")
{
    for (size_t n : {1, 31, 93, 201, 693, 2087, 7358, 88289, 345897, 6094798, 5238082, 10043898})
        check_sum_against_accumulate(n);
}

// Not in the reference: every operator on full-range values (wrap-around mod 2^32) and on signed ints.
TEST_CASE("Reduce-operators-wraparound", "")
{
    std::mt19937 engine(7);
    std::vector<uint32_t> data(1000003);
    for (uint32_t& x : data)
        x = engine();
    for (int op = ReduceOperator_Sum; op <= ReduceOperator_Max; op++)
    {
        uint32_t expected = data[0];
        for (size_t i = 1; i < data.size(); i++)
            expected = op == ReduceOperator_Sum   ? expected + data[i]
                       : op == ReduceOperator_Mul ? expected * data[i]
                       : op == ReduceOperator_Min ? std::min(expected, data[i])
                                                  : std::max(expected, data[i]);
        DeviceBuffer buffer(data);
        Reduce reduce(DataType_Uint, ReduceOperator(op));
        reduce(buffer.handle(), data.size());
        CHECK(buffer.get_data<uint32_t>()[0] == expected);

        std::vector<int32_t> sdata(data.begin(), data.end());
        int32_t sexpected = sdata[0];
        for (size_t i = 1; i < sdata.size(); i++)
            sexpected = op == ReduceOperator_Sum   ? int32_t(uint32_t(sexpected) + uint32_t(sdata[i]))
                        : op == ReduceOperator_Mul ? int32_t(uint32_t(sexpected) * uint32_t(sdata[i]))
                        : op == ReduceOperator_Min ? std::min(sexpected, sdata[i])
                                                   : std::max(sexpected, sdata[i]);
        DeviceBuffer sbuffer(sdata);
        Reduce sreduce(DataType_Int, ReduceOperator(op));
        sreduce(sbuffer.handle(), sdata.size());
        CHECK(sbuffer.get_data<int32_t>()[0] == sexpected);
    }
}

TEST_CASE("Reduce-benchmark", "[.][benchmark]")
{
    for (size_t k_num_elements : {1024, 16384, 65536, 131072, 524288, 1048576, 16777216, 67108864, 134217728, 268435456})
    {
        std::vector<uint32_t> data(k_num_elements); // zero-filled, like the reference's benchmark
        DeviceBuffer buffer(data);
        Reduce reduce(DataType_Uint, ReduceOperator_Sum);
        reduce(buffer.handle(), k_num_elements); // warm-up (the reference times one cold run)
        buffer.write_data(data.data(), data.size() * sizeof(uint32_t));
        const uint64_t ns = measure_elapsed_time([&]() { reduce(buffer.handle(), k_num_elements); });
        std::printf("Reduce; Num elements: %zu, Elapsed: %s\n", k_num_elements, ns_to_human_string(ns).c_str());
    }
}

// The reference's error convention: a failed argument check prints and exits with status 1 (glu/errors.hpp:8-18).
// Hidden: run by name from tests/test_cpp_runner_gpu.py, which expects the process to die here.
TEST_CASE("Errors-null-buffer-exits", "[.]")
{
    Reduce reduce(DataType_Uint, ReduceOperator_Sum);
    reduce(nullptr, 10); // "Invalid buffer"
    CHECK(false);        // not reached
}
