// reduce_tests.cpp — the reference's Reduce cases (test/reduce_tests.cpp:14-209) against device pointers:
// same case names, inputs, known answers and std::accumulate oracle; plus the operator/type combinations the
// reference never exercises (SURVEY.md §4 "gaps").
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <vector>

#include "glu/Reduce.hpp"
#include "harness.hpp"
#include "util/Random.hpp"
#include "util/StopWatch.hpp"

using namespace glu;

namespace
{
    struct Vec2
    {
        float x, y;
    };
    struct Vec4
    {
        float x, y, z, w;
    };
    struct IVec2
    {
        int32_t x, y;
    };
    struct IVec4
    {
        int32_t x, y, z, w;
    };

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
    // the 100-value table of test/reduce_tests.cpp:16-20
    const uint32_t k_data[]{32, 35, 1,  3,  95, 10, 22, 24, 44, 37, 7,  80, 33, 54, 46, 23, 14, 84, 11, 67,
                            4,  58, 70, 61, 16, 36, 83, 9,  56, 99, 28, 98, 69, 21, 51, 34, 48, 91, 62, 19,
                            59, 79, 39, 92, 97, 78, 52, 40, 66, 47, 89, 88, 74, 49, 31, 20, 45, 13, 26, 72,
                            43, 30, 65, 94, 63, 8,  60, 15, 93, 86, 41, 75, 12, 73, 55, 90, 64, 96, 53, 1,
                            57, 71, 50, 42, 29, 2,  77, 25, 82, 18, 81, 85, 27, 5,  6,  68, 17, 38, 87, 76};
    const size_t k_data_length = sizeof(k_data) / sizeof(k_data[0]);
    struct Section
    {
        ReduceOperator op;
        size_t count;
        uint32_t expected;
    };
    for (const Section& s : {Section{ReduceOperator_Sum, k_data_length, 4951u}, Section{ReduceOperator_Mul, 5, 319200u},
                             Section{ReduceOperator_Min, k_data_length, 1u}, Section{ReduceOperator_Max, k_data_length, 99u}})
    {
        DeviceBuffer buffer(k_data, sizeof(k_data)); // a fresh buffer per section, as Catch2 SECTIONs re-run the setup
        Reduce reduce(DataType_Uint, s.op);
        reduce(buffer.handle(), s.count);
        CHECK(buffer.get_data<uint32_t>()[0] == s.expected);
    }
}

TEST_CASE("Reduce-all", "")
{
    { // uint
        const std::vector<uint32_t> k_data{1, 11, 80, 73, 48, 40, 89, 36, 70, 57};
        Reduce reduce(DataType_Uint, ReduceOperator_Sum);
        DeviceBuffer buffer(k_data);
        reduce(buffer.handle(), k_data.size());
        CHECK(buffer.get_data<uint32_t>()[0] == 505);
    }
    { // float
        const std::vector<float> k_data{42.138f, 18.228f, -19.127f, 86.564f,  11.904f,
                                        48.538f, 30.606f, 11.338f,  -32.699f, -29.587f};
        Reduce reduce(DataType_Float, ReduceOperator_Sum);
        DeviceBuffer buffer(k_data);
        reduce(buffer.handle(), k_data.size());
        CHECK_WITHIN_ABS(buffer.get_data<float>()[0], 167.9f, 0.1f);
    }
    { // double
        const std::vector<double> k_data{-6.20, -56.02, 49.42, 52.38, -23.81, -29.72, 95.46, 77.37, -85.00, 81.74};
        Reduce reduce(DataType_Double, ReduceOperator_Sum);
        DeviceBuffer buffer(k_data);
        reduce(buffer.handle(), k_data.size());
        CHECK_WITHIN_ABS(buffer.get_data<double>()[0], 155.6, 0.1);
    }
    { // vec2
        const std::vector<Vec2> k_data{{-77.08f, 19.54f}, {98.89f, -16.09f},  {10.53f, 91.17f}, {43.06f, -94.18f},
                                       {-19.18f, 0.86f},  {-49.99f, -92.53f}, {-4.68f, 42.34f}, {2.79f, -4.26f},
                                       {-17.49f, 43.99f}, {79.45f, -14.58f}};
        Reduce reduce(DataType_Vec2, ReduceOperator_Sum);
        DeviceBuffer buffer(k_data);
        reduce(buffer.handle(), k_data.size());
        const Vec2 sum = buffer.get_data<Vec2>()[0];
        CHECK_WITHIN_ABS(sum.x, 66.29f, 0.1f);
        CHECK_WITHIN_ABS(sum.y, -23.75f, 0.1f);
    }
    { // vec4
        const std::vector<Vec4> k_data{{-17.04f, 1.79f, 82.67f, 39.72f},    {52.66f, 24.75f, -19.05f, 91.92f},
                                       {19.15f, 44.93f, -52.13f, 18.85f},   {-84.25f, 69.53f, -11.43f, 33.17f},
                                       {19.46f, -14.30f, -15.20f, -63.83f}, {-20.51f, -56.75f, -2.70f, 82.66f},
                                       {3.86f, 55.48f, -12.37f, -11.02f},   {-30.62f, -67.54f, -29.89f, -77.30f},
                                       {-21.55f, 50.46f, 39.34f, 81.08f},   {-56.40f, 84.61f, 90.26f, 13.35f}};
        Reduce reduce(DataType_Vec4, ReduceOperator_Sum);
        DeviceBuffer buffer(k_data);
        reduce(buffer.handle(), k_data.size());
        const Vec4 sum = buffer.get_data<Vec4>()[0];
        CHECK_WITHIN_ABS(sum.x, -135.24f, 0.1f);
        CHECK_WITHIN_ABS(sum.y, 192.97f, 0.1f);
        CHECK_WITHIN_ABS(sum.z, 69.49f, 0.1f);
        CHECK_WITHIN_ABS(sum.w, 208.59f, 0.1f);
    }
    { // ivec2
        const std::vector<IVec2> k_data{{-38, -88}, {57, -34}, {61, 60},  {-90, 73}, {-23, -17},
                                        {34, -79},  {-80, 53}, {24, -23}, {-88, 69}, {-83, -67}};
        Reduce reduce(DataType_IVec2, ReduceOperator_Sum);
        DeviceBuffer buffer(k_data);
        reduce(buffer.handle(), k_data.size());
        const IVec2 sum = buffer.get_data<IVec2>()[0];
        CHECK(sum.x == -226);
        CHECK(sum.y == -53);
    }
    { // ivec4
        const std::vector<IVec4> k_data{{-95, 99, -30, 2},   {-69, 33, 78, 20},  {33, -43, -38, -26}, {69, -67, -17, -57},
                                        {18, -23, -2, -53},  {88, -96, 40, -48}, {-93, -47, -91, 59}, {-89, 82, 10, 94},
                                        {-15, 7, 41, 14},    {63, 53, -40, 53}};
        Reduce reduce(DataType_IVec4, ReduceOperator_Sum);
        DeviceBuffer buffer(k_data);
        reduce(buffer.handle(), k_data.size());
        const IVec4 sum = buffer.get_data<IVec4>()[0];
        CHECK(sum.x == -90);
        CHECK(sum.y == -2);
        CHECK(sum.z == -49);
        CHECK(sum.w == 58);
    }
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
