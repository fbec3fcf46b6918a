// reference_vectors.hpp — the fixed inputs and known answers of the reference's Reduce tests
// (test/reduce_tests.cpp:16-20 and :58-143), kept as data so that the test code reads on its own.
#pragma once

#include <cstdint>
#include <vector>

namespace reference_vectors
{
    struct Vec2 { float x, y; };
    struct Vec4 { float x, y, z, w; };
    struct IVec2 { int32_t x, y; };
    struct IVec4 { int32_t x, y, z, w; };

    // Reduce-simple-uint: 100 values; sum 4951, product of the first five 319200, min 1, max 99
    inline const std::vector<uint32_t> k_simple_uint = {
        32, 35, 1, 3, 95, 10, 22, 24, 44, 37, 7, 80, 33, 54, 46, 23, 14, 84, 11, 67, 4, 58, 70, 61, 16,
        36, 83, 9, 56, 99, 28, 98, 69, 21, 51, 34, 48, 91, 62, 19, 59, 79, 39, 92, 97, 78, 52, 40, 66, 47,
        89, 88, 74, 49, 31, 20, 45, 13, 26, 72, 43, 30, 65, 94, 63, 8, 60, 15, 93, 86, 41, 75, 12, 73, 55,
        90, 64, 96, 53, 1, 57, 71, 50, 42, 29, 2, 77, 25, 82, 18, 81, 85, 27, 5, 6, 68, 17, 38, 87, 76};

    // Reduce-all: ten elements per type, Sum
    inline const std::vector<uint32_t> k_all_uint = {1, 11, 80, 73, 48, 40, 89, 36, 70, 57}; // 505
    inline const std::vector<float> k_all_float = {42.138f, 18.228f, -19.127f, 86.564f, 11.904f,
                                                   48.538f, 30.606f, 11.338f,  -32.699f, -29.587f}; // 167.9
    inline const std::vector<double> k_all_double = {-6.20, -56.02, 49.42, 52.38, -23.81,
                                                     -29.72, 95.46, 77.37, -85.00, 81.74}; // 155.6
    inline const std::vector<Vec2> k_all_vec2 = {{-77.08f, 19.54f}, {98.89f, -16.09f}, {10.53f, 91.17f}, {43.06f, -94.18f},
                                                 {-19.18f, 0.86f}, {-49.99f, -92.53f}, {-4.68f, 42.34f}, {2.79f, -4.26f},
                                                 {-17.49f, 43.99f}, {79.45f, -14.58f}}; // (66.29, -23.75)
    inline const std::vector<Vec4> k_all_vec4 = {
        {-17.04f, 1.79f, 82.67f, 39.72f},   {52.66f, 24.75f, -19.05f, 91.92f},  {19.15f, 44.93f, -52.13f, 18.85f},
        {-84.25f, 69.53f, -11.43f, 33.17f}, {19.46f, -14.30f, -15.20f, -63.83f}, {-20.51f, -56.75f, -2.70f, 82.66f},
        {3.86f, 55.48f, -12.37f, -11.02f},  {-30.62f, -67.54f, -29.89f, -77.30f}, {-21.55f, 50.46f, 39.34f, 81.08f},
        {-56.40f, 84.61f, 90.26f, 13.35f}}; // (-135.24, 192.97, 69.49, 208.59)
    inline const std::vector<IVec2> k_all_ivec2 = {{-38, -88}, {57, -34}, {61, 60}, {-90, 73}, {-23, -17},
                                                   {34, -79}, {-80, 53}, {24, -23}, {-88, 69}, {-83, -67}}; // (-226, -53)
    inline const std::vector<IVec4> k_all_ivec4 = {{-95, 99, -30, 2}, {-69, 33, 78, 20}, {33, -43, -38, -26},
                                                   {69, -67, -17, -57}, {18, -23, -2, -53}, {88, -96, 40, -48},
                                                   {-93, -47, -91, 59}, {-89, 82, 10, 94}, {-15, 7, 41, 14},
                                                   {63, 53, -40, 53}}; // (-90, -2, -49, 58)
} // namespace reference_vectors
