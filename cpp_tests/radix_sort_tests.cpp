// radix_sort_tests.cpp — the reference's RadixSort cases (test/radix_sort_tests.cpp:54-193) against device
// pointers.  The reference checks keys only (permutation + sortedness, all-zero values); every case here also
// sorts values = input index and requires keys AND values to equal std::stable_sort of the pairs.
#include <algorithm>
#include <cstdint>
#include <numeric>
#include <unordered_map>
#include <utility>
#include <vector>

#include "glu/RadixSort.hpp"
#include "harness.hpp"
#include "util/Random.hpp"
#include "util/StopWatch.hpp"

using namespace glu;

namespace
{
    /// vector1 is a permutation of vector2 (multiset equality, test/radix_sort_tests.cpp:20-43)
    template<typename T> void check_permutation(const std::vector<T>& vector1, const std::vector<T>& vector2)
    {
        CHECK(vector1.size() == vector2.size());
        std::unordered_map<T, long> balance;
        for (const T& v : vector1)
            balance[v]++;
        for (const T& v : vector2)
            balance[v]--;
        CHECK(std::all_of(balance.begin(), balance.end(), [](const auto& kv) { return kv.second == 0; }));
    }

    template<typename T> void check_sorted(const std::vector<T>& vector) { CHECK(std::is_sorted(vector.begin(), vector.end())); }

    /// Sorts (keys, index) on the device and compares both arrays with std::stable_sort of the pairs.
    void sort_and_check(const std::vector<uint32_t>& keys, size_t num_steps = 0)
    {
        const size_t n = keys.size();
        std::vector<uint32_t> vals(n);
        std::iota(vals.begin(), vals.end(), 0u);
        DeviceBuffer key_buffer(keys), val_buffer(vals);
        RadixSort radix_sort;
        radix_sort(key_buffer.handle(), val_buffer.handle(), n, num_steps);
        const std::vector<uint32_t> sorted_keys = key_buffer.get_data<uint32_t>();
        const std::vector<uint32_t> sorted_vals = val_buffer.get_data<uint32_t>();

        if (num_steps == 0)
        { // the reference's own assertions
            check_permutation(keys, sorted_keys);
            check_sorted(sorted_keys);
        }
        const uint32_t mask = (num_steps == 0 || num_steps >= 8) ? 0xffffffffu : ((1u << (4 * num_steps)) - 1u);
        std::vector<std::pair<uint32_t, uint32_t>> pairs(n);
        for (size_t i = 0; i < n; i++)
            pairs[i] = {keys[i], vals[i]};
        std::stable_sort(pairs.begin(), pairs.end(),
                         [mask](const auto& a, const auto& b) { return (a.first & mask) < (b.first & mask); });
        bool equal = true;
        for (size_t i = 0; i < n && equal; i++)
            equal = pairs[i].first == sorted_keys[i] && pairs[i].second == sorted_vals[i];
        CHECK(equal);
    }
} // namespace

TEST_CASE("RadixSort-simple", "[.]")
{
    Random random(1);
    const std::vector<uint32_t> keys = random.sample_int_vector<uint32_t>(10, 0, UINT32_MAX);
    std::vector<uint32_t> vals(keys.size(), 0);
    DeviceBuffer key_buffer(keys), val_buffer(vals);
    RadixSort radix_sort;
    radix_sort(key_buffer.handle(), val_buffer.handle(), keys.size());
    print_buffer_hex(key_buffer);
    const std::vector<uint32_t> sorted_keys = key_buffer.get_data<uint32_t>();
    check_permutation(keys, sorted_keys);
    check_sorted(sorted_keys);
}

TEST_CASE("RadixSort-128-256-512-1024", "")
{
    for (size_t k_num_elements : {128, 256, 512, 1024})
    {
        Random random(1);
        sort_and_check(random.sample_int_vector<uint32_t>(k_num_elements, 0, UINT32_MAX));
    }
}

TEST_CASE("RadixSort-2048", "")
{
    Random random(1);
    sort_and_check(random.sample_int_vector<uint32_t>(2048, 0, 10)); // heavy duplicates: stability matters
}

TEST_CASE("RadixSort-multiple-sizes", "")
{
    for (size_t k_num_elements : {10993, 14978, 16243, 18985, 23857, 27865, 33363, 41298, 45821, 47487})
    {
        Random random(1);
        sort_and_check(random.sample_int_vector<uint32_t>(k_num_elements, 0, UINT32_MAX));
    }
}

// BASELINE.json configs[0]: 1,048,576 pairs — only a benchmark size in the reference, a checked case here;
// once with the reference's 31-bit generator and once with true 32-bit keys (bit 31 set).
TEST_CASE("RadixSort-1048576", "")
{
    Random random(1);
    sort_and_check(random.sample_int_vector<uint32_t>(1048576, 0, UINT32_MAX));
    std::mt19937 engine(1);
    std::vector<uint32_t> keys(1048576);
    for (uint32_t& k : keys)
        k = engine();
    sort_and_check(keys);
}

// num_steps is never exercised by the reference's tests: low 4*num_steps bits only, stable.
TEST_CASE("RadixSort-num-steps", "")
{
    std::mt19937 engine(5);
    std::vector<uint32_t> keys(50021);
    for (uint32_t& k : keys)
        k = engine();
    for (size_t num_steps : {1, 2, 3, 4, 5, 7, 8, 9})
        sort_and_check(keys, num_steps);
}

// Beyond the reference: key-only, bit-range and descending sorts through RadixSort::sort_ex.
TEST_CASE("RadixSort-ex-keys-only-bit-range-descending", "")
{
    std::mt19937 engine(9);
    for (size_t n : {3000, 100003, 2500001})
    {
        std::vector<uint32_t> keys(n), vals(n);
        for (uint32_t& k : keys)
            k = engine();
        std::iota(vals.begin(), vals.end(), 0u);
        RadixSort radix_sort;
        { // keys only, ascending and descending
            DeviceBuffer key_buffer(keys);
            radix_sort.sort_ex(key_buffer.handle(), nullptr, n);
            std::vector<uint32_t> expected = keys;
            std::sort(expected.begin(), expected.end());
            CHECK(key_buffer.get_data<uint32_t>() == expected);
            key_buffer.write_data(keys.data(), n * sizeof(uint32_t));
            radix_sort.sort_ex(key_buffer.handle(), nullptr, n, 0, 32, true);
            std::reverse(expected.begin(), expected.end());
            CHECK(key_buffer.get_data<uint32_t>() == expected);
        }
        for (bool descending : {false, true})
        { // pairs, key bits [6, 19) only
            const unsigned begin_bit = 6, end_bit = 19;
            DeviceBuffer key_buffer(keys), val_buffer(vals);
            radix_sort.sort_ex(key_buffer.handle(), val_buffer.handle(), n, begin_bit, end_bit, descending);
            const uint32_t mask = (1u << (end_bit - begin_bit)) - 1u;
            std::vector<std::pair<uint32_t, uint32_t>> pairs(n);
            for (size_t i = 0; i < n; i++)
                pairs[i] = {keys[i], vals[i]};
            std::stable_sort(pairs.begin(), pairs.end(), [=](const auto& a, const auto& b) {
                const uint32_t fa = (a.first >> begin_bit) & mask, fb = (b.first >> begin_bit) & mask;
                return descending ? fa > fb : fa < fb;
            });
            const std::vector<uint32_t> sorted_keys = key_buffer.get_data<uint32_t>();
            const std::vector<uint32_t> sorted_vals = val_buffer.get_data<uint32_t>();
            bool equal = true;
            for (size_t i = 0; i < n && equal; i++)
                equal = pairs[i].first == sorted_keys[i] && pairs[i].second == sorted_vals[i];
            CHECK(equal);
        }
    }
}

// Beyond the reference: 64-bit keys and payloads wider than 32 bits through RadixSort::sort_wide.
TEST_CASE("RadixSort-wide-u64-keys-and-wide-values", "")
{
    std::mt19937_64 engine(11);
    for (size_t n : {2500, 400003})
    {
        std::vector<uint64_t> keys(n);
        for (uint64_t& k : keys)
            k = engine() >> (engine() % 3 == 0 ? 40 : 0); // mixed magnitudes, many equal high words
        struct Wide
        {
            uint64_t a, b;
            bool operator==(const Wide& o) const { return a == o.a && b == o.b; }
        };
        std::vector<Wide> vals(n);
        std::vector<uint32_t> vals32(n);
        for (size_t i = 0; i < n; i++)
        {
            vals[i] = {i, ~uint64_t(i)};
            vals32[i] = uint32_t(i);
        }
        RadixSort radix_sort;
        for (bool descending : {false, true})
        {
            std::vector<size_t> order(n);
            std::iota(order.begin(), order.end(), size_t(0));
            std::stable_sort(order.begin(), order.end(),
                             [&](size_t x, size_t y) { return descending ? keys[x] > keys[y] : keys[x] < keys[y]; });
            { // 8-byte keys + 16-byte values (general permutation path)
                DeviceBuffer key_buffer(keys), val_buffer(vals);
                radix_sort.sort_wide(key_buffer.handle(), 8, val_buffer.handle(), 16, n, descending);
                const std::vector<uint64_t> k = key_buffer.get_data<uint64_t>();
                const std::vector<Wide> v = val_buffer.get_data<Wide>();
                bool equal = true;
                for (size_t i = 0; i < n && equal; i++)
                    equal = k[i] == keys[order[i]] && v[i] == vals[order[i]];
                CHECK(equal);
            }
            { // 8-byte keys + 4-byte values (four-sort path), 8-byte keys alone (two-sort path)
                DeviceBuffer key_buffer(keys), val_buffer(vals32), key_only(keys);
                radix_sort.sort_wide(key_buffer.handle(), 8, val_buffer.handle(), 4, n, descending);
                radix_sort.sort_wide(key_only.handle(), 8, nullptr, 0, n, descending);
                const std::vector<uint64_t> k = key_buffer.get_data<uint64_t>();
                const std::vector<uint64_t> k2 = key_only.get_data<uint64_t>();
                const std::vector<uint32_t> v = val_buffer.get_data<uint32_t>();
                bool equal = true;
                for (size_t i = 0; i < n && equal; i++)
                    equal = k[i] == keys[order[i]] && k2[i] == k[i] && v[i] == vals32[order[i]];
                CHECK(equal);
            }
        }
    }
}

// Beyond the reference: the segmented sort (the local step of the multi-GPU sort) through RadixSort::sort_segmented —
// segments at tile boundaries, and the same segments given as runs placed anywhere (one run per "source rank").
TEST_CASE("RadixSort-segmented-and-runs", "")
{
    const size_t tile = RadixSort::segment_tile();
    Random rng(5);
    const std::vector<std::vector<size_t>> run_counts = {{tile + 3, 17}, {0, 2 * tile}, {5, 0}, {tile - 1, tile + 1}};
    const unsigned begin_bit = 0, end_bit = 24;
    const uint32_t mask = (1u << end_bit) - 1u;
    std::vector<size_t> seg_count, flat, seg_of, seg_first_run;
    for (size_t s = 0; s < run_counts.size(); s++)
    {
        seg_first_run.push_back(flat.size());
        size_t c = 0;
        for (size_t r : run_counts[s])
        {
            flat.push_back(r);
            seg_of.push_back(s);
            c += r;
        }
        seg_count.push_back(c);
    }
    auto tiles_of = [&](size_t c) { return (c + tile - 1) / tile; };
    // expected result: every segment stable-sorted on the key bits that take part, segments compact one after the other
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> seg_pairs(run_counts.size());
    std::vector<std::vector<std::pair<uint32_t, uint32_t>>> run_pairs(flat.size());
    uint32_t next_val = 0;
    for (size_t r = 0; r < flat.size(); r++)
        for (size_t i = 0; i < flat[r]; i++)
        {
            const std::pair<uint32_t, uint32_t> p{rng.sample_int<uint32_t>(0, 0xffffffffu) & 0x00ff00ffu, next_val++};
            run_pairs[r].push_back(p);
            seg_pairs[seg_of[r]].push_back(p);
        }
    std::vector<std::pair<uint32_t, uint32_t>> expected;
    for (auto& sp : seg_pairs)
    {
        std::stable_sort(sp.begin(), sp.end(), [mask](const auto& a, const auto& b) { return (a.first & mask) < (b.first & mask); });
        expected.insert(expected.end(), sp.begin(), sp.end());
    }
    std::vector<uint32_t> counts32(seg_count.begin(), seg_count.end());
    size_t run_tiles = 0, seg_tiles = 0;
    for (size_t c : flat)
        run_tiles += tiles_of(c);
    for (size_t c : seg_count)
        seg_tiles += tiles_of(c);
    const size_t max_tiles = std::max(run_tiles, seg_tiles) + 2;
    auto check = [&](DeviceBuffer& ka, DeviceBuffer& va, DeviceBuffer& kb, DeviceBuffer& vb, bool in_b) {
        CHECK(in_b == (((end_bit - begin_bit + 7) / 8) % 2 == 1));
        const std::vector<uint32_t> k = (in_b ? kb : ka).get_data<uint32_t>();
        const std::vector<uint32_t> v = (in_b ? vb : va).get_data<uint32_t>();
        bool equal = true;
        for (size_t i = 0; i < expected.size() && equal; i++)
            equal = k[i] == expected[i].first && v[i] == expected[i].second;
        CHECK(equal);
    };
    { // segments at tile boundaries
        std::vector<uint32_t> ak(max_tiles * tile, 0x0badf00du), av(max_tiles * tile, 0xdeaddeadu);
        size_t first = 0;
        for (size_t s = 0, r = 0; s < run_counts.size(); s++)
        {
            size_t at = first * tile;
            for (size_t j = 0; j < run_counts[s].size(); j++, r++)
                for (const auto& p : run_pairs[r])
                {
                    ak[at] = p.first;
                    av[at++] = p.second;
                }
            first += tiles_of(seg_count[s]);
        }
        DeviceBuffer ka(ak), va(av), kb(ak), vb(av), cnt(counts32);
        RadixSort radix_sort;
        const bool in_b = radix_sort.sort_segmented(ka.handle(), va.handle(), kb.handle(), vb.handle(), cnt.handle(),
                                                    seg_count.size(), max_tiles, begin_bit, end_bit);
        check(ka, va, kb, vb, in_b);
    }
    { // the same segments as runs, laid out in REVERSE run order
        std::vector<uint32_t> ak(max_tiles * tile, 0x0badf00du), av(max_tiles * tile, 0xdeaddeadu);
        const size_t R = flat.size();
        std::vector<uint32_t> runs(5 * (R + 1), 0u);
        std::vector<size_t> first(R + 1, 0), phys(R, 0);
        for (size_t r = 0; r < R; r++)
            first[r + 1] = first[r] + tiles_of(flat[r]);
        size_t at = 1;
        for (size_t r = R; r-- > 0;)
        {
            phys[r] = at;
            at += tiles_of(flat[r]);
        }
        for (size_t r = 0; r < R; r++)
        {
            size_t e = phys[r] * tile;
            for (const auto& p : run_pairs[r])
            {
                ak[e] = p.first;
                av[e++] = p.second;
            }
            runs[0 * (R + 1) + r] = uint32_t(first[r]);
            runs[1 * (R + 1) + r] = uint32_t(phys[r]);
            runs[2 * (R + 1) + r] = uint32_t(flat[r]);
            runs[3 * (R + 1) + r] = uint32_t(seg_of[r]);
            runs[4 * (R + 1) + r] = uint32_t(first[seg_first_run[seg_of[r]]]);
        }
        runs[R] = uint32_t(first[R]);
        DeviceBuffer ka(ak), va(av), kb(ak), vb(av), cnt(counts32), run_buffer(runs);
        RadixSort radix_sort;
        const bool in_b = radix_sort.sort_segmented(ka.handle(), va.handle(), kb.handle(), vb.handle(), cnt.handle(),
                                                    seg_count.size(), max_tiles, begin_bit, end_bit, run_buffer.handle(), R);
        check(ka, va, kb, vb, in_b);
    }
}

TEST_CASE("RadixSort-benchmark", "[.][benchmark]")
{
    for (size_t k_num_elements : {1024, 16384, 65536, 131072, 524288, 1048576, 2097152, 4194304, 8388608, 16777216,
                                  33554432, 67108864, 134217728, 268435456})
    {
        // zero-filled like the reference's benchmark, and a uniform-random run beside it (all-equal keys put
        // every pair in one digit bin, which is not what a sort usually sees)
        std::vector<uint32_t> zeros(k_num_elements), vals(k_num_elements), uniform(k_num_elements);
        std::mt19937 engine(1);
        for (uint32_t& k : uniform)
            k = engine();
        DeviceBuffer key_buffer(zeros), val_buffer(vals);
        RadixSort radix_sort;
        radix_sort.prepare_internal_buffers(k_num_elements);
        radix_sort(key_buffer.handle(), val_buffer.handle(), k_num_elements); // warm-up
        const uint64_t ns_zero =
            measure_elapsed_time([&]() { radix_sort(key_buffer.handle(), val_buffer.handle(), k_num_elements); });
        key_buffer.write_data(uniform.data(), uniform.size() * sizeof(uint32_t));
        const uint64_t ns_uniform =
            measure_elapsed_time([&]() { radix_sort(key_buffer.handle(), val_buffer.handle(), k_num_elements); });
        std::printf("Radix sort; Num elements: %zu, Elapsed: %s (uniform keys: %s)\n", k_num_elements,
                    ns_to_human_string(ns_zero).c_str(), ns_to_human_string(ns_uniform).c_str());
    }
}
