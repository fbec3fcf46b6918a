// harness.hpp — the few Catch2 features the reference's test-suite uses (TEST_CASE with tags, CHECK /
// REQUIRE, hidden "[.]" cases, "[benchmark]" selection; test/*.cpp), as a ~100-line registry so that
// glu_test builds with nothing but g++.  Parametrised cases (Catch2 GENERATE) are plain loops here.
#pragma once

#include <cmath>
#include <cstdio>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

namespace harness
{
    struct TestCase
    {
        const char* name;
        const char* tags; // "" | "[.]" | "[.][benchmark]"
        std::function<void()> body;
    };

    inline std::vector<TestCase>& registry()
    {
        static std::vector<TestCase> cases;
        return cases;
    }

    struct Registrar
    {
        Registrar(const char* name, const char* tags, std::function<void()> body)
        {
            registry().push_back({name, tags, std::move(body)});
        }
    };

    struct Counters
    {
        size_t checks = 0, failed = 0;
        bool abort_case = false;
    };
    inline Counters& counters()
    {
        static Counters c;
        return c;
    }

    struct RequireFailed
    {
    };

    inline void report(bool ok, bool fatal, const char* expr, const char* file, int line)
    {
        counters().checks++;
        if (ok)
            return;
        counters().failed++;
        std::printf("%s:%d: FAILED: %s( %s )\n", file, line, fatal ? "REQUIRE" : "CHECK", expr);
        if (fatal)
            throw RequireFailed{};
    }

    inline bool within_abs(double value, double target, double margin) { return std::fabs(value - target) <= margin; }

    /// Runs the cases selected by the command line; returns the process exit code.
    ///   (no argument)   every case that is not hidden ("[.]")
    ///   [tag]           every case carrying the tag, hidden or not (e.g. "[benchmark]")
    ///   name            the case with that exact name
    inline int run(int argc, char* argv[])
    {
        std::vector<std::string> filters(argv + 1, argv + argc);
        size_t ran = 0, failed_cases = 0;
        for (const TestCase& tc : registry())
        {
            bool selected;
            if (filters.empty())
                selected = std::strstr(tc.tags, "[.]") == nullptr;
            else
            {
                selected = false;
                for (const std::string& f : filters)
                    selected = selected || (f.size() && f[0] == '[' ? std::strstr(tc.tags, f.c_str()) != nullptr : f == tc.name);
            }
            if (!selected)
                continue;
            ran++;
            const size_t failed_before = counters().failed;
            std::printf("---- %s %s\n", tc.name, tc.tags);
            try
            {
                tc.body();
            }
            catch (const RequireFailed&)
            {
            }
            if (counters().failed != failed_before)
                failed_cases++;
            std::fflush(stdout);
        }
        std::printf("===============================================================================\n");
        if (counters().failed == 0)
            std::printf("All tests passed (%zu assertions in %zu test cases)\n", counters().checks, ran);
        else
            std::printf("test cases: %zu | %zu failed\nassertions: %zu | %zu failed\n", ran, failed_cases, counters().checks,
                        counters().failed);
        return counters().failed == 0 ? 0 : 1;
    }
} // namespace harness

#define HARNESS_CAT2(a, b) a##b
#define HARNESS_CAT(a, b) HARNESS_CAT2(a, b)
#define TEST_CASE(name_, tags_)                                                                                        \
    static void HARNESS_CAT(test_body_, __LINE__)();                                                                   \
    static harness::Registrar HARNESS_CAT(test_reg_, __LINE__)(name_, tags_, HARNESS_CAT(test_body_, __LINE__));       \
    static void HARNESS_CAT(test_body_, __LINE__)()
#define CHECK(...) harness::report(bool(__VA_ARGS__), false, #__VA_ARGS__, __FILE__, __LINE__)
#define REQUIRE(...) harness::report(bool(__VA_ARGS__), true, #__VA_ARGS__, __FILE__, __LINE__)
#define CHECK_WITHIN_ABS(value_, target_, margin_)                                                                     \
    harness::report(harness::within_abs(double(value_), double(target_), double(margin_)), false,                      \
                    #value_ " within " #margin_ " of " #target_, __FILE__, __LINE__)
