// util/StopWatch.hpp — nanoseconds as "x.xxx s" / "x.xxx ms" / "n ns", the format of the reference's
// benchmark lines (test/util/StopWatch.hpp:11-32), so README-style tables can be regenerated.
#pragma once

#include <cstdint>
#include <cstdio>
#include <string>

namespace glu
{
    inline std::string ns_to_human_string(uint64_t ns)
    {
        const double ms = double(ns) / 1e6, s = double(ns) / 1e9;
        char text[64];
        if (s >= 0.1)
            std::snprintf(text, sizeof text, "%.3f s", s);
        else if (ms >= 0.001)
            std::snprintf(text, sizeof text, "%.3f ms", ms);
        else
            std::snprintf(text, sizeof text, "%llu ns", (unsigned long long) ns);
        return text;
    }
} // namespace glu
