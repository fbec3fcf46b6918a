// glu/errors.hpp — the reference's error convention (glu/errors.hpp:8-18): a failed check prints its
// message to stderr and terminates the process with exit(1).  There are no exceptions and no error
// codes at the C++ class level; the C ABI underneath (glu_b200.h) returns glu_status codes instead.
#ifndef GLU_B200_ERRORS_HPP
#define GLU_B200_ERRORS_HPP

#include <cstdio>
#include <cstdlib>

#include "../glu_b200.h"

#define GLU_CHECK_STATE(condition_, ...)                                                                               \
    do                                                                                                                 \
    {                                                                                                                  \
        if (__builtin_expect(!(condition_), 0))                                                                        \
        {                                                                                                              \
            std::fprintf(stderr, __VA_ARGS__);                                                                         \
            std::fputc('\n', stderr);                                                                                  \
            std::exit(1);                                                                                              \
        }                                                                                                              \
    } while (0)

#define GLU_CHECK_ARGUMENT(condition_, ...) GLU_CHECK_STATE(condition_, __VA_ARGS__)
#define GLU_FAIL(...) GLU_CHECK_STATE(false, __VA_ARGS__)

// A non-zero glu_status from the C ABI is fatal at this level, like a GL error would be for the reference.
#define GLU_CHECK_STATUS(call_)                                                                                        \
    do                                                                                                                 \
    {                                                                                                                  \
        const int glu_status__ = (call_);                                                                              \
        if (__builtin_expect(glu_status__ != GLU_SUCCESS, 0))                                                          \
        {                                                                                                              \
            std::fprintf(stderr, "%s failed: %s%s%s\n", #call_, glu_status_string(glu_status__),                      \
                         glu_status__ == GLU_ERROR_CUDA ? ": " : "",                                                   \
                         glu_status__ == GLU_ERROR_CUDA ? glu_last_cuda_error() : "");                                 \
            std::exit(1);                                                                                              \
        }                                                                                                              \
    } while (0)

#endif // GLU_B200_ERRORS_HPP
