// main.cpp — glu_test: device banner + test runner (the role of the reference's test/main.cpp:60-109, which
// creates a hidden GLFW window for a GL 4.6 context, prints the device limits and hands over to Catch2).
//   ./glu_test                 all regular cases
//   ./glu_test [benchmark]     the hidden benchmark cases (README.md:136-140)
//   ./glu_test <case name>     one case
#include <cstdio>

#include "glu/device_utils.hpp"
#include "harness.hpp"

int main(int argc, char* argv[])
{
    int device_count = 0;
    if (glu_device_count(&device_count) != GLU_SUCCESS || device_count == 0)
    {
        std::fprintf(stderr, "glu_test: no CUDA device (%s); there is no CPU fallback\n", glu_last_cuda_error());
        return 2;
    }
    GLU_CHECK_STATUS(glu_set_device(0));
    char name[256];
    int sm_count = 0, cc_major = 0, cc_minor = 0, warp_size = 0;
    size_t total_mem = 0;
    GLU_CHECK_STATUS(glu_device_info(0, name, sizeof name, &sm_count, &cc_major, &cc_minor, &total_mem, &warp_size));
    std::printf("Device: %s\nCompute capability: %d.%d\nSMs: %d\nWarp size: %d\nGlobal memory: %.1f GiB\nglu_b200 version: %d\n",
                name, cc_major, cc_minor, sm_count, warp_size, double(total_mem) / double(1ull << 30), glu_version());
    return harness::run(argc, argv);
}
