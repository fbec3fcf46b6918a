// Compile check: the umbrella of public headers can be included together, twice, in one translation unit
// (the role of the reference's test/generated/test_include_*.cpp for its single-header dist/ copies).
#include "glu/RadixSort.hpp"
#include "glu/BlellochScan.hpp"
#include "glu/Reduce.hpp"
#include "glu/data_types.hpp"
#include "glu/device_utils.hpp"
#include "glu/errors.hpp"
#include "glu/RadixSort.hpp"
