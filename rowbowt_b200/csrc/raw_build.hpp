// From rb_build's raw inputs (.bwt/.ssa/.esa/.ma) to the flat arrays of formats.hpp; see raw_build.cu.
#pragma once
#include <cstdint>
#include <string>

#include "formats.hpp"

namespace rbg {

struct RawBuildStats {
    uint64_t bwt_bytes = 0, runs = 0;
    double s_read = 0;          // fread of the .bwt into pinned memory
    double s_gpu_wait = 0;      // host time blocked on the device
    float ms_kernels = 0;       // H2D + run-length kernels, CUDA events
    uint32_t launches = 0;
};

RunsBwt rle_bwt_gpu(const std::string& bwt_path, int device, RawBuildStats* stats);
ToeholdArrays toehold_from_raw(const std::string& ssa_path, const std::string& esa_path, uint64_t n, uint64_t r, int device);
MarkerArrays markers_from_ma(const std::string& ma_path);

}  // namespace rbg
