// Wire forms of per-read results (not part of the search / locate kernels: their sources are hashed for the ncu records).
//   narrow_ranges_kernel   lo / hi (u64) -> two u32 planes, for RBG_NARROW_RANGES on an index with n <= 2^32:
//                          8 instead of 16 bytes per read leave the device.  With eight ranks on one host the combined
//                          H2D + D2H volume is what bounds the end-to-end step (profiles/r2_pcie_sweep_8gpu.jsonl).
#include "kernels.cuh"

namespace rbg {

namespace {
__global__ void __launch_bounds__(256) narrow_ranges_kernel(const uint64_t* __restrict__ lo, const uint64_t* __restrict__ hi,
                                                            uint32_t* __restrict__ lo32, uint32_t* __restrict__ hi32, uint64_t r0, uint64_t r1) {
    for (uint64_t i = r0 + (uint64_t) blockIdx.x * blockDim.x + threadIdx.x; i < r1; i += (uint64_t) gridDim.x * blockDim.x) {
        lo32[i] = (uint32_t) lo[i];
        hi32[i] = (uint32_t) hi[i];
    }
}
}  // namespace

// Test scaffolding (RBG_TEST_STALL, api.cu): one thread that keeps its stream busy for `ns` nanoseconds, to force an order of
// events between the two search streams that real timing produces only rarely.
namespace {
__global__ void stall_kernel(unsigned long long ns) {
    unsigned long long t0, t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    do {
        __nanosleep(1000);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    } while (t - t0 < ns);
}
}  // namespace
void launch_stall(unsigned long long ns, cudaStream_t st) { stall_kernel<<<1, 1, 0, st>>>(ns); }

int launch_narrow_ranges(const uint64_t* lo, const uint64_t* hi, uint32_t* lo32, uint32_t* hi32, uint64_t r0, uint64_t r1, cudaStream_t st) {
    if (r1 <= r0) return 0;
    narrow_ranges_kernel<<<grid_for(r1 - r0, 256, 8), 256, 0, st>>>(lo, hi, lo32, hi32, r0, r1);
    return 1;
}

}  // namespace rbg
