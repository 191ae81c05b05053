// Experiment: does the allocation method change the TLB reach for random 64-byte gathers?
// (a) cudaMalloc, (b) cuMemCreate + cuMemMap on a 512 MB-aligned VA with 512 MB-multiple size.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256) gather(const uint32_t* buf, uint64_t n_lines, int iters, unsigned long long* sink) {
    uint64_t state = mix64(((uint64_t) blockIdx.x * blockDim.x + threadIdx.x) * 2654435761ull + 12345);
    uint32_t acc = 0;
    for (int it = 0; it < iters; it += 4) {
        uint32_t w[4][16];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const uint64_t line = __umul64hi(mix64(state + u), n_lines);
            const uint32_t* p = buf + line * 16;
#pragma unroll
            for (int h = 0; h < 2; ++h)
                asm volatile("ld.global.nc.L1::no_allocate.v8.u32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                             : "=r"(w[u][8 * h + 0]), "=r"(w[u][8 * h + 1]), "=r"(w[u][8 * h + 2]), "=r"(w[u][8 * h + 3]),
                               "=r"(w[u][8 * h + 4]), "=r"(w[u][8 * h + 5]), "=r"(w[u][8 * h + 6]), "=r"(w[u][8 * h + 7])
                             : "l"(p + 8 * h));
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 16; ++i) acc ^= w[u][i];
        state = mix64(state + 4 + acc);
    }
    if (acc == 0x12345678u) atomicAdd(sink, 1ull);
}

static double run(const uint32_t* buf, size_t bytes) {
    unsigned long long* sink;
    cudaMalloc(&sink, 8);
    const int grid = 148 * 8, iters = 256;
    gather<<<grid, 256>>>(buf, bytes / 64, 8, sink);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    gather<<<grid, 256>>>(buf, bytes / 64, iters, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(sink);
    return (double) grid * 256 * iters / (ms * 1e-3) / 1e9;   // G lines/s
}

int main() {
    cudaSetDevice(0);
    cudaFree(0);
    const size_t HUGE = 512ull << 20;
    size_t sizes[] = {128ull << 20, 256ull << 20, 384ull << 20, 512ull << 20, 1024ull << 20, 2048ull << 20, 8192ull << 20};
    // (a) cudaMalloc
    for (size_t sz : sizes) {
        void* p; if (cudaMalloc(&p, sz) != cudaSuccess) { printf("cudaMalloc fail\n"); return 1; }
        cudaMemset(p, 0x5A, sz);
        printf("{\"alloc\":\"cudaMalloc\",\"MB\":%zu,\"glines_per_s\":%.2f}\n", sz >> 20, run((const uint32_t*) p, sz));
        cudaFree(p);
    }
    // (b) VMM: physical handle of k*512MB mapped at a 512MB-aligned VA
    CUmemAllocationProp prop = {};
    prop.type = CU_MEM_ALLOCATION_TYPE_PINNED;
    prop.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
    prop.location.id = 0;
    size_t gmin = 0, grec = 0;
    cuMemGetAllocationGranularity(&gmin, &prop, CU_MEM_ALLOC_GRANULARITY_MINIMUM);
    cuMemGetAllocationGranularity(&grec, &prop, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED);
    printf("{\"granularity_min\":%zu,\"granularity_recommended\":%zu}\n", gmin, grec);
    for (size_t sz : sizes) {
        size_t padded = (sz + HUGE - 1) / HUGE * HUGE;
        CUmemGenericAllocationHandle h;
        if (cuMemCreate(&h, padded, &prop, 0) != CUDA_SUCCESS) { printf("cuMemCreate fail\n"); continue; }
        CUdeviceptr va;
        if (cuMemAddressReserve(&va, padded, HUGE, 0, 0) != CUDA_SUCCESS) { printf("reserve fail\n"); continue; }
        if (cuMemMap(va, padded, 0, h, 0) != CUDA_SUCCESS) { printf("map fail\n"); continue; }
        CUmemAccessDesc ad = {};
        ad.location = prop.location;
        ad.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
        cuMemSetAccess(va, padded, &ad, 1);
        cudaMemset((void*) va, 0x5A, sz);
        printf("{\"alloc\":\"vmm512\",\"MB\":%zu,\"va_mod_512M\":%llu,\"glines_per_s\":%.2f}\n", sz >> 20,
               (unsigned long long) (va % HUGE), run((const uint32_t*) va, sz));
        cuMemUnmap(va, padded);
        cuMemAddressFree(va, padded);
        cuMemRelease(h);
    }
    return 0;
}
