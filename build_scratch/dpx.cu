#include <cstdint>
__global__ void k(const uint32_t* a, const uint32_t* b, const uint32_t* c, uint32_t* o) {
    uint32_t x = a[threadIdx.x], y = b[threadIdx.x], z = c[threadIdx.x];
    o[threadIdx.x] = __viaddmin_s16x2_relu(x, y, z);
    o[threadIdx.x + 32] = __viaddmin_s32_relu((int) x, (int) y, (int) z);
    o[threadIdx.x + 64] = __vminu2(x, y);
    o[threadIdx.x + 96] = __dp2a_lo(x, y, z);
    o[threadIdx.x + 128] = __vimin3_s16x2(x, y, z);
    o[threadIdx.x + 160] = __popc(x);
    o[threadIdx.x + 192] = __vcmpeq4(x, y);
    o[threadIdx.x + 224] = __vsadu4(x, y);
    o[threadIdx.x + 256] = __dp4a(x, y, z);
}
