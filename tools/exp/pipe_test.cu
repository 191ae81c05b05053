// Experiment: per-SM throughput of the integer ops the leaf decode is made of, alone and mixed,
// to learn which issue on the ALU pipe and which on the FMA pipe (sm_100a).
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

#define ITER 4096
template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t y, uint32_t z) {
    uint32_t x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x * 8 + i;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) x[i] = __vminu2(x[i], y);
            if (OP == 1) x[i] = __dp2a_lo(x[i], y, z);
            if (OP == 2) x[i] = __viaddmin_s16x2_relu(x[i], y, z);
            if (OP == 3) x[i] = x[i] * y + z;                       // IMAD
            if (OP == 4) x[i] = __byte_perm(x[i], y, z);            // PRMT
            if (OP == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y), "r"(z));
            if (OP == 6) x[i] = __funnelshift_l(x[i], y, 8);        // SHF
            if (OP == 7) { if (i & 1) x[i] = __dp2a_lo(x[i], y, z); else x[i] = __vminu2(x[i], y); }          // mix IDP + VIMNMX
            if (OP == 8) { if (i & 1) x[i] = x[i] * y + z; else x[i] = __vminu2(x[i], y); }                  // mix IMAD + VIMNMX
            if (OP == 9) { if (i & 1) x[i] = __dp2a_lo(x[i], y, z); else x[i] = x[i] * y + z; }              // mix IDP + IMAD
            if (OP == 10) x[i] = __popc(x[i]) + y;
            if (OP == 11) x[i] = __dp4a(x[i], y, z);
            if (OP == 12) x[i] = x[i] + y;                          // IADD
            if (OP == 13) x[i] = (x[i] < y) ? z : x[i];             // ISETP+SEL
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s ^= x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, uint32_t* out) {
    const int grid = 148 * 4;
    k<OP><<<grid, 256>>>(out, 0x00030007u, 0x00010001u);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<grid, 256>>>(out, 0x00030007u, 0x00010001u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double) grid * 8 * ITER * 8;       // 8 warps per CTA
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("{\"op\":\"%s\",\"ms\":%.3f,\"warp_instr_per_clk_per_sm\":%.3f}\n", name, ms, warp_instr / (ms * 1e-3) / (clk * 1e3) / 148);
}

int main() {
    uint32_t* out; cudaMalloc(&out, 148 * 4 * 256 * 4);
    run<0>("VIMNMX.U16x2", out); run<1>("IDP.2A", out); run<2>("VIADDMNMX.S16x2.RELU", out); run<3>("IMAD", out);
    run<4>("PRMT", out); run<5>("LOP3", out); run<6>("SHF", out); run<7>("mix IDP+VIMNMX", out); run<8>("mix IMAD+VIMNMX", out);
    run<9>("mix IDP+IMAD", out); run<10>("POPC+IADD", out); run<11>("IDP.4A", out); run<12>("IADD", out); run<13>("ISETP+SEL", out);
    return 0;
}
