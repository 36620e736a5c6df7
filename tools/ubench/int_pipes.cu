// Micro-benchmark: issue rate of the integer instructions the stencil kernels are made of (sm_100a).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipes int_pipes.cu && ./int_pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 4096
#define NACC 8

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t a, uint32_t b) {
    uint32_t r[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) r[i] = threadIdx.x * 7 + i + a;
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (OP == 0) r[i] = __byte_perm(r[i], b, 0x5432);                 // PRMT 2 regs
            if (OP == 1) r[i] = __byte_perm(r[i], 0, 0x4240);                 // PRMT with RZ
            if (OP == 2) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(r[i]) : "r"(a), "r"(b));   // LOP3
            if (OP == 3) asm volatile("shf.r.clamp.b32 %0, %0, %1, 3;" : "+r"(r[i]) : "r"(b));           // SHF
            if (OP == 4) r[i] = r[i] * 6u + b;                                 // IMAD imm
            if (OP == 5) asm volatile("add.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));                      // IADD (compiler's choice)
            if (OP == 6) r[i] = __vimin3_u16x2(r[i], a, b);                    // VIMNMX3
            if (OP == 7) r[i] = __umulhi(r[i] & 0x01000100u, 1u << 24) + r[i]; // LOP3 + LEA.HI
            if (OP == 8) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(r[i]) : "r"(a), "r"(b));       // IMAD 3 regs
            if (OP == 9) r[i] = __dp4a(r[i], a, b);                            // IDP.4A
            if (OP == 10) asm volatile("and.b32 %0, %0, 0x00FF00FF;" : "+r"(r[i]));                      // LOP3 imm
            if (OP == 11) asm volatile("add.u32 %0, %0, 0x70007;" : "+r"(r[i]));                         // add imm
            if (OP == 12) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b));                  // IMAD.HI, register multiplier
            if (OP == 13) { if (i & 1) r[i] = __byte_perm(r[i], b, 0x5432); else r[i] = r[i] * 6u + b; } // PRMT / IMAD alternating
            if (OP == 14) { if (i & 1) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(r[i]) : "r"(b)); else r[i] = __byte_perm(r[i], b, 0x5432); } // IMAD.HI / PRMT
            if (OP == 15) { if ((i & 3) == 0) r[i] = __byte_perm(r[i], b, 0x5432); else r[i] = r[i] * 6u + b; } // 1 PRMT : 3 IMAD

        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s ^= r[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
void run(const char* name, int ops_per_iter) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8;
    uint32_t* out;
    cudaMalloc(&out, blocks * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, 256>>>(out, 3, 1u << 24);
    cudaEventRecord(e0);
    k<OP><<<blocks, 256>>>(out, 3, 1u << 24);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const double warp_instr = (double)blocks * 8 * ITER * NACC * ops_per_iter;   // per launch
    const double per_sm_per_us = warp_instr / sms / (ms * 1e3);
    printf("%-22s %8.3f ms  %7.1f warp-instr/us/SM  (= %.2f per cycle per SM at 1.9 GHz, %.2f per SMSP)\n", name, ms, per_sm_per_us,
           per_sm_per_us / 1900.0, per_sm_per_us / 1900.0 / 4);
    cudaFree(out);
}

int main() {
    run<0>("PRMT r,r", 1);
    run<1>("PRMT r,RZ", 1);
    run<2>("LOP3 3 regs", 1);
    run<10>("LOP3 imm", 1);
    run<3>("SHF", 1);
    run<4>("IMAD imm", 1);
    run<8>("IMAD 3 regs", 1);
    run<5>("IADD", 1);
    run<11>("IADD imm", 1);
    run<6>("VIMNMX3.U16x2", 1);
    run<7>("LOP3+LEA.HI (2 instr)", 2);
    run<9>("IDP.4A", 1);
    run<12>("IMAD.HI (reg mult)", 1);
    run<13>("PRMT/IMAD 1:1", 1);
    run<14>("PRMT/IMAD.HI 1:1", 1);
    run<15>("PRMT/IMAD 1:3", 1);
    return 0;
}
