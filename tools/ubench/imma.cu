// Micro-benchmark: issue rate of the legacy integer tensor-core instruction (mma.sync m16n8k32 u8 x s8 -> s32,
// SASS IMMA.16832) on sm_100a, alone and mixed with the integer ALU / FMA pipe instructions of the stencil
// kernels.  Question it answers: can the horizontal / vertical taps of the 5-tap Gaussians
// run as banded-Toeplitz MMAs at full HBM rate?
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o imma imma.cu && ./imma
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITER 2048
#define NACC 8

__device__ __forceinline__ void imma(int (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.u8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(c[0]), "+r"(c[1]), "+r"(c[2]), "+r"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// MIX: 0 = MMA only; k > 0 = k PRMT + k IMAD per MMA (independent registers)
template <int MIX>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t s0, uint32_t s1) {
    int c[NACC][4];
    uint32_t a[4], b[2];
    uint32_t r[NACC], q[NACC];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = threadIdx.x * 3 + i + s0;
    b[0] = s1; b[1] = s1 * 5;
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
        r[i] = threadIdx.x + i; q[i] = threadIdx.x * 9 + i;
#pragma unroll
        for (int j = 0; j < 4; ++j) c[i][j] = i + j;
    }
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            imma(c[i], a, b);
#pragma unroll
            for (int m = 0; m < MIX; ++m) {
                r[i] = __byte_perm(r[i], s1, 0x5432);
                q[i] = q[i] * 6u + s0;
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s ^= r[i] ^ q[i] ^ c[i][0] ^ c[i][1] ^ c[i][2] ^ c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MIX>
void run(const char* name, int warps_per_block) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 4;
    uint32_t* out;
    cudaMalloc(&out, blocks * 256 * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MIX><<<blocks, warps_per_block * 32>>>(out, 3, 1u << 24);
    cudaEventRecord(e0);
    k<MIX><<<blocks, warps_per_block * 32>>>(out, 3, 1u << 24);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double mma = (double)blocks * warps_per_block * ITER * NACC;
    const double per_sm_per_us = mma / sms / (ms * 1e3);
    printf("%-34s warps/block %d  %8.3f ms  %7.1f MMA/us/SM (= %.3f per clk per SM at 1.9 GHz, one per %.1f clk per SMSP); other instr/clk/SM %.2f\n",
           name, warps_per_block, ms, per_sm_per_us, per_sm_per_us / 1900.0, 4.0 * 1900.0 / per_sm_per_us, per_sm_per_us / 1900.0 * 2 * MIX);
    cudaFree(out);
}

int main() {
    for (int w : {4, 8}) {
        run<0>("IMMA.16832 only", w);
        run<1>("IMMA + 1 PRMT + 1 IMAD", w);
        run<2>("IMMA + 2 PRMT + 2 IMAD", w);
        run<4>("IMMA + 4 PRMT + 4 IMAD", w);
        run<8>("IMMA + 8 PRMT + 8 IMAD", w);
    }
    return 0;
}
