// Micro-benchmark: packed-half arithmetic (HFMA2 / HADD2 / HMUL2) on sm_100a as a replacement for the packed
// u16x2 integer arithmetic of the Gaussian kernels.  Two questions:
//  (1) issue rate of HFMA2 alone, with SUBNORMAL operands (a u8 in a 16-bit lane IS the half value b * 2^-24),
//      and mixed with PRMT / IMAD;
//  (2) exactness: fma.rn.f16x2(S, 2^-4, 0) on a subnormal-scaled integer S must equal round-half-even(S / 16)
//      (the subnormal ulp is 2^-24, so the FMA's single rounding IS the reference's cvRound), for every S.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o half_pipes half_pipes.cu && ./half_pipes
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#define ITER 4096
#define NACC 8

__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm volatile("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm volatile("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hmul2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm volatile("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}

template <int OP>
__global__ void __launch_bounds__(256) k(uint32_t* out, uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r[NACC];
#pragma unroll
    for (int i = 0; i < NACC; ++i) r[i] = ((threadIdx.x * 7 + i) & 0xFF) * 0x00010001u;   // subnormal halves
    for (int it = 0; it < ITER; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) {
            if (OP == 0) r[i] = hfma2(r[i], a, b);                    // HFMA2, stays subnormal (a = 1.0, b = 0)
            if (OP == 1) r[i] = hadd2(r[i], b);                       // HADD2
            if (OP == 2) r[i] = hmul2(r[i], a);                       // HMUL2
            if (OP == 3) { if (i & 1) r[i] = __byte_perm(r[i], c, 0x5432); else r[i] = hfma2(r[i], a, b); }   // PRMT / HFMA2 1:1
            if (OP == 4) { if ((i & 3) == 0) r[i] = __byte_perm(r[i], c, 0x5432); else r[i] = hfma2(r[i], a, b); } // 1:3
            if (OP == 5) { if (i & 1) r[i] = r[i] * 6u + c; else r[i] = hfma2(r[i], a, b); }                   // IMAD / HFMA2 1:1
            if (OP == 6) { if ((i & 3) == 0) r[i] = r[i] * 6u + c; else r[i] = hfma2(r[i], a, b); }             // IMAD / HFMA2 1:3
            if (OP == 7) r[i] = hfma2(r[i] | 0x3C003C00u, a, b);      // LOP3 + HFMA2 on NORMAL operands (compare with OP 0)
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
    const uint32_t one = 0x3C003C00u;  // half2(1, 1)
    k<OP><<<blocks, 256>>>(out, one, 0u, 1u << 24);
    cudaEventRecord(e0);
    k<OP><<<blocks, 256>>>(out, one, 0u, 1u << 24);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double warp_instr = (double)blocks * 8 * ITER * NACC * ops_per_iter;
    const double per_sm_per_us = warp_instr / sms / (ms * 1e3);
    printf("%-28s %8.3f ms  %7.1f warp-instr/us/SM  (= %.2f per cycle per SM at 1.9 GHz, %.2f per SMSP)\n", name, ms, per_sm_per_us,
           per_sm_per_us / 1900.0, per_sm_per_us / 1900.0 / 4);
    cudaFree(out);
}

// exactness: every S in [-2048, 2047] (centred 5-tap sums) and [0, 1020] (3-tap sums)
__global__ void exact(int* bad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;   // 0 .. 4095
    const int s = i - 2048;                                // centred sum S' = S - 2048, S in [0, 4095]
    // S' * 2^-24 as a half: exact for |S'| <= 2048
    const __half hs = __float2half_rn((float)s * 5.9604644775390625e-08f);
    const uint32_t x = (uint32_t)__half_as_ushort(hs) * 0x00010001u;
    const uint32_t sixteenth = 0x2C002C00u;                // half2(2^-4)
    const uint32_t c128 = 0x00800080u;                     // 128 * 2^-24 (subnormal)
    const uint32_t r = hfma2(x, sixteenth, c128);          // RNE(S'/16) + 128 = RNE(S/16)
    const int S = s + 2048;
    const int want = (S + 7 + ((S >> 4) & 1)) >> 4;         // the oracle's integer round-half-even, 0 .. 256
    if (want <= 255 && ((r & 0xFFFFu) != (uint32_t)want || (r >> 16) != (uint32_t)want)) atomicAdd(bad, 1);
    // 3-tap: S in [0, 1023], RNE(S / 4), no centring
    if (i < 1024) {
        const uint32_t y = (uint32_t)i * 0x00010001u;      // S * 2^-24, subnormal or (>= 1024: not reached) normal
        const uint32_t q = hmul2(y, 0x34003400u);          // * 0.25
        const int w4 = (i + 1 + ((i >> 2) & 1)) >> 2;
        if ((q & 0xFFFFu) != (uint32_t)w4) atomicAdd(bad + 1, 1);
    }
    // centred T' = RNE(S'/16) as a signed subnormal, then back: T' + 128
    {
        const uint32_t t = hmul2(x, sixteenth);            // RNE(S'/16), in [-128, 128]
        const uint32_t back = hadd2(t, c128);
        if (want <= 255 && (back & 0xFFFFu) != (uint32_t)want) atomicAdd(bad + 2, 1);
    }
}

int main() {
    run<0>("HFMA2 subnormal", 1);
    run<7>("LOP3+HFMA2 normal (2 instr)", 2);
    run<1>("HADD2 subnormal", 1);
    run<2>("HMUL2 subnormal", 1);
    run<3>("PRMT/HFMA2 1:1", 1);
    run<4>("PRMT/HFMA2 1:3", 1);
    run<5>("IMAD/HFMA2 1:1", 1);
    run<6>("IMAD/HFMA2 1:3", 1);
    int* bad; cudaMalloc(&bad, 12); cudaMemset(bad, 0, 12);
    exact<<<16, 256>>>(bad);
    int h[3]; cudaMemcpy(h, bad, 12, cudaMemcpyDeviceToHost);
    printf("exactness: fma(S', 2^-4, 128u) vs integer RNE: %d mismatches of 4096; mul(S, 1/4): %d of 1024; centred round trip: %d\n", h[0], h[1], h[2]);
    return 0;
}
