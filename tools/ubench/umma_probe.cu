// Probe: one tcgen05.mma.kind::i8 tile (M = 128, N = 128 or 96, K = 128) with hand-built operands, checked against the CPU.
// Pins down, on the real chip, everything the tensor-core Gaussian kernel relies on:
//   * the K-major SWIZZLE_128B shared-memory layout as written by ordinary threads (chunk ^= row & 7),
//   * the shared-memory matrix descriptor (start >> 4, SBO = 1024 B, version 1, layout type 2) and its K advance (+32 B),
//   * the instruction descriptor for s8 (A) x u8 (B) -> s32,
//   * D's placement in tensor memory (lane = m, column = n) and tcgen05.ld.32x32b.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_probe umma_probe.cu && ./umma_probe
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_desc_sw128(const void* smem) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF);        // start address
    d |= (uint64_t)1 << 16;                                  // leading byte offset (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                        // stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                                  // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                                  // SWIZZLE_128B
    return d;
}

template <int N>
__global__ void __launch_bounds__(128) probe(const int8_t* A, const uint8_t* B, int32_t* D) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                  // 128 rows x 128 B
    uint8_t* sB = smem + 16384;          // N rows x 128 B
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 32768 + 16);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // operands: logical [row][k] -> swizzled 128-byte rows
    for (int i = tid; i < 128 * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 128 + c * 16);
    }
    for (int i = tid; i < N * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 128 + c * 16);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;

    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);  // s32 = s8 (A) x u8 (B)
        const uint64_t da = make_desc_sw128(sA), db = make_desc_sw128(sB);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t acc = k > 0;
            asm volatile(
                "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                ::"r"(tmem), "l"(da + (uint64_t)(k * 2)), "l"(db + (uint64_t)(k * 2)), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0));
        }
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
    }
    // everybody waits for the MMAs
    asm volatile(
        "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
        ::"r"(smem_u32(bar)), "r"(0) : "memory");
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

    // D: lane = m, column = n
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t v[32];
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
              "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
              "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]),
              "=r"(v[30]), "=r"(v[31])
            : "r"(taddr));
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        const int m = warp * 32 + lane;
        for (int j = 0; j < 32; ++j) D[m * N + c0 + j] = (int32_t)v[j];
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

// Throughput: REPS x (4 MMAs of M 128 x N x K 32) back to back, one commit; clocks from first issue to completion.
template <int N>
__global__ void __launch_bounds__(128) rate(const int8_t* A, const uint8_t* B, long long* clocks, int reps) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;
    uint8_t* sB = smem + 16384;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 32768);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + 32768 + 16);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 128 * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sA + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(A + r * 128 + c * 16);
    }
    for (int i = tid; i < N * 8; i += 128) {
        const int r = i >> 3, c = i & 7;
        *reinterpret_cast<uint4*>(sB + r * 128 + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(B + r * 128 + c * 16);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (tid == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = (2u << 4) | (1u << 7) | ((uint32_t)(N >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t da = make_desc_sw128(sA), db = make_desc_sw128(sB);
        const long long c0 = clock64();
        for (int r = 0; r < reps; ++r) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint32_t acc = (r | k) > 0;
                asm volatile(
                    "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                    "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
                    ::"r"(tmem + (uint32_t)((r & 1) * 128)), "l"(da + (uint64_t)(k * 2)), "l"(db + (uint64_t)(k * 2)), "r"(idesc), "r"(acc), "r"(0), "r"(0), "r"(0), "r"(0));
            }
        }
        const long long c1 = clock64();
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
        asm volatile(
            "{\n\t.reg .pred p;\n\tWAIT_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra DONE_%=;\n\tbra WAIT_%=;\n\tDONE_%=:\n\t}\n"
            ::"r"(smem_u32(bar)), "r"(0) : "memory");
        const long long c2 = clock64();
        clocks[blockIdx.x * 2] = c1 - c0;
        clocks[blockIdx.x * 2 + 1] = c2 - c0;
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem));
}

template <int N>
void run_rate(int blocks) {
    std::vector<int8_t> A(128 * 128, 1);
    std::vector<uint8_t> B(N * 128, 1);
    int8_t* dA; uint8_t* dB; long long* dC;
    cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dC, blocks * 16);
    cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
    const int smem = 32768 + 64 + 1024, reps = 256;
    cudaFuncSetAttribute(rate<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    rate<N><<<blocks, 128, smem>>>(dA, dB, dC, reps);
    cudaDeviceSynchronize();
    std::vector<long long> C(blocks * 2);
    cudaMemcpy(C.data(), dC, blocks * 16, cudaMemcpyDeviceToHost);
    printf("N=%d, %d CTA(s): issue %.1f clk per MMA, issue..completion %.1f clk per MMA (M 128, K 32; %d MMAs)\n", N, blocks,
           (double)C[0] / (reps * 4), (double)C[1] / (reps * 4), reps * 4);
}

template <int N>
int run() {
    std::vector<int8_t> A(128 * 128);
    std::vector<uint8_t> B(N * 128);
    srand(7 + N);
    for (auto& a : A) a = (int8_t)(rand() % 255 - 127);
    for (auto& b : B) b = (uint8_t)(rand() & 255);
    int8_t* dA; uint8_t* dB; int32_t* dD;
    cudaMalloc(&dA, A.size()); cudaMalloc(&dB, B.size()); cudaMalloc(&dD, 128 * N * 4);
    cudaMemcpy(dA, A.data(), A.size(), cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size(), cudaMemcpyHostToDevice);
    cudaMemset(dD, 0xCD, 128 * N * 4);
    const int smem = 32768 + 64 + 1024;
    cudaFuncSetAttribute(probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    probe<N><<<1, 128, smem>>>(dA, dB, dD);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d: CUDA error %s\n", N, cudaGetErrorString(e)); return 1; }
    std::vector<int32_t> D(128 * N);
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N; ++n) {
            int32_t s = 0;
            for (int k = 0; k < 128; ++k) s += (int32_t)A[m * 128 + k] * (int32_t)B[n * 128 + k];
            if (s != D[m * N + n]) { if (bad < 5) printf("  N=%d mismatch m=%d n=%d want %d got %d\n", N, m, n, s, D[m * N + n]); ++bad; }
        }
    printf("N=%d: %d mismatches of %d\n", N, bad, 128 * N);
    return bad != 0;
}

int main() { int r = run<128>(); r |= run<96>(); run_rate<128>(1); run_rate<96>(1); run_rate<128>(148); run_rate<128>(296); return r; }
