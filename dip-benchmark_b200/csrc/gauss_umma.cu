// The separable 5-tap Gaussian (1x5 then 5x1, [1 4 6 4 1]/16, u8 intermediate like the reference's `aux` image,
// opencv:88-90 / sycl:353-356) on the 5th-generation tensor cores: tcgen05.mma.kind::i8 with TMEM accumulators.
//
// Why tensor cores for a memory-bound op: the packed-SIMD kernel (stencil_tiled.cuh) needs ~180 warp instructions per
// 512 output bytes for the 5+5 taps and two exact roundings, which is what bounds it once the 1 kW power cap holds the
// SMs near 1.5 GHz (0.66 of the copy bandwidth).  Both passes are matrix products with a banded Toeplitz matrix:
//
//   pass 1   D1[m][r] = sum_k A1[m][k] * img[r][k]      m: output byte of the tile row (96), r: image row of the tile (128),
//                                                       k: byte of the 128-byte window; A1[m][16 + m + (j-2) C] = 16 w[j]
//   pass 2   D2[y][x] = sum_r A2[y][r] * T[x][r]        y: output row (124), x: output byte (96), r: row of T (128);
//                                                       A2[y][y + j] = 16 w[j],  T = the u8 rounded pass-1 result
//
// so the multiply-adds AND the neighbour gathering (the PRMT funnels of the SIMD kernel) move to the tensor pipe, and
// the CUDA cores only round: D holds 16 * S, so round-half-even(S / 16) is the high byte of (16 S + 0x7F + bit 8), the
// same LOP3 + LEA.HI pair as everywhere else in the library, on two values per register.  ~75 thread instructions
// per 16 output bytes instead of ~180.
//
// One persistent, warp-specialised CTA per SM (16 warps) walks tiles of 96 bytes x 124 rows; every hand-over is an mbarrier,
// nothing in the steady state is a block-wide barrier, and TMEM / the staging buffers are double-buffered so that the
// stages of consecutive tiles overlap:
//   TMA warp   the 128-byte x 128-row window of the source into a 3-deep ring, SWIZZLE_128B = the K-major canonical
//              layout of the B operand
//   MMA warp   patches the reflect-101 halo of windows on the image rim, then one lane issues pass 1 of tile i
//              (4 x tcgen05.mma M 128, N 128, K 32; D1[i&1] in tensor memory) and pass 2 of tile i-1 (N 96; D2)
//   E1 warps   (6) thread m < 96 reads ITS TMEM lane = the 128 row sums of output byte m (tcgen05.ld), rounds them to u8 and
//              writes them as row m of the next B operand T -- K-major for pass 2 is "rows contiguous per byte column",
//              which is exactly what one TMEM lane holds
//   E2 warps   (8) thread y < 124 reads its lane = the 96 column sums of output row y, rounds, stages the row in shared
//              memory; one thread sends the 96 x 124 tile to the destination with a TMA store (full sectors, clipped
//              at the image edge by the tensor map)
// tools/ubench/umma_probe.cu is the unit test of the descriptor encodings used here.
#include "stencil_tiled.cuh"

namespace dip {

namespace {

constexpr int kUTX = 96;      // output bytes per tile row
constexpr int kUTY = 124;     // output rows per tile
constexpr int kUK = 128;      // window bytes = K of pass 1; rows of the window = N of pass 1 = K of pass 2
constexpr int kUHaloX = 16;   // window starts 16 bytes left of the first output byte (TMA needs 16-byte aligned boxes)
constexpr int kUThreads = 544;   // warps 0-2, 4-6: E1 (two per TMEM lane quarter), 3: TMA, 7: pass-1 issuer, 8-15: E2, 16: pass-2 issuer
constexpr int kUStages = 8;      // source windows in flight: 128 KiB per SM, what HBM latency x bandwidth asks for

constexpr int kWin = kUK * kUK;                    // 16 KiB: a source window, a weight matrix
constexpr int kOffImg = 0;                         // kUStages windows
constexpr int kOffA1 = kUStages * kWin;
constexpr int kOffA2 = kOffA1 + kWin;
constexpr int kOffT = kOffA2 + kWin;               // 2 x (96 rows x 128 B)
constexpr int kTBytes = kUTX * kUK;
constexpr int kOffOut = kOffT + 2 * kTBytes;       // 2 x (124 rows x 96 B), padded to 12 KiB
constexpr int kOutBytes = 12288;
constexpr int kOffBar = kOffOut + 2 * kOutBytes;   // the mbarriers and the TMEM base slot
constexpr int kUSmem = kOffBar + 512 + 1024;       // + the slack for the manual 1 KiB alignment

struct UmmaParams {
    ImgView dst;
    int rowbytes, width;
    int valid_y0, valid_y1, out_y0, out_y1, dst_y0;
    int tiles_x, tiles_y;
    int ntiles;              // < 2^31 (checked by the launcher): tile arithmetic stays 32-bit in the kernel
    bool wait_prev;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int x, int y, int z) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"(map), "r"(smem_u32(src)), "r"(x), "r"(y), "r"(z) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void named_bar_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

// byte k of row r of a 128-byte-row SWIZZLE_128B tile (16-byte chunk index XOR row & 7)
__device__ __forceinline__ uint8_t* sw128(uint8_t* tile, int r, int k) { return tile + r * kUK + ((((k >> 4) ^ (r & 7)) << 4) | (k & 15)); }

__device__ __forceinline__ uint64_t umma_desc_sw128(const void* smem) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_u32(smem) >> 4) & 0x3FFF);   // start address >> 4
    d |= (uint64_t)1 << 16;                              // leading byte offset: unused for swizzled K-major operands
    d |= (uint64_t)(1024 >> 4) << 32;                    // stride byte offset: one 8-row group
    d |= (uint64_t)1 << 46;                              // descriptor version
    d |= (uint64_t)2 << 61;                              // SWIZZLE_128B
    return d;
}
// s32 (D) = s8 (A, weights) x u8 (B, pixels), both K-major, M = 128
__host__ __device__ constexpr uint32_t umma_idesc(int n) { return (2u << 4) | (1u << 7) | (0u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24); }

__device__ __forceinline__ void umma_i8(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}\n"
        ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// 32 consecutive TMEM columns of this thread's lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
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
}

// 16 s32 sums scaled by 16 (<= 65280) -> 16 round-half-even bytes, in order.  Two sums are packed into the halves of a
// register and the rounding bias added with IMADs (the FMA pipe is otherwise idle here), then LOP3 + LEA.HI round both.
__constant__ uint32_t c_64k = 65536u;   // a multiplier ptxas cannot turn into a shift: the packing stays on the FMA pipe
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint4 round16(const uint32_t* v) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t a = rne_finish(v[4 * q + 1] * c_64k + padd(v[4 * q], 0x007F007Fu));       // results in the high byte of each half
        const uint32_t b = rne_finish(v[4 * q + 3] * c_64k + padd(v[4 * q + 2], 0x007F007Fu));
        w[q] = __byte_perm(a, b, 0x7531);
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// reflect-101 columns, then rows, of a swizzled window whose row 0 is source row ytop and byte 0 source byte x0 - 16
// Run by ONE warp (the MMA warp, before it issues pass 1 of the tile).
template <int C>
__device__ void patch_umma(uint8_t* tile, const UmmaParams& p, int x0, int ytop) {
    const int tid = threadIdx.x & 31;
    constexpr int kPatchThreads = 32;
    const bool left = x0 == 0;
    const bool right = x0 + kUTX + 2 * C > p.rowbytes;
    if (left || right) {
        constexpr int per_row = 4 * C;   // two pixels either side
        for (int i = tid; i < kUK * per_row; i += kPatchThreads) {
            const int r = i / per_row;
            int k = i - r * per_row;
            const int y = ytop + r;
            if (y < p.valid_y0 || y >= p.valid_y1) continue;
            const bool is_right = k >= 2 * C;
            if (is_right) k -= 2 * C;
            if (is_right ? !right : !left) continue;
            const int d = k / C, ch = k - d * C;
            const int px = is_right ? p.width + d : -1 - d;
            const int lb = px * C + ch - (x0 - kUHaloX);
            if (lb < 0 || lb >= kUK) continue;
            const int sb = reflect101(px, p.width) * C + ch - (x0 - kUHaloX);
            *sw128(tile, r, lb) = (sb >= 0 && sb < kUK) ? *sw128(tile, r, sb) : (uint8_t)0;
        }
        __syncwarp();
    }
    if (ytop < p.valid_y0 || ytop + kUK > p.valid_y1) {
        for (int i = tid; i < kUK * 8; i += kPatchThreads) {
            const int r = i >> 3, c = i & 7;
            const int y = ytop + r;
            if (y >= p.valid_y0 && y < p.valid_y1) continue;
            if (y < p.valid_y0 - 2 || y >= p.valid_y1 + 2) continue;   // never read by a stored output
            const int rs = p.valid_y0 + reflect101(y - p.valid_y0, p.valid_y1 - p.valid_y0) - ytop;
            if (rs < 0 || rs >= kUK) continue;
            *reinterpret_cast<uint4*>(tile + r * kUK + ((c ^ (r & 7)) << 4)) = *reinterpret_cast<const uint4*>(tile + rs * kUK + ((c ^ (rs & 7)) << 4));
        }
        __syncwarp();
    }
}

template <int C>
__global__ void __launch_bounds__(kUThreads, 1) gauss5sep_umma_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap dmap,
                                                                      const UmmaParams p) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    // SWIZZLE_128B operands want a 1024-byte aligned base: align by hand (the launch asks for 1 KiB more than it uses)
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* a1 = smem + kOffA1;
    uint8_t* a2 = smem + kOffA2;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
    uint64_t* full_img = bars;                      // [kUStages] TMA -> pass 1
    uint64_t* empty_img = bars + kUStages;          // [kUStages] pass 1 finished reading the window -> TMA
    uint64_t* d1_full = bars + 2 * kUStages;        // [2] pass 1 done -> E1
    uint64_t* d1_empty = d1_full + 2;               // [2] E1 has read D1 -> pass 1
    uint64_t* t_full = d1_full + 4;                 // [2] E1 has written T -> pass 2
    uint64_t* t_empty = d1_full + 6;                // [2] pass 2 finished reading T -> E1
    uint64_t* d2_full = d1_full + 8;                // [2] pass 2 done -> E2
    uint64_t* d2_empty = d1_full + 10;              // [2] E2 has read D2 -> pass 2
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(d1_full + 12);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int quarter = warp & 3;                    // the TMEM lane quarter this warp may read
    const int lane128 = quarter * 32 + lane;         // the TMEM lane (= matrix row) of this thread

    pdl_launch_dependents();
    // ---- the two banded weight matrices, K-major SWIZZLE_128B: threads 0..127 one row of A1, 128..255 one row of A2 ---
    if (tid < 256) {
        const int r = tid & 127;
        uint8_t* a = tid < 128 ? a1 : a2;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int c = 0; c < 8; ++c) *reinterpret_cast<uint4*>(a + r * kUK + (c << 4)) = z;
        constexpr int w16[5] = {16, 64, 96, 64, 16};   // 16 x [1 4 6 4 1]: D holds 16 S, whose high byte is S / 16
#pragma unroll
        for (int j = 0; j < 5; ++j) {
            if (tid < 128 && r < kUTX) *sw128(a1, r, kUHaloX + r + (j - 2) * C) = (uint8_t)w16[j];
            if (tid >= 128 && r < kUTY) *sw128(a2, r, r + j) = (uint8_t)w16[j];
        }
    }
    if (tid == 0) {
        for (int i = 0; i < kUStages; ++i) { mbar_init(&full_img[i], 1); mbar_init(&empty_img[i], 1); }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&d1_full[i], 1); mbar_init(&d1_empty[i], 6);
            mbar_init(&t_full[i], 6);  mbar_init(&t_empty[i], 1);
            mbar_init(&d2_full[i], 1); mbar_init(&d2_empty[i], 8);
        }
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    fence_async_smem();   // the weights, written by ordinary stores, are read by the tensor core's async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const uint32_t tlane = tmem + ((uint32_t)(quarter * 32) << 16);   // this warp's quarter of the 128 lanes
    // tensor memory: D1[b] at column 128 b, D2[b] at column 256 + 128 b

    struct Tile { int x0, y0, f; };
    auto tile_of = [&](int t) {
        const unsigned q = (unsigned)t / (unsigned)p.tiles_x;
        const unsigned tx = (unsigned)t - q * (unsigned)p.tiles_x;
        const unsigned f = q / (unsigned)p.tiles_y;
        return Tile{(int)tx * kUTX, p.out_y0 + (int)(q - f * (unsigned)p.tiles_y) * kUTY, (int)f};
    };
    const int t0 = blockIdx.x, stride = gridDim.x;
    const int my_tiles = t0 < p.ntiles ? (p.ntiles - t0 + stride - 1) / stride : 0;

    pdl_wait(p.wait_prev);

    if (warp == 3) {
        // ================================ TMA producer (one lane) =================================
        if (lane == 0) {
            for (int k = 0; k < my_tiles; ++k) {
                const int s = k % kUStages;
                mbar_wait(&empty_img[s], ((k / kUStages) & 1) ^ 1);
                const Tile tl = tile_of(t0 + k * stride);
                mbar_expect_tx(&full_img[s], kWin);
                tma_load_3d(smem + kOffImg + s * kWin, &tmap, &full_img[s], (tl.x0 - kUHaloX) / 4, tl.y0 - 2, tl.f);
            }
        }
    } else if (warp == 16) {
        // ================================ pass-2 issuer (one lane) ================================
        if (lane == 0) {
            const uint64_t desc_a2 = umma_desc_sw128(a2), desc_t = umma_desc_sw128(smem + kOffT);
            for (int j = 0; j < my_tiles; ++j) {
                const int b = j & 1, n = j >> 1;
                mbar_wait(&t_full[b], n & 1);
                mbar_wait(&d2_empty[b], (n & 1) ^ 1);
                tc_fence_after();
                const uint64_t db = desc_t + (uint64_t)(b * (kTBytes >> 4));
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) umma_i8(tmem + 256 + 128 * b, desc_a2 + 2 * kk, db + 2 * kk, umma_idesc(kUTX), kk > 0);
                umma_commit(&d2_full[b]);
                umma_commit(&t_empty[b]);
            }
        }
    } else if (warp == 7) {
        // ================================ pass-1 issuer (and halo patcher) =========================
        const uint64_t desc_a1 = umma_desc_sw128(a1), desc_img = umma_desc_sw128(smem + kOffImg);
        for (int i = 0; i < my_tiles; ++i) {
            const int s = i % kUStages, b = i & 1, n = i >> 1;
            const Tile tl = tile_of(t0 + i * stride);
            mbar_wait(&full_img[s], (i / kUStages) & 1);
            const int ytop = tl.y0 - 2;
            const bool border = tl.x0 == 0 || tl.x0 + kUTX + 2 * C > p.rowbytes || ytop < p.valid_y0 || ytop + kUK > p.valid_y1;
            if (border) {   // warp-uniform
                patch_umma<C>(smem + kOffImg + s * kWin, p, tl.x0, ytop);
                fence_async_smem();
                __syncwarp();
            }
            mbar_wait(&d1_empty[b], (n & 1) ^ 1);
            tc_fence_after();
            if (lane == 0) {
                const uint64_t db = desc_img + (uint64_t)(s * (kWin >> 4));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_i8(tmem + 128 * b, desc_a1 + 2 * k, db + 2 * k, umma_idesc(kUK), k > 0);
                umma_commit(&d1_full[b]);
                umma_commit(&empty_img[s]);
            }
            __syncwarp();
        }
    } else if (warp < 8) {
        // ================================ E1: round pass 1, write T ===================
        const int half = warp >> 2;   // which 64 of the 128 rows
        for (int i = 0; i < my_tiles; ++i) {
            const int b = i & 1, n = i >> 1;
            uint8_t* tb = smem + kOffT + b * kTBytes;
            mbar_wait(&d1_full[b], n & 1);
            tc_fence_after();
            uint32_t v0[32], v1[32];
            tmem_ld32(tlane + 128 * b + 64 * half, v0);
            tmem_ld32(tlane + 128 * b + 64 * half + 32, v1);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d1_empty[b]);   // the sums are in registers: pass 1 of tile i+2 may overwrite D1[b]
            const uint4 r0 = round16(v0), r1 = round16(v0 + 16), r2 = round16(v1), r3 = round16(v1 + 16);
            mbar_wait(&t_empty[b], (n & 1) ^ 1);        // pass 2 of tile i-2 has finished reading T[b]
            const int ch = 4 * half, sw = lane128 & 7;
            *reinterpret_cast<uint4*>(tb + lane128 * kUK + (((ch + 0) ^ sw) << 4)) = r0;
            *reinterpret_cast<uint4*>(tb + lane128 * kUK + (((ch + 1) ^ sw) << 4)) = r1;
            *reinterpret_cast<uint4*>(tb + lane128 * kUK + (((ch + 2) ^ sw) << 4)) = r2;
            *reinterpret_cast<uint4*>(tb + lane128 * kUK + (((ch + 3) ^ sw) << 4)) = r3;
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&t_full[b]);
        }
    } else {
        // ================================ E2: round pass 2, store the tile ============
        const int half = (warp >> 2) & 1;   // which 48 of the 96 columns
        const int cbase = 48 * half;
        for (int i = 0; i < my_tiles; ++i) {
            const int b = i & 1, n = i >> 1;
            const Tile tl = tile_of(t0 + i * stride);
            uint8_t* ob = smem + kOffOut + b * kOutBytes;
            mbar_wait(&d2_full[b], n & 1);
            tc_fence_after();
            uint32_t v[32], u[16];
            tmem_ld32(tlane + 256 + 128 * b + cbase + (half ? 16 : 0), v);
            tmem_ld16(tlane + 256 + 128 * b + cbase + (half ? 0 : 32), u);
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&d2_empty[b]);
            const uint4 r0 = round16(v), r1 = round16(v + 16), r2 = round16(u);
#ifndef DIP_UMMA_DIRECT_STORES   // (direct per-thread stores of 3 x 16 B to 128 different rows measured 21 us per frame against 15.9)
            named_bar_sync(1, 256);   // the store that last read this staging buffer (tile i-2) has finished reading it
            if (lane128 < kUTY) {
                uint8_t* orow = ob + lane128 * kUTX;
                const int xv = cbase + (half ? 16 : 0), xu = cbase + (half ? 0 : 32);
                *reinterpret_cast<uint4*>(orow + xv) = r0;
                *reinterpret_cast<uint4*>(orow + xv + 16) = r1;
                *reinterpret_cast<uint4*>(orow + xu) = r2;
            }
            fence_async_smem();
            named_bar_sync(2, 256);
            if (tid == 256) {
                tma_store_3d(&dmap, ob, tl.x0 / 4, p.dst_y0 + (tl.y0 - p.out_y0), tl.f);
                tma_store_wait_read<1>();   // the PREVIOUS tile's store has read its buffer: that one is written next
            }
#else
            // every thread stores its own 48 bytes of output row y (three 16-byte vectors; L2 merges the half sectors)
            (void)ob;
            const int y = tl.y0 + lane128;
            if (lane128 < kUTY && y < p.out_y1) {
                uint8_t* drow = p.dst.base + (size_t)tl.f * p.dst.frame_stride + (size_t)(p.dst_y0 + (y - p.out_y0)) * p.dst.pitch + tl.x0;
                const int row_padded = (p.rowbytes + 15) & ~15;
                const int xv = cbase + (half ? 16 : 0), xu = cbase + (half ? 0 : 32);
                if (tl.x0 + xv < row_padded) __stcs(reinterpret_cast<uint4*>(drow + xv), r0);
                if (tl.x0 + xv + 16 < row_padded) __stcs(reinterpret_cast<uint4*>(drow + xv + 16), r1);
                if (tl.x0 + xu < row_padded) __stcs(reinterpret_cast<uint4*>(drow + xu), r2);
            }
#endif
        }
        if (tid == 256) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

template <int C>
int launch_umma_c(dip_ctx* ctx, const dip_image* src, dip_image* dst, const RowGeom& g) {
    const CUtensorMap* map = nullptr;
    int rc = get_tmap(ctx, const_cast<dip_image*>(src), kUK / 4, kUK, &map, /*swizzle128=*/true);
    if (rc != DIP_OK) return rc;
    const CUtensorMap smap = *map;   // (the next get_tmap may move the cache)
    // destination: the rows up to the last one this launch produces, so that the TMA store clips the bottom tile there
    dip_image dview;
    dview.v = dst->v;
    dview.v.height = g.dst_y0 + (g.out_y1 - g.out_y0);
    dview.parent = root_of(dst);
    if ((rc = get_tmap(ctx, &dview, kUTX / 4, kUTY, &map)) != DIP_OK) return rc;
    UmmaParams p{};
    p.dst = dst->v;
    p.rowbytes = src->v.rowbytes;
    p.width = src->v.width;
    p.valid_y0 = g.valid_y0; p.valid_y1 = g.valid_y1;
    p.out_y0 = g.out_y0; p.out_y1 = g.out_y1; p.dst_y0 = g.dst_y0;
    p.tiles_x = (src->v.rowbytes + kUTX - 1) / kUTX;
    p.tiles_y = (g.out_y1 - g.out_y0 + kUTY - 1) / kUTY;
    const long long ntiles = (long long)p.tiles_x * p.tiles_y * src->v.frames;
    if (ntiles <= 0) return DIP_OK;
    if (ntiles > 0x7FFFFFFFll) return fail(ctx, DIP_ERR_UNSUPPORTED, "image stack too large for one launch");
    p.ntiles = (int)ntiles;
    static std::atomic<uint64_t> attr_set{0};
    const uint64_t dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_set.load(std::memory_order_relaxed) & dev_bit)) {
        DIP_CUDA(ctx, cudaFuncSetAttribute(gauss5sep_umma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, kUSmem));
        attr_set.fetch_or(dev_bit, std::memory_order_relaxed);
    }
    const LaunchMode mode = plan_launch(ctx, src->v, dst->v);
    p.wait_prev = mode != kNoWait;
    const unsigned grid = (unsigned)std::min(p.ntiles, ctx->sm_count);
    DIP_CUDA(ctx, launch_kernel(mode, gauss5sep_umma_kernel<C>, dim3(grid), dim3(kUThreads), kUSmem, ctx->compute, smap, *map, p));
    DIP_CHECK_LAUNCH(ctx, "gauss5sep_umma_kernel");
    return DIP_OK;
}

}  // namespace

// Can the tensor-core kernel take this launch?  Library-owned rows (16-byte padded, TMA-aligned), at least a few rows
// and pixels so that reflected border sources lie inside the window, no neighbour-GPU rows.
bool umma_gauss5sep_ok(const dip_image* src, const dip_image* dst, const RowGeom& g) {
    const ImgView& v = src->v;
    if (g.peer_top || g.peer_bot) return false;
    if (exact_rows(dst->v) || (reinterpret_cast<uintptr_t>(dst->v.base) & 15) || (dst->v.pitch & 15) || (dst->v.frame_stride & 15)) return false;
    if ((reinterpret_cast<uintptr_t>(v.base) & 15) || (v.pitch & 15) || (v.frame_stride & 15)) return false;
    if (v.width < 8 || g.valid_y1 - g.valid_y0 < 8) return false;
    return true;
}

int launch_gauss5sep_umma(dip_ctx* ctx, const dip_image* src, dip_image* dst, const RowGeom& g) {
    switch (src->v.channels) {
        case 1: return launch_umma_c<1>(ctx, src, dst, g);
        case 3: return launch_umma_c<3>(ctx, src, dst, g);
        case 4: return launch_umma_c<4>(ctx, src, dst, g);
    }
    return fail(ctx, DIP_ERR_ARG, "bad channel count");
}

}  // namespace dip
