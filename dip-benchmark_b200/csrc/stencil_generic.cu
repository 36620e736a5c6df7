// Tiled kernels for ARBITRARY small masks / kernels (up to 7 x 7): SURVEY.md 8f rank 3 -- the reference's operations
// take the mask / kernel as a runtime argument (sycl/benchmark.cpp:123,166, visiongl:125-131) even though the benchmark
// passes only seven fixed ones, which have kernels of their own (stencil.cu).
//
// Same frame as those: one TMA tile (512 bytes + 16-byte halo, 32 rows + kh/2 halo rows) per CTA, borders patched in
// shared memory, a thread owns a 16-byte column of 8 rows.  The taps travel by value in the kernel parameters (the
// constant bank), not through global memory as the reference's do.  The arithmetic is generic: for every tap row the
// thread converts the 16 + 2*RX*C bytes it needs to fp32 once (PRMT into the mantissa of 2^23, one FADD) and runs
// (2*RX+1) x 16 FFMAs (convolution: fp32 accumulation in row-major tap order, cvRound, saturate -- bit-exact for
// dyadic kernels, within 1 LSB otherwise, like the direct kernel in stencil.cu) or FMNMXs (erosion / dilation; the
// bytes are exact in fp32).  Larger masks keep the direct kernel.
// BPS = 2 instantiates the same kernel for 16-bit samples (depth16.cu): a thread's 16 bytes are 8 samples, the pixel
// stride in the tile is C * 2 bytes, results saturate at 65535, and mode 3 rounds half up (GaussianBlur).
#include "stencil_tiled.cuh"

namespace dip {

namespace {

constexpr int kGenWarps = 4, kGenRows = 8, kGenTH = kGenWarps * kGenRows;

// reflect-101 / constant border patch with runtime radii (the tiled kernels' patch_borders has them as template constants)
template <int C, int kThreads>
__device__ void patch_borders_rt(uint8_t* tile, const StencilParams& p, int bx0, int ytop, int rx, int ry, int rows, bool morph, uint8_t fill) {
    const int tid = threadIdx.x;
    const bool left = rx > 0 && bx0 <= 0;
    const bool right = rx > 0 && bx0 + kTileBytes + rx * C > p.rowbytes;
    if (left || right) {
        const int per_row = 2 * rx * C;
        for (int i = tid; i < rows * per_row; i += kThreads) {
            const int t = i / per_row;
            int k = i - t * per_row;
            const int y = ytop + t;
            if (y < p.valid_y0 || y >= p.valid_y1) continue;
            const bool is_right = k >= rx * C;
            if (is_right) k -= rx * C;
            const int d = k / C, ch = k - d * C;
            if (is_right ? !right : !left) continue;
            const int px = is_right ? p.width + d : -1 - d;
            const int lb = px * C + ch - (bx0 - kHalo);
            if (lb < 0 || lb >= kRowBytes) continue;
            uint8_t v = fill;
            if (!morph) {
                const int sb = reflect101(px, p.width) * C + ch - (bx0 - kHalo);
                v = (sb >= 0 && sb < kRowBytes) ? tile[t * kRowBytes + sb] : 0;  // a reflected source outside this tile: rows narrower than the radius
            }
            tile[t * kRowBytes + lb] = v;
        }
        __syncthreads();
    }
    const bool top = ry > 0 && ytop < p.valid_y0;
    const bool bottom = ry > 0 && ytop + rows > p.valid_y1;
    if (top || bottom) {
        constexpr int vec_per_row = kRowBytes / 16;
        for (int i = tid; i < rows * vec_per_row; i += kThreads) {
            const int t = i / vec_per_row, vx = i - t * vec_per_row;
            const int y = ytop + t;
            if (y >= p.valid_y0 && y < p.valid_y1) continue;
            if (y < p.valid_y0 - ry || y >= p.valid_y1 + ry) continue;
            uint4 v = make_uint4(fill * 0x01010101u, fill * 0x01010101u, fill * 0x01010101u, fill * 0x01010101u);
            if (!morph) {
                const int ts = p.valid_y0 + reflect101(y - p.valid_y0, p.valid_y1 - p.valid_y0) - ytop;
                if (ts < 0 || ts >= rows) continue;
                v = *reinterpret_cast<const uint4*>(tile + ts * kRowBytes + vx * 16);
            }
            *reinterpret_cast<uint4*>(tile + t * kRowBytes + vx * 16) = v;
        }
        __syncthreads();
    }
}

// byte `B` of the 12-word window (word 4 = the thread's first own word) as fp32
template <int B>
__device__ __forceinline__ float win_byte(const uint32_t (&w)[12]) {
    static_assert(B >= 0 && B < 48, "window");
    return __uint_as_float(__byte_perm(w[B / 4], 0x4B000000u, 0x7650 + (B & 3))) - 8388608.0f;
}

// 16-bit sample starting at byte `B` (even) of the window as fp32
template <int B>
__device__ __forceinline__ float win_half(const uint32_t (&w)[12]) {
    static_assert(B >= 0 && B < 48 && (B & 1) == 0, "window");
    return __uint_as_float(__byte_perm(w[B / 4], 0x4B000000u, (B & 2) ? 0x7632 : 0x7610)) - 8388608.0f;
}

template <int MODE, int C, int RX, int BPS, int N, int I = 0>
struct RowTaps {
    // acc[b] (op)= wt[k] * f[b + k*C] for every tap column k and own sample b; f[q] = the sample at window byte
    // 16 - RX*C*BPS + q*BPS
    static __device__ __forceinline__ void fill(const uint32_t (&w)[12], float (&f)[N]) {
        if constexpr (I < N) {
            if constexpr (BPS == 1) f[I] = win_byte<16 - RX * C + I>(w);
            else f[I] = win_half<16 - RX * C * 2 + 2 * I>(w);
            RowTaps<MODE, C, RX, BPS, N, I + 1>::fill(w, f);
        }
    }
};

template <int MODE, int C, int RX, int BPS = 1>  // MODE 0 erode, 1 dilate, 2 correlation (RNE), 3 correlation (round half up)
__global__ void __launch_bounds__(kGenWarps * 32) generic_tiled_kernel(const __grid_constant__ CUtensorMap tmap, const StencilParams p, const SmallTaps taps) {
    constexpr int kThreads = kGenWarps * 32;
    constexpr int CB = C * BPS;              // bytes per pixel
    constexpr int NOUT = 16 / BPS;           // samples a thread produces per row
    constexpr int N = NOUT + 2 * RX * C;     // samples of a tap row one thread reads
    constexpr float kTop = BPS == 1 ? 255.0f : 65535.0f;
    static_assert(RX * CB <= 16, "horizontal reach beyond the 16-byte halo");
    extern __shared__ __align__(128) uint8_t smem[];
    const int ry = taps.kh / 2;
    const int rows = kGenTH + 2 * ry;
    uint8_t* tile = smem;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + rows * kRowBytes);
    const int tx = blockIdx.x, ty = blockIdx.y, f = blockIdx.z;
    const int bx0 = tx * kTileBytes;
    const int y0 = p.out_y0 + ty * kGenTH;
    const int ytop = y0 - ry;

    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        mbar_expect_tx(bar, rows * kRowBytes);
    }
    pdl_wait(p.wait_prev);
    if (threadIdx.x == 0) tma_load_3d(tile, &tmap, bar, (bx0 - kHalo) / 4, ytop, f);
    __syncthreads();
    mbar_wait(bar, 0);
    const bool border = bx0 == 0 || bx0 + kTileBytes + RX * CB > p.rowbytes || ytop < p.valid_y0 || ytop + rows > p.valid_y1;
    if (border) {
        __syncthreads();
        patch_borders_rt<CB, kThreads>(tile, p, bx0, ytop, RX, ry, rows, MODE < 2, MODE == 1 ? 0 : 255);   // per byte: 0xFFFF for 16-bit erosion
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cx = bx0 + 16 * lane;
    const int yw = y0 + warp * kGenRows;
    if (cx >= ((p.rowbytes + 15) & ~15) || yw >= p.out_y1) return;
    const int nrows = min(kGenRows, p.out_y1 - yw);
    const uint8_t* colp = tile + (warp * kGenRows) * kRowBytes + kHalo + 16 * lane;  // tile row of source row yw - ry
    uint8_t* dptr = p.dst.base + (size_t)f * p.dst.frame_stride + (size_t)(p.dst_y0 + (yw - p.out_y0)) * p.dst.pitch + cx;
    const int valid = p.dst_exact ? min(16, p.rowbytes - cx) : 16;

    // kM output rows per pass share their source rows (kM + kh - 1 conversions to fp32 for kM results instead of kM * kh);
    // every result still accumulates its taps in row-major order, so the fp32 sums are what one row at a time gives
    constexpr int kM = 32 / NOUT;   // 2 rows of 16 bytes, 4 rows of 8 16-bit samples: 32 accumulators
    static_assert(kGenRows % kM == 0, "rows per warp");
    for (int i = 0; i < nrows; i += kM) {
        float acc[kM][NOUT];
#pragma unroll
        for (int m = 0; m < kM; ++m)
#pragma unroll
            for (int b = 0; b < NOUT; ++b) acc[m][b] = MODE == 0 ? kTop : 0.0f;
        for (int t = 0; t < taps.kh + kM - 1; ++t) {   // tile row i + t; rows past the band's end exist in the tile and feed no stored result
            const uint8_t* rp = colp + (i + t) * kRowBytes;
            const uint4 wa = *reinterpret_cast<const uint4*>(rp - 16);
            const uint4 wb = *reinterpret_cast<const uint4*>(rp);
            const uint4 wc = *reinterpret_cast<const uint4*>(rp + 16);
            const uint32_t w[12] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x, wc.y, wc.z, wc.w};
            float fw[N];
            RowTaps<MODE, C, RX, BPS, N>::fill(w, fw);
#pragma unroll
            for (int m = 0; m < kM; ++m) {
                const int dy = t - m;
                if (dy < 0 || dy >= taps.kh) continue;   // uniform over the launch
#pragma unroll
                for (int k = 0; k <= 2 * RX; ++k) {
                    const float wt = taps.w[dy][k];
                    if (MODE >= 2) {
#pragma unroll
                        for (int b = 0; b < NOUT; ++b) acc[m][b] = __fmaf_rn(wt, fw[b + k * C], acc[m][b]);
                    } else if (wt != 0.0f) {   // uniform: is this position of the mask set?
#pragma unroll
                        for (int b = 0; b < NOUT; ++b) acc[m][b] = MODE == 0 ? fminf(acc[m][b], fw[b + k * C]) : fmaxf(acc[m][b], fw[b + k * C]);
                    }
                }
            }
        }
        constexpr int kPerWord = 4 / BPS;
#pragma unroll
        for (int m = 0; m < kM; ++m) {
            if (i + m >= nrows) break;
            uint32_t o[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                uint32_t word = 0;
#pragma unroll
                for (int b = 0; b < kPerWord; ++b) {
                    const float a = acc[m][kPerWord * j + b];
                    // cvRound (or round half up) + saturate_cast
                    const int r = MODE == 2 ? min(max(__float2int_rn(a), 0), (int)kTop) : MODE == 3 ? min(max(__float2int_rd(a + 0.5f), 0), (int)kTop) : (int)a;
                    word |= (uint32_t)r << (8 * BPS * b);
                }
                o[j] = word;
            }
            store16(dptr + (size_t)(i + m) * p.dst.pitch, make_uint4(o[0], o[1], o[2], o[3]), valid);
        }
    }
}

// ---- arbitrary masks in packed integer arithmetic (8-bit samples) ---------------------------------------------------
// Erosion / dilation need no arithmetic but minima, so an arbitrary mask can run on the packed u16x2 lanes of the
// dedicated kernels instead of fp32: a row is unpacked once per (output row, mask row) into even / odd byte streams
// (load_row), the tap at column k is the compile-time selection pair_at<4j + k*C> of neighbouring packed words, and only
// WHETHER it takes part is a run-time (launch-uniform) test of the mask.  ~150 instructions per 16 output bytes for a full
// 3x3 mask against ~330 in fp32.  Reach of the 8-word window: RX * C <= 8 bytes.
template <int C, int K, bool MAX>
__device__ __forceinline__ void morph_tap(const Row8& r, Acc8& a) {
#define DIP_MT(j)                                                                                          \
    a.v[2 * j] = MAX ? __vmaxu2(a.v[2 * j], pair_at<4 * j + K * C>(r)) : __vminu2(a.v[2 * j], pair_at<4 * j + K * C>(r));          \
    a.v[2 * j + 1] = MAX ? __vmaxu2(a.v[2 * j + 1], pair_at<4 * j + 1 + K * C>(r)) : __vminu2(a.v[2 * j + 1], pair_at<4 * j + 1 + K * C>(r));
    DIP_MT(0) DIP_MT(1) DIP_MT(2) DIP_MT(3)
#undef DIP_MT
}
template <int C, int RX, bool MAX, int K = -RX>
__device__ __forceinline__ void morph_row(const Row8& r, Acc8& a, const float* wrow) {
    if constexpr (K <= RX) {
        if (wrow[K + RX] != 0.0f) morph_tap<C, K, MAX>(r, a);   // uniform over the launch
        morph_row<C, RX, MAX, K + 1>(r, a, wrow);
    }
}

template <int C, int RX, bool MAX>
__global__ void __launch_bounds__(kGenWarps * 32) generic_morph_kernel(const __grid_constant__ CUtensorMap tmap, const StencilParams p, const SmallTaps taps) {
    constexpr int kThreads = kGenWarps * 32;
    static_assert(RX * C <= 8, "tap outside the 8-word window");
    extern __shared__ __align__(128) uint8_t smem[];
    const int ry = taps.kh / 2;
    const int rows = kGenTH + 2 * ry;
    uint8_t* tile = smem;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + rows * kRowBytes);
    const int tx = blockIdx.x, ty = blockIdx.y, f = blockIdx.z;
    const int bx0 = tx * kTileBytes;
    const int y0 = p.out_y0 + ty * kGenTH;
    const int ytop = y0 - ry;

    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        mbar_expect_tx(bar, rows * kRowBytes);
    }
    pdl_wait(p.wait_prev);
    if (threadIdx.x == 0) tma_load_3d(tile, &tmap, bar, (bx0 - kHalo) / 4, ytop, f);
    __syncthreads();
    mbar_wait(bar, 0);
    const bool border = bx0 == 0 || bx0 + kTileBytes + RX * C > p.rowbytes || ytop < p.valid_y0 || ytop + rows > p.valid_y1;
    if (border) {
        __syncthreads();
        patch_borders_rt<C, kThreads>(tile, p, bx0, ytop, RX, ry, rows, true, MAX ? 0 : 255);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cx = bx0 + 16 * lane;
    const int yw = y0 + warp * kGenRows;
    if (cx >= ((p.rowbytes + 15) & ~15) || yw >= p.out_y1) return;
    const int nrows = min(kGenRows, p.out_y1 - yw);
    const uint8_t* colp = tile + (warp * kGenRows) * kRowBytes + kHalo + 16 * lane;
    uint8_t* dptr = p.dst.base + (size_t)f * p.dst.frame_stride + (size_t)(p.dst_y0 + (yw - p.out_y0)) * p.dst.pitch + cx;
    const int valid = p.dst_exact ? min(16, p.rowbytes - cx) : 16;
    // kM output rows per pass share their source rows: kM + kh - 1 rows are unpacked for kM results instead of kM * kh
    constexpr int kM = 4;
    static_assert(kGenRows % kM == 0, "rows per warp");
    for (int i = 0; i < nrows; i += kM) {
        Acc8 a[kM];
#pragma unroll
        for (int m = 0; m < kM; ++m)
#pragma unroll
            for (int k = 0; k < 8; ++k) a[m].v[k] = MAX ? 0u : 0x00FF00FFu;
        for (int t = 0; t < taps.kh + kM - 1; ++t) {   // tile row i + t; rows past the band's end exist in the tile and feed no stored result
            Row8 row;
            load_row(colp + (i + t) * kRowBytes, row);
#pragma unroll
            for (int m = 0; m < kM; ++m) {
                const int dy = t - m;
                if (dy >= 0 && dy < taps.kh) morph_row<C, RX, MAX>(row, a[m], taps.w[dy]);
            }
        }
#pragma unroll
        for (int m = 0; m < kM; ++m)
            if (i + m < nrows)
                store16(dptr + (size_t)(i + m) * p.dst.pitch,
                        make_uint4(pack_lo(a[m].v[0], a[m].v[1]), pack_lo(a[m].v[2], a[m].v[3]), pack_lo(a[m].v[4], a[m].v[5]), pack_lo(a[m].v[6], a[m].v[7])), valid);
    }
}

template <int MODE, int C, int RX, int BPS = 1>
int launch_generic_tiled_t(dip_ctx* ctx, const dip_image* src, dip_image* dst, const RowGeom& g, const SmallTaps& taps) {
    const int rows = kGenTH + 2 * (taps.kh / 2);
    const int smem_bytes = rows * kRowBytes + 16;
    const CUtensorMap* map = nullptr;
    int rc = get_tmap(ctx, const_cast<dip_image*>(src), kBoxWords, rows, &map);
    if (rc != DIP_OK) return rc;
    StencilParams p{};
    p.dst = dst->v;
    p.rowbytes = src->v.rowbytes;
    p.width = src->v.width;
    p.valid_y0 = g.valid_y0; p.valid_y1 = g.valid_y1;
    p.out_y0 = g.out_y0; p.out_y1 = g.out_y1; p.dst_y0 = g.dst_y0;
    p.tiles_x = (src->v.rowbytes + kTileBytes - 1) / kTileBytes;
    p.tiles_y = (g.out_y1 - g.out_y0 + kGenTH - 1) / kGenTH;
    p.dst_exact = exact_rows(dst->v);
    if (p.tiles_x <= 0 || p.tiles_y <= 0) return DIP_OK;
    if (p.tiles_y > 65535 || src->v.frames > 65535) return fail(ctx, DIP_ERR_UNSUPPORTED, "image stack too large for one launch");
    const LaunchMode mode = plan_tiled_launch(ctx, src, dst, g);
    p.wait_prev = mode != kNoWait;
    DIP_CUDA(ctx, launch_kernel(mode, generic_tiled_kernel<MODE, C, RX, BPS>, dim3((unsigned)p.tiles_x, (unsigned)p.tiles_y, (unsigned)src->v.frames),
                                dim3(kGenWarps * 32), smem_bytes, ctx->compute, *map, p, taps));
    DIP_CHECK_LAUNCH(ctx, "generic_tiled_kernel");
    return DIP_OK;
}

// The same for 16-bit samples: two neighbouring samples ARE a packed u16x2 word, so there is nothing to unpack -- a tap at
// q samples is a word of the row (q even) or one funnel of two (q odd).  The thread's 16 bytes are 8 samples = 4 words;
// the 12-word window (3 x 16 bytes) reaches 8 samples either side: RX * C <= 8.
template <int Q>   // the word holding samples (Q, Q + 1) relative to the thread's first own sample (word 4 of the window)
__device__ __forceinline__ uint32_t pair16_at(const uint32_t (&w)[12]) {
    static_assert(Q >= -8 && Q + 1 < 16, "tap outside the 12-word window");
    constexpr int q = Q + 8;
    if constexpr ((q & 1) == 0) return w[q / 2];
    else return __byte_perm(w[(q - 1) / 2], w[(q + 1) / 2], 0x5432);
}
template <int C, int K, bool MAX>
__device__ __forceinline__ void morph16_tap(const uint32_t (&w)[12], uint32_t (&a)[4]) {
#define DIP_MT16(j) a[j] = MAX ? __vmaxu2(a[j], pair16_at<2 * j + K * C>(w)) : __vminu2(a[j], pair16_at<2 * j + K * C>(w));
    DIP_MT16(0) DIP_MT16(1) DIP_MT16(2) DIP_MT16(3)
#undef DIP_MT16
}
template <int C, int RX, bool MAX, int K = -RX>
__device__ __forceinline__ void morph16_row(const uint32_t (&w)[12], uint32_t (&a)[4], const float* wrow) {
    if constexpr (K <= RX) {
        if (wrow[K + RX] != 0.0f) morph16_tap<C, K, MAX>(w, a);   // uniform over the launch
        morph16_row<C, RX, MAX, K + 1>(w, a, wrow);
    }
}

template <int C, int RX, bool MAX>
__global__ void __launch_bounds__(kGenWarps * 32) generic_morph16_kernel(const __grid_constant__ CUtensorMap tmap, const StencilParams p, const SmallTaps taps) {
    constexpr int kThreads = kGenWarps * 32;
    constexpr int CB = 2 * C;
    static_assert(RX * C <= 8, "tap outside the 12-word window");
    extern __shared__ __align__(128) uint8_t smem[];
    const int ry = taps.kh / 2;
    const int rows = kGenTH + 2 * ry;
    uint8_t* tile = smem;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + rows * kRowBytes);
    const int tx = blockIdx.x, ty = blockIdx.y, f = blockIdx.z;
    const int bx0 = tx * kTileBytes;
    const int y0 = p.out_y0 + ty * kGenTH;
    const int ytop = y0 - ry;

    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        mbar_expect_tx(bar, rows * kRowBytes);
    }
    pdl_wait(p.wait_prev);
    if (threadIdx.x == 0) tma_load_3d(tile, &tmap, bar, (bx0 - kHalo) / 4, ytop, f);
    __syncthreads();
    mbar_wait(bar, 0);
    const bool border = bx0 == 0 || bx0 + kTileBytes + RX * CB > p.rowbytes || ytop < p.valid_y0 || ytop + rows > p.valid_y1;
    if (border) {
        __syncthreads();
        patch_borders_rt<CB, kThreads>(tile, p, bx0, ytop, RX, ry, rows, true, MAX ? 0 : 255);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cx = bx0 + 16 * lane;
    const int yw = y0 + warp * kGenRows;
    if (cx >= ((p.rowbytes + 15) & ~15) || yw >= p.out_y1) return;
    const int nrows = min(kGenRows, p.out_y1 - yw);
    const uint8_t* colp = tile + (warp * kGenRows) * kRowBytes + kHalo + 16 * lane;
    uint8_t* dptr = p.dst.base + (size_t)f * p.dst.frame_stride + (size_t)(p.dst_y0 + (yw - p.out_y0)) * p.dst.pitch + cx;
    const int valid = p.dst_exact ? min(16, p.rowbytes - cx) : 16;
    constexpr int kM = 4;
    for (int i = 0; i < nrows; i += kM) {
        uint32_t a[kM][4];
#pragma unroll
        for (int m = 0; m < kM; ++m)
#pragma unroll
            for (int k = 0; k < 4; ++k) a[m][k] = MAX ? 0u : 0xFFFFFFFFu;
        for (int t = 0; t < taps.kh + kM - 1; ++t) {
            const uint8_t* rp = colp + (i + t) * kRowBytes;
            const uint4 wa = *reinterpret_cast<const uint4*>(rp - 16);
            const uint4 wb = *reinterpret_cast<const uint4*>(rp);
            const uint4 wc = *reinterpret_cast<const uint4*>(rp + 16);
            const uint32_t w[12] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w, wc.x, wc.y, wc.z, wc.w};
#pragma unroll
            for (int m = 0; m < kM; ++m) {
                const int dy = t - m;
                if (dy >= 0 && dy < taps.kh) morph16_row<C, RX, MAX>(w, a[m], taps.w[dy]);
            }
        }
#pragma unroll
        for (int m = 0; m < kM; ++m)
            if (i + m < nrows) store16(dptr + (size_t)(i + m) * p.dst.pitch, make_uint4(a[m][0], a[m][1], a[m][2], a[m][3]), valid);
    }
}

template <int C, int RX, bool MAX, int BPS = 1>
int launch_generic_morph_t(dip_ctx* ctx, const dip_image* src, dip_image* dst, const RowGeom& g, const SmallTaps& taps) {
    const int rows = kGenTH + 2 * (taps.kh / 2);
    const int smem_bytes = rows * kRowBytes + 16;
    const CUtensorMap* map = nullptr;
    int rc = get_tmap(ctx, const_cast<dip_image*>(src), kBoxWords, rows, &map);
    if (rc != DIP_OK) return rc;
    StencilParams p{};
    p.dst = dst->v;
    p.rowbytes = src->v.rowbytes;
    p.width = src->v.width;
    p.valid_y0 = g.valid_y0; p.valid_y1 = g.valid_y1;
    p.out_y0 = g.out_y0; p.out_y1 = g.out_y1; p.dst_y0 = g.dst_y0;
    p.tiles_x = (src->v.rowbytes + kTileBytes - 1) / kTileBytes;
    p.tiles_y = (g.out_y1 - g.out_y0 + kGenTH - 1) / kGenTH;
    p.dst_exact = exact_rows(dst->v);
    if (p.tiles_x <= 0 || p.tiles_y <= 0) return DIP_OK;
    if (p.tiles_y > 65535 || src->v.frames > 65535) return fail(ctx, DIP_ERR_UNSUPPORTED, "image stack too large for one launch");
    const LaunchMode mode = plan_tiled_launch(ctx, src, dst, g);
    p.wait_prev = mode != kNoWait;
    if constexpr (BPS == 1)
        DIP_CUDA(ctx, launch_kernel(mode, generic_morph_kernel<C, RX, MAX>, dim3((unsigned)p.tiles_x, (unsigned)p.tiles_y, (unsigned)src->v.frames),
                                    dim3(kGenWarps * 32), smem_bytes, ctx->compute, *map, p, taps));
    else
        DIP_CUDA(ctx, launch_kernel(mode, generic_morph16_kernel<C, RX, MAX>, dim3((unsigned)p.tiles_x, (unsigned)p.tiles_y, (unsigned)src->v.frames),
                                    dim3(kGenWarps * 32), smem_bytes, ctx->compute, *map, p, taps));
    DIP_CHECK_LAUNCH(ctx, "generic_morph_kernel");
    return DIP_OK;
}

// the packed-integer form where the 8-word window reaches (RX * C <= 8), else -1: the caller takes the fp32 kernel
template <bool MAX>
int launch_generic_morph(dip_ctx* ctx, int rx, const dip_image* src, dip_image* dst, const RowGeom& g, const SmallTaps& taps) {
    const int c = src->v.channels;
    if (c == 1 && rx == 1) return launch_generic_morph_t<1, 1, MAX>(ctx, src, dst, g, taps);
    if (c == 1 && rx == 2) return launch_generic_morph_t<1, 2, MAX>(ctx, src, dst, g, taps);
    if (c == 1 && rx == 3) return launch_generic_morph_t<1, 3, MAX>(ctx, src, dst, g, taps);
    if (c == 3 && rx == 1) return launch_generic_morph_t<3, 1, MAX>(ctx, src, dst, g, taps);
    if (c == 3 && rx == 2) return launch_generic_morph_t<3, 2, MAX>(ctx, src, dst, g, taps);
    if (c == 4 && rx == 1) return launch_generic_morph_t<4, 1, MAX>(ctx, src, dst, g, taps);
    if (c == 4 && rx == 2) return launch_generic_morph_t<4, 2, MAX>(ctx, src, dst, g, taps);
    return -1;
}
template <bool MAX>
int launch_generic_morph16(dip_ctx* ctx, int rx, const dip_image* src, dip_image* dst, const RowGeom& g, const SmallTaps& taps) {
    const int c = src->v.channels;
    if (c == 1 && rx == 1) return launch_generic_morph_t<1, 1, MAX, 2>(ctx, src, dst, g, taps);
    if (c == 1 && rx == 2) return launch_generic_morph_t<1, 2, MAX, 2>(ctx, src, dst, g, taps);
    if (c == 1 && rx == 3) return launch_generic_morph_t<1, 3, MAX, 2>(ctx, src, dst, g, taps);
    if (c == 3 && rx == 1) return launch_generic_morph_t<3, 1, MAX, 2>(ctx, src, dst, g, taps);
    if (c == 3 && rx == 2) return launch_generic_morph_t<3, 2, MAX, 2>(ctx, src, dst, g, taps);
    if (c == 4 && rx == 1) return launch_generic_morph_t<4, 1, MAX, 2>(ctx, src, dst, g, taps);
    if (c == 4 && rx == 2) return launch_generic_morph_t<4, 2, MAX, 2>(ctx, src, dst, g, taps);
    return -1;
}

template <int MODE, int C>
int launch_generic_tiled_c(dip_ctx* ctx, int rx, const dip_image* src, dip_image* dst, const RowGeom& g, const SmallTaps& taps) {
    switch (rx) {
        case 1: return launch_generic_tiled_t<MODE, C, 1>(ctx, src, dst, g, taps);
        case 2: return launch_generic_tiled_t<MODE, C, 2>(ctx, src, dst, g, taps);
        case 3: return launch_generic_tiled_t<MODE, C, 3>(ctx, src, dst, g, taps);
    }
    return fail(ctx, DIP_ERR_ARG, "bad horizontal radius %d", rx);
}

// 16-bit samples: the 16-byte halo reaches 2 pixels of 3 or 4 channels, 3 pixels of one channel (generic_tiled_reach)
template <int MODE>
int launch_generic_tiled_16(dip_ctx* ctx, int rx, const dip_image* src, dip_image* dst, const RowGeom& g, const SmallTaps& taps) {
    const int c = src->v.channels;
    if (rx == 1) {
        if (c == 1) return launch_generic_tiled_t<MODE, 1, 1, 2>(ctx, src, dst, g, taps);
        if (c == 3) return launch_generic_tiled_t<MODE, 3, 1, 2>(ctx, src, dst, g, taps);
        if (c == 4) return launch_generic_tiled_t<MODE, 4, 1, 2>(ctx, src, dst, g, taps);
    } else if (rx == 2) {
        if (c == 1) return launch_generic_tiled_t<MODE, 1, 2, 2>(ctx, src, dst, g, taps);
        if (c == 3) return launch_generic_tiled_t<MODE, 3, 2, 2>(ctx, src, dst, g, taps);
        if (c == 4) return launch_generic_tiled_t<MODE, 4, 2, 2>(ctx, src, dst, g, taps);
    } else if (rx == 3 && c == 1) {
        return launch_generic_tiled_t<MODE, 1, 3, 2>(ctx, src, dst, g, taps);
    }
    return fail(ctx, DIP_ERR_ARG, "bad horizontal radius %d for %d 16-bit channels", rx, c);
}

template <int MODE>
int launch_generic_tiled_m(dip_ctx* ctx, int rx, const dip_image* src, dip_image* dst, const RowGeom& g, const SmallTaps& taps) {
    switch (src->v.channels) {
        case 1: return launch_generic_tiled_c<MODE, 1>(ctx, rx, src, dst, g, taps);
        case 3: return launch_generic_tiled_c<MODE, 3>(ctx, rx, src, dst, g, taps);
        case 4: return launch_generic_tiled_c<MODE, 4>(ctx, rx, src, dst, g, taps);
    }
    return fail(ctx, DIP_ERR_ARG, "bad channel count");
}

}  // namespace

bool g_generic_float_morph = false;   // development knob "float_morph": arbitrary masks on the fp32 kernel (A/B)

bool generic_tiled_reach(const dip_image* src, int kw, int kh) {
    if (kw > 7 || kh > 7) return false;
    return std::max(1, kw / 2) * src->v.channels * src->v.bps <= 16;
}

int launch_generic_tiled(dip_ctx* ctx, int mode, int kw, const dip_image* src, dip_image* dst, const RowGeom& g, const SmallTaps& taps) {
    const int rx = std::max(1, kw / 2);
    if (src->v.bps == 2) {
        if (mode < 2 && !g_generic_float_morph) {   // minima straight on the packed 16-bit words
            const int rc = mode == 0 ? launch_generic_morph16<false>(ctx, rx, src, dst, g, taps) : launch_generic_morph16<true>(ctx, rx, src, dst, g, taps);
            if (rc != -1) return rc;
        }
        switch (mode) {
            case 0: return launch_generic_tiled_16<0>(ctx, rx, src, dst, g, taps);
            case 1: return launch_generic_tiled_16<1>(ctx, rx, src, dst, g, taps);
            case 2: return launch_generic_tiled_16<2>(ctx, rx, src, dst, g, taps);
            case 3: return launch_generic_tiled_16<3>(ctx, rx, src, dst, g, taps);
        }
        return fail(ctx, DIP_ERR_ARG, "bad mode");
    }
    if (mode < 2 && !g_generic_float_morph) {   // erosion / dilation of 8-bit samples: packed minima where the window reaches
        const int rc = mode == 0 ? launch_generic_morph<false>(ctx, rx, src, dst, g, taps) : launch_generic_morph<true>(ctx, rx, src, dst, g, taps);
        if (rc != -1) return rc;
    }
    switch (mode) {
        case 0: return launch_generic_tiled_m<0>(ctx, rx, src, dst, g, taps);
        case 1: return launch_generic_tiled_m<1>(ctx, rx, src, dst, g, taps);
        case 2: return launch_generic_tiled_m<2>(ctx, rx, src, dst, g, taps);
    }
    return fail(ctx, DIP_ERR_ARG, "bad mode");
}

}  // namespace dip
