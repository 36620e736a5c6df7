// Stencil operations of libdip_b200.so: erosion / dilation, Gaussian convolutions (full and
// fused separable) and the 3x3 Gaussian blur, for sm_100a.
//
// Replaces ErosionKernel / ConvolutionKernel / GaussianBlur3x3Kernel
// (/root/reference/benchmarks/sycl/benchmark.cpp:116-252: one thread per pixel, byte loads of
// every tap and of the mask from global memory, two launches and an `aux` round trip for the
// separable variants :341-356) with the semantics of cv.erode / cv.filter2D / cv.GaussianBlur
// (/root/reference/benchmarks/opencv/benchmark.py:98-105).
//
// Design
//  * The image is a 3-D u32 tensor (pitch/4 words, rows, frames).  One CTA owns a tile of
//    512 interior bytes x TH rows; one elected thread issues a single TMA tiled load
//    (cp.async.bulk.tensor.3d) of the tile plus a 16-byte / r-row halo into shared memory
//    and everybody waits on its mbarrier.  TMA zero-fills outside the tensor; CTAs that
//    touch the image border patch the halo in shared memory (reflect-101 for convolutions,
//    +inf / -inf for erosion / dilation), so the arithmetic below has no border branches.
//  * A thread owns one 16-byte column of the tile (lane = column, warp = band of R rows) and
//    walks down its rows keeping the vertical window in registers.  Bytes are unpacked once
//    per row into packed u16x2 lanes (even / odd bytes); a tap at +-C bytes is a compile-time
//    selection of neighbouring packed words (at most one PRMT funnel).  Minima use the
//    VIMNMX3.U16x2 instruction, Gaussian weights are integers evaluated with packed IMADs (the
//    FMA pipe; the integer ALU pipe issues at half rate on sm_100, so PRMT/LOP3/LEA are the
//    expensive instructions here), and rounding reproduces OpenCV exactly: round-half-even of the exact sum for filter2D
//    (a fused separable pass rounds its u8 intermediate just as the reference's `aux` image
//    is rounded), round-half-up for GaussianBlur.
//  * Outputs leave as 16-byte coalesced stores; nothing but the source tile and the result
//    ever crosses HBM (2 bytes of traffic per output byte).
//  * Arbitrary masks / kernels (anything but the benchmark's seven) run through a simple
//    direct kernel further down (integer min/max, or fp32 accumulation within 1 LSB).
#include <algorithm>
#include <atomic>
#include <cmath>

#include "dip_internal.h"

namespace dip {

namespace {

constexpr int kTileBytes = 512;                       // interior bytes per tile row
constexpr int kHalo = 16;                             // halo bytes either side
constexpr int kRowBytes = kTileBytes + 2 * kHalo;     // 544 bytes per shared-memory row
constexpr int kBoxWords = kRowBytes / 4;

enum StencilOp {
    kErodeSq3 = 0,    // 3x3 all-ones min (also the fused 1x3+3x1)
    kErodeCross3 = 1, // 3x3 cross min
    kDilateSq3 = 2,
    kDilateCross3 = 3,
    kConv3 = 4,       // [1 2 1]x[1 2 1]/16, single RNE
    kConv3Sep = 5,    // RNE([1 2 1]/4) rows -> u8 -> RNE([1 2 1]/4) columns
    kGauss3 = 6,      // [1 2 1]x[1 2 1]/16, round half up
    kConv5 = 7,       // [1 4 6 4 1]x[1 4 6 4 1]/256, single RNE
    kConv5Sep = 8,    // RNE(/16) rows -> u8 -> RNE(/16) columns
    // one-dimensional passes (the reference's native harness runs the separable ops as two calls
    // through an aux image, sycl/benchmark.cpp:341-356)
    kErodeRow3 = 9,   // 1x3 all-ones min
    kErodeCol3 = 10,  // 3x1
    kDilateRow3 = 11,
    kDilateCol3 = 12,
    kConvRow3 = 13,   // RNE([1 2 1]/4) along x
    kConvCol3 = 14,   // along y
    kConvRow5 = 15,   // RNE([1 4 6 4 1]/16) along x
    kConvCol5 = 16,   // along y
};

__host__ __device__ constexpr bool op_is_row_only(int op) { return op == kErodeRow3 || op == kDilateRow3 || op == kConvRow3 || op == kConvRow5; }
__host__ __device__ constexpr bool op_is_col_only(int op) { return op == kErodeCol3 || op == kDilateCol3 || op == kConvCol3 || op == kConvCol5; }
__host__ __device__ constexpr int op_taps(int op) { return (op == kConv5 || op == kConv5Sep || op == kConvRow5 || op == kConvCol5) ? 2 : 1; }
__host__ __device__ constexpr int op_radius_x(int op) { return op_is_col_only(op) ? 0 : op_taps(op); }
__host__ __device__ constexpr int op_radius_y(int op) { return op_is_row_only(op) ? 0 : op_taps(op); }
__host__ __device__ constexpr bool op_is_dilate(int op) { return op == kDilateSq3 || op == kDilateCross3 || op == kDilateRow3 || op == kDilateCol3; }
__host__ __device__ constexpr bool op_is_morph(int op) { return op <= kDilateCross3 || (op >= kErodeRow3 && op <= kDilateCol3); }

struct StencilParams {
    ImgView dst;
    int rowbytes;            // valid bytes per source row
    int width;               // pixels per row
    int valid_y0, valid_y1;  // source rows that exist for border purposes
    int out_y0, out_y1;      // rows to produce (source row coordinates)
    int dst_y0;              // dst row receiving out_y0
    int tiles_x, tiles_y;
    bool dst_exact;          // dst rows are not padded to 16 bytes (wrapped region of interest)
};

// ---- PTX wrappers ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int x, int y, int z) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}

// ---- packed u16x2 helpers -------------------------------------------------------------------
// Pipe balancing.  On sm_100 the integer ALU pipe (PRMT, LOP3, SHF, LEA, IADD3, VIMNMX) and the FMA pipe
// (IMAD) each accept one warp instruction every 2 cycles per scheduler and issue in parallel
// (tools/ubench/int_pipes.cu: 0.51 / 0.51 / 0.98 instr/clk alone / alone / mixed 1:1; IMAD.HI is 0.26).
// The 5-tap and fused 3-tap kernels are ALU-pipe bound, so work that can be phrased as a multiply-add is
// phrased that way: additions as a*1+b with the 1 read from constant memory (ptxas cannot fold it and
// emits IMAD), and the even-byte unpack as w - 256*odd instead of a second PRMT.
// A constant the compiler must keep in a register (so that `x * imm + K` is ONE IMAD with K as its addend
// instead of a shift plus an ALU-pipe add).  Call once, outside the row loop.
template <uint32_t V>
__device__ __forceinline__ uint32_t reg_const() {
    uint32_t r;
    asm volatile("mov.u32 %0, %1;" : "=r"(r) : "n"(V));
    return r;
}
__constant__ uint32_t c_one = 1u;
__device__ __forceinline__ uint32_t padd(uint32_t a, uint32_t b) { return a * c_one + b; }  // add on the FMA pipe

__device__ __forceinline__ uint32_t odd_bytes(uint32_t w) { return __byte_perm(w, 0u, 0x4341); }   // (b1, b3): PRMT
__device__ __forceinline__ uint32_t even_bytes(uint32_t w, uint32_t odd) { return w - odd * 256u; } // (b0, b2): IMAD
__device__ __forceinline__ uint32_t funnel16(uint32_t lo, uint32_t hi) { return __byte_perm(lo, hi, 0x5432); }

// A row of 8 words (2 halo, 4 own, 2 halo) unpacked into even / odd byte streams.
struct Row8 {
    uint32_t e[8], o[8];
};

// Packed (byte[P], byte[P+2]) with P relative to the first own byte (word 2); P in [-8, 22].
template <int P>
__device__ __forceinline__ uint32_t pair_at(const Row8& r) {
    constexpr int PP = P + 8;
    static_assert(PP >= 0 && PP + 2 < 32, "tap outside the 8-word window");
    if constexpr ((PP & 1) == 0) {
        constexpr int q = PP / 2;
        if constexpr ((q & 1) == 0) return r.e[q / 2];
        else return funnel16(r.e[(q - 1) / 2], r.e[(q + 1) / 2]);
    } else {
        constexpr int q = (PP - 1) / 2;
        if constexpr ((q & 1) == 0) return r.o[q / 2];
        else return funnel16(r.o[(q - 1) / 2], r.o[(q + 1) / 2]);
    }
}

template <bool MAX>
__device__ __forceinline__ uint32_t mnmx3(uint32_t a, uint32_t b, uint32_t c) {
    if constexpr (MAX) return __vimax3_u16x2(a, b, c);
    else return __vimin3_u16x2(a, b, c);
}

// 8 packed results for the thread's 4 own words: index 2j = even lanes, 2j+1 = odd lanes.
struct Acc8 {
    uint32_t v[8];
};

// horizontal [1 2 1] (sum <= 1020 per lane) plus a per-lane rounding bias folded into the first add
template <int C>
__device__ __forceinline__ void hsum3(const Row8& r, Acc8& h, uint32_t bias) {
#define DIP_H3(j)                                                                                     \
    h.v[2 * j] = padd(pair_at<4 * j - C>(r), padd(pair_at<4 * j + C>(r), 2u * r.e[2 + j] + bias));    \
    h.v[2 * j + 1] = padd(pair_at<4 * j + 1 - C>(r), padd(pair_at<4 * j + 1 + C>(r), 2u * r.o[2 + j] + bias));
    DIP_H3(0) DIP_H3(1) DIP_H3(2) DIP_H3(3)
#undef DIP_H3
}

// horizontal [1 4 6 4 1] (sum <= 4080 per lane) plus a per-lane rounding bias folded into the first add
template <int C>
__device__ __forceinline__ void hsum5(const Row8& r, Acc8& h, uint32_t bias) {
#define DIP_H5(j)                                                                                     \
    h.v[2 * j] = (pair_at<4 * j - 2 * C>(r) + pair_at<4 * j + 2 * C>(r)) +                            \
                 (4u * padd(pair_at<4 * j - C>(r), pair_at<4 * j + C>(r)) + (6u * r.e[2 + j] + bias));     \
    h.v[2 * j + 1] = (pair_at<4 * j + 1 - 2 * C>(r) + pair_at<4 * j + 1 + 2 * C>(r)) +                \
                     (4u * padd(pair_at<4 * j + 1 - C>(r), pair_at<4 * j + 1 + C>(r)) + (6u * r.o[2 + j] + bias));
    DIP_H5(0) DIP_H5(1) DIP_H5(2) DIP_H5(3)
#undef DIP_H5
}

template <int C, bool MAX>
__device__ __forceinline__ void hmnmx3(const Row8& r, Acc8& h) {
#define DIP_M3(j)                                                                                     \
    h.v[2 * j] = mnmx3<MAX>(pair_at<4 * j - C>(r), r.e[2 + j], pair_at<4 * j + C>(r));                \
    h.v[2 * j + 1] = mnmx3<MAX>(pair_at<4 * j + 1 - C>(r), r.o[2 + j], pair_at<4 * j + 1 + C>(r));
    DIP_M3(0) DIP_M3(1) DIP_M3(2) DIP_M3(3)
#undef DIP_M3
}

// Lanes hold X = value * 256 (plus nothing); returns the lanes whose HIGH byte is
// round-half-even(X / 256).  X + 0x7F must already be folded in by the caller.
__device__ __forceinline__ uint32_t rne_finish(uint32_t x_plus_7f) {
    // bit 8 of each lane added at bit 0.  Written as the high half of a multiply by 2^24 so that ptxas
    // emits LOP3 + LEA.HI (2 ALU-pipe ops) instead of SHF + LOP3 + IADD (measured 4-7 % faster).
    return __umulhi(x_plus_7f & 0x01000100u, 1u << 24) + x_plus_7f;
}
// results in the high byte of each lane -> 4 output bytes
__device__ __forceinline__ uint32_t pack_hi(uint32_t ev, uint32_t od) { return __byte_perm(ev, od, 0x7351); }
// results in the low byte of each lane -> 4 output bytes
__device__ __forceinline__ uint32_t pack_lo(uint32_t ev, uint32_t od) { return __byte_perm(ev, od, 0x6240); }

// Round a packed horizontal sum to u8 exactly like the reference's intermediate image and
// keep it scaled by 2^S in the lane (T = t << S), so the vertical pass needs no re-shift.
// `r` = sum + (2^(S-1) - 1) per lane (the bias is folded into the horizontal sum by the caller).
template <int S>
__device__ __forceinline__ uint32_t round_keep_scaled(uint32_t r) {
    constexpr uint32_t keep = ~(((1u << S) - 1u) * 0x00010001u);
    return (__umulhi(r & (0x00010001u << S), 1u << (32 - S)) + r) & keep;  // LOP3 + LEA.HI + LOP3
}

__device__ __forceinline__ void load_row(const uint8_t* rowp, Row8& r) {
    const uint2 a = *reinterpret_cast<const uint2*>(rowp - 8);
    const uint4 b = *reinterpret_cast<const uint4*>(rowp);
    const uint2 c = *reinterpret_cast<const uint2*>(rowp + 16);
    const uint32_t w[8] = {a.x, a.y, b.x, b.y, b.z, b.w, c.x, c.y};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r.o[i] = odd_bytes(w[i]);
        r.e[i] = even_bytes(w[i], r.o[i]);
    }
}

// ---- border patching in shared memory --------------------------------------------------------
__device__ __forceinline__ int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// Rewrites the halo bytes/rows of the tile that lie outside the image.  `tile` row t holds
// source row (ytop + t); byte l of a row holds source byte (bx0 - kHalo + l).
template <int OP, int C, int ROWS, int kThreads>
__device__ void patch_borders(uint8_t* tile, const StencilParams& p, int bx0, int ytop) {
    constexpr int rx = op_radius_x(OP), ry = op_radius_y(OP);
    constexpr bool morph = op_is_morph(OP);
    constexpr uint8_t fill = op_is_dilate(OP) ? 0 : 255;
    const int tid = threadIdx.x;
    const bool left = rx > 0 && bx0 <= 0;
    const bool right = rx > 0 && bx0 + kTileBytes + rx * C > p.rowbytes;
    if (left || right) {
        // columns: (row, side, d, ch)
        constexpr int r = rx > 0 ? rx : 1;
        const int per_row = 2 * r * C;
        for (int i = tid; i < ROWS * per_row; i += kThreads) {
            const int t = i / per_row;
            int k = i - t * per_row;
            const int y = ytop + t;
            if (y < p.valid_y0 || y >= p.valid_y1) continue;
            const bool is_right = k >= r * C;
            if (is_right) k -= r * C;
            const int d = k / C, ch = k - d * C;
            int px, spx;
            if (!is_right) {
                if (!left) continue;
                px = -1 - d;
            } else {
                if (!right) continue;
                px = p.width + d;
            }
            const int lb = px * C + ch - (bx0 - kHalo);
            if (lb < 0 || lb >= kRowBytes) continue;
            uint8_t v = fill;
            if (!morph) {
                spx = reflect101(px, p.width);
                v = tile[t * kRowBytes + (spx * C + ch - (bx0 - kHalo))];
            }
            tile[t * kRowBytes + lb] = v;
        }
        __syncthreads();
    }
    const bool top = ry > 0 && ytop < p.valid_y0;
    const bool bottom = ry > 0 && ytop + ROWS > p.valid_y1;
    if (top || bottom) {
        constexpr int r = ry;
        constexpr int vec_per_row = kRowBytes / 16;
        for (int i = tid; i < ROWS * vec_per_row; i += kThreads) {
            const int t = i / vec_per_row;
            const int vx = i - t * vec_per_row;
            const int y = ytop + t;
            if (y >= p.valid_y0 && y < p.valid_y1) continue;
            if (y < p.valid_y0 - r || y >= p.valid_y1 + r) continue;  // never read by a valid output
            uint4 v = make_uint4(fill * 0x01010101u, fill * 0x01010101u, fill * 0x01010101u, fill * 0x01010101u);
            if (!morph) {
                const int ys = p.valid_y0 + reflect101(y - p.valid_y0, p.valid_y1 - p.valid_y0);
                const int ts = ys - ytop;
                if (ts < 0 || ts >= ROWS) continue;
                v = *reinterpret_cast<const uint4*>(tile + ts * kRowBytes + vx * 16);
            }
            *reinterpret_cast<uint4*>(tile + t * kRowBytes + vx * 16) = v;
        }
        __syncthreads();
    }
}

// ---- the arithmetic of one warp band: the thread walks down its 16-byte column --------------------
// FULL = the band has all R rows (compile-time trip count: the row loops unroll by the period of the
// register window without a remainder loop); otherwise `nrows_rt` rows.
// CLIP = this warp holds the row's last, partial 16-byte chunk of a destination whose rows are NOT padded (a wrapped
// region of interest): its stores write only the valid bytes.  Every other warp keeps plain vector stores.
template <bool CLIP>
__device__ __forceinline__ void put16(uint8_t* p, const uint4& v, int valid) {
    if constexpr (CLIP) store16(p, v, valid);
    else __stcs(reinterpret_cast<uint4*>(p), v);
}

template <int OP, int C, int R, bool FULL, bool CLIP>
__device__ __forceinline__ void compute_band(const uint8_t* colp, uint8_t* dptr, const StencilParams& p, int nrows_rt, int p_cx) {
    const int nrows = FULL ? R : nrows_rt;
    const int valid = CLIP ? p.rowbytes - p_cx : 16;  // bytes of this thread's 16-byte chunk that may be stored
    if constexpr (op_is_row_only(OP)) {
        // one-dimensional pass along x: no vertical window
        Row8 row;
        Acc8 h;
#pragma unroll 3
        for (int i = 0; i < nrows; ++i) {
            load_row(colp + i * kRowBytes, row);
            uint32_t out[4];
            if constexpr (op_is_morph(OP)) {
                hmnmx3<C, op_is_dilate(OP)>(row, h);
#pragma unroll
                for (int j = 0; j < 4; ++j) out[j] = pack_lo(h.v[2 * j], h.v[2 * j + 1]);
            } else {
                constexpr uint32_t scale = OP == kConvRow3 ? 64u : 16u;  // sum * scale has the result in the high byte
                if constexpr (OP == kConvRow3) hsum3<C>(row, h, 0u);
                else hsum5<C>(row, h, 0u);
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    out[j] = pack_hi(rne_finish(h.v[2 * j] * scale + 0x007F007Fu), rne_finish(h.v[2 * j + 1] * scale + 0x007F007Fu));
            }
            put16<CLIP>(dptr + (size_t)i * p.dst.pitch, make_uint4(out[0], out[1], out[2], out[3]), valid);
        }
    } else if constexpr (op_is_col_only(OP)) {
        // one-dimensional pass along y: only the thread's own 4 words, vertical window of raw rows
        constexpr int taps = op_taps(OP);
        Acc8 w0, w1, w2, w3, w4;
        auto load_own = [&](const uint8_t* rp, Acc8& a) {
            const uint4 b = *reinterpret_cast<const uint4*>(rp);
            const uint32_t w[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { a.v[2 * j + 1] = odd_bytes(w[j]); a.v[2 * j] = even_bytes(w[j], a.v[2 * j + 1]); }
        };
        load_own(colp, w0);
        load_own(colp + kRowBytes, w1);
        if constexpr (taps == 2) { load_own(colp + 2 * kRowBytes, w2); load_own(colp + 3 * kRowBytes, w3); }
#pragma unroll (op_taps(OP) == 2 ? 5 : 3)
        for (int i = 0; i < nrows; ++i) {
            uint32_t res[8];
            if constexpr (taps == 1) {
                load_own(colp + (i + 2) * kRowBytes, w2);
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    if constexpr (op_is_morph(OP)) res[k] = mnmx3<op_is_dilate(OP)>(w0.v[k], w1.v[k], w2.v[k]);
                    else res[k] = rne_finish((w0.v[k] + w2.v[k]) * 64u + w1.v[k] * 128u + 0x007F007Fu);
                }
                w0 = w1; w1 = w2;
            } else {
                load_own(colp + (i + 4) * kRowBytes, w4);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    res[k] = rne_finish((w0.v[k] + w4.v[k]) * 16u + (w1.v[k] + w3.v[k]) * 64u + w2.v[k] * 96u + 0x007F007Fu);
                w0 = w1; w1 = w2; w2 = w3; w3 = w4;
            }
            uint4 o;
            if constexpr (op_is_morph(OP)) o = make_uint4(pack_lo(res[0], res[1]), pack_lo(res[2], res[3]), pack_lo(res[4], res[5]), pack_lo(res[6], res[7]));
            else o = make_uint4(pack_hi(res[0], res[1]), pack_hi(res[2], res[3]), pack_hi(res[4], res[5]), pack_hi(res[6], res[7]));
            put16<CLIP>(dptr + (size_t)i * p.dst.pitch, o, valid);
        }
    } else if constexpr (op_is_morph(OP)) {
        constexpr bool MAX = op_is_dilate(OP);
        constexpr bool cross = OP == kErodeCross3 || OP == kDilateCross3;
        Acc8 h0, h1, h2;   // square: horizontal extrema of rows i-1, i, i+1
        Acc8 c0, c1, c2;   // cross: centre columns of rows i-1, i, i+1; h1 = horizontal of row i
        Row8 row;
        load_row(colp, row);
        if constexpr (cross) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { c0.v[2 * j] = row.e[2 + j]; c0.v[2 * j + 1] = row.o[2 + j]; }
        } else {
            hmnmx3<C, MAX>(row, h0);
        }
        load_row(colp + kRowBytes, row);
        hmnmx3<C, MAX>(row, h1);
        if constexpr (cross) {
#pragma unroll
            for (int j = 0; j < 4; ++j) { c1.v[2 * j] = row.e[2 + j]; c1.v[2 * j + 1] = row.o[2 + j]; }
        }
#pragma unroll 3
        for (int i = 0; i < nrows; ++i) {
            load_row(colp + (i + 2) * kRowBytes, row);
            hmnmx3<C, MAX>(row, h2);
            uint32_t out[4];
            if constexpr (cross) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { c2.v[2 * j] = row.e[2 + j]; c2.v[2 * j + 1] = row.o[2 + j]; }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    out[j] = pack_lo(mnmx3<MAX>(c0.v[2 * j], h1.v[2 * j], c2.v[2 * j]),
                                     mnmx3<MAX>(c0.v[2 * j + 1], h1.v[2 * j + 1], c2.v[2 * j + 1]));
                c0 = c1; c1 = c2;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    out[j] = pack_lo(mnmx3<MAX>(h0.v[2 * j], h1.v[2 * j], h2.v[2 * j]),
                                     mnmx3<MAX>(h0.v[2 * j + 1], h1.v[2 * j + 1], h2.v[2 * j + 1]));
                h0 = h1;
            }
            h1 = h2;
            put16<CLIP>(dptr + (size_t)i * p.dst.pitch, make_uint4(out[0], out[1], out[2], out[3]), valid);
        }
    } else if constexpr (op_taps(OP) == 1) {
        // 3-tap Gaussians.  h* hold the horizontal pass of rows i-1, i, i+1.
        const uint32_t hbias = OP == kConv3Sep ? reg_const<0x00010001u>() : 0u;
        const uint32_t vbias = reg_const<(OP == kGauss3 ? 0x00800080u : 0x007F007Fu)>();
        Acc8 h0, h1, h2;
        Row8 row;
        auto hpass = [&](const uint8_t* rp, Acc8& h) {
            load_row(rp, row);
            hsum3<C>(row, h, hbias);
            if constexpr (OP == kConv3Sep) {
#pragma unroll
                for (int k = 0; k < 8; ++k) h.v[k] = round_keep_scaled<2>(h.v[k]);  // 4 * RNE(sum / 4)
            }
        };
        hpass(colp, h0);
        hpass(colp + kRowBytes, h1);
#pragma unroll 3
        for (int i = 0; i < nrows; ++i) {
            hpass(colp + (i + 2) * kRowBytes, h2);
            uint32_t res[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                // V = h0 + 2 h1 + h2 <= 4080 (both variants); X = 16 V has the result in the high byte
                const uint32_t u = padd(h0.v[k], h2.v[k]);
                if constexpr (OP == kGauss3) {
                    res[k] = u * 16u + (h1.v[k] * 32u + vbias);              // (S + 8) >> 4
                } else {
                    res[k] = rne_finish(u * 16u + (h1.v[k] * 32u + vbias));
                }
            }
            put16<CLIP>(dptr + (size_t)i * p.dst.pitch,
                    make_uint4(pack_hi(res[0], res[1]), pack_hi(res[2], res[3]), pack_hi(res[4], res[5]), pack_hi(res[6], res[7])), valid);
            h0 = h1; h1 = h2;
        }
    } else {
        // 5-tap Gaussians.  h* hold the horizontal pass of rows i-2 .. i+2.
        const uint32_t hbias = OP == kConv5Sep ? reg_const<0x00070007u>() : 0u;
        const uint32_t vbias = reg_const<0x007F007Fu>();
        Acc8 h0, h1, h2, h3, h4;
        Row8 row;
        auto hpass = [&](const uint8_t* rp, Acc8& h) {
            load_row(rp, row);
            hsum5<C>(row, h, hbias);
            if constexpr (OP == kConv5Sep) {
#pragma unroll
                for (int k = 0; k < 8; ++k) h.v[k] = round_keep_scaled<4>(h.v[k]);  // 16 * RNE(sum / 16)
            }
        };
        hpass(colp, h0);
        hpass(colp + kRowBytes, h1);
        hpass(colp + 2 * kRowBytes, h2);
        hpass(colp + 3 * kRowBytes, h3);
#pragma unroll 5
        for (int i = 0; i < nrows; ++i) {
            hpass(colp + (i + 4) * kRowBytes, h4);
            uint32_t res[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                // X = sum of [1 4 6 4 1] x h <= 65280: result in the high byte after RNE
                // the separable variant (two roundings) is the most ALU-bound: all of its vertical adds go to the FMA pipe
                const uint32_t x = OP == kConv5Sep
                                       ? padd(h0.v[k], padd(h4.v[k], 4u * padd(h1.v[k], h3.v[k]) + (6u * h2.v[k] + vbias)))
                                       : (h0.v[k] + h4.v[k]) + (4u * padd(h1.v[k], h3.v[k]) + (6u * h2.v[k] + vbias));
                res[k] = rne_finish(x);
            }
            put16<CLIP>(dptr + (size_t)i * p.dst.pitch,
                    make_uint4(pack_hi(res[0], res[1]), pack_hi(res[2], res[3]), pack_hi(res[4], res[5]), pack_hi(res[6], res[7])), valid);
            h0 = h1; h1 = h2; h2 = h3; h3 = h4;
        }
    }
}

// ---- the tiled kernel -------------------------------------------------------------------------
template <int OP, int C, int R, int kWarps>
__global__ void __launch_bounds__(kWarps * 32) stencil_kernel(const __grid_constant__ CUtensorMap tmap, const StencilParams p) {
    constexpr int r = op_radius_y(OP);   // halo rows
    constexpr int rx = op_radius_x(OP);  // halo pixels
    constexpr int kThreads = kWarps * 32;
    constexpr int TH = kWarps * R;
    constexpr int ROWS = TH + 2 * r;
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* tile = smem;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + ROWS * kRowBytes);

    const int tx = blockIdx.x, ty = blockIdx.y, f = blockIdx.z;  // x fastest: neighbouring CTAs share halos in L2
    const int bx0 = tx * kTileBytes;
    const int y0 = p.out_y0 + ty * TH;  // first output row of the tile
    const int ytop = y0 - r;            // source row held by tile row 0

    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        mbar_expect_tx(bar, ROWS * kRowBytes);
    }
    pdl_wait();  // the previous kernel's writes are visible from here on
    if (threadIdx.x == 0) tma_load_3d(tile, &tmap, bar, (bx0 - kHalo) / 4, ytop, f);
    __syncthreads();  // barrier init visible to the waiters
    mbar_wait(bar, 0);

    const bool border = (rx > 0 && (bx0 == 0 || bx0 + kTileBytes + rx * C > p.rowbytes)) ||
                        (r > 0 && (ytop < p.valid_y0 || ytop + ROWS > p.valid_y1));
    if (border) {
        __syncthreads();
        patch_borders<OP, C, ROWS, kThreads>(tile, p, bx0, ytop);
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cx = bx0 + 16 * lane;  // first byte of this thread's column
    if (cx >= ((p.rowbytes + 15) & ~15)) return;
    const int yw = y0 + warp * R;    // first output row of this warp
    if (yw >= p.out_y1) return;
    const int nrows = min(R, p.out_y1 - yw);
    const uint8_t* colp = tile + (warp * R) * kRowBytes + kHalo + 16 * lane;  // tile row of source row yw - r
    uint8_t* dptr = p.dst.base + (size_t)f * p.dst.frame_stride + (size_t)(p.dst_y0 + (yw - p.out_y0)) * p.dst.pitch + cx;

    // warp-uniform: does this tile column contain the partial last chunk of an unpadded destination row?
    const bool clip = p.dst_exact && bx0 + kTileBytes > p.rowbytes && (p.rowbytes & 15) != 0;
    if (clip) compute_band<OP, C, R, false, true>(colp, dptr, p, nrows, cx);
    else if (nrows == R) compute_band<OP, C, R, true, false>(colp, dptr, p, R, cx);
    else compute_band<OP, C, R, false, false>(colp, dptr, p, nrows, cx);
}

// Development knobs (dip_debug_set): "tile_rows" 0 = automatic, 64 or 32 forces the tile height.
struct Knobs { int tile_rows = 0, fused_rows = 0, extra_smem = 0; };
Knobs g_knobs;

template <int OP, int C, int R, int kWarps>
int launch_tiled(dip_ctx* ctx, const dip_image* src, dip_image* dst, const RowGeom& g) {
    constexpr int r = op_radius_y(OP);
    constexpr int kThreads = kWarps * 32;
    constexpr int TH = kWarps * R;
    constexpr int ROWS = TH + 2 * r;
    // Resident CTAs per SM.  The tile alone allows 6 (which the issue-bound Gaussians want: measured +11 % time at 5);
    // the HBM-bound ops run faster with fewer tiles in flight (256-frame stack, us per frame: 3x3 erosions 7.30 at 6,
    // 7.30 at 5, 7.07 at 4, 7.8 at 3; GaussianBlur 7.29 / 7.18 / 7.28), so their launches ask for more dynamic shared
    // memory than the tile needs.  Only the stack shape (R = 15); profiles/r1_occupancy_sweep.md.
    constexpr int kCtasPerSm = R != 15 ? 0 : (OP == kErodeSq3 || OP == kErodeCross3 || OP == kDilateSq3 || OP == kDilateCross3) ? 4 : OP == kGauss3 ? 5 : 0;
    constexpr int kTileSmem = ROWS * kRowBytes + 16;
    constexpr int kPadded = kCtasPerSm ? (227 * 1024) / kCtasPerSm - 1024 - 512 : 0;  // 1 KiB per CTA is reserved by the driver
    const int smem_bytes = (kPadded > kTileSmem ? kPadded : kTileSmem) + g_knobs.extra_smem;
    const CUtensorMap* map = nullptr;
    int rc = get_tmap(ctx, const_cast<dip_image*>(src), kBoxWords, ROWS, &map);
    if (rc != DIP_OK) return rc;
    StencilParams p{};
    p.dst = dst->v;
    p.rowbytes = src->v.rowbytes;
    p.width = src->v.width;
    p.valid_y0 = g.valid_y0; p.valid_y1 = g.valid_y1;
    p.out_y0 = g.out_y0; p.out_y1 = g.out_y1; p.dst_y0 = g.dst_y0;
    p.tiles_x = (src->v.rowbytes + kTileBytes - 1) / kTileBytes;
    p.tiles_y = (g.out_y1 - g.out_y0 + TH - 1) / TH;
    p.dst_exact = exact_rows(dst->v);
    if (p.tiles_x <= 0 || p.tiles_y <= 0) return DIP_OK;
    if (p.tiles_y > 65535 || src->v.frames > 65535) return fail(ctx, DIP_ERR_UNSUPPORTED, "image stack too large for one launch");
    static std::atomic<uint64_t> attr_set{0};  // per instantiation, one bit per device
    const uint64_t dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_set.load(std::memory_order_relaxed) & dev_bit) || g_knobs.extra_smem) {
        DIP_CUDA(ctx, cudaFuncSetAttribute(stencil_kernel<OP, C, R, kWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr_set.fetch_or(dev_bit, std::memory_order_relaxed);
    }
    DIP_CUDA(ctx, launch_kernel(stencil_kernel<OP, C, R, kWarps>, dim3((unsigned)p.tiles_x, (unsigned)p.tiles_y, (unsigned)src->v.frames), dim3(kThreads), smem_bytes, ctx->compute, *map, p));
    DIP_CHECK_LAUNCH(ctx, "stencil_kernel");
    return DIP_OK;
}

// =================================================================================================
// Fused chain: 5x5 Gaussian (filter2D semantics, reflect-101, RNE) followed by 3x3 square erosion
// (+inf border) in ONE kernel -- BASELINE.json config 5's pipeline at 2 bytes of HBM traffic per pixel
// instead of 4.  Everything after the source tile stays in registers: a thread computes
// the Gaussian of its 16-byte column row by row (as the 5x5 kernel above), keeps the result in its
// packed u16x2 form (value in the high byte of each lane; the low byte is rounding residue, which cannot
// change which HIGH byte a 16-bit minimum selects), gets the neighbouring columns' edge words with four
// warp shuffles, takes the horizontal 3-minimum and keeps the vertical 3-window in registers.  Lanes 0
// and 31 have only one neighbour, so a tile delivers 480 of its 512 byte columns and tiles overlap by
// one lane; a warp computes R+2 Gaussian rows for R output rows.  No second shared-memory tile, no
// __syncthreads after the load.
// =================================================================================================
template <int C, int R, int kWarps>
__global__ void __launch_bounds__(kWarps * 32) fused_g5e3_kernel(const __grid_constant__ CUtensorMap tmap, const StencilParams p) {
    constexpr int kThreads = kWarps * 32;
    constexpr int TO = kWarps * R;        // output rows per tile
    constexpr int ROWS = TO + 6;          // source rows per tile (Gaussian rows TO + 2, each needing +-2)
    constexpr int kOutBytes = kTileBytes - 32;  // 480 output bytes per tile row (lanes 1..30)
    static_assert(R % 15 == 0, "R must be a multiple of both register-window periods (5 and 3)");
    extern __shared__ __align__(128) uint8_t smem[];
    uint8_t* tile = smem;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + ROWS * kRowBytes);

    const int tx = blockIdx.x, ty = blockIdx.y, f = blockIdx.z;
    const int gx0 = tx * kOutBytes - 16;  // first byte column of the tile's 512-byte region
    const int y0 = p.out_y0 + ty * TO;    // first output row
    const int ytop = y0 - 3;              // source row held by tile row 0
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        mbar_expect_tx(bar, ROWS * kRowBytes);
    }
    pdl_wait();
    if (threadIdx.x == 0) tma_load_3d(tile, &tmap, bar, (gx0 - kHalo) / 4, ytop, f);
    __syncthreads();
    mbar_wait(bar, 0);
    const bool border = gx0 <= 0 || gx0 + kTileBytes + 2 * C > p.rowbytes || ytop < p.valid_y0 || ytop + ROWS > p.valid_y1;
    if (border) {
        __syncthreads();
        patch_borders<kConv5, C, ROWS, kThreads>(tile, p, gx0, ytop);
    }

    const int cx = gx0 + 16 * lane;       // first byte of this thread's column
    const int yw = y0 + warp * R;         // first output row of this warp
    if (yw >= p.out_y1) return;           // whole warp
    const int n_out = min(R, p.out_y1 - yw);
    // Gaussian bytes outside [0, rowbytes) count as 255 for the erosion: per-lane masks in packed form
    uint32_t colmask[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const int b0 = cx + 4 * j;
        auto out = [&](int pos) { return (pos < 0 || pos >= p.rowbytes) ? 0xFF00u : 0u; };
        colmask[2 * j] = out(b0) | (out(b0 + 2) << 16);
        colmask[2 * j + 1] = out(b0 + 1) | (out(b0 + 3) << 16);
    }
    const uint8_t* colp = tile + (warp * R) * kRowBytes + kHalo + 16 * lane;  // source row yw - 3
    const bool store_lane = lane != 0 && lane != 31 && cx < ((p.rowbytes + 15) & ~15);
    const bool clip = p.dst_exact && gx0 + kTileBytes > p.rowbytes && (p.rowbytes & 15) != 0;  // block-uniform
    const int valid = clip ? p.rowbytes - cx : 16;
    uint8_t* dptr = p.dst.base + (size_t)f * p.dst.frame_stride + (size_t)(p.dst_y0 + (yw - p.out_y0)) * p.dst.pitch + cx;

    Acc8 h0, h1, h2, h3, h4;  // horizontal Gaussian passes of five consecutive source rows
    Acc8 m0, m1;              // horizontal minima of the two previous Gaussian rows
    Row8 row;
    auto hpass = [&](const uint8_t* rp, Acc8& h) {
        load_row(rp, row);
        hsum5<C>(row, h, 0u);
    };
    // Gaussian row g (image row yw - 1 + g) from the window, then its horizontal 3-minimum
    auto gauss_hmin = [&](int g, Acc8& hm) {
        const int yg = yw - 1 + g;
        const uint32_t rowmask = (yg >= p.valid_y0 && yg < p.valid_y1) ? 0u : 0xFF00FF00u;
        Row8 gr;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t v = rne_finish((h0.v[k] + h4.v[k] + 0x007F007Fu) + 4u * (h1.v[k] + h3.v[k]) + 6u * h2.v[k]) | colmask[k] | rowmask;
            if (k & 1) gr.o[2 + (k >> 1)] = v; else gr.e[2 + (k >> 1)] = v;
        }
        gr.e[1] = __shfl_up_sync(0xFFFFFFFFu, gr.e[5], 1);
        gr.o[1] = __shfl_up_sync(0xFFFFFFFFu, gr.o[5], 1);
        gr.e[6] = __shfl_down_sync(0xFFFFFFFFu, gr.e[2], 1);
        gr.o[6] = __shfl_down_sync(0xFFFFFFFFu, gr.o[2], 1);
        gr.e[0] = gr.o[0] = gr.e[7] = gr.o[7] = 0u;  // outside a 3-tap's reach
        hmnmx3<C, false>(gr, hm);
    };
    hpass(colp, h0);
    hpass(colp + kRowBytes, h1);
    hpass(colp + 2 * kRowBytes, h2);
    hpass(colp + 3 * kRowBytes, h3);
    hpass(colp + 4 * kRowBytes, h4);
    gauss_hmin(0, m0);
    h0 = h1; h1 = h2; h2 = h3; h3 = h4;
    hpass(colp + 5 * kRowBytes, h4);
    gauss_hmin(1, m1);
    h0 = h1; h1 = h2; h2 = h3; h3 = h4;
    const int n_loop = (n_out == R) ? R : n_out;
#pragma unroll 5
    for (int i = 0; i < n_loop; ++i) {
        hpass(colp + (i + 6) * kRowBytes, h4);
        Acc8 m2;
        gauss_hmin(i + 2, m2);
        if (store_lane) {
            uint32_t out[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                out[j] = pack_hi(mnmx3<false>(m0.v[2 * j], m1.v[2 * j], m2.v[2 * j]), mnmx3<false>(m0.v[2 * j + 1], m1.v[2 * j + 1], m2.v[2 * j + 1]));
            store16(dptr + (size_t)i * p.dst.pitch, make_uint4(out[0], out[1], out[2], out[3]), valid);
        }
        m0 = m1; m1 = m2;
        h0 = h1; h1 = h2; h2 = h3; h3 = h4;
    }
}

template <int C, int R, int kWarps>
int launch_fused_g5e3(dip_ctx* ctx, const dip_image* src, dip_image* dst, const RowGeom& g) {
    constexpr int TO = kWarps * R, ROWS = TO + 6;
    constexpr int smem_bytes = ROWS * kRowBytes + 16;
    const CUtensorMap* map = nullptr;
    int rc = get_tmap(ctx, const_cast<dip_image*>(src), kBoxWords, ROWS, &map);
    if (rc != DIP_OK) return rc;
    StencilParams p{};
    p.dst = dst->v;
    p.rowbytes = src->v.rowbytes;
    p.width = src->v.width;
    p.valid_y0 = g.valid_y0; p.valid_y1 = g.valid_y1;
    p.out_y0 = g.out_y0; p.out_y1 = g.out_y1; p.dst_y0 = g.dst_y0;
    p.tiles_x = (src->v.rowbytes + (kTileBytes - 32) - 1) / (kTileBytes - 32);
    p.tiles_y = (g.out_y1 - g.out_y0 + TO - 1) / TO;
    p.dst_exact = exact_rows(dst->v);
    if (p.tiles_x <= 0 || p.tiles_y <= 0) return DIP_OK;
    if (p.tiles_y > 65535 || src->v.frames > 65535) return fail(ctx, DIP_ERR_UNSUPPORTED, "image stack too large for one launch");
    static std::atomic<uint64_t> attr_set{0};
    const uint64_t dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_set.load(std::memory_order_relaxed) & dev_bit)) {
        DIP_CUDA(ctx, cudaFuncSetAttribute(fused_g5e3_kernel<C, R, kWarps>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr_set.fetch_or(dev_bit, std::memory_order_relaxed);
    }
    DIP_CUDA(ctx, launch_kernel(fused_g5e3_kernel<C, R, kWarps>, dim3((unsigned)p.tiles_x, (unsigned)p.tiles_y, (unsigned)src->v.frames),
                                dim3(kWarps * 32), smem_bytes, ctx->compute, *map, p));
    DIP_CHECK_LAUNCH(ctx, "fused_g5e3_kernel");
    return DIP_OK;
}


template <int OP>
int launch_tiled_c(dip_ctx* ctx, const dip_image* src, dip_image* dst, const RowGeom& g) {
    switch (src->v.channels) {
        case 1: return launch_tiled<OP, 1, 15, 4>(ctx, src, dst, g);
        case 4: return launch_tiled<OP, 4, 15, 4>(ctx, src, dst, g);
        case 3: {
            // Few tiles (a single image): halve the tile height so the launch is about two waves and
            // late CTAs' loads overlap early CTAs' arithmetic; stacks keep the 64-row tile.
            const long long tiles64 = (long long)((src->v.rowbytes + kTileBytes - 1) / kTileBytes) * ((g.out_y1 - g.out_y0 + 63) / 64) * src->v.frames;
            const bool small = g_knobs.tile_rows ? g_knobs.tile_rows < 60 : tiles64 < 2ll * 6 * ctx->sm_count;
            // multiples of the window period (5 / 3 rows); measured best for one 3504x2336 image: 40- and 24-row tiles
            constexpr int kSmallR = op_taps(OP) == 2 ? 10 : 6;
            return small ? launch_tiled<OP, 3, kSmallR, 4>(ctx, src, dst, g) : launch_tiled<OP, 3, 15, 4>(ctx, src, dst, g);
        }
    }
    return fail(ctx, DIP_ERR_ARG, "bad channel count");
}

// ---- direct kernels for arbitrary masks / kernels ------------------------------------------------
constexpr int kMaxTaps = 225;  // up to 15 x 15

struct GenericParams {
    ImgView src, dst;
    int valid_y0, valid_y1, out_y0, out_y1, dst_y0;
    int kw, kh, ntaps;
    int8_t dx[kMaxTaps], dy[kMaxTaps];
    float wt[kMaxTaps];
};

template <int MODE>  // 0 erode, 1 dilate, 2 fp32 correlation
__global__ void __launch_bounds__(256) generic_kernel(const __grid_constant__ GenericParams p) {
    const size_t total = (size_t)p.src.rowbytes * (size_t)(p.out_y1 - p.out_y0) * (size_t)p.src.frames;
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (i >= total) return;
    const int C = p.src.channels;
    const int xb = (int)(i % (size_t)p.src.rowbytes);
    const size_t q = i / (size_t)p.src.rowbytes;
    const int yo = (int)(q % (size_t)(p.out_y1 - p.out_y0));
    const size_t f = q / (size_t)(p.out_y1 - p.out_y0);
    const int y = p.out_y0 + yo;
    const int x = xb / C, ch = xb - x * C;
    const uint8_t* sbase = p.src.base + f * p.src.frame_stride;
    const int vh = p.valid_y1 - p.valid_y0;
    uint8_t result;
    if (MODE == 2) {
        float acc = 0.0f;
        for (int t = 0; t < p.ntaps; ++t) {
            const int xx = reflect101(x + p.dx[t], p.src.width);
            const int yy = p.valid_y0 + reflect101(y + p.dy[t] - p.valid_y0, vh);
            acc = __fmaf_rn(p.wt[t], (float)sbase[(size_t)yy * p.src.pitch + xx * C + ch], acc);
        }
        // cvRound + saturate_cast<uchar>
        const int rr = __float2int_rn(acc);
        result = (uint8_t)min(max(rr, 0), 255);
    } else {
        int m = MODE == 0 ? 255 : 0;
        for (int t = 0; t < p.ntaps; ++t) {
            const int xx = x + p.dx[t], yy = y + p.dy[t];
            if (xx < 0 || xx >= p.src.width || yy < p.valid_y0 || yy >= p.valid_y1) continue;
            const int v = sbase[(size_t)yy * p.src.pitch + xx * C + ch];
            m = MODE == 0 ? min(m, v) : max(m, v);
        }
        result = (uint8_t)m;
    }
    p.dst.base[f * p.dst.frame_stride + (size_t)(p.dst_y0 + yo) * p.dst.pitch + xb] = result;
}

template <int MODE>
int launch_generic(dip_ctx* ctx, const GenericParams& p) {
    const size_t total = (size_t)p.src.rowbytes * (size_t)(p.out_y1 - p.out_y0) * (size_t)p.src.frames;
    if (total == 0) return DIP_OK;
    const size_t blocks = (total + 255) / 256;
    if (blocks > 0x7FFFFFFFull) return fail(ctx, DIP_ERR_UNSUPPORTED, "image stack too large for one launch");
    DIP_CUDA(ctx, launch_kernel(generic_kernel<MODE>, dim3((unsigned)blocks), dim3(256), 0, ctx->compute, p));
    DIP_CHECK_LAUNCH(ctx, "generic_kernel");
    return DIP_OK;
}

void fill_generic(GenericParams& p, const dip_image* src, const dip_image* dst, const RowGeom& g, int kw, int kh) {
    p.src = src->v; p.dst = dst->v;
    p.valid_y0 = g.valid_y0; p.valid_y1 = g.valid_y1; p.out_y0 = g.out_y0; p.out_y1 = g.out_y1; p.dst_y0 = g.dst_y0;
    p.kw = kw; p.kh = kh; p.ntaps = 0;
}

int check_geom(dip_ctx* ctx, const dip_image* src, const dip_image* dst, const RowGeom& g) {
    if (g.valid_y0 < 0 || g.valid_y1 > src->v.height || g.valid_y0 >= g.valid_y1)
        return fail(ctx, DIP_ERR_ARG, "valid rows [%d,%d) outside source of %d rows", g.valid_y0, g.valid_y1, src->v.height);
    if (g.out_y0 < g.valid_y0 || g.out_y1 > g.valid_y1 || g.out_y0 > g.out_y1)
        return fail(ctx, DIP_ERR_ARG, "output rows [%d,%d) outside valid rows [%d,%d)", g.out_y0, g.out_y1, g.valid_y0, g.valid_y1);
    if (g.dst_y0 < 0 || g.dst_y0 + (g.out_y1 - g.out_y0) > dst->v.height)
        return fail(ctx, DIP_ERR_SHAPE, "destination has too few rows");
    if (src->v.width != dst->v.width || src->v.channels != dst->v.channels || src->v.frames != dst->v.frames)
        return fail(ctx, DIP_ERR_SHAPE, "src and dst differ in width / channels / frames");
    // overlap check on the underlying memory
    const uint8_t* s0 = src->v.base;
    const uint8_t* s1 = s0 + src->v.frame_stride * (size_t)(src->v.frames - 1) + src->v.pitch * (size_t)src->v.height;
    const uint8_t* d0 = dst->v.base;
    const uint8_t* d1 = d0 + dst->v.frame_stride * (size_t)(dst->v.frames - 1) + dst->v.pitch * (size_t)dst->v.height;
    if (s0 < d1 && d0 < s1) return fail(ctx, DIP_ERR_SHAPE, "stencil operations cannot run in place (src overlaps dst)");
    return DIP_OK;
}

// weights * scale are integers matching `pattern`?
bool matches(const float* k, int n, const int* pattern, float scale) {
    for (int i = 0; i < n; ++i)
        if (k[i] * scale != (float)pattern[i]) return false;
    return true;
}

}  // namespace

int stencil_morph(dip_ctx* ctx, Morph m, const dip_image* src, dip_image* dst, const uint8_t* mask, int mw, int mh,
                  const RowGeom& g) {
    if (!mask || mw < 1 || mh < 1) return fail(ctx, DIP_ERR_ARG, "bad mask");
    if (mw * mh > kMaxTaps || mw > 127 || mh > 127) return fail(ctx, DIP_ERR_UNSUPPORTED, "mask larger than %d taps", kMaxTaps);
    int rc = check_geom(ctx, src, dst, g);
    if (rc != DIP_OK) return rc;
    int nz = 0;
    for (int i = 0; i < mw * mh; ++i) nz += mask[i] != 0;
    if (nz == 0) return fail(ctx, DIP_ERR_ARG, "mask has no non-zero tap");
    const bool dil = m == Morph::Dilate;
    if (mw == 3 && mh == 1 && nz == 3) return dil ? launch_tiled_c<kDilateRow3>(ctx, src, dst, g) : launch_tiled_c<kErodeRow3>(ctx, src, dst, g);
    if (mw == 1 && mh == 3 && nz == 3) return dil ? launch_tiled_c<kDilateCol3>(ctx, src, dst, g) : launch_tiled_c<kErodeCol3>(ctx, src, dst, g);
    if (mw == 3 && mh == 3) {
        if (nz == 9) return dil ? launch_tiled_c<kDilateSq3>(ctx, src, dst, g) : launch_tiled_c<kErodeSq3>(ctx, src, dst, g);
        const bool cross = nz == 5 && mask[1] && mask[3] && mask[4] && mask[5] && mask[7];
        if (cross) return dil ? launch_tiled_c<kDilateCross3>(ctx, src, dst, g) : launch_tiled_c<kErodeCross3>(ctx, src, dst, g);
    }
    GenericParams p;
    fill_generic(p, src, dst, g, mw, mh);
    for (int j = 0; j < mh; ++j)
        for (int i = 0; i < mw; ++i)
            if (mask[j * mw + i]) {
                p.dx[p.ntaps] = (int8_t)(i - mw / 2);
                p.dy[p.ntaps] = (int8_t)(j - mh / 2);
                p.wt[p.ntaps] = 1.0f;
                ++p.ntaps;
            }
    return dil ? launch_generic<1>(ctx, p) : launch_generic<0>(ctx, p);
}

int stencil_convolve(dip_ctx* ctx, const dip_image* src, dip_image* dst, const float* k, int kw, int kh, const RowGeom& g) {
    if (!k || kw < 1 || kh < 1) return fail(ctx, DIP_ERR_ARG, "bad kernel");
    if (kw * kh > kMaxTaps || kw > 127 || kh > 127) return fail(ctx, DIP_ERR_UNSUPPORTED, "kernel larger than %d taps", kMaxTaps);
    int rc = check_geom(ctx, src, dst, g);
    if (rc != DIP_OK) return rc;
    static const int g3[9] = {1, 2, 1, 2, 4, 2, 1, 2, 1};
    static const int g5[25] = {1, 4, 6, 4, 1, 4, 16, 24, 16, 4, 6, 24, 36, 24, 6, 4, 16, 24, 16, 4, 1, 4, 6, 4, 1};
    static const int b3[3] = {1, 2, 1};
    static const int b5[5] = {1, 4, 6, 4, 1};
    if (kw == 3 && kh == 1 && matches(k, 3, b3, 4.0f)) return launch_tiled_c<kConvRow3>(ctx, src, dst, g);
    if (kw == 1 && kh == 3 && matches(k, 3, b3, 4.0f)) return launch_tiled_c<kConvCol3>(ctx, src, dst, g);
    if (kw == 5 && kh == 1 && matches(k, 5, b5, 16.0f)) return launch_tiled_c<kConvRow5>(ctx, src, dst, g);
    if (kw == 1 && kh == 5 && matches(k, 5, b5, 16.0f)) return launch_tiled_c<kConvCol5>(ctx, src, dst, g);
    if (kw == 3 && kh == 3 && matches(k, 9, g3, 16.0f)) return launch_tiled_c<kConv3>(ctx, src, dst, g);
    if (kw == 5 && kh == 5 && matches(k, 25, g5, 256.0f)) return launch_tiled_c<kConv5>(ctx, src, dst, g);
    GenericParams p;
    fill_generic(p, src, dst, g, kw, kh);
    for (int j = 0; j < kh; ++j)
        for (int i = 0; i < kw; ++i)
            if (k[j * kw + i] != 0.0f) {
                p.dx[p.ntaps] = (int8_t)(i - kw / 2);
                p.dy[p.ntaps] = (int8_t)(j - kh / 2);
                p.wt[p.ntaps] = k[j * kw + i];
                ++p.ntaps;
            }
    return launch_generic<2>(ctx, p);
}

int stencil_convolve_sep(dip_ctx* ctx, const dip_image* src, dip_image* dst, const float* krow, int kw, const float* kcol,
                         int kh, const RowGeom& g) {
    if (!krow || !kcol || kw < 1 || kh < 1) return fail(ctx, DIP_ERR_ARG, "bad kernel");
    int rc = check_geom(ctx, src, dst, g);
    if (rc != DIP_OK) return rc;
    static const int b3[3] = {1, 2, 1};
    static const int b5[5] = {1, 4, 6, 4, 1};
    if (kw == 3 && kh == 3 && matches(krow, 3, b3, 4.0f) && matches(kcol, 3, b3, 4.0f))
        return launch_tiled_c<kConv3Sep>(ctx, src, dst, g);
    if (kw == 5 && kh == 5 && matches(krow, 5, b5, 16.0f) && matches(kcol, 5, b5, 16.0f))
        return launch_tiled_c<kConv5Sep>(ctx, src, dst, g);
    return fail(ctx, DIP_ERR_UNSUPPORTED,
                "fused separable convolution supports the binomial [1 2 1]/4 and [1 4 6 4 1]/16 kernels; "
                "run two dip_convolve passes for other kernels");
}

int stencil_gauss5_erode3(dip_ctx* ctx, const dip_image* src, dip_image* dst, const RowGeom& g) {
    int rc = check_geom(ctx, src, dst, g);
    if (rc != DIP_OK) return rc;
    const bool two = g_knobs.fused_rows == 2;  // development: 2-warp CTAs
    switch (src->v.channels) {
        case 1: return two ? launch_fused_g5e3<1, 15, 2>(ctx, src, dst, g) : launch_fused_g5e3<1, 15, 4>(ctx, src, dst, g);
        case 3: return two ? launch_fused_g5e3<3, 15, 2>(ctx, src, dst, g) : launch_fused_g5e3<3, 15, 4>(ctx, src, dst, g);
        case 4: return two ? launch_fused_g5e3<4, 15, 2>(ctx, src, dst, g) : launch_fused_g5e3<4, 15, 4>(ctx, src, dst, g);
    }
    return fail(ctx, DIP_ERR_ARG, "bad channel count");
}

int stencil_gaussian3(dip_ctx* ctx, const dip_image* src, dip_image* dst, const RowGeom& g) {
    int rc = check_geom(ctx, src, dst, g);
    if (rc != DIP_OK) return rc;
    return launch_tiled_c<kGauss3>(ctx, src, dst, g);
}

}  // namespace dip

using namespace dip;

extern "C" {

int dip_debug_set(const char* key, int value) {
    if (!key) return fail(nullptr, DIP_ERR_ARG, "null key");
    if (!strcmp(key, "extra_smem")) { g_knobs.extra_smem = value; return DIP_OK; }
    if (!strcmp(key, "tile_rows")) g_knobs.tile_rows = value;
    else if (!strcmp(key, "fused_rows")) g_knobs.fused_rows = value;
    else return fail(nullptr, DIP_ERR_ARG, "unknown knob %s", key);
    return DIP_OK;
}

int dip_erode(dip_ctx* ctx, const dip_image* src, dip_image* dst, const uint8_t* mask, int mw, int mh) {
    int rc = check_pair(ctx, src, dst, false);
    if (rc != DIP_OK) return rc;
    if ((rc = begin_op(ctx, src, dst)) != DIP_OK) return rc;
    return stencil_morph(ctx, Morph::Erode, src, dst, mask, mw, mh, full_geom(src));
}

int dip_dilate(dip_ctx* ctx, const dip_image* src, dip_image* dst, const uint8_t* mask, int mw, int mh) {
    int rc = check_pair(ctx, src, dst, false);
    if (rc != DIP_OK) return rc;
    if ((rc = begin_op(ctx, src, dst)) != DIP_OK) return rc;
    return stencil_morph(ctx, Morph::Dilate, src, dst, mask, mw, mh, full_geom(src));
}

int dip_erode_separable(dip_ctx* ctx, const dip_image* src, dip_image* dst, int mw, int mh) {
    int rc = check_pair(ctx, src, dst, false);
    if (rc != DIP_OK) return rc;
    if (mw < 1 || mh < 1 || mw * mh > 225) return fail(ctx, DIP_ERR_ARG, "bad separable mask %dx%d", mw, mh);
    if ((rc = begin_op(ctx, src, dst)) != DIP_OK) return rc;
    // A row minimum followed by a column minimum over all-ones masks is the minimum over the
    // mh x mw rectangle (the +inf border keeps this exact at the image edges), so the fused
    // kernel is the rectangle erosion: one pass, no intermediate image.
    std::vector<uint8_t> ones((size_t)mw * mh, 1);
    return stencil_morph(ctx, Morph::Erode, src, dst, ones.data(), mw, mh, full_geom(src));
}

int dip_convolve(dip_ctx* ctx, const dip_image* src, dip_image* dst, const float* kernel, int kw, int kh) {
    int rc = check_pair(ctx, src, dst, false);
    if (rc != DIP_OK) return rc;
    if ((rc = begin_op(ctx, src, dst)) != DIP_OK) return rc;
    return stencil_convolve(ctx, src, dst, kernel, kw, kh, full_geom(src));
}

int dip_convolve_separable(dip_ctx* ctx, const dip_image* src, dip_image* dst, const float* krow, int kw,
                           const float* kcol, int kh) {
    int rc = check_pair(ctx, src, dst, false);
    if (rc != DIP_OK) return rc;
    if ((rc = begin_op(ctx, src, dst)) != DIP_OK) return rc;
    return stencil_convolve_sep(ctx, src, dst, krow, kw, kcol, kh, full_geom(src));
}

int dip_gauss5x5_erode3x3(dip_ctx* ctx, const dip_image* src, dip_image* dst) {
    int rc = check_pair(ctx, src, dst, false);
    if (rc != DIP_OK) return rc;
    if ((rc = begin_op(ctx, src, dst)) != DIP_OK) return rc;
    return stencil_gauss5_erode3(ctx, src, dst, full_geom(src));
}

int dip_gaussian_blur3x3(dip_ctx* ctx, const dip_image* src, dip_image* dst) {
    int rc = check_pair(ctx, src, dst, false);
    if (rc != DIP_OK) return rc;
    if ((rc = begin_op(ctx, src, dst)) != DIP_OK) return rc;
    return stencil_gaussian3(ctx, src, dst, full_geom(src));
}

}  // extern "C"
