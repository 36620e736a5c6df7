// The tiled stencil kernels of libdip_b200.so (TMA tile in shared memory, packed u16x2 arithmetic in registers).
// Included by stencil.cu (plain kernels) and stencil_mapped.cu (the same kernels with a per-byte map applied to
// the tile before and / or to the result after the stencil, for dip_apply_chain); everything lives in an
// anonymous namespace, so each translation unit instantiates only what it launches.
#pragma once

#include <algorithm>
#include <atomic>

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
    // the packed-integer forms of kConv3Sep / kConv5Sep (the shipped ones compute in packed half precision, see
    // "packed-half arithmetic" below); kept for A/B measurements behind the development knob "int_math"
    kConv3SepI = 17,
    kConv5SepI = 18,
};

__host__ __device__ constexpr bool op_is_row_only(int op) { return op == kErodeRow3 || op == kDilateRow3 || op == kConvRow3 || op == kConvRow5; }
__host__ __device__ constexpr bool op_is_col_only(int op) { return op == kErodeCol3 || op == kDilateCol3 || op == kConvCol3 || op == kConvCol5; }
__host__ __device__ constexpr int op_taps(int op) { return (op == kConv5 || op == kConv5Sep || op == kConv5SepI || op == kConvRow5 || op == kConvCol5) ? 2 : 1; }
__host__ __device__ constexpr int op_radius_x(int op) { return op_is_col_only(op) ? 0 : op_taps(op); }
__host__ __device__ constexpr int op_radius_y(int op) { return op_is_row_only(op) ? 0 : op_taps(op); }
__host__ __device__ constexpr bool op_is_dilate(int op) { return op == kDilateSq3 || op == kDilateCross3 || op == kDilateRow3 || op == kDilateCol3; }
__host__ __device__ constexpr bool op_is_morph(int op) { return op <= kDilateCross3 || (op >= kErodeRow3 && op <= kDilateCol3); }

// Resident CTAs per SM the register allocation must allow (0 = whatever ptxas picks).  The packed-half 1x5+5x1 kernel
// comes out at 89 registers (5 CTAs of 4 warps) without the bound; it is issue-bound and wants the 6 the tile allows.
__host__ __device__ constexpr int op_min_ctas(int op) { return op == kConv5Sep ? 6 : 0; }  // 0 = no bound (a bound of 1 makes ptxas spend registers freely)
#ifndef DIP_EVEN_PRMT5
#define DIP_EVEN_PRMT5 8
#endif
constexpr int kEvenPrmt5 = DIP_EVEN_PRMT5;  // packed-half 1x5+5x1: words per row whose even bytes come from a PRMT
#ifndef DIP_CASCADE_HALF
#define DIP_CASCADE_HALF 6
#endif
constexpr int kCascadeHalf = DIP_CASCADE_HALF;  // packed-half 1x5+5x1: packed lanes (of 8) whose vertical pass is the all-half cascade

struct StencilParams {
    ImgView dst;
    int rowbytes;            // valid bytes per source row
    int width;               // pixels per row
    int valid_y0, valid_y1;  // source rows that exist for border purposes
    int out_y0, out_y1;      // rows to produce (source row coordinates)
    int dst_y0;              // dst row receiving out_y0
    int tiles_x, tiles_y;
    bool dst_exact;          // dst rows are not padded to 16 bytes (wrapped region of interest)
    bool wait_prev;          // execute griddepcontrol.wait (false: proven independent of every unfinished kernel, plan_launch)
    // Mosaic bands (dip_apply_rows_peer): source rows [peer_top_y0, peer_top_y1) / [peer_bot_y0, peer_bot_y1) are not in
    // this GPU's band buffer but in the neighbour's, which is mapped into this process (CUDA IPC, NVLink): the CTAs whose
    // tile covers such rows read them straight from the peer.  peer_* point at the peer byte that corresponds to byte 0
    // of source row 0 of frame 0 (same pitch and frame stride as the source), so the address arithmetic is the source's.
    const uint8_t* peer_top;
    const uint8_t* peer_bot;
    int peer_top_y0, peer_top_y1, peer_bot_y0, peer_bot_y1;
    size_t src_pitch, src_frame_stride;
    // MAPPED kernels (dip_apply_chain): per-byte map of the source words as they are unpacked / of the result words
    bool has_pre;
    MapParams pre, post;
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
// x * 2^k + y as ONE IMAD.  With an immediate power of two ptxas prefers LEA / IADD3 (ALU pipe) or IMAD.SHL + VIADD
// (two instructions) whatever the source says; a multiplier it cannot see keeps the multiply-add on the FMA pipe.
__constant__ uint32_t c_pow2[6] = {1u, 2u, 4u, 8u, 16u, 32u};
template <int LOG2>
__device__ __forceinline__ uint32_t shl_add(uint32_t x, uint32_t y) { return x * c_pow2[LOG2] + y; }

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
    h.v[2 * j] = padd(pair_at<4 * j - C>(r), padd(pair_at<4 * j + C>(r), shl_add<1>(r.e[2 + j], bias)));    \
    h.v[2 * j + 1] = padd(pair_at<4 * j + 1 - C>(r), padd(pair_at<4 * j + 1 + C>(r), shl_add<1>(r.o[2 + j], bias)));
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

// ---- packed-half arithmetic ---------------------------------------------------------------------------------------
// A u8 in a 16-bit lane IS the half-precision number b * 2^-24 (a subnormal), so the unpacked even / odd byte streams
// can be fed to HFMA2 as they are.  Two properties make the separable Gaussians exact in half precision:
//   * every integer of magnitude <= 2048 (in units of u = 2^-24) is a half, so a [1 4 6 4 1] sum (<= 4080) is exact once
//     it is centred: S' = S - 2048 = ((6 c - 2048) + 4 (b + d)) + (a + e), each step inside [-2048, 2047];
//   * the subnormal ulp is exactly u, so ONE fused multiply-add  S' * 2^-4 + 128 u  rounds S / 16 to an integer with
//     round-half-even -- the reference's cvRound -- and leaves the u8 result as the lane's bit pattern.
// That replaces the LOP3 + LEA.HI (+ LOP3) rounding sequences on the integer ALU pipe, which is what bounds these kernels
// (tools/ubench/half_pipes.cu: HFMA2 issues at the IMAD rate on the FMA pipe, also on subnormals, and the rounding
// identity holds for all 4096 sums; the small non-negative sums a + e, b + d stay integer adds on the ALU pipe).
__device__ __forceinline__ uint32_t hfma2(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t d;
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t hadd2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
__device__ __forceinline__ uint32_t hmul2(uint32_t a, uint32_t b) {
    uint32_t d;
    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
constexpr uint32_t kH2 = 0x40004000u;        // half2(2)
constexpr uint32_t kH4 = 0x44004400u;        // half2(4)
constexpr uint32_t kH6 = 0x46004600u;        // half2(6)
constexpr uint32_t kHQuarter = 0x34003400u;  // half2(1/4)
constexpr uint32_t kH16th = 0x2C002C00u;     // half2(1/16)
constexpr uint32_t kHNeg2048u = 0x88008800u; // half2(-2048 * 2^-24) = -2^-13
constexpr uint32_t kH128u = 0x00800080u;     // half2(128 * 2^-24): the bit pattern of the byte 128
// RNE((a + 4 b + 6 c + 4 d + e) / 16) per lane; ae = a + e and bd = b + d are plain integer sums (<= 510)
// `addend` = 128 u gives the u8 result as the lane's bit pattern; 0 gives the half (result - 128) u (see "vertical cascade")
__device__ __forceinline__ uint32_t half_gauss5(uint32_t ae, uint32_t bd, uint32_t c, uint32_t addend = kH128u) {
    const uint32_t s = hadd2(hfma2(bd, kH4, hfma2(c, kH6, kHNeg2048u)), ae);  // S - 2048, exact
    return hfma2(s, kH16th, addend);
}
// The same in FIVE instructions from the five byte patterns: S - 2048 = (a + e + 2 c - 2048) + 4 (b + c + d).  The first
// bracket needs no half arithmetic: for 0 <= t <= 1020 the half (t - 2048) u has the bit pattern 0x8800 - t (sign, exponent 1,
// mantissa 1024 - t: the binade [1024, 2048) u has ulp u), so K = 0x8800 - a - e - 2 c is one 3-input add with two negated
// operands and one multiply-add per packed pair (32-bit arithmetic is linear over the two lanes and every lane's final value is
// in range, so transient borrows cancel).  v = b + c + d <= 765 is one 3-input add and still a (subnormal) half; 4 v + K lies in
// [-2048, 2032] and is exact.  Used by the 2-warp stack shape only: under the 80-register cap of the 4-warp shapes it spills
// (one image per call 3.5 % slower, profiles/r2_cascade_ab.md); on stacks it is neutral in time and 4 % fewer instructions.
__device__ __forceinline__ uint32_t sub3(uint32_t x, uint32_t a, uint32_t b) {   // x - a - b (written in PTX: the front end re-associates the C expression)
    uint32_t r;
    asm("{\n.reg .u32 t;\nsub.u32 t, %1, %2;\nsub.u32 %0, t, %3;\n}" : "=r"(r) : "r"(x), "r"(a), "r"(b));
    return r;
}
__device__ __forceinline__ uint32_t half_gauss5_v2(uint32_t a, uint32_t e, uint32_t b, uint32_t d, uint32_t c, uint32_t addend) {
    uint32_t k;
    asm("mad.lo.u32 %0, %1, 0xFFFFFFFE, %2;" : "=r"(k) : "r"(c), "r"(sub3(kHNeg2048u, a, e)));   // kHNeg2048u = 0x88008800
    const uint32_t v = b + c + d;
    return hfma2(hfma2(v, kH4, k), kH16th, addend);
}

// RNE((a + 2 b + c) / 4) per lane; ac = a + c (<= 510)
__device__ __forceinline__ uint32_t half_gauss3(uint32_t ac, uint32_t b) { return hmul2(hfma2(b, kH2, ac), kHQuarter); }

// CENTRED: the first CENTRED packed results come out as halves (result - 128) u instead of byte patterns
// FIVE: the five-instruction form (half_gauss5_v2)
template <int C, int CENTRED = 0, bool FIVE = false>
__device__ __forceinline__ void hgauss5_half(const Row8& r, Acc8& h) {
    if constexpr (FIVE) {
#define DIP_H5V(j)                                                                                                   \
    h.v[2 * j] = half_gauss5_v2(pair_at<4 * j - 2 * C>(r), pair_at<4 * j + 2 * C>(r), pair_at<4 * j - C>(r), pair_at<4 * j + C>(r), r.e[2 + j], 2 * j < CENTRED ? 0u : kH128u); \
    h.v[2 * j + 1] = half_gauss5_v2(pair_at<4 * j + 1 - 2 * C>(r), pair_at<4 * j + 1 + 2 * C>(r), pair_at<4 * j + 1 - C>(r), pair_at<4 * j + 1 + C>(r), r.o[2 + j], 2 * j + 1 < CENTRED ? 0u : kH128u);
        DIP_H5V(0) DIP_H5V(1) DIP_H5V(2) DIP_H5V(3)
#undef DIP_H5V
        return;
    }
#define DIP_H5H(j)                                                                                                   \
    h.v[2 * j] = half_gauss5(pair_at<4 * j - 2 * C>(r) + pair_at<4 * j + 2 * C>(r), pair_at<4 * j - C>(r) + pair_at<4 * j + C>(r), r.e[2 + j], 2 * j < CENTRED ? 0u : kH128u); \
    h.v[2 * j + 1] = half_gauss5(pair_at<4 * j + 1 - 2 * C>(r) + pair_at<4 * j + 1 + 2 * C>(r), pair_at<4 * j + 1 - C>(r) + pair_at<4 * j + 1 + C>(r), r.o[2 + j], 2 * j + 1 < CENTRED ? 0u : kH128u);
    DIP_H5H(0) DIP_H5H(1) DIP_H5H(2) DIP_H5H(3)
#undef DIP_H5H
}
template <int C>
__device__ __forceinline__ void hgauss3_half(const Row8& r, Acc8& h) {
#define DIP_H3H(j)                                                                                   \
    h.v[2 * j] = half_gauss3(pair_at<4 * j - C>(r) + pair_at<4 * j + C>(r), r.e[2 + j]);               \
    h.v[2 * j + 1] = half_gauss3(pair_at<4 * j + 1 - C>(r) + pair_at<4 * j + 1 + C>(r), r.o[2 + j]);
    DIP_H3H(0) DIP_H3H(1) DIP_H3H(2) DIP_H3H(3)
#undef DIP_H3H
}

// PRE: the chain's map in front of a convolution is applied to the words as they are unpacked (halo words are
// mapped by both neighbours; cheaper than a pass over the tile with its barrier).  Erosion / dilation never carry one:
// launch_tiled_op moves it behind the op.
// EVEN_PRMT: how many of the 8 words get their even bytes by a second PRMT (integer ALU pipe) instead of the IMAD (FMA
// pipe) -- the packed-half kernels keep the FMA pipe busier than the ALU pipe, the integer ones the other way round.
template <bool PRE = false, int EVEN_PRMT = 0>
__device__ __forceinline__ void load_row(const uint8_t* rowp, Row8& r, const StencilParams* sp = nullptr) {
    const uint2 a = *reinterpret_cast<const uint2*>(rowp - 8);
    const uint4 b = *reinterpret_cast<const uint4*>(rowp);
    const uint2 c = *reinterpret_cast<const uint2*>(rowp + 16);
    uint32_t w[8] = {a.x, a.y, b.x, b.y, b.z, b.w, c.x, c.y};
    if constexpr (PRE) map_words(w, sp->pre);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        r.o[i] = odd_bytes(w[i]);
        r.e[i] = i < EVEN_PRMT ? __byte_perm(w[i], 0u, 0x4240) : even_bytes(w[i], r.o[i]);
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

// Tile rows that belong to a neighbour GPU's band: 16-byte loads from the peer mapping (ld.global.cv: never a stale
// line of an earlier step), stored over the tile rows the TMA load filled from the local buffer's unused halo rows.
// Columns outside [0, padded row) are zero like TMA's out-of-bounds fill; the column patch that follows reflects them.
template <int ROWS, int kThreads>
__device__ __forceinline__ void patch_peer_rows(uint8_t* tile, const StencilParams& p, int bx0, int ytop, int f) {
    const int row_padded = (p.rowbytes + 15) & ~15;
    constexpr int vec_per_row = kRowBytes / 16;
    auto pull = [&](const uint8_t* peer, int y0, int y1) {
        const int t0 = max(y0 - ytop, 0), t1 = min(y1 - ytop, ROWS);
        for (int i = threadIdx.x; i < (t1 - t0) * vec_per_row; i += kThreads) {
            const int t = t0 + i / vec_per_row, vx = i % vec_per_row;
            const int x = bx0 - kHalo + 16 * vx;
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (x >= 0 && x < row_padded)
                v = __ldcv(reinterpret_cast<const uint4*>(peer + (size_t)f * p.src_frame_stride + (size_t)(ytop + t) * p.src_pitch + x));
            *reinterpret_cast<uint4*>(tile + t * kRowBytes + 16 * vx) = v;
        }
    };
    if (p.peer_top && ytop < p.peer_top_y1 && ytop + ROWS > p.peer_top_y0) pull(p.peer_top, p.peer_top_y0, p.peer_top_y1);
    if (p.peer_bot && ytop < p.peer_bot_y1 && ytop + ROWS > p.peer_bot_y0) pull(p.peer_bot, p.peer_bot_y0, p.peer_bot_y1);
}
// block-uniform: does this tile hold rows that live on a neighbour GPU?
template <int ROWS>
__device__ __forceinline__ bool tile_has_peer_rows(const StencilParams& p, int ytop) {
    return (p.peer_top && ytop < p.peer_top_y1 && ytop + ROWS > p.peer_top_y0) ||
           (p.peer_bot && ytop < p.peer_bot_y1 && ytop + ROWS > p.peer_bot_y0);
}

// ---- the arithmetic of one warp band: the thread walks down its 16-byte column --------------------
// FULL = the band has all R rows (compile-time trip count: the row loops unroll by the period of the
// register window without a remainder loop); otherwise `nrows_rt` rows.
// CLIP = this warp holds the row's last, partial 16-byte chunk of a destination whose rows are NOT padded (a wrapped
// region of interest): its stores write only the valid bytes.  Every other warp keeps plain vector stores.
template <bool CLIP, bool MAPPED>
__device__ __forceinline__ void put16(uint8_t* p, uint4 v, int valid, const StencilParams& sp) {
    if constexpr (MAPPED) v = map_vec(v, sp.post);
    if constexpr (CLIP) store16(p, v, valid);
    else __stcs(reinterpret_cast<uint4*>(p), v);
}

// MAPPED: `post` is applied to the result words; PRE (only with MAPPED): `pre` is applied to the source words.  The
// caller picks PRE outside the row loops: a uniform test between a row's shared-memory loads and their first use
// costs far more than its instructions (measured +1.5 us per frame on the 5-tap kernels).
template <int OP, int C, int R, bool FULL, bool CLIP, bool MAPPED, bool PRE>
__device__ __forceinline__ void compute_band(const uint8_t* colp, uint8_t* dptr, const StencilParams& p, int nrows_rt, int p_cx) {
    const int nrows = FULL ? R : nrows_rt;
    const int valid = CLIP ? p.rowbytes - p_cx : 16;  // bytes of this thread's 16-byte chunk that may be stored
    if constexpr (op_is_row_only(OP)) {
        // one-dimensional pass along x: no vertical window
        Row8 row;
        Acc8 h;
#pragma unroll 3
        for (int i = 0; i < nrows; ++i) {
            load_row<PRE && !op_is_morph(OP)>(colp + i * kRowBytes, row, &p);
            uint32_t out[4];
            if constexpr (op_is_morph(OP)) {
                hmnmx3<C, op_is_dilate(OP)>(row, h);
#pragma unroll
                for (int j = 0; j < 4; ++j) out[j] = pack_lo(h.v[2 * j], h.v[2 * j + 1]);
            } else {
                if constexpr (OP == kConvRow3) hgauss3_half<C>(row, h);
                else hgauss5_half<C>(row, h);
#pragma unroll
                for (int j = 0; j < 4; ++j) out[j] = pack_lo(h.v[2 * j], h.v[2 * j + 1]);
            }
            put16<CLIP, MAPPED>(dptr + (size_t)i * p.dst.pitch, make_uint4(out[0], out[1], out[2], out[3]), valid, p);
        }
    } else if constexpr (op_is_col_only(OP)) {
        // one-dimensional pass along y: only the thread's own 4 words, vertical window of raw rows
        constexpr int taps = op_taps(OP);
        Acc8 w0, w1, w2, w3, w4;
        auto load_own = [&](const uint8_t* rp, Acc8& a) {
            const uint4 b = *reinterpret_cast<const uint4*>(rp);
            uint32_t w[4] = {b.x, b.y, b.z, b.w};
            if constexpr (PRE && !op_is_morph(OP)) map_words(w, p.pre);
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
                    else res[k] = half_gauss3(w0.v[k] + w2.v[k], w1.v[k]);
                }
                w0 = w1; w1 = w2;
            } else {
                load_own(colp + (i + 4) * kRowBytes, w4);
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    res[k] = half_gauss5(w0.v[k] + w4.v[k], w1.v[k] + w3.v[k], w2.v[k]);
                w0 = w1; w1 = w2; w2 = w3; w3 = w4;
            }
            uint4 o;
            o = make_uint4(pack_lo(res[0], res[1]), pack_lo(res[2], res[3]), pack_lo(res[4], res[5]), pack_lo(res[6], res[7]));
            put16<CLIP, MAPPED>(dptr + (size_t)i * p.dst.pitch, o, valid, p);
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
            put16<CLIP, MAPPED>(dptr + (size_t)i * p.dst.pitch, make_uint4(out[0], out[1], out[2], out[3]), valid, p);
        }
    } else if constexpr (OP == kConv3Sep) {
        // RNE([1 2 1]/4) rows -> u8 -> RNE([1 2 1]/4) columns in packed half precision: h* hold the u8 rows i-1, i, i+1
        Acc8 h0, h1, h2;
        Row8 row;
        auto hpass = [&](const uint8_t* rp, Acc8& h) {
            load_row<PRE>(rp, row, &p);
            hgauss3_half<C>(row, h);
        };
        hpass(colp, h0);
        hpass(colp + kRowBytes, h1);
#pragma unroll 3
        for (int i = 0; i < nrows; ++i) {
            hpass(colp + (i + 2) * kRowBytes, h2);
            uint32_t res[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) res[k] = half_gauss3(h0.v[k] + h2.v[k], h1.v[k]);
            put16<CLIP, MAPPED>(dptr + (size_t)i * p.dst.pitch,
                    make_uint4(pack_lo(res[0], res[1]), pack_lo(res[2], res[3]), pack_lo(res[4], res[5]), pack_lo(res[6], res[7])), valid, p);
            h0 = h1; h1 = h2;
        }
    } else if constexpr (OP == kConv5Sep) {
        // RNE([1 4 6 4 1]/16) rows -> u8 -> RNE([1 4 6 4 1]/16) columns in packed half precision.
        // Vertical cascade: [1 4 6 4 1] = [1 1]^4, so with the rounded rows h_k kept as running pair sums
        //   p_k = h_k + h_k+1,  v_k = p_k + p_k+1,  q_k = v_k + v_k+1,  o_k = q_k + q_k+1  (= the column sum of output row k+2)
        // a new row costs 4 additions per packed lane, the state is 4 registers per lane (H = h, P = p, V = v, Q = q of
        // the newest complete index) and nothing is multiplied.  o reaches 4080 and must be centred to be a half:
        //   * lanes < kCascadeHalf take the rows as halves (h - 128) u -- the horizontal pass's rounding FMA gives that
        //     for free with a zero addend (RNE(x) - 128 = RNE(x - 128)) -- so p, v, q, o are exact halves in
        //     [-2048, 2047] u: 4 HADD2 + the rounding HFMA2, all on the FMA pipe;
        //   * the other lanes keep byte patterns, add p, v, q (<= 2040: still bit patterns of halves) on the integer ALU
        //     pipe and centre q with one HADD2: 3 IADD + 2 HADD2 + HFMA2, for the balance between the two pipes.
        // (The window form before this was 2 IADD + 3 HFMA2/HADD2 + the rounding HFMA2 per lane.)
        constexpr bool kFive = R == 30;               // the 2-warp stack shape (launch_tiled_c): five-instruction horizontal pass, all lanes on the half cascade
        constexpr int NH = kFive ? 8 : kCascadeHalf;
        Acc8 H, P, V, Q, h;
        Row8 row;
        auto hpass = [&](const uint8_t* rp, Acc8& o) {
            load_row<PRE, kEvenPrmt5>(rp, row, &p);
            hgauss5_half<C, NH, kFive>(row, o);
        };
        auto add = [&](int k, uint32_t a, uint32_t b) { return k < NH ? hadd2(a, b) : a + b; };
        hpass(colp, H);
        hpass(colp + kRowBytes, h);
#pragma unroll
        for (int k = 0; k < 8; ++k) { P.v[k] = add(k, H.v[k], h.v[k]); H.v[k] = h.v[k]; }
        hpass(colp + 2 * kRowBytes, h);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t pn = add(k, H.v[k], h.v[k]);
            V.v[k] = add(k, P.v[k], pn); P.v[k] = pn; H.v[k] = h.v[k];
        }
        hpass(colp + 3 * kRowBytes, h);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t pn = add(k, H.v[k], h.v[k]);
            const uint32_t vn = add(k, P.v[k], pn);
            const uint32_t qn = add(k, V.v[k], vn);
            Q.v[k] = k < NH ? qn : hadd2(qn, kHNeg2048u);  // integer lanes keep Q centred
            V.v[k] = vn; P.v[k] = pn; H.v[k] = h.v[k];
        }
#pragma unroll 5
        for (int i = 0; i < nrows; ++i) {
            hpass(colp + (i + 4) * kRowBytes, h);
            uint32_t res[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const uint32_t pn = add(k, H.v[k], h.v[k]);
                const uint32_t vn = add(k, P.v[k], pn);
                const uint32_t qn = add(k, V.v[k], vn);
                res[k] = hfma2(hadd2(Q.v[k], qn), kH16th, kH128u);
                Q.v[k] = k < NH ? qn : hadd2(qn, kHNeg2048u);
                V.v[k] = vn; P.v[k] = pn; H.v[k] = h.v[k];
            }
            put16<CLIP, MAPPED>(dptr + (size_t)i * p.dst.pitch,
                    make_uint4(pack_lo(res[0], res[1]), pack_lo(res[2], res[3]), pack_lo(res[4], res[5]), pack_lo(res[6], res[7])), valid, p);
        }
    } else if constexpr (op_taps(OP) == 1) {
        // 3-tap Gaussians.  h* hold the horizontal pass of rows i-1, i, i+1.
        const uint32_t hbias = OP == kConv3SepI ? reg_const<0x00010001u>() : 0u;
        const uint32_t vbias = reg_const<(OP == kGauss3 ? 0x00800080u : 0x007F007Fu)>();
        Acc8 h0, h1, h2;
        Row8 row;
        auto hpass = [&](const uint8_t* rp, Acc8& h) {
            load_row<PRE>(rp, row, &p);
            hsum3<C>(row, h, hbias);
            if constexpr (OP == kConv3SepI) {
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
                const uint32_t x = shl_add<4>(u, shl_add<5>(h1.v[k], vbias));
                if constexpr (OP == kGauss3) res[k] = x;                     // (S + 8) >> 4
                else res[k] = rne_finish(x);
            }
            put16<CLIP, MAPPED>(dptr + (size_t)i * p.dst.pitch,
                    make_uint4(pack_hi(res[0], res[1]), pack_hi(res[2], res[3]), pack_hi(res[4], res[5]), pack_hi(res[6], res[7])), valid, p);
            h0 = h1; h1 = h2;
        }
    } else {
        // 5-tap Gaussians.  h* hold the horizontal pass of rows i-2 .. i+2.
        const uint32_t hbias = OP == kConv5SepI ? reg_const<0x00070007u>() : 0u;
        const uint32_t vbias = reg_const<0x007F007Fu>();
        Acc8 h0, h1, h2, h3, h4;
        Row8 row;
        auto hpass = [&](const uint8_t* rp, Acc8& h) {
            load_row<PRE>(rp, row, &p);
            hsum5<C>(row, h, hbias);
            if constexpr (OP == kConv5SepI) {
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
                const uint32_t x = OP == kConv5SepI
                                       ? padd(h0.v[k], padd(h4.v[k], 4u * padd(h1.v[k], h3.v[k]) + (6u * h2.v[k] + vbias)))
                                       : (h0.v[k] + h4.v[k]) + (4u * padd(h1.v[k], h3.v[k]) + (6u * h2.v[k] + vbias));
                res[k] = rne_finish(x);
            }
            put16<CLIP, MAPPED>(dptr + (size_t)i * p.dst.pitch,
                    make_uint4(pack_hi(res[0], res[1]), pack_hi(res[2], res[3]), pack_hi(res[4], res[5]), pack_hi(res[6], res[7])), valid, p);
            h0 = h1; h1 = h2; h2 = h3; h3 = h4;
        }
    }
}

// Everything after the tile has landed in shared memory: border patching, then each warp's band.  Block-uniform
// control flow around the barriers (threads without a column / band skip the arithmetic, not the function).
template <int OP, int C, int R, int kWarps, bool MAPPED>
__device__ __forceinline__ void process_tile(uint8_t* tile, const StencilParams& p, int tx, int ty, int f) {
    constexpr int r = op_radius_y(OP);   // halo rows
    constexpr int rx = op_radius_x(OP);  // halo pixels
    constexpr int kThreads = kWarps * 32;
    constexpr int TH = kWarps * R;
    constexpr int ROWS = TH + 2 * r;
    const int bx0 = tx * kTileBytes;
    const int y0 = p.out_y0 + ty * TH;  // first output row of the tile
    const int ytop = y0 - r;            // source row held by tile row 0
    const bool border = (rx > 0 && (bx0 == 0 || bx0 + kTileBytes + rx * C > p.rowbytes)) ||
                        (r > 0 && (ytop < p.valid_y0 || ytop + ROWS > p.valid_y1));
    if (tile_has_peer_rows<ROWS>(p, ytop)) {
        __syncthreads();
        patch_peer_rows<ROWS, kThreads>(tile, p, bx0, ytop, f);
        if (!border) __syncthreads();
    }
    if (border) {
        __syncthreads();
        patch_borders<OP, C, ROWS, kThreads>(tile, p, bx0, ytop);
    }

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int cx = bx0 + 16 * lane;  // first byte of this thread's column
    const int yw = y0 + warp * R;    // first output row of this warp
    if (cx >= ((p.rowbytes + 15) & ~15) || yw >= p.out_y1) return;
    const int nrows = min(R, p.out_y1 - yw);
    const uint8_t* colp = tile + (warp * R) * kRowBytes + kHalo + 16 * lane;  // tile row of source row yw - r
    uint8_t* dptr = p.dst.base + (size_t)f * p.dst.frame_stride + (size_t)(p.dst_y0 + (yw - p.out_y0)) * p.dst.pitch + cx;

    // warp-uniform: does this tile column contain the partial last chunk of an unpadded destination row?
    const bool clip = p.dst_exact && bx0 + kTileBytes > p.rowbytes && (p.rowbytes & 15) != 0;
    if (MAPPED && !op_is_morph(OP) && p.has_pre) {
        if (clip) compute_band<OP, C, R, false, true, MAPPED, MAPPED && !op_is_morph(OP)>(colp, dptr, p, nrows, cx);
        else if (nrows == R) compute_band<OP, C, R, true, false, MAPPED, MAPPED && !op_is_morph(OP)>(colp, dptr, p, R, cx);
        else compute_band<OP, C, R, false, false, MAPPED, MAPPED && !op_is_morph(OP)>(colp, dptr, p, nrows, cx);
        return;
    }
    if (clip) compute_band<OP, C, R, false, true, MAPPED, false>(colp, dptr, p, nrows, cx);
    else if (nrows == R) compute_band<OP, C, R, true, false, MAPPED, false>(colp, dptr, p, R, cx);
    else compute_band<OP, C, R, false, false, MAPPED, false>(colp, dptr, p, nrows, cx);
}

// ---- the tiled kernel -------------------------------------------------------------------------
template <int OP, int C, int R, int kWarps, bool MAPPED>
__global__ void __launch_bounds__(kWarps * 32, op_min_ctas(OP)) stencil_kernel(const __grid_constant__ CUtensorMap tmap, const StencilParams p) {
    constexpr int r = op_radius_y(OP);   // halo rows
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
    pdl_wait(p.wait_prev);  // the previous kernels' writes are visible from here on
    if (threadIdx.x == 0) tma_load_3d(tile, &tmap, bar, (bx0 - kHalo) / 4, ytop, f);
    __syncthreads();  // barrier init visible to the waiters
    mbar_wait(bar, 0);

    process_tile<OP, C, R, kWarps, MAPPED>(tile, p, tx, ty, f);
}

bool g_no_occupancy_pad = false;  // development knob "no_pad"

// a tiled launch reads the source view and writes the destination view; rows that come from a neighbour GPU are memory
// this process knows nothing about (ordered by the caller's flags / barriers), so such a launch is never a no-wait one
inline LaunchMode plan_tiled_launch(dip_ctx* ctx, const dip_image* src, const dip_image* dst, const RowGeom& g) {
    if (g.peer_top || g.peer_bot) return plan_launch(ctx, nullptr, 0);
    return plan_launch(ctx, src->v, dst->v);
}

// fills the source-side fields every tiled launch shares
inline void fill_peer_params(StencilParams& p, const dip_image* src, const RowGeom& g) {
    p.src_pitch = src->v.pitch;
    p.src_frame_stride = src->v.frame_stride;
    p.peer_top = g.peer_top; p.peer_top_y0 = g.peer_top_y0; p.peer_top_y1 = g.peer_top_y1;
    p.peer_bot = g.peer_bot; p.peer_bot_y0 = g.peer_bot_y0; p.peer_bot_y1 = g.peer_bot_y1;
}

template <int OP, int C, int R, int kWarps, bool MAPPED>
int launch_tiled(dip_ctx* ctx, const dip_image* src, dip_image* dst, const RowGeom& g, const ChainMaps* maps) {
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
    const int smem_bytes = (kPadded > kTileSmem && !g_no_occupancy_pad) ? kPadded : kTileSmem;
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
    fill_peer_params(p, src, g);
    if (MAPPED && maps) {
        p.has_pre = !maps->pre.identity();
        p.pre = map_params(maps->pre);
        p.post = map_params(maps->post);
    }
    if (p.tiles_x <= 0 || p.tiles_y <= 0) return DIP_OK;
    if (p.tiles_y > 65535 || src->v.frames > 65535) return fail(ctx, DIP_ERR_UNSUPPORTED, "image stack too large for one launch");
    static std::atomic<uint64_t> attr_set{0};  // per instantiation, one bit per device
    const uint64_t dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_set.load(std::memory_order_relaxed) & dev_bit)) {
        // the limit covers both settings of the "no_pad" knob, whichever the first launch used
        DIP_CUDA(ctx, cudaFuncSetAttribute(stencil_kernel<OP, C, R, kWarps, MAPPED>, cudaFuncAttributeMaxDynamicSharedMemorySize, std::max(kPadded, kTileSmem)));
        attr_set.fetch_or(dev_bit, std::memory_order_relaxed);
    }
    const LaunchMode mode = plan_tiled_launch(ctx, src, dst, g);
    p.wait_prev = mode != kNoWait;
    DIP_CUDA(ctx, launch_kernel(mode, stencil_kernel<OP, C, R, kWarps, MAPPED>, dim3((unsigned)p.tiles_x, (unsigned)p.tiles_y, (unsigned)src->v.frames), dim3(kThreads), smem_bytes, ctx->compute, *map, p));
    DIP_CHECK_LAUNCH(ctx, "stencil_kernel");
    return DIP_OK;
}

}  // namespace

}  // namespace dip
