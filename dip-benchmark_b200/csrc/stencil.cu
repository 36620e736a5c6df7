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
#include "stencil_tiled.cuh"

namespace dip {

namespace {

// Development knobs (dip_debug_set): "tile_rows" 0 = automatic, 64 or 32 forces the tile height.
// "int_math" 1 = the packed-integer forms of the two separable Gaussians instead of the packed-half ones (A/B only).
// "direct_generic" 1 = arbitrary small masks / kernels through the direct kernel instead of the tiled generic one (A/B only).
// "umma" 1 = the fused 1x5+5x1 Gaussian on the tensor cores (gauss_umma.cu: tcgen05.mma.kind::i8, bit-exact, but 15.9 us
// per frame against 10.4 for the packed-half SIMD kernel so far -- profiles/r2_umma_experiment.md) instead of the SIMD kernel.
// "fused_shape": band rows x warps (and a register cap) of the fused Gaussian -> erosion kernel on one-channel images (A/B only).
// "band_shape": rows per warp x warps per CTA of the 5-tap tiled kernels on 3-channel stacks (A/B only): 0 = 15 x 4.
struct Knobs { int tile_rows = 0; int int_math = 0; int direct_generic = 0; int umma = 0; int fused_shape = 0; int band_shape = 0; };
Knobs g_knobs;

// One warp band of the fused Gaussian -> erosion kernel below.  EDGE: the tile touches the image rim, so Gaussian
// values outside the image count as 255 (+inf) for the erosion; CLIP: partial last chunk of an unpadded row.
template <int C, int R, bool EDGE, bool CLIP, bool FULL, bool MAPPED, bool PRE, int kUnroll = 5>
__device__ __forceinline__ void fused_band(const uint8_t* colp, const StencilParams& p, int cx, int yw, int f, int lane) {
    const int n_out = FULL ? R : min(R, p.out_y1 - yw);  // FULL: compile-time trip count (no remainder loop)
    // Gaussian bytes outside [0, rowbytes) count as 255 for the erosion: per-lane masks in packed form
    uint32_t colmask[8];
    if constexpr (EDGE) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int b0 = cx + 4 * j;
            auto out = [&](int pos) { return (pos < 0 || pos >= p.rowbytes) ? 0xFF00u : 0u; };
            colmask[2 * j] = out(b0) | (out(b0 + 2) << 16);
            colmask[2 * j + 1] = out(b0 + 1) | (out(b0 + 3) << 16);
        }
    }
    const bool store_lane = lane != 0 && lane != 31 && cx < ((p.rowbytes + 15) & ~15);
    const int valid = CLIP ? p.rowbytes - cx : 16;
    uint8_t* dptr = p.dst.base + (size_t)f * p.dst.frame_stride + (size_t)(p.dst_y0 + (yw - p.out_y0)) * p.dst.pitch + cx;
    const uint32_t vbias = reg_const<0x007F007Fu>();

    Acc8 h0, h1, h2, h3, h4;  // horizontal Gaussian passes of five consecutive source rows
    Acc8 m0, m1;              // horizontal minima of the two previous Gaussian rows
    Row8 row;
    auto hpass = [&](const uint8_t* rp, Acc8& h) {
        load_row<PRE>(rp, row, &p);  // PRE: the chain's map in front of the Gaussian, applied as the words are unpacked
        hsum5<C>(row, h, 0u);
    };
    // Gaussian row g (image row yw - 1 + g) from the window, then its horizontal 3-minimum
    auto gauss_hmin = [&](int g, Acc8& hm) {
        uint32_t rowmask = 0u;
        if constexpr (EDGE) {
            const int yg = yw - 1 + g;
            rowmask = (yg >= p.valid_y0 && yg < p.valid_y1) ? 0u : 0xFF00FF00u;
        }
        Row8 gr;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            // one ALU-pipe add, three IMADs (this kernel is ALU-pipe bound: PRMT funnels, VIMNMX3, rounding)
            uint32_t v = rne_finish((h0.v[k] + h4.v[k]) + (4u * padd(h1.v[k], h3.v[k]) + (6u * h2.v[k] + vbias)));
            if constexpr (EDGE) v |= colmask[k] | rowmask;
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
#pragma unroll kUnroll
    for (int i = 0; i < n_out; ++i) {
        hpass(colp + (i + 6) * kRowBytes, h4);
        Acc8 m2;
        gauss_hmin(i + 2, m2);
        if (store_lane) {
            uint32_t out[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                out[j] = pack_hi(mnmx3<false>(m0.v[2 * j], m1.v[2 * j], m2.v[2 * j]), mnmx3<false>(m0.v[2 * j + 1], m1.v[2 * j + 1], m2.v[2 * j + 1]));
            uint4 o = make_uint4(out[0], out[1], out[2], out[3]);
            if constexpr (MAPPED) o = map_vec(o, p.post);  // the chain's map behind the erosion
            if constexpr (CLIP) store16(dptr + (size_t)i * p.dst.pitch, o, valid);
            else __stcs(reinterpret_cast<uint4*>(dptr + (size_t)i * p.dst.pitch), o);
        }
        m0 = m1; m1 = m2;
        h0 = h1; h1 = h2; h2 = h3; h3 = h4;
    }
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
template <int C, int R, int kWarps, bool MAPPED, int kMinCtas = 0, int kUnroll = 5>
__global__ void __launch_bounds__(kWarps * 32, kMinCtas) fused_g5e3_kernel(const __grid_constant__ CUtensorMap tmap, const StencilParams p) {
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
    pdl_wait(p.wait_prev);
    if (threadIdx.x == 0) tma_load_3d(tile, &tmap, bar, (gx0 - kHalo) / 4, ytop, f);
    __syncthreads();
    mbar_wait(bar, 0);
    const bool border = gx0 <= 0 || gx0 + kTileBytes + 2 * C > p.rowbytes || ytop < p.valid_y0 || ytop + ROWS > p.valid_y1;
    if (tile_has_peer_rows<ROWS>(p, ytop)) {  // rows of a neighbour GPU's band: read over NVLink, no exchange step
        __syncthreads();
        patch_peer_rows<ROWS, kThreads>(tile, p, gx0, ytop, f);
        if (!border) __syncthreads();
    }
    if (border) {
        __syncthreads();
        patch_borders<kConv5, C, ROWS, kThreads>(tile, p, gx0, ytop);
    }

    const int cx = gx0 + 16 * lane;       // first byte of this thread's column
    const int yw = y0 + warp * R;         // first output row of this warp
    if (yw >= p.out_y1) return;           // whole warp
    const uint8_t* colp = tile + (warp * R) * kRowBytes + kHalo + 16 * lane;  // source row yw - 3
    const bool clip = p.dst_exact && gx0 + kTileBytes > p.rowbytes && (p.rowbytes & 15) != 0;  // block-uniform, implies border
    // interior tiles (all but the image's rim) run the variant without the out-of-image masks
    if constexpr (MAPPED) {  // chains: two general variants per choice of PRE (the test stays outside the row loops)
        if (p.has_pre) {
            if (clip) fused_band<C, R, true, true, false, true, true>(colp, p, cx, yw, f, lane);
            else fused_band<C, R, true, false, false, true, true>(colp, p, cx, yw, f, lane);
        } else {
            if (clip) fused_band<C, R, true, true, false, true, false>(colp, p, cx, yw, f, lane);
            else fused_band<C, R, true, false, false, true, false>(colp, p, cx, yw, f, lane);
        }
    } else {
        if (clip) fused_band<C, R, true, true, false, false, false>(colp, p, cx, yw, f, lane);
        else if (border) fused_band<C, R, true, false, false, false, false>(colp, p, cx, yw, f, lane);
        else if (yw + R <= p.out_y1) fused_band<C, R, false, false, true, false, false, kUnroll>(colp, p, cx, yw, f, lane);
        else fused_band<C, R, false, false, false, false, false>(colp, p, cx, yw, f, lane);
    }
}

template <int C, int R, int kWarps, bool MAPPED, int kMinCtas = 0, int kUnroll = 5>
int launch_fused_g5e3(dip_ctx* ctx, const dip_image* src, dip_image* dst, const RowGeom& g, const ChainMaps* maps) {
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
    fill_peer_params(p, src, g);
    if (MAPPED && maps) {
        p.has_pre = !maps->pre.identity();
        p.pre = map_params(maps->pre);
        p.post = map_params(maps->post);
    }
    if (p.tiles_x <= 0 || p.tiles_y <= 0) return DIP_OK;
    if (p.tiles_y > 65535 || src->v.frames > 65535) return fail(ctx, DIP_ERR_UNSUPPORTED, "image stack too large for one launch");
    static std::atomic<uint64_t> attr_set{0};
    const uint64_t dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_set.load(std::memory_order_relaxed) & dev_bit)) {
        DIP_CUDA(ctx, cudaFuncSetAttribute(fused_g5e3_kernel<C, R, kWarps, MAPPED, kMinCtas, kUnroll>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
        attr_set.fetch_or(dev_bit, std::memory_order_relaxed);
    }
    const LaunchMode mode = plan_tiled_launch(ctx, src, dst, g);
    p.wait_prev = mode != kNoWait;
    DIP_CUDA(ctx, launch_kernel(mode, fused_g5e3_kernel<C, R, kWarps, MAPPED, kMinCtas, kUnroll>, dim3((unsigned)p.tiles_x, (unsigned)p.tiles_y, (unsigned)src->v.frames),
                                dim3(kWarps * 32), smem_bytes, ctx->compute, *map, p));
    DIP_CHECK_LAUNCH(ctx, "fused_g5e3_kernel");
    return DIP_OK;
}


template <int OP>
int launch_tiled_c(dip_ctx* ctx, const dip_image* src, dip_image* dst, const RowGeom& g) {
    switch (src->v.channels) {
        case 1: return launch_tiled<OP, 1, 15, 4, false>(ctx, src, dst, g, nullptr);
        case 4: return launch_tiled<OP, 4, 15, 4, false>(ctx, src, dst, g, nullptr);
        case 3: {
            // Few tiles (a single image): halve the tile height so the launch is about two waves and
            // late CTAs' loads overlap early CTAs' arithmetic; stacks keep the 64-row tile.
            const long long tiles64 = (long long)((src->v.rowbytes + kTileBytes - 1) / kTileBytes) * ((g.out_y1 - g.out_y0 + 63) / 64) * src->v.frames;
            const bool small = g_knobs.tile_rows ? g_knobs.tile_rows < 60 : tiles64 < 2ll * 6 * ctx->sm_count;
            // multiples of the window period (5 / 3 rows); measured best for one 3504x2336 image: 40- and 24-row tiles
            constexpr int kSmallR = op_taps(OP) == 2 ? 10 : 6;
            if constexpr (OP == kConv5Sep || OP == kConv5) {
                // Stacks run in the sustained power state (1 kW cap), where time follows the instructions executed, not
                // the resident warps: 30-row bands on 2-warp CTAs (same 64-row tile) recompute 4 horizontal passes per 30
                // rows instead of per 15 and are 3 % faster for the 1x5+5x1 kernel there (tools/sustained_op.py,
                // profiles/r2_sustained_shapes.md) although slower in a cold burst; 5x5 is best at 15 x 4.
                if (!small && OP == kConv5Sep && g_knobs.band_shape == 0) return launch_tiled<OP, 3, 30, 2, false>(ctx, src, dst, g, nullptr);
                if (!small) switch (g_knobs.band_shape) {
                    case 1: return launch_tiled<OP, 3, 30, 2, false>(ctx, src, dst, g, nullptr);
                    case 7: return launch_tiled<OP, 3, 15, 4, false>(ctx, src, dst, g, nullptr);
                    case 2: return launch_tiled<OP, 3, 30, 4, false>(ctx, src, dst, g, nullptr);
                    case 3: return launch_tiled<OP, 3, 45, 2, false>(ctx, src, dst, g, nullptr);
                    case 4: return launch_tiled<OP, 3, 60, 2, false>(ctx, src, dst, g, nullptr);
                    case 5: return launch_tiled<OP, 3, 20, 3, false>(ctx, src, dst, g, nullptr);
                    case 6: return launch_tiled<OP, 3, 60, 1, false>(ctx, src, dst, g, nullptr);
                    default: break;
                }
            }
            return small ? launch_tiled<OP, 3, kSmallR, 4, false>(ctx, src, dst, g, nullptr) : launch_tiled<OP, 3, 15, 4, false>(ctx, src, dst, g, nullptr);
        }
    }
    return fail(ctx, DIP_ERR_ARG, "bad channel count");
}

// the tiled kernel `op` (a StencilOp), plain or with the chain's per-byte maps fused in
int flip_morph(int op) {
    switch (op) {
        case kErodeSq3: return kDilateSq3;
        case kDilateSq3: return kErodeSq3;
        case kErodeCross3: return kDilateCross3;
        case kDilateCross3: return kErodeCross3;
        case kErodeRow3: return kDilateRow3;
        case kDilateRow3: return kErodeRow3;
        case kErodeCol3: return kDilateCol3;
        case kDilateCol3: return kErodeCol3;
    }
    return op;
}

int launch_tiled_op(dip_ctx* ctx, int op, const dip_image* src, dip_image* dst, const RowGeom& g, const ChainMaps* maps) {
    ChainMaps moved;
    if (maps && op_is_morph(op) && !maps->pre.identity()) {
        // Every ByteMap is monotone and min / max commute with monotone maps (out-of-image taps are absent, not a
        // value): min f(x_i) = f(min x_i) for a non-decreasing f and f(max x_i) for a non-increasing one.  So a map
        // in front of an erosion / dilation moves behind it -- flipping the op if it decreases -- and costs no
        // pass over the tile in shared memory.
        const ByteMap& f = maps->pre;
        if (f.kind == 1 || (f.kind == 2 && f.hi < f.lo)) op = flip_morph(op);
        moved.post = f.then(maps->post);
        maps = &moved;
    }
    if (maps && !(maps->pre.identity() && maps->post.identity())) return launch_tiled_mapped(ctx, op, src, dst, g, *maps);
    if (op == kConv5Sep && g_knobs.umma && !g_knobs.int_math && umma_gauss5sep_ok(src, dst, g)) return launch_gauss5sep_umma(ctx, src, dst, g);
    if (g_knobs.int_math) op = op == kConv3Sep ? kConv3SepI : op == kConv5Sep ? kConv5SepI : op;
    switch (op) {
#define DIP_T(OP) case OP: return launch_tiled_c<OP>(ctx, src, dst, g);
        DIP_T(kErodeSq3) DIP_T(kErodeCross3) DIP_T(kDilateSq3) DIP_T(kDilateCross3) DIP_T(kConv3) DIP_T(kConv3Sep) DIP_T(kGauss3)
        DIP_T(kConv5) DIP_T(kConv5Sep) DIP_T(kErodeRow3) DIP_T(kErodeCol3) DIP_T(kDilateRow3) DIP_T(kDilateCol3) DIP_T(kConvRow3)
        DIP_T(kConvCol3) DIP_T(kConvRow5) DIP_T(kConvCol5) DIP_T(kConv3SepI) DIP_T(kConv5SepI)
#undef DIP_T
    }
    return fail(ctx, DIP_ERR_ARG, "bad tiled op %d", op);
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
    DIP_CUDA(ctx, launch_kernel(plan_launch(ctx, nullptr, 0), generic_kernel<MODE>, dim3((unsigned)blocks), dim3(256), 0, ctx->compute, p));
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

// can this view be a TMA tensor (16-byte aligned base, pitch and frame stride)?  Library-owned images always can.
bool tma_ok(const dip_image* im) {
    return !((reinterpret_cast<uintptr_t>(im->v.base) & 15) || (im->v.pitch & 15) || (im->v.frame_stride & 15));
}

// weights * scale are integers matching `pattern`?
bool matches(const float* k, int n, const int* pattern, float scale) {
    for (int i = 0; i < n; ++i)
        if (k[i] * scale != (float)pattern[i]) return false;
    return true;
}

// Which tiled kernel computes this mask / kernel (a StencilOp), or -1 for the direct kernels.
int classify_morph(Morph m, const uint8_t* mask, int mw, int mh) {
    int nz = 0;
    for (int i = 0; i < mw * mh; ++i) nz += mask[i] != 0;
    const bool dil = m == Morph::Dilate;
    if (mw == 3 && mh == 1 && nz == 3) return dil ? kDilateRow3 : kErodeRow3;
    if (mw == 1 && mh == 3 && nz == 3) return dil ? kDilateCol3 : kErodeCol3;
    if (mw == 3 && mh == 3) {
        if (nz == 9) return dil ? kDilateSq3 : kErodeSq3;
        if (nz == 5 && mask[1] && mask[3] && mask[4] && mask[5] && mask[7]) return dil ? kDilateCross3 : kErodeCross3;
    }
    return -1;
}

const int kB3[3] = {1, 2, 1};
const int kB5[5] = {1, 4, 6, 4, 1};

int classify_convolve(const float* k, int kw, int kh) {
    static const int g3[9] = {1, 2, 1, 2, 4, 2, 1, 2, 1};
    static const int g5[25] = {1, 4, 6, 4, 1, 4, 16, 24, 16, 4, 6, 24, 36, 24, 6, 4, 16, 24, 16, 4, 1, 4, 6, 4, 1};
    if (kw == 3 && kh == 1 && matches(k, 3, kB3, 4.0f)) return kConvRow3;
    if (kw == 1 && kh == 3 && matches(k, 3, kB3, 4.0f)) return kConvCol3;
    if (kw == 5 && kh == 1 && matches(k, 5, kB5, 16.0f)) return kConvRow5;
    if (kw == 1 && kh == 5 && matches(k, 5, kB5, 16.0f)) return kConvCol5;
    if (kw == 3 && kh == 3 && matches(k, 9, g3, 16.0f)) return kConv3;
    if (kw == 5 && kh == 5 && matches(k, 25, g5, 256.0f)) return kConv5;
    return -1;
}

int classify_convolve_sep(const float* krow, int kw, const float* kcol, int kh) {
    if (kw == 3 && kh == 3 && matches(krow, 3, kB3, 4.0f) && matches(kcol, 3, kB3, 4.0f)) return kConv3Sep;
    if (kw == 5 && kh == 5 && matches(krow, 5, kB5, 16.0f) && matches(kcol, 5, kB5, 16.0f)) return kConv5Sep;
    return -1;
}

}  // namespace

int stencil_tiled_op(const dip_op& op) {
    switch (op.kind) {
        case DIP_OP_ERODE:
        case DIP_OP_DILATE:
            if (!op.mask || op.mask_w < 1 || op.mask_h < 1 || op.mask_w * op.mask_h > kMaxTaps) return -1;
            return classify_morph(op.kind == DIP_OP_DILATE ? Morph::Dilate : Morph::Erode, op.mask, op.mask_w, op.mask_h);
        case DIP_OP_ERODE_SEPARABLE:
            return op.mask_w == 3 && op.mask_h == 3 ? kErodeSq3 : op.mask_w == 3 && op.mask_h == 1 ? kErodeRow3 : op.mask_w == 1 && op.mask_h == 3 ? kErodeCol3 : -1;
        case DIP_OP_CONVOLVE:
            if (!op.kernel || op.kw < 1 || op.kh < 1 || op.kw * op.kh > kMaxTaps) return -1;
            return classify_convolve(op.kernel, op.kw, op.kh);
        case DIP_OP_CONVOLVE_SEPARABLE:
            if (!op.kernel || !op.kernel_col || op.kw < 1 || op.kh < 1 || op.kw > 5 || op.kh > 5) return -1;
            return classify_convolve_sep(op.kernel, op.kw, op.kernel_col, op.kh);
        case DIP_OP_GAUSSIAN_BLUR3X3: return kGauss3;
    }
    return -1;
}

bool stencil_takes_maps(const dip_op& op) {
    const int t = stencil_tiled_op(op);
    return t >= 0 && t <= kConv5Sep;
}
bool stencil_is_gauss5(const dip_op& op) { return stencil_tiled_op(op) == kConv5; }
bool stencil_is_erode_sq3(const dip_op& op) { return stencil_tiled_op(op) == kErodeSq3; }

int stencil_morph(dip_ctx* ctx, Morph m, const dip_image* src, dip_image* dst, const uint8_t* mask, int mw, int mh,
                  const RowGeom& g, const ChainMaps* maps) {
    if (!mask || mw < 1 || mh < 1) return fail(ctx, DIP_ERR_ARG, "bad mask");
    if (mw * mh > kMaxTaps || mw > 127 || mh > 127) return fail(ctx, DIP_ERR_UNSUPPORTED, "mask larger than %d taps", kMaxTaps);
    int rc = check_geom(ctx, src, dst, g);
    if (rc != DIP_OK) return rc;
    int nz = 0;
    for (int i = 0; i < mw * mh; ++i) nz += mask[i] != 0;
    if (nz == 0) return fail(ctx, DIP_ERR_ARG, "mask has no non-zero tap");
    const bool dil = m == Morph::Dilate;
    const bool wide = src->v.bps == 2;   // 16-bit samples: the benchmark's dedicated kernels are 8-bit only
    if (wide && (g.peer_top || g.peer_bot)) return fail(ctx, DIP_ERR_UNSUPPORTED, "16-bit images cannot read peer rows");
    const int top = wide ? -1 : classify_morph(m, mask, mw, mh);
    if (top >= 0) return launch_tiled_op(ctx, top, src, dst, g, maps);
    if (maps) return fail(ctx, DIP_ERR_UNSUPPORTED, "per-byte maps fuse into the tiled kernels only");
    if (generic_tiled_reach(src, mw, mh) && !g_knobs.direct_generic && tma_ok(src)) {  // small masks: the tiled generic kernel
        SmallTaps t{};
        t.kh = mh;
        for (int j = 0; j < mh; ++j)
            for (int i = 0; i < mw; ++i) t.w[j][i + std::max(1, mw / 2) - mw / 2] = mask[j * mw + i] ? 1.0f : 0.0f;
        return launch_generic_tiled(ctx, dil ? 1 : 0, mw, src, dst, g, t);
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
    if (wide) return launch_direct16(ctx, dil ? 1 : 0, src, dst, g.valid_y0, g.valid_y1, g.out_y0, g.out_y1, g.dst_y0, p.dx, p.dy, p.wt, p.ntaps);
    return dil ? launch_generic<1>(ctx, p) : launch_generic<0>(ctx, p);
}

int stencil_convolve(dip_ctx* ctx, const dip_image* src, dip_image* dst, const float* k, int kw, int kh, const RowGeom& g,
                     const ChainMaps* maps) {
    if (!k || kw < 1 || kh < 1) return fail(ctx, DIP_ERR_ARG, "bad kernel");
    if (kw * kh > kMaxTaps || kw > 127 || kh > 127) return fail(ctx, DIP_ERR_UNSUPPORTED, "kernel larger than %d taps", kMaxTaps);
    int rc = check_geom(ctx, src, dst, g);
    if (rc != DIP_OK) return rc;
    const bool wide = src->v.bps == 2;
    if (wide && (g.peer_top || g.peer_bot)) return fail(ctx, DIP_ERR_UNSUPPORTED, "16-bit images cannot read peer rows");
    const int top = wide ? -1 : classify_convolve(k, kw, kh);
    if (top >= 0) return launch_tiled_op(ctx, top, src, dst, g, maps);
    if (maps) return fail(ctx, DIP_ERR_UNSUPPORTED, "per-byte maps fuse into the tiled kernels only");
    if (generic_tiled_reach(src, kw, kh) && !g_knobs.direct_generic && tma_ok(src)) {  // small kernels: the tiled generic kernel
        SmallTaps t{};
        t.kh = kh;
        for (int j = 0; j < kh; ++j)
            for (int i = 0; i < kw; ++i) t.w[j][i + std::max(1, kw / 2) - kw / 2] = k[j * kw + i];
        return launch_generic_tiled(ctx, 2, kw, src, dst, g, t);
    }
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
    if (p.ntaps == 0) { p.dx[0] = 0; p.dy[0] = 0; p.wt[0] = 0.0f; p.ntaps = 1; }   // an all-zero kernel: zeros out
    if (wide) return launch_direct16(ctx, 2, src, dst, g.valid_y0, g.valid_y1, g.out_y0, g.out_y1, g.dst_y0, p.dx, p.dy, p.wt, p.ntaps);
    return launch_generic<2>(ctx, p);
}

int stencil_convolve_sep(dip_ctx* ctx, const dip_image* src, dip_image* dst, const float* krow, int kw, const float* kcol,
                         int kh, const RowGeom& g, const ChainMaps* maps) {
    if (!krow || !kcol || kw < 1 || kh < 1) return fail(ctx, DIP_ERR_ARG, "bad kernel");
    int rc = check_geom(ctx, src, dst, g);
    if (rc != DIP_OK) return rc;
    const int top = src->v.bps == 2 ? -1 : classify_convolve_sep(krow, kw, kcol, kh);
    if (top >= 0) return launch_tiled_op(ctx, top, src, dst, g, maps);
    // Any other pair of kernels: the reference's two calls through `aux` (opencv:84-90, sycl:347-356) -- a 1 x kw pass
    // whose result is rounded and saturated to u8 exactly as `aux` is, then a kh x 1 pass -- with the ctx's scratch stack
    // as `aux`.  The row pass also produces the rows the column pass reads above and below the requested range.
    if (maps) return fail(ctx, DIP_ERR_UNSUPPORTED, "per-byte maps fuse into the tiled kernels only");
    if (kw * 1 > kMaxTaps || kh * 1 > kMaxTaps || kw > 127 || kh > 127) return fail(ctx, DIP_ERR_UNSUPPORTED, "kernel longer than %d taps", kMaxTaps);
    if (g.peer_top || g.peer_bot) return fail(ctx, DIP_ERR_UNSUPPORTED, "two-pass separable convolution cannot read peer rows");
    dip_image* aux = nullptr;
    if ((rc = scratch_view(ctx, 1, src, &aux)) != DIP_OK) return rc;
    const int ry = kh / 2;
    RowGeom g1 = g;
    g1.out_y0 = std::max(g.valid_y0, g.out_y0 - ry);
    g1.out_y1 = std::min(g.valid_y1, g.out_y1 + (kh - 1 - ry));
    g1.dst_y0 = g1.out_y0;
    rc = stencil_convolve(ctx, src, aux, krow, kw, 1, g1, nullptr);
    if (rc == DIP_OK) rc = stencil_convolve(ctx, aux, dst, kcol, 1, kh, g, nullptr);
    dip_image_destroy(aux);
    return rc;
}

int stencil_gauss5_erode3(dip_ctx* ctx, const dip_image* src, dip_image* dst, const RowGeom& g, const ChainMaps* maps) {
    int rc = check_geom(ctx, src, dst, g);
    if (rc != DIP_OK) return rc;
    if (src->v.bps == 2) {   // 16-bit samples: the two operations one after the other through the ctx's `aux` stack
        if (maps || g.peer_top || g.peer_bot) return fail(ctx, DIP_ERR_UNSUPPORTED, "16-bit images: no fused maps, no peer rows");
        static const float b5[5] = {1 / 16.f, 4 / 16.f, 6 / 16.f, 4 / 16.f, 1 / 16.f};
        float k5[25];
        for (int j = 0; j < 5; ++j)
            for (int i = 0; i < 5; ++i) k5[5 * j + i] = b5[j] * b5[i];
        static const uint8_t ones[9] = {1, 1, 1, 1, 1, 1, 1, 1, 1};
        dip_image* aux = nullptr;
        if ((rc = scratch_view(ctx, 1, src, &aux)) != DIP_OK) return rc;
        RowGeom g1 = g;
        g1.out_y0 = std::max(g.valid_y0, g.out_y0 - 1);
        g1.out_y1 = std::min(g.valid_y1, g.out_y1 + 1);
        g1.dst_y0 = g1.out_y0;
        rc = stencil_convolve(ctx, src, aux, k5, 5, 5, g1, nullptr);
        if (rc == DIP_OK) rc = stencil_morph(ctx, Morph::Erode, aux, dst, ones, 3, 3, g, nullptr);
        dip_image_destroy(aux);
        return rc;
    }
    const bool mapped = maps && !(maps->pre.identity() && maps->post.identity());
    switch (src->v.channels) {
        case 1:
            if (!mapped) switch (g_knobs.fused_shape) {
                case 1: return launch_fused_g5e3<1, 30, 2, false>(ctx, src, dst, g, nullptr);
                case 2: return launch_fused_g5e3<1, 30, 4, false>(ctx, src, dst, g, nullptr);
                case 3: return launch_fused_g5e3<1, 15, 4, false, 5>(ctx, src, dst, g, nullptr);
                case 4: return launch_fused_g5e3<1, 15, 2, false>(ctx, src, dst, g, nullptr);
                case 5: return launch_fused_g5e3<1, 30, 3, false>(ctx, src, dst, g, nullptr);
                case 6: return launch_fused_g5e3<1, 15, 4, false, 6>(ctx, src, dst, g, nullptr);
                case 7: return launch_fused_g5e3<1, 45, 2, false>(ctx, src, dst, g, nullptr);
                case 8: return launch_fused_g5e3<1, 15, 4, false, 0, 1>(ctx, src, dst, g, nullptr);
                case 9: return launch_fused_g5e3<1, 15, 4, false, 0, 15>(ctx, src, dst, g, nullptr);
                case 10: return launch_fused_g5e3<1, 15, 4, false, 0, 3>(ctx, src, dst, g, nullptr);
                case 11: return launch_fused_g5e3<1, 60, 2, false>(ctx, src, dst, g, nullptr);
                case 12: return launch_fused_g5e3<1, 60, 1, false>(ctx, src, dst, g, nullptr);
                case 13: return launch_fused_g5e3<1, 45, 3, false>(ctx, src, dst, g, nullptr);
                default: break;
            }
            // 30-row bands on 2-warp CTAs: 36 horizontal passes per 30 output rows instead of 21 per 15; 4.6 % faster in the
            // sustained power state (profiles/r2_sustained_shapes.md); fused_shape 14 = the 15 x 4 shape for A/B
            if (!mapped && g_knobs.fused_shape == 14) return launch_fused_g5e3<1, 15, 4, false>(ctx, src, dst, g, nullptr);
            return mapped ? launch_fused_g5e3<1, 15, 4, true>(ctx, src, dst, g, maps) : launch_fused_g5e3<1, 30, 2, false>(ctx, src, dst, g, nullptr);
        case 3:
            if (!mapped && g_knobs.fused_shape == 1) return launch_fused_g5e3<3, 30, 2, false>(ctx, src, dst, g, nullptr);   // A/B
            return mapped ? launch_fused_g5e3<3, 15, 4, true>(ctx, src, dst, g, maps) : launch_fused_g5e3<3, 15, 4, false>(ctx, src, dst, g, nullptr);
        case 4: return mapped ? launch_fused_g5e3<4, 15, 4, true>(ctx, src, dst, g, maps) : launch_fused_g5e3<4, 15, 4, false>(ctx, src, dst, g, nullptr);
    }
    return fail(ctx, DIP_ERR_ARG, "bad channel count");
}

int stencil_gaussian3(dip_ctx* ctx, const dip_image* src, dip_image* dst, const RowGeom& g, const ChainMaps* maps) {
    int rc = check_geom(ctx, src, dst, g);
    if (rc != DIP_OK) return rc;
    if (src->v.bps == 2) {   // 16-bit samples: ([1 2 1] x [1 2 1] sum + 8) >> 4, exact in fp32 (20 bits), rounded half up
        if (maps) return fail(ctx, DIP_ERR_UNSUPPORTED, "per-byte maps fuse into the tiled kernels only");
        if (g.peer_top || g.peer_bot) return fail(ctx, DIP_ERR_UNSUPPORTED, "16-bit images cannot read peer rows");
        static const float b3[3] = {0.25f, 0.5f, 0.25f};
        if (!g_knobs.direct_generic && tma_ok(src)) {
            SmallTaps t{};
            t.kh = 3;
            for (int j = 0; j < 3; ++j)
                for (int i = 0; i < 3; ++i) t.w[j][i] = b3[j] * b3[i];
            return launch_generic_tiled(ctx, 3, 3, src, dst, g, t);
        }
        int8_t dx[9], dy[9];
        float wt[9];
        for (int j = 0; j < 3; ++j)
            for (int i = 0; i < 3; ++i) { dx[3 * j + i] = (int8_t)(i - 1); dy[3 * j + i] = (int8_t)(j - 1); wt[3 * j + i] = b3[j] * b3[i]; }
        return launch_direct16(ctx, 3, src, dst, g.valid_y0, g.valid_y1, g.out_y0, g.out_y1, g.dst_y0, dx, dy, wt, 9);
    }
    return launch_tiled_op(ctx, kGauss3, src, dst, g, maps);
}

}  // namespace dip

using namespace dip;

extern "C" {

int dip_debug_set(const char* key, int value) {
    if (!key) return fail(nullptr, DIP_ERR_ARG, "null key");
    if (!strcmp(key, "tile_rows")) g_knobs.tile_rows = value;
    else if (!strcmp(key, "int_math")) g_knobs.int_math = value;
    else if (!strcmp(key, "direct_generic")) g_knobs.direct_generic = value;
    else if (!strcmp(key, "umma")) g_knobs.umma = value;
    else if (!strcmp(key, "fused_shape")) g_knobs.fused_shape = value;
    else if (!strcmp(key, "band_shape")) g_knobs.band_shape = value;
    else if (!strcmp(key, "float_morph")) g_generic_float_morph = value != 0;
    else if (!strcmp(key, "no_pad")) g_no_occupancy_pad = value != 0;
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
