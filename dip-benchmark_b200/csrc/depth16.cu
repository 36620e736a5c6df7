// 16-bit images (last item): the operations of the benchmark on CV_16U data with the semantics
// OpenCV gives them -- probed with cv2 4.13 (tests/golden/make_golden_u16.py): the same formulas as for 8-bit samples
// with the wider type, i.e.
//   threshold   x > t ? maxval : 0 per sample                                  (cv.threshold, opencv/benchmark.py:97)
//   grayscale   Y = (3735 B + 19235 G + 9798 R + 16384) >> 15, alpha := 65535  (cv.cvtColor, opencv:96)
//   erosion     minimum over the in-image taps (+inf border)                   (cv.erode, opencv:98-100)
//   filter2D    correlation accumulated in fp32, cvRound, saturated to 65535   (opencv:101-104; exact for the
//               benchmark's kernels: weights k/256 times 16-bit samples need 24 bits, an fp32 significand)
//   GaussianBlur 3x3  ([1 2 1]x[1 2 1] sum + 8) >> 4                           (opencv:105)
// Copy, inversion, upload and download are byte-wise and run through the 8-bit paths unchanged (~x on the two bytes of a
// sample is 65535 - x).  The reference itself only ever feeds 8-bit images (cv.imread without flags), so these are not
// benchmark kernels: the stencils run on the tiled generic kernel (stencil_generic.cu, up to 7x7 taps where the
// 16-byte tile halo reaches; erosions / dilations as packed minima over the raw words) and on the direct kernel below
// for anything larger; threshold is a variant of the flat point kernel (point_ops.cu), grayscale a 16-byte-vector kernel here.
#include "dip_internal.h"

namespace dip {

namespace {

constexpr int kThreads = 256;

// BT.601 in 15-bit fixed point like cv.cvtColor's 16-bit path, gray replicated into the colour channels.  A thread owns a
// group of 8 (BGR) or 4 (BGRA) pixels = 48 / 32 bytes = whole 16-byte vectors; the last, partial group of a row goes
// sample by sample.
template <int C>
__global__ void __launch_bounds__(kThreads) gray16_kernel(ImgView src, ImgView dst, int groups_per_row, size_t total) {
    constexpr int P = C == 3 ? 8 : 4;          // pixels per thread
    constexpr int V = P * C * 2 / 16;          // 16-byte vectors per thread
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (i >= total) return;
    const int gx = (int)(i % (size_t)groups_per_row);
    const size_t q = i / (size_t)groups_per_row;
    const int y = (int)(q % (size_t)src.height);
    const size_t f = q / (size_t)src.height;
    const int x0 = gx * P;
    const uint8_t* srow = src.base + f * src.frame_stride + (size_t)y * src.pitch;
    uint8_t* drow = dst.base + f * dst.frame_stride + (size_t)y * dst.pitch;
    auto luma = [](uint32_t b, uint32_t g, uint32_t r) { return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15; };
    if (x0 + P <= src.width) {
        uint32_t w[4 * V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const uint4 t = *reinterpret_cast<const uint4*>(srow + (size_t)x0 * C * 2 + 16 * v);
            w[4 * v] = t.x; w[4 * v + 1] = t.y; w[4 * v + 2] = t.z; w[4 * v + 3] = t.w;
        }
        auto sample = [&](int k) { return (w[k >> 1] >> (16 * (k & 1))) & 0xFFFFu; };
        uint32_t o[4 * V];
#pragma unroll
        for (int k = 0; k < 4 * V; ++k) o[k] = 0;
#pragma unroll
        for (int px = 0; px < P; ++px) {
            const uint32_t yv = luma(sample(px * C), sample(px * C + 1), sample(px * C + 2));
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                const int k = px * C + ch;
                o[k >> 1] |= (ch == 3 ? 65535u : yv) << (16 * (k & 1));
            }
        }
#pragma unroll
        for (int v = 0; v < V; ++v)
            __stcs(reinterpret_cast<uint4*>(drow + (size_t)x0 * C * 2 + 16 * v), make_uint4(o[4 * v], o[4 * v + 1], o[4 * v + 2], o[4 * v + 3]));
    } else {
        const uint16_t* sp = reinterpret_cast<const uint16_t*>(srow);
        uint16_t* dp = reinterpret_cast<uint16_t*>(drow);
        for (int x = x0; x < src.width; ++x) {
            const uint16_t yv = (uint16_t)luma(sp[x * C], sp[x * C + 1], sp[x * C + 2]);
            dp[x * C] = yv; dp[x * C + 1] = yv; dp[x * C + 2] = yv;
            if (C == 4) dp[x * C + 3] = 65535;
        }
    }
}

constexpr int kMaxTaps16 = 225;
struct Direct16Params {
    ImgView src, dst;
    int valid_y0, valid_y1, out_y0, out_y1, dst_y0;
    int ntaps;
    int8_t dx[kMaxTaps16], dy[kMaxTaps16];
    float wt[kMaxTaps16];
};

__device__ __forceinline__ int reflect101_16(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) p = p < 0 ? -p : 2 * (len - 1) - p;
    return p;
}

// one thread per sample; taps in the kernel parameters.  MODE: 0 erode, 1 dilate, 2 correlation + RNE, 3 correlation + half up
template <int MODE>
__global__ void __launch_bounds__(kThreads) direct16_kernel(const __grid_constant__ Direct16Params p) {
    const int spr = p.src.width * p.src.channels;   // samples per row
    const size_t total = (size_t)spr * (size_t)(p.out_y1 - p.out_y0) * (size_t)p.src.frames;
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    pdl_launch_dependents();
    pdl_wait();
    if (i >= total) return;
    const int C = p.src.channels;
    const int xs = (int)(i % (size_t)spr);
    const size_t q = i / (size_t)spr;
    const int yo = (int)(q % (size_t)(p.out_y1 - p.out_y0));
    const size_t f = q / (size_t)(p.out_y1 - p.out_y0);
    const int y = p.out_y0 + yo;
    const int x = xs / C, ch = xs - x * C;
    const uint8_t* sbase = p.src.base + f * p.src.frame_stride;
    const int vh = p.valid_y1 - p.valid_y0;
    auto at = [&](int xx, int yy) { return (int)reinterpret_cast<const uint16_t*>(sbase + (size_t)yy * p.src.pitch)[xx * C + ch]; };
    int result;
    if (MODE >= 2) {
        float acc = 0.0f;
        for (int t = 0; t < p.ntaps; ++t) {
            const int xx = reflect101_16(x + p.dx[t], p.src.width);
            const int yy = p.valid_y0 + reflect101_16(y + p.dy[t] - p.valid_y0, vh);
            acc = __fmaf_rn(p.wt[t], (float)at(xx, yy), acc);
        }
        const int rr = MODE == 2 ? __float2int_rn(acc) : __float2int_rd(acc + 0.5f);
        result = min(max(rr, 0), 65535);
    } else {
        int m = MODE == 0 ? 65535 : 0;
        for (int t = 0; t < p.ntaps; ++t) {
            const int xx = x + p.dx[t], yy = y + p.dy[t];
            if (xx < 0 || xx >= p.src.width || yy < p.valid_y0 || yy >= p.valid_y1) continue;
            const int v = at(xx, yy);
            m = MODE == 0 ? min(m, v) : max(m, v);
        }
        result = m;
    }
    reinterpret_cast<uint16_t*>(p.dst.base + f * p.dst.frame_stride + (size_t)(p.dst_y0 + yo) * p.dst.pitch)[xs] = (uint16_t)result;
}

template <int MODE>
int launch_direct16_t(dip_ctx* ctx, const Direct16Params& p) {
    const size_t total = (size_t)p.src.width * p.src.channels * (size_t)(p.out_y1 - p.out_y0) * (size_t)p.src.frames;
    if (total == 0) return DIP_OK;
    const size_t blocks = (total + kThreads - 1) / kThreads;
    if (blocks > 0x7FFFFFFFull) return fail(ctx, DIP_ERR_UNSUPPORTED, "image stack too large for one launch");
    DIP_CUDA(ctx, launch_kernel(plan_launch(ctx, nullptr, 0), direct16_kernel<MODE>, dim3((unsigned)blocks), dim3(kThreads), 0, ctx->compute, p));
    DIP_CHECK_LAUNCH(ctx, "direct16_kernel");
    return DIP_OK;
}

}  // namespace

int launch_grayscale16(dip_ctx* ctx, const ImgView& src, const ImgView& dst) {
    if (src.channels == 1) return launch_point(ctx, kCopy, src, dst, 0, 0);   // already gray: a byte-wise copy
    const int per = src.channels == 3 ? 8 : 4;
    const int groups = (src.width + per - 1) / per;
    const size_t total = (size_t)groups * src.height * src.frames;
    if (total == 0) return DIP_OK;
    const size_t blocks = (total + kThreads - 1) / kThreads;
    if (blocks > 0x7FFFFFFFull) return fail(ctx, DIP_ERR_UNSUPPORTED, "image stack too large for one launch");
    const LaunchMode mode = plan_launch(ctx, src, dst);
    if (src.channels == 3) DIP_CUDA(ctx, launch_kernel(mode, gray16_kernel<3>, dim3((unsigned)blocks), dim3(kThreads), 0, ctx->compute, src, dst, groups, total));
    else DIP_CUDA(ctx, launch_kernel(mode, gray16_kernel<4>, dim3((unsigned)blocks), dim3(kThreads), 0, ctx->compute, src, dst, groups, total));
    DIP_CHECK_LAUNCH(ctx, "gray16_kernel");
    return DIP_OK;
}

int launch_direct16(dip_ctx* ctx, int mode, const dip_image* src, dip_image* dst, int valid_y0, int valid_y1, int out_y0, int out_y1,
                    int dst_y0, const int8_t* dx, const int8_t* dy, const float* wt, int ntaps) {
    if (ntaps < 1 || ntaps > kMaxTaps16) return fail(ctx, DIP_ERR_UNSUPPORTED, "mask / kernel of %d taps (limit %d)", ntaps, kMaxTaps16);
    Direct16Params p;
    p.src = src->v; p.dst = dst->v;
    p.valid_y0 = valid_y0; p.valid_y1 = valid_y1; p.out_y0 = out_y0; p.out_y1 = out_y1; p.dst_y0 = dst_y0;
    p.ntaps = ntaps;
    for (int t = 0; t < ntaps; ++t) { p.dx[t] = dx[t]; p.dy[t] = dy[t]; p.wt[t] = wt[t]; }
    switch (mode) {
        case 0: return launch_direct16_t<0>(ctx, p);
        case 1: return launch_direct16_t<1>(ctx, p);
        case 2: return launch_direct16_t<2>(ctx, p);
        case 3: return launch_direct16_t<3>(ctx, p);
    }
    return fail(ctx, DIP_ERR_ARG, "bad mode");
}

}  // namespace dip
