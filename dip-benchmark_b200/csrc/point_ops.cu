// Point operations of libdip_b200.so: copy, inversion, threshold, grayscale, plus the
// synthetic-input generator and the checksum used by the batch / mosaic configurations.
//
// Replaces InvertKernel / GrayscaleKernel / ThresholdKernel
// (/root/reference/benchmarks/sycl/benchmark.cpp:54-114: one thread per pixel, byte loads) and
// cv.bitwise_not / cvtColor / threshold (/root/reference/benchmarks/opencv/benchmark.py:95-97).
//
// Design: the image stack is a dense byte stream (pitch = rowbytes rounded to 16), so the
// byte-wise ops run over it as 16-byte vectors: every thread has two independent
// 128-bit loads in flight before its first store, streaming (evict-first) cache policy on
// both sides because every byte is touched exactly once.  Grayscale works on 48-byte groups
// (16 BGR pixels = 3 vectors) with OpenCV's 15-bit BT.601 weights evaluated by two dp4a per
// pixel (weights split into high and low bytes), exact for all 2^24 colours.
#include <algorithm>

#include "dip_internal.h"

namespace dip {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint4 ld_stream(const uint4* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) { __stcs(p, v); }

// BT.601 weights in 15-bit fixed point (OpenCV: B 3735, G 19235, R 9798), split in bytes so
// that Y = (256*dp4a(px, hi) + dp4a(px, lo) + 16384) >> 15.
//   3735 = 14*256 + 151, 19235 = 75*256 + 35, 9798 = 38*256 + 70
constexpr uint32_t kWLo0 = 151u | (35u << 8) | (70u << 16);        // pixel in bytes 0..2
constexpr uint32_t kWHi0 = 14u | (75u << 8) | (38u << 16);
constexpr uint32_t kWLo1 = kWLo0 << 8;                               // pixel in bytes 1..3
constexpr uint32_t kWHi1 = kWHi0 << 8;

__device__ __forceinline__ uint32_t luma(uint32_t px, uint32_t wlo, uint32_t whi) {
    const uint32_t lo = __dp4a(px, wlo, 16384u);
    const uint32_t hi = __dp4a(px, whi, 0u);
    return (hi * 256u + lo) >> 15;
}

constexpr int kGray4 = 3;       // BGRA -> Y,Y,Y,255
constexpr int kMap = 4;         // any composition of the byte-wise ops (dip_internal.h: ByteMap)
constexpr int kMapGray4 = 5;    // post(gray(pre(x))) on BGRA
constexpr int kThreshold16 = 6; // x > t ? maxval : 0 on 16-bit samples (t4 / flip4 hold the value in both halves)

// `m`: the threshold (t4, flip4 = maxval) or the map; `post`: second map of kMapGray4
template <int OP>
__device__ __forceinline__ uint32_t point_word(uint32_t x, const MapParams& m, const MapParams& post) {
    if (OP == kCopy) return x;
    if (OP == kInvert) return ~x;
    if (OP == kThreshold) return __vcmpgtu4(x, m.t4) & m.flip4;
    if (OP == kThreshold16) return __vcmpgtu2(x, m.t4) & m.flip4;
    if (OP == kMap) return map_word(x, m);
    if (OP == kMapGray4) x = map_word(x, m);
    const uint32_t y = luma(x, kWLo0, kWHi0);
    const uint32_t g = y * 0x00010101u | 0xFF000000u;
    return OP == kMapGray4 ? map_word(g, post) : g;
}

template <int OP>
__device__ __forceinline__ uint4 point_vec(uint4 v, const MapParams& m, const MapParams& post) {
    v.x = point_word<OP>(v.x, m, post);
    v.y = point_word<OP>(v.y, m, post);
    v.z = point_word<OP>(v.z, m, post);
    v.w = point_word<OP>(v.w, m, post);
    return v;
}

// Dense stream: one CTA per chunk of kThreads*UNROLL consecutive vectors.  Concurrently running
// CTAs touch neighbouring chunks, which keeps DRAM pages open; a machine-sized persistent grid
// with one long span per CTA measured 12 % slower on B200 (1184 scattered streams).
template <int OP, int UNROLL>
__global__ void __launch_bounds__(kThreads) point_flat_kernel(const uint4* __restrict__ src,
                                                              uint4* __restrict__ dst, size_t nvec,
                                                              const MapParams m, const MapParams post, const bool wait_prev) {
    pdl_launch_dependents();
    pdl_wait(wait_prev);
    const size_t base = (size_t)blockIdx.x * (kThreads * UNROLL) + threadIdx.x;
    uint4 v[UNROLL];
    if (base + (size_t)(UNROLL - 1) * kThreads < nvec) {
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) v[j] = ld_stream(src + base + (size_t)j * kThreads);
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) st_stream(dst + base + (size_t)j * kThreads, point_vec<OP>(v[j], m, post));
    } else {
#pragma unroll
        for (int j = 0; j < UNROLL; ++j) {
            const size_t i = base + (size_t)j * kThreads;
            if (i < nvec) st_stream(dst + i, point_vec<OP>(ld_stream(src + i), m, post));
        }
    }
}

// Strided layouts (wrapped memory whose pitches differ): one vector per thread.
template <int OP>
__global__ void __launch_bounds__(kThreads) point_rows_kernel(ImgView src, ImgView dst, int vec_per_row,
                                                              size_t total, const MapParams m, const MapParams post) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= total) return;
    const size_t row = i / (size_t)vec_per_row;
    const int vx = (int)(i - row * (size_t)vec_per_row);
    const size_t f = row / (size_t)src.height;
    const size_t y = row - f * (size_t)src.height;
    // strided layouts are wrapped memory: neither read nor write past the row's last byte
    const uint8_t* s = src.base + f * src.frame_stride + y * src.pitch + 16 * vx;
    uint8_t* d = dst.base + f * dst.frame_stride + y * dst.pitch + 16 * vx;
    const int valid = min(16, src.rowbytes - 16 * vx);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (valid >= 16) {
        v = *reinterpret_cast<const uint4*>(s);
    } else {
        uint32_t w[4] = {0, 0, 0, 0};
        for (int i = 0; i < valid; ++i) w[i >> 2] |= (uint32_t)s[i] << (8 * (i & 3));
        v = make_uint4(w[0], w[1], w[2], w[3]);
    }
    store16(d, point_vec<OP>(v, m, post), valid);
}

// BGR grayscale: a thread converts one 48-byte group = 16 pixels = 3 vectors.
// 4 BGR pixels in 3 words -> 4 x (Y,Y,Y); MAPPED: post(gray(pre(x)))
template <bool MAPPED>
__device__ __forceinline__ void gray3_words(const uint32_t* w, uint32_t* o, const MapParams& pre, const MapParams& post) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {  // 4 pixels per 3 words
        uint32_t a = w[3 * q], b = w[3 * q + 1], c = w[3 * q + 2];
        if (MAPPED) { a = map_word(a, pre); b = map_word(b, pre); c = map_word(c, pre); }
        const uint32_t y0 = luma(a, kWLo0, kWHi0);                          // bytes 0,1,2 of a
        const uint32_t y1 = luma(__byte_perm(a, b, 0x7543), kWLo0, kWHi0);  // a.3, b.0, b.1
        const uint32_t y2 = luma(__byte_perm(b, c, 0x7432), kWLo0, kWHi0);  // b.2, b.3, c.0
        const uint32_t y3 = luma(c, kWLo1, kWHi1);                          // bytes 1,2,3 of c
        o[3 * q + 0] = y0 * 0x00010101u | (y1 << 24);
        o[3 * q + 1] = y1 * 0x00000101u | (y2 * 0x01010000u);
        o[3 * q + 2] = y2 | (y3 * 0x01010100u);
        if (MAPPED) { o[3 * q] = map_word(o[3 * q], post); o[3 * q + 1] = map_word(o[3 * q + 1], post); o[3 * q + 2] = map_word(o[3 * q + 2], post); }
    }
}

template <bool exact, bool MAPPED>  // exact: rows are not padded to 16 bytes (wrapped regions of interest)
__global__ void __launch_bounds__(kThreads) gray3_kernel(ImgView src, ImgView dst, int groups_per_row, size_t total,
                                                         const MapParams pre, const MapParams post) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= total) return;
    const size_t row = i / (size_t)groups_per_row;
    const int g = (int)(i - row * (size_t)groups_per_row);
    const size_t f = row / (size_t)src.height;
    const size_t y = row - f * (size_t)src.height;
    const uint8_t* srow = src.base + f * src.frame_stride + y * src.pitch;
    uint8_t* drow = dst.base + f * dst.frame_stride + y * dst.pitch;
    const int off = g * 48;
    // vectors past the padded row end are neither read nor written
    const int padded = (src.rowbytes + 15) & ~15;
    const int nv = min(3, (padded - off) >> 4);
    uint32_t w[12];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (j < nv) {
            const int valid = exact ? src.rowbytes - (off + 16 * j) : 16;
            if (valid >= 16) {
                v = __ldg(reinterpret_cast<const uint4*>(srow + off) + j);
            } else {  // wrapped region of interest: do not read past the row
                uint32_t t[4] = {0, 0, 0, 0};
                for (int i = 0; i < valid; ++i) t[i >> 2] |= (uint32_t)srow[off + 16 * j + i] << (8 * (i & 3));
                v = make_uint4(t[0], t[1], t[2], t[3]);
            }
        }
        w[4 * j + 0] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
    }
    uint32_t o[12];
    gray3_words<MAPPED>(w, o, pre, post);
#pragma unroll
    for (int j = 0; j < 3; ++j)
        if (j < nv) store16(drow + off + 16 * j, make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]), chunk_valid(off + 16 * j, src.rowbytes, exact));
}

// Dense BGR stacks whose rows are whole 48-byte groups (W % 16 == 0, e.g. 3504): the stack is one flat array of
// groups.  A CTA moves 256 groups = 12 KB through shared memory so that global loads and stores are perfectly
// coalesced 16-byte vectors (the direct kernel above reads 48-byte-strided vectors, 3x the L1 wavefronts, which
// shows once the SM clock drops under the power cap); each thread still converts its own 48-byte group, read from
// shared memory at a 48-byte stride (12 banks apart: conflict-free for 128-bit accesses).
template <bool MAPPED>
__global__ void __launch_bounds__(kThreads) gray3_flat_kernel(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t nvec,
                                                              const MapParams pre, const MapParams post, const bool wait_prev) {
    __shared__ uint4 stage[kThreads * 3];
    pdl_launch_dependents();
    pdl_wait(wait_prev);
    const size_t base = (size_t)blockIdx.x * (kThreads * 3);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const size_t i = base + threadIdx.x + (size_t)j * kThreads;
        if (i < nvec) stage[threadIdx.x + j * kThreads] = ld_stream(src + i);
    }
    __syncthreads();
    uint32_t w[12];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const uint4 v = stage[threadIdx.x * 3 + j];
        w[4 * j + 0] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
    }
    uint32_t o[12];
    gray3_words<MAPPED>(w, o, pre, post);
#pragma unroll
    for (int j = 0; j < 3; ++j) stage[threadIdx.x * 3 + j] = make_uint4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const size_t i = base + threadIdx.x + (size_t)j * kThreads;
        if (i < nvec) st_stream(dst + i, stage[threadIdx.x + j * kThreads]);
    }
}

// Same integer mix as oracle/dip_oracle.c:orc_mix32 (kept in sync by tests/test_parity_gpu.py).
__host__ __device__ __forceinline__ uint32_t mix32(uint64_t idx, uint32_t seed) {
    uint32_t x = (uint32_t)idx ^ ((uint32_t)(idx >> 32) * 0x9E3779B9u) ^ (seed * 0x85EBCA6Bu + 0xC2B2AE35u);
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(kThreads) synth_kernel(ImgView dst, uint32_t seed, uint64_t first_row, int vec_per_row,
                                                         size_t total) {
    pdl_launch_dependents();
    pdl_wait();
    const size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x;
    if (i >= total) return;
    const size_t row = i / (size_t)vec_per_row;
    const int vx = (int)(i - row * (size_t)vec_per_row);
    const size_t f = row / (size_t)dst.height;
    const size_t y = row - f * (size_t)dst.height;
    const uint64_t rowbase = (first_row + f * (uint64_t)dst.height + y) * (uint64_t)dst.rowbytes;
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        uint32_t word = 0;
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const int x = vx * 16 + k * 4 + b;
            const uint32_t v = x < dst.rowbytes ? (mix32(rowbase + (uint64_t)x, seed) >> 24) : 0u;
            word |= v << (8 * b);
        }
        w[k] = word;
    }
    store16(dst.base + f * dst.frame_stride + y * dst.pitch + 16 * vx, make_uint4(w[0], w[1], w[2], w[3]),
            chunk_valid(16 * vx, dst.rowbytes, !dst.pad_writable));
}

__global__ void __launch_bounds__(kThreads) checksum_kernel(ImgView src, int vec_per_row, size_t vec_per_frame, uint64_t first_row,
                                                            unsigned long long* __restrict__ out) {
    pdl_launch_dependents();
    pdl_wait();
    // grid.y = frame; grid-stride over the frame's vectors
    const size_t f = blockIdx.y;
    unsigned long long acc = 0;
    for (size_t i = (size_t)blockIdx.x * kThreads + threadIdx.x; i < vec_per_frame; i += (size_t)gridDim.x * kThreads) {
        const size_t y = i / (size_t)vec_per_row;
        const int vx = (int)(i - y * (size_t)vec_per_row);
        const uint4 v = *(reinterpret_cast<const uint4*>(src.base + f * src.frame_stride + y * src.pitch) + vx);
        const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int x = vx * 16 + k * 4 + b;
                if (x < src.rowbytes) {
                    const uint64_t pos = (first_row + (uint64_t)y) * (uint64_t)src.rowbytes + (uint64_t)x;
                    acc += (unsigned long long)(((w[k] >> (8 * b)) & 0xFFu) + 1u) * (unsigned long long)mix32(pos, 0x1234567u);
                }
            }
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) acc += __shfl_xor_sync(0xFFFFFFFFu, acc, s);
    if ((threadIdx.x & 31) == 0) atomicAdd(out + f, acc);
}

bool dense_pair(const ImgView& a, const ImgView& b) {
    const size_t padded = ((size_t)a.rowbytes + 15) & ~(size_t)15;
    auto dense = [&](const ImgView& v) {  // one flat byte stream whose row padding may be written
        return v.pitch == padded && !exact_rows(v) && (v.frames == 1 || v.frame_stride == v.pitch * (size_t)v.height);
    };
    return dense(a) && dense(b);
}

template <int OP>
int launch_point_t(dip_ctx* ctx, const ImgView& src, const ImgView& dst, const MapParams& m, const MapParams& post = MapParams{}) {
    if (dense_pair(src, dst)) {
        const size_t nvec = src.pitch * (size_t)src.height * (size_t)src.frames / 16;
        // 2 x 16 B per thread measured best on B200 for one image (8.8 us) and for stacks (6.98 TB/s); 1, 4 and 8 were slower
        constexpr int unroll = 2;
        const size_t chunk = (size_t)kThreads * unroll;
        const size_t blocks = (nvec + chunk - 1) / chunk;
        if (blocks > 0x7FFFFFFFull) return fail(ctx, DIP_ERR_UNSUPPORTED, "image stack too large for one launch");
        const LaunchMode mode = plan_launch(ctx, src, dst);
        DIP_CUDA(ctx, launch_kernel(mode, point_flat_kernel<OP, unroll>, dim3((unsigned)blocks), dim3(kThreads), 0, ctx->compute,
                                    reinterpret_cast<const uint4*>(src.base), reinterpret_cast<uint4*>(dst.base), nvec, m, post, mode != kNoWait));
    } else {
        const int vpr = (src.rowbytes + 15) / 16;
        const size_t total = (size_t)vpr * (size_t)src.height * (size_t)src.frames;
        const size_t blocks = (total + kThreads - 1) / kThreads;
        if (blocks > 0x7FFFFFFFull) return fail(ctx, DIP_ERR_UNSUPPORTED, "image stack too large for one launch");
        DIP_CUDA(ctx, launch_kernel(plan_launch(ctx, nullptr, 0), point_rows_kernel<OP>, dim3((unsigned)blocks), dim3(kThreads), 0, ctx->compute, src, dst, vpr, total, m, post));
    }
    DIP_CHECK_LAUNCH(ctx, "point kernel");
    return DIP_OK;
}

}  // namespace

int launch_threshold16(dip_ctx* ctx, const ImgView& src, const ImgView& dst, uint16_t thresh, uint16_t maxval) {
    MapParams m;
    m.t4 = (uint32_t)thresh * 0x00010001u;
    m.flip4 = (uint32_t)maxval * 0x00010001u;
    return launch_point_t<kThreshold16>(ctx, src, dst, m);
}

int launch_point(dip_ctx* ctx, PointOp op, const ImgView& src, const ImgView& dst, uint8_t thresh, uint8_t maxval) {
    switch (op) {
        case kCopy: return launch_point_t<kCopy>(ctx, src, dst, MapParams{});
        case kInvert: return launch_point_t<kInvert>(ctx, src, dst, MapParams{});
        case kThreshold: return launch_point_t<kThreshold>(ctx, src, dst, map_params(ByteMap::step(thresh, maxval, 0)));
    }
    return fail(ctx, DIP_ERR_ARG, "bad point op");
}

namespace {

template <bool MAPPED>
int launch_gray3(dip_ctx* ctx, const ImgView& src, const ImgView& dst, const MapParams& pre, const MapParams& post) {
    if (dense_pair(src, dst) && src.rowbytes % 48 == 0) {
        const size_t nvec = src.pitch * (size_t)src.height * (size_t)src.frames / 16;
        const size_t blocks = (nvec + kThreads * 3 - 1) / (kThreads * 3);
        if (blocks > 0x7FFFFFFFull) return fail(ctx, DIP_ERR_UNSUPPORTED, "image stack too large for one launch");
        const LaunchMode mode = plan_launch(ctx, src, dst);
        DIP_CUDA(ctx, launch_kernel(mode, gray3_flat_kernel<MAPPED>, dim3((unsigned)blocks), dim3(kThreads), 0, ctx->compute,
                                    reinterpret_cast<const uint4*>(src.base), reinterpret_cast<uint4*>(dst.base), nvec, pre, post, mode != kNoWait));
        DIP_CHECK_LAUNCH(ctx, "gray3_flat_kernel");
        return DIP_OK;
    }
    const int groups = (src.width + 15) / 16;
    const size_t total = (size_t)groups * (size_t)src.height * (size_t)src.frames;
    const size_t blocks = (total + kThreads - 1) / kThreads;
    if (blocks > 0x7FFFFFFFull) return fail(ctx, DIP_ERR_UNSUPPORTED, "image stack too large for one launch");
    if (exact_rows(src) || exact_rows(dst))
        DIP_CUDA(ctx, launch_kernel(plan_launch(ctx, nullptr, 0), gray3_kernel<true, MAPPED>, dim3((unsigned)blocks), dim3(kThreads), 0, ctx->compute, src, dst, groups, total, pre, post));
    else
        DIP_CUDA(ctx, launch_kernel(plan_launch(ctx, nullptr, 0), gray3_kernel<false, MAPPED>, dim3((unsigned)blocks), dim3(kThreads), 0, ctx->compute, src, dst, groups, total, pre, post));
    DIP_CHECK_LAUNCH(ctx, "gray3_kernel");
    return DIP_OK;
}

}  // namespace

int launch_grayscale(dip_ctx* ctx, const ImgView& src, const ImgView& dst) {
    if (src.channels == 1) return launch_point_t<kCopy>(ctx, src, dst, MapParams{});
    if (src.channels == 4) return launch_point_t<kGray4>(ctx, src, dst, MapParams{});
    return launch_gray3<false>(ctx, src, dst, MapParams{}, MapParams{});
}

int launch_point_chain(dip_ctx* ctx, const ImgView& src, const ImgView& dst, const ByteMap& pre, bool gray, const ByteMap& post) {
    if (!gray || src.channels == 1) {  // one-channel grayscale is the identity (cvtColor GRAY2GRAY)
        const ByteMap m = pre.then(post);
        if (m.identity()) return launch_point_t<kCopy>(ctx, src, dst, MapParams{});
        return launch_point_t<kMap>(ctx, src, dst, map_params(m));
    }
    if (pre.identity() && post.identity()) return launch_grayscale(ctx, src, dst);
    if (src.channels == 4) return launch_point_t<kMapGray4>(ctx, src, dst, map_params(pre), map_params(post));
    return launch_gray3<true>(ctx, src, dst, map_params(pre), map_params(post));
}

int launch_synth(dip_ctx* ctx, const ImgView& dst, uint32_t seed, uint64_t first_row) {
    const int vpr = (dst.rowbytes + 15) / 16;
    const size_t total = (size_t)vpr * (size_t)dst.height * (size_t)dst.frames;
    const size_t blocks = (total + kThreads - 1) / kThreads;
    if (blocks > 0x7FFFFFFFull) return fail(ctx, DIP_ERR_UNSUPPORTED, "image stack too large for one launch");
    DIP_CUDA(ctx, launch_kernel(plan_launch(ctx, nullptr, 0), synth_kernel, dim3((unsigned)blocks), dim3(kThreads), 0, ctx->compute, dst, seed, first_row, vpr, total));
    DIP_CHECK_LAUNCH(ctx, "synth_kernel");
    return DIP_OK;
}

int launch_checksum(dip_ctx* ctx, const ImgView& src, uint64_t first_row, uint64_t* d_out) {
    const int vpr = (src.rowbytes + 15) / 16;
    const size_t vpf = (size_t)vpr * (size_t)src.height;
    if (src.frames > 65535) return fail(ctx, DIP_ERR_UNSUPPORTED, "more than 65535 frames");
    const unsigned bx = (unsigned)std::min<size_t>((vpf + kThreads - 1) / kThreads, (size_t)ctx->sm_count * 8);
    dim3 grid(bx ? bx : 1, (unsigned)src.frames);
    DIP_CUDA(ctx, launch_kernel(plan_launch(ctx, nullptr, 0), checksum_kernel, grid, dim3(kThreads), 0, ctx->compute, src, vpr, vpf, first_row,
                                reinterpret_cast<unsigned long long*>(d_out)));
    DIP_CHECK_LAUNCH(ctx, "checksum_kernel");
    return DIP_OK;
}

int launch_fill(dip_ctx* ctx, uint8_t* p, size_t bytes, uint32_t value) {
    dep_touch(ctx);
    DIP_CUDA(ctx, cudaMemsetAsync(p, (int)value, bytes, ctx->compute));
    return DIP_OK;
}

}  // namespace dip

using namespace dip;

extern "C" {

int dip_copy(dip_ctx* ctx, const dip_image* src, dip_image* dst) {
    int rc = check_pair(ctx, src, dst, true);
    if (rc != DIP_OK) return rc;
    if ((rc = begin_op(ctx, src, dst)) != DIP_OK) return rc;
    return launch_point(ctx, kCopy, src->v, dst->v, 0, 0);
}

int dip_invert(dip_ctx* ctx, const dip_image* src, dip_image* dst) {
    int rc = check_pair(ctx, src, dst, true);
    if (rc != DIP_OK) return rc;
    if ((rc = begin_op(ctx, src, dst)) != DIP_OK) return rc;
    return launch_point(ctx, kInvert, src->v, dst->v, 0, 0);
}

int dip_threshold(dip_ctx* ctx, const dip_image* src, dip_image* dst, uint8_t thresh, uint8_t maxval) {
    int rc = check_pair(ctx, src, dst, true);
    if (rc != DIP_OK) return rc;
    if ((rc = begin_op(ctx, src, dst)) != DIP_OK) return rc;
    if (src->v.bps == 2) return launch_threshold16(ctx, src->v, dst->v, thresh, maxval);   // the same numbers on 16-bit samples, like cv.threshold
    return launch_point(ctx, kThreshold, src->v, dst->v, thresh, maxval);
}

int dip_threshold16(dip_ctx* ctx, const dip_image* src, dip_image* dst, uint16_t thresh, uint16_t maxval) {
    int rc = check_pair(ctx, src, dst, true);
    if (rc != DIP_OK) return rc;
    if (src->v.bps != 2) return fail(ctx, DIP_ERR_ARG, "dip_threshold16 needs 16-bit images (dip_image_create16)");
    if ((rc = begin_op(ctx, src, dst)) != DIP_OK) return rc;
    return launch_threshold16(ctx, src->v, dst->v, thresh, maxval);
}

int dip_grayscale(dip_ctx* ctx, const dip_image* src, dip_image* dst) {
    int rc = check_pair(ctx, src, dst, true);
    if (rc != DIP_OK) return rc;
    if ((rc = begin_op(ctx, src, dst)) != DIP_OK) return rc;
    if (src->v.bps == 2) return launch_grayscale16(ctx, src->v, dst->v);
    return launch_grayscale(ctx, src->v, dst->v);
}

}  // extern "C"
