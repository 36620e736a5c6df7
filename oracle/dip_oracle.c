/*
 * dip_oracle.c -- CPU restatement of the reference's image-op hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing shipped in dip-benchmark_b200/ may link,
 * import or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * What it restates: the 12 device operations of the reference's OpenCV
 * backend, /root/reference/benchmarks/opencv/benchmark.py:94-105, with the
 * masks and kernels of :46-72.  The arithmetic itself lives in a third-party
 * dependency that is NOT vendored in the reference: OpenCV (opencv-python
 * pinned 4.11.0.86 in benchmarks/opencv/uv.lock:62-63, ">=4.10.0.84" in
 * pyproject.toml:8; README ran 4.10.0).  The formulas below are OpenCV 4.x's
 * published algorithms for u8 images:
 *   - bitwise_not, threshold(THRESH_BINARY): per-byte.
 *   - cvtColor BGR2GRAY: 15-bit fixed point BT.601,
 *       Y = (B*3735 + G*19235 + R*9798 + 16384) >> 15, then GRAY2BGR(A)
 *       replicates Y (alpha = 255).
 *   - erode: min over the mask's non-zero taps, anchor = centre,
 *       BORDER_CONSTANT with morphologyDefaultBorderValue (+inf): taps outside
 *       the image do not take part.
 *   - filter2D (ddepth -1, float kernel): correlation, anchor = centre,
 *       BORDER_REFLECT_101, float accumulation over the kernel's non-zero
 *       coefficients in row-major order, saturate_cast<uchar>(cvRound(sum))
 *       (round-half-to-even).
 *   - GaussianBlur(ksize 3x3, sigma 0): the bit-exact fixed-point path,
 *       taps [1/4 1/2 1/4] as 8.8 fixed point, horizontal pass kept in 16 bit,
 *       vertical pass rounded half-up: out = (S + 8) >> 4 with S the
 *       [1 2 1]x[1 2 1] integer sum; BORDER_REFLECT_101.
 *
 * Parity pin: the reference has no tests and no golden vectors
 * (SURVEY.md section 4), so this file is pinned against outputs of the
 * reference's own cv2 calls run in the build container (cv2 4.13.0):
 * tests/golden/make_golden.py generated tests/golden/ (npz files) and kat.json, and
 * tests/test_oracle.py checks every function here against them.
 *
 * Layout everywhere: row-major H x W x C interleaved u8, `pitch` bytes per row.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#if defined(_OPENMP)
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

ORC_API int orc_max_threads(void) {
#if defined(_OPENMP)
    return omp_get_max_threads();
#else
    return 1;
#endif
}

ORC_API void orc_set_threads(int n) {
#if defined(_OPENMP)
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}

/* cv::borderInterpolate(p, len, BORDER_REFLECT_101): -1 -> 1, len -> len-2,
 * repeated until in range; a length-1 axis maps everything to 0. */
static inline int reflect101(int p, int len) {
    if (len == 1) return 0;
    while (p < 0 || p >= len) {
        if (p < 0) p = -p;
        else p = 2 * (len - 1) - p;
    }
    return p;
}

/* opencv/benchmark.py:94 (intended semantics, sycl/benchmark.cpp:333): dst = src */
ORC_API void orc_copy(const uint8_t* src, size_t spitch, uint8_t* dst, size_t dpitch,
                      int w, int h, int c) {
    const size_t rowb = (size_t)w * c;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) memcpy(dst + y * dpitch, src + y * spitch, rowb);
}

/* opencv/benchmark.py:95  cv.bitwise_not */
ORC_API void orc_invert(const uint8_t* src, size_t spitch, uint8_t* dst, size_t dpitch,
                        int w, int h, int c) {
    const size_t rowb = (size_t)w * c;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src + y * spitch;
        uint8_t* d = dst + y * dpitch;
        for (size_t i = 0; i < rowb; ++i) d[i] = (uint8_t)(255 - s[i]);
    }
}

/* opencv/benchmark.py:97  cv.threshold(image, 127, 255, THRESH_BINARY) */
ORC_API void orc_threshold(const uint8_t* src, size_t spitch, uint8_t* dst, size_t dpitch,
                           int w, int h, int c, uint8_t thresh, uint8_t maxval) {
    const size_t rowb = (size_t)w * c;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src + y * spitch;
        uint8_t* d = dst + y * dpitch;
        for (size_t i = 0; i < rowb; ++i) d[i] = s[i] > thresh ? maxval : 0;
    }
}

/* opencv/benchmark.py:96  cvtColor(cvtColor(BGR2GRAY), GRAY2BGR).
 * c == 3: BGR -> Y,Y,Y ; c == 4: BGRA -> Y,Y,Y,255 ; c == 1: identity. */
ORC_API void orc_grayscale(const uint8_t* src, size_t spitch, uint8_t* dst, size_t dpitch,
                           int w, int h, int c) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src + y * spitch;
        uint8_t* d = dst + y * dpitch;
        if (c == 1) { memcpy(d, s, (size_t)w); continue; }
        for (int x = 0; x < w; ++x) {
            const uint32_t b = s[x * c + 0], g = s[x * c + 1], r = s[x * c + 2];
            const uint8_t yv = (uint8_t)((b * 3735u + g * 19235u + r * 9798u + 16384u) >> 15);
            d[x * c + 0] = yv; d[x * c + 1] = yv; d[x * c + 2] = yv;
            if (c == 4) d[x * c + 3] = 255;
        }
    }
}

/* opencv/benchmark.py:98-100 (and :76-78)  cv.erode(image, mask): anchor centre,
 * one iteration, out-of-image taps ignored (+inf border). mask is mh x mw u8. */
ORC_API int orc_erode(const uint8_t* src, size_t spitch, uint8_t* dst, size_t dpitch,
                      int w, int h, int c, const uint8_t* mask, int mw, int mh) {
    if (mw < 1 || mh < 1) return -1;
    const int ax = mw / 2, ay = mh / 2;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        uint8_t* d = dst + y * dpitch;
        for (int x = 0; x < w; ++x) {
            for (int ch = 0; ch < c; ++ch) {
                uint8_t m = 255;
                for (int j = 0; j < mh; ++j) {
                    const int yy = y + j - ay;
                    if (yy < 0 || yy >= h) continue;
                    const uint8_t* s = src + yy * spitch;
                    for (int i = 0; i < mw; ++i) {
                        if (!mask[j * mw + i]) continue;
                        const int xx = x + i - ax;
                        if (xx < 0 || xx >= w) continue;
                        const uint8_t v = s[xx * c + ch];
                        if (v < m) m = v;
                    }
                }
                d[x * c + ch] = m;
            }
        }
    }
    return 0;
}

/* same structure as orc_erode with max and a -inf (0) border; cv.dilate is
 * defined but unused by the reference (opencv/benchmark.py:80-82). */
ORC_API int orc_dilate(const uint8_t* src, size_t spitch, uint8_t* dst, size_t dpitch,
                       int w, int h, int c, const uint8_t* mask, int mw, int mh) {
    if (mw < 1 || mh < 1) return -1;
    const int ax = mw / 2, ay = mh / 2;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        uint8_t* d = dst + y * dpitch;
        for (int x = 0; x < w; ++x) {
            for (int ch = 0; ch < c; ++ch) {
                uint8_t m = 0;
                for (int j = 0; j < mh; ++j) {
                    const int yy = y + j - ay;
                    if (yy < 0 || yy >= h) continue;
                    const uint8_t* s = src + yy * spitch;
                    for (int i = 0; i < mw; ++i) {
                        if (!mask[j * mw + i]) continue;
                        const int xx = x + i - ax;
                        if (xx < 0 || xx >= w) continue;
                        const uint8_t v = s[xx * c + ch];
                        if (v > m) m = v;
                    }
                }
                d[x * c + ch] = m;
            }
        }
    }
    return 0;
}

/* opencv/benchmark.py:101-104 (and :84-90)  cv.filter2D(image, -1, kernel):
 * correlation with a kh x kw float kernel, anchor centre, BORDER_REFLECT_101,
 * float accumulate over non-zero coefficients in row-major order,
 * cvRound (round-half-even; lrintf under the default rounding mode) + saturate. */
ORC_API int orc_filter2d(const uint8_t* src, size_t spitch, uint8_t* dst, size_t dpitch,
                         int w, int h, int c, const float* kernel, int kw, int kh) {
    if (kw < 1 || kh < 1) return -1;
    const int ax = kw / 2, ay = kh / 2;
    int* xmap = (int*)malloc(sizeof(int) * (size_t)(w + kw));
    if (!xmap) return -2;
    for (int i = 0; i < w + kw; ++i) xmap[i] = reflect101(i - ax, w);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        uint8_t* d = dst + y * dpitch;
        for (int x = 0; x < w; ++x) {
            for (int ch = 0; ch < c; ++ch) {
                float acc = 0.0f;
                for (int j = 0; j < kh; ++j) {
                    const uint8_t* s = src + (size_t)reflect101(y + j - ay, h) * spitch;
                    for (int i = 0; i < kw; ++i) {
                        const float kf = kernel[j * kw + i];
                        if (kf == 0.0f) continue;
                        acc += kf * (float)s[xmap[x + i] * c + ch];
                    }
                }
                long r = lrintf(acc);
                d[x * c + ch] = (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
            }
        }
    }
    free(xmap);
    return 0;
}

/* opencv/benchmark.py:105  cv.GaussianBlur(image, (3,3), 0): OpenCV's bit-exact
 * fixed-point smoothing.  getGaussianKernel(3, 0) = [1/4 1/2 1/4] -> 8.8 fixed
 * point {64,128,64}; the horizontal pass is kept as a 16-bit 8.8 value, the
 * vertical pass accumulates in 32 bit and rounds half-up back to u8. */
ORC_API void orc_gaussian_blur3x3(const uint8_t* src, size_t spitch, uint8_t* dst, size_t dpitch,
                                  int w, int h, int c) {
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        const uint8_t* r0 = src + (size_t)reflect101(y - 1, h) * spitch;
        const uint8_t* r1 = src + (size_t)y * spitch;
        const uint8_t* r2 = src + (size_t)reflect101(y + 1, h) * spitch;
        uint8_t* d = dst + y * dpitch;
        for (int x = 0; x < w; ++x) {
            const int xl = reflect101(x - 1, w) * c, xc = x * c, xr = reflect101(x + 1, w) * c;
            for (int ch = 0; ch < c; ++ch) {
                /* 8.8 fixed point rows: (a*64 + b*128 + c*64) fits 16 bit */
                const uint32_t h0 = r0[xl + ch] * 64u + r0[xc + ch] * 128u + r0[xr + ch] * 64u;
                const uint32_t h1 = r1[xl + ch] * 64u + r1[xc + ch] * 128u + r1[xr + ch] * 64u;
                const uint32_t h2 = r2[xl + ch] * 64u + r2[xc + ch] * 128u + r2[xr + ch] * 64u;
                const uint32_t v = h0 * 64u + h1 * 128u + h2 * 64u; /* 16.16 */
                d[xc + ch] = (uint8_t)((v + (1u << 15)) >> 16);
            }
        }
    }
}

/* Synthetic image generator shared (by formula) with the CUDA library's
 * dip_synth_fill: byte i of frame f is the top byte of a 32-bit mix of the
 * 64-bit byte index and the seed.  Used by bench.py and the parity tests so
 * that device-generated batches can be spot-checked on the host. */
static inline uint32_t orc_mix32(uint64_t idx, uint32_t seed) {
    uint32_t x = (uint32_t)idx ^ ((uint32_t)(idx >> 32) * 0x9E3779B9u) ^ (seed * 0x85EBCA6Bu + 0xC2B2AE35u);
    x ^= x >> 16; x *= 0x7FEB352Du;
    x ^= x >> 15; x *= 0x846CA68Bu;
    x ^= x >> 16;
    return x;
}

ORC_API void orc_synth_fill_rows(uint8_t* dst, size_t pitch, int w, int h, int c, uint32_t seed, uint64_t first_row);

ORC_API void orc_synth_fill(uint8_t* dst, size_t pitch, int w, int h, int c,
                            uint32_t seed, uint64_t frame) {
    orc_synth_fill_rows(dst, pitch, w, h, c, seed, frame * (uint64_t)h);
}

/* rows addressed globally: a row band of a tall mosaic fills only its own rows */
ORC_API void orc_synth_fill_rows(uint8_t* dst, size_t pitch, int w, int h, int c, uint32_t seed, uint64_t first_row) {
    const size_t rowb = (size_t)w * c;
    const uint64_t base = first_row * rowb;
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; ++y) {
        uint8_t* d = dst + y * pitch;
        for (size_t i = 0; i < rowb; ++i)
            d[i] = (uint8_t)(orc_mix32(base + (uint64_t)y * rowb + i, seed) >> 24);
    }
}

/* 64-bit FNV-1a style checksum over the valid bytes of an image, used for
 * "checksum of checksums" comparisons at sizes too big to ship goldens for. */
ORC_API uint64_t orc_checksum_rows(const uint8_t* src, size_t pitch, int w, int h, int c, uint64_t first_row);

ORC_API uint64_t orc_checksum(const uint8_t* src, size_t pitch, int w, int h, int c) {
    return orc_checksum_rows(src, pitch, w, h, c, 0);
}

/* rows weighted as global rows first_row + y: band checksums of a mosaic add up mod 2^64 */
ORC_API uint64_t orc_checksum_rows(const uint8_t* src, size_t pitch, int w, int h, int c, uint64_t first_row) {
    const size_t rowb = (size_t)w * c;
    uint64_t total = 0;
#pragma omp parallel for schedule(static) reduction(+ : total)
    for (int y = 0; y < h; ++y) {
        const uint8_t* s = src + y * pitch;
        uint64_t acc = 0;
        for (size_t i = 0; i < rowb; ++i) {
            const uint64_t pos = (first_row + (uint64_t)y) * rowb + i;
            acc += (uint64_t)(s[i] + 1u) * (uint64_t)orc_mix32(pos, 0x1234567u);
        }
        total += acc;
    }
    return total;
}
