// JPEG in and out on the GPU side of upload / download.
//
// The reference's harnesses read their input with cv.imread / cv::imread (opencv/benchmark.py:14, sycl/benchmark.cpp:445)
// and save every result with imwrite (opencv:114, sycl:374): a CPU decode of 8 Mpixel, an upload of 24.5 MB of raw
// pixels, and the way back.  Here the compressed bytes are what crosses PCIe: dip_jpeg_decode lands the pixels directly
// in a frame of a device stack, dip_jpeg_encode reads a frame of a device stack and returns the bitstream.
//
// The codec itself is nvJPEG (CUDA toolkit LIBRARY code, not one of this library's kernels, and not part of the
// operation hot path).  It is bound at run time (dlopen) so that libdip_b200.so has no load-time dependency on it: a
// box without libnvjpeg gets DIP_ERR_UNSUPPORTED from these four entry points and everything else works.
// Batches go through nvjpegDecodeBatched, on the hardware JPEG engines (NVJPEG_BACKEND_HARDWARE) when the GPU and the
// streams allow it (baseline Huffman only), else on the default GPU-assisted backend.
#include <dlfcn.h>
#include <nvjpeg.h>

#include <mutex>

#include "dip_internal.h"

namespace dip {

namespace {

#define DIP_NVJPEG_FUNCS(X)              \
    X(nvjpegCreateSimple)                \
    X(nvjpegCreateEx)                    \
    X(nvjpegDestroy)                     \
    X(nvjpegJpegStateCreate)             \
    X(nvjpegJpegStateDestroy)            \
    X(nvjpegGetImageInfo)                \
    X(nvjpegDecode)                      \
    X(nvjpegDecodeBatchedInitialize)     \
    X(nvjpegDecodeBatched)               \
    X(nvjpegEncoderStateCreate)          \
    X(nvjpegEncoderStateDestroy)         \
    X(nvjpegEncoderParamsCreate)         \
    X(nvjpegEncoderParamsDestroy)        \
    X(nvjpegEncoderParamsSetQuality)     \
    X(nvjpegEncoderParamsSetSamplingFactors) \
    X(nvjpegEncoderParamsSetOptimizedHuffman) \
    X(nvjpegEncodeImage)                 \
    X(nvjpegEncodeRetrieveBitstream)

struct NvjpegApi {
    void* so = nullptr;
#define X(name) decltype(&::name) name = nullptr;
    DIP_NVJPEG_FUNCS(X)
#undef X
    std::string error;
};

// process-wide, loaded once
const NvjpegApi* nvjpeg_api() {
    static NvjpegApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("DIP_NVJPEG_LIB");
        const char* names[] = {env, "libnvjpeg.so.12", "libnvjpeg.so", "/usr/local/cuda/lib64/libnvjpeg.so.12", "/usr/local/cuda/lib64/libnvjpeg.so"};
        for (const char* n : names) {
            if (!n || !n[0]) continue;
            api.so = dlopen(n, RTLD_NOW | RTLD_LOCAL);
            if (api.so) break;
        }
        if (!api.so) {
            api.error = "libnvjpeg not found (set DIP_NVJPEG_LIB)";
            return;
        }
#define X(name)                                                                \
    api.name = reinterpret_cast<decltype(&::name)>(dlsym(api.so, #name));      \
    if (!api.name && api.error.empty()) api.error = "libnvjpeg lacks " #name;
        DIP_NVJPEG_FUNCS(X)
#undef X
    });
    return &api;
}

}  // namespace

// per-ctx codec state, created on first use, released by dip_shutdown (jpeg_release)
struct JpegState {
    nvjpegHandle_t handle = nullptr;          // default backend: single decodes, encodes, batched fallback
    nvjpegJpegState_t state = nullptr;
    nvjpegHandle_t hw_handle = nullptr;       // NVJPEG_BACKEND_HARDWARE, if the GPU has the engines
    nvjpegJpegState_t hw_state = nullptr;
    nvjpegJpegState_t batch_state = nullptr;  // default backend, batched
    int hw_batch = 0, hw_fmt = -1, def_batch = 0, def_fmt = -1;  // what the batched states are initialised for
    nvjpegEncoderState_t enc_state = nullptr;
    nvjpegEncoderParams_t enc_params = nullptr;
    const char* last_backend = "none";
};

void jpeg_release(dip_ctx* ctx) {
    JpegState* s = static_cast<JpegState*>(ctx->jpeg);
    if (!s) return;
    const NvjpegApi* a = nvjpeg_api();
    if (s->enc_params) a->nvjpegEncoderParamsDestroy(s->enc_params);
    if (s->enc_state) a->nvjpegEncoderStateDestroy(s->enc_state);
    if (s->batch_state) a->nvjpegJpegStateDestroy(s->batch_state);
    if (s->hw_state) a->nvjpegJpegStateDestroy(s->hw_state);
    if (s->hw_handle) a->nvjpegDestroy(s->hw_handle);
    if (s->state) a->nvjpegJpegStateDestroy(s->state);
    if (s->handle) a->nvjpegDestroy(s->handle);
    delete s;
    ctx->jpeg = nullptr;
}

namespace {

int nvfail(dip_ctx* ctx, nvjpegStatus_t st, const char* what) {
    const int code = (st == NVJPEG_STATUS_ALLOCATOR_FAILURE) ? DIP_ERR_NOMEM
                   : (st == NVJPEG_STATUS_BAD_JPEG || st == NVJPEG_STATUS_INVALID_PARAMETER) ? DIP_ERR_ARG
                   : (st == NVJPEG_STATUS_JPEG_NOT_SUPPORTED || st == NVJPEG_STATUS_ARCH_MISMATCH) ? DIP_ERR_UNSUPPORTED
                   : DIP_ERR_CUDA;
    return fail(ctx, code, "%s failed with nvjpeg status %d", what, (int)st);
}
#define DIP_NVJPEG(ctx, expr)                                          \
    do {                                                               \
        nvjpegStatus_t _s = (expr);                                    \
        if (_s != NVJPEG_STATUS_SUCCESS) return nvfail((ctx), _s, #expr); \
    } while (0)

int jpeg_state(dip_ctx* ctx, JpegState** out) {
    if (!ctx) return fail(nullptr, DIP_ERR_ARG, "null ctx");
    if (ctx->capturing) return fail(ctx, DIP_ERR_UNSUPPORTED, "JPEG calls cannot be recorded into a graph");
    const NvjpegApi* a = nvjpeg_api();
    if (!a->error.empty()) return fail(ctx, DIP_ERR_UNSUPPORTED, "nvJPEG unavailable: %s", a->error.c_str());
    DIP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->jpeg) {
        JpegState* s = new JpegState();
        ctx->jpeg = s;
        DIP_NVJPEG(ctx, a->nvjpegCreateSimple(&s->handle));
        DIP_NVJPEG(ctx, a->nvjpegJpegStateCreate(s->handle, &s->state));
        // the hardware engines are optional: absent on some parts, and refused when the driver has none to give
        const char* no_hw = getenv("DIP_JPEG_NO_HW");
        if (!(no_hw && no_hw[0] == '1') &&
            a->nvjpegCreateEx(NVJPEG_BACKEND_HARDWARE, nullptr, nullptr, NVJPEG_FLAGS_DEFAULT, &s->hw_handle) == NVJPEG_STATUS_SUCCESS) {
            if (a->nvjpegJpegStateCreate(s->hw_handle, &s->hw_state) != NVJPEG_STATUS_SUCCESS) {
                a->nvjpegDestroy(s->hw_handle);
                s->hw_handle = nullptr;
                s->hw_state = nullptr;
            }
        } else {
            s->hw_handle = nullptr;
        }
    }
    *out = static_cast<JpegState*>(ctx->jpeg);
    return DIP_OK;
}

int check_target(dip_ctx* ctx, const dip_image* img, int first, int count, const char* what) {
    if (!img) return fail(ctx, DIP_ERR_ARG, "%s: null image", what);
    if (img->device != ctx->device) return fail(ctx, DIP_ERR_ARG, "%s: the image lives on another device", what);
    if (img->v.channels != 3 && img->v.channels != 1) return fail(ctx, DIP_ERR_UNSUPPORTED, "%s: 1- or 3-channel images only", what);
    if (first < 0 || count < 0 || first + count > img->v.frames) return fail(ctx, DIP_ERR_ARG, "%s: frames [%d, %d) outside the stack of %d", what, first, first + count, img->v.frames);
    return DIP_OK;
}

// the compute stream is about to get work from another library: order it after outstanding copies on the image
int order_after_copies(dip_ctx* ctx, const dip_image* img) {
    const dip_image* root = root_of(const_cast<dip_image*>(img));
    if (root->copy_pending && root->copy_done) DIP_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, root->copy_done, 0));
    dep_touch(ctx);
    return DIP_OK;
}

}  // namespace

}  // namespace dip

using namespace dip;

extern "C" {

int dip_jpeg_info(dip_ctx* ctx, const uint8_t* data, size_t len, int* width, int* height, int* components) {
    JpegState* s = nullptr;
    int rc = jpeg_state(ctx, &s);
    if (rc != DIP_OK) return rc;
    if (!data || !len) return fail(ctx, DIP_ERR_ARG, "dip_jpeg_info: empty stream");
    int ncomp = 0;
    nvjpegChromaSubsampling_t ss;
    int w[NVJPEG_MAX_COMPONENT] = {0}, h[NVJPEG_MAX_COMPONENT] = {0};
    DIP_NVJPEG(ctx, nvjpeg_api()->nvjpegGetImageInfo(s->handle, data, len, &ncomp, &ss, w, h));
    if (w[0] <= 0 || h[0] <= 0) return fail(ctx, DIP_ERR_ARG, "dip_jpeg_info: empty image");
    if (width) *width = w[0];
    if (height) *height = h[0];
    if (components) *components = ncomp;
    return DIP_OK;
}

int dip_jpeg_decode(dip_ctx* ctx, const uint8_t* data, size_t len, dip_image* dst, int frame) {
    JpegState* s = nullptr;
    int rc = jpeg_state(ctx, &s);
    if (rc != DIP_OK) return rc;
    if ((rc = check_target(ctx, dst, frame, 1, "dip_jpeg_decode")) != DIP_OK) return rc;
    int w = 0, h = 0;
    if ((rc = dip_jpeg_info(ctx, data, len, &w, &h, nullptr)) != DIP_OK) return rc;
    if (w != dst->v.width || h != dst->v.height) return fail(ctx, DIP_ERR_SHAPE, "dip_jpeg_decode: the stream is %dx%d, the image %dx%d", w, h, dst->v.width, dst->v.height);
    if ((rc = order_after_copies(ctx, dst)) != DIP_OK) return rc;
    nvjpegImage_t out{};
    out.channel[0] = dst->v.base + (size_t)frame * dst->v.frame_stride;
    out.pitch[0] = dst->v.pitch;
    DIP_NVJPEG(ctx, nvjpeg_api()->nvjpegDecode(s->handle, s->state, data, len, dst->v.channels == 3 ? NVJPEG_OUTPUT_BGRI : NVJPEG_OUTPUT_Y, &out, ctx->compute));
    s->last_backend = "gpu-hybrid";
    return DIP_OK;
}

int dip_jpeg_decode_batch(dip_ctx* ctx, const uint8_t* const* data, const size_t* len, int count, dip_image* dst, int first_frame) {
    JpegState* s = nullptr;
    int rc = jpeg_state(ctx, &s);
    if (rc != DIP_OK) return rc;
    if (!data || !len) return fail(ctx, DIP_ERR_ARG, "dip_jpeg_decode_batch: null stream table");
    if ((rc = check_target(ctx, dst, first_frame, count, "dip_jpeg_decode_batch")) != DIP_OK) return rc;
    if (count == 0) return DIP_OK;
    const NvjpegApi* a = nvjpeg_api();
    for (int i = 0; i < count; ++i) {
        int w = 0, h = 0;
        if ((rc = dip_jpeg_info(ctx, data[i], len[i], &w, &h, nullptr)) != DIP_OK) return rc;
        if (w != dst->v.width || h != dst->v.height) return fail(ctx, DIP_ERR_SHAPE, "dip_jpeg_decode_batch: stream %d is %dx%d, the image %dx%d", i, w, h, dst->v.width, dst->v.height);
    }
    if ((rc = order_after_copies(ctx, dst)) != DIP_OK) return rc;
    const nvjpegOutputFormat_t fmt = dst->v.channels == 3 ? NVJPEG_OUTPUT_BGRI : NVJPEG_OUTPUT_Y;
    std::vector<nvjpegImage_t> outs((size_t)count);
    for (int i = 0; i < count; ++i) {
        outs[i] = nvjpegImage_t{};
        outs[i].channel[0] = dst->v.base + (size_t)(first_frame + i) * dst->v.frame_stride;
        outs[i].pitch[0] = dst->v.pitch;
    }
    // 1. the hardware engines (baseline streams only; anything they refuse falls through)
    if (s->hw_handle) {
        bool ok = true;
        if (s->hw_batch != count || s->hw_fmt != (int)fmt) {
            ok = a->nvjpegDecodeBatchedInitialize(s->hw_handle, s->hw_state, count, 1, fmt) == NVJPEG_STATUS_SUCCESS;
            s->hw_batch = ok ? count : 0;
            s->hw_fmt = (int)fmt;
        }
        if (ok && a->nvjpegDecodeBatched(s->hw_handle, s->hw_state, data, len, outs.data(), ctx->compute) == NVJPEG_STATUS_SUCCESS) {
            s->last_backend = "hardware";
            return DIP_OK;
        }
        s->hw_batch = 0;
    }
    // 2. the default backend, batched
    if (!s->batch_state && a->nvjpegJpegStateCreate(s->handle, &s->batch_state) != NVJPEG_STATUS_SUCCESS) s->batch_state = nullptr;
    if (s->batch_state) {
        bool ok = true;
        if (s->def_batch != count || s->def_fmt != (int)fmt) {
            ok = a->nvjpegDecodeBatchedInitialize(s->handle, s->batch_state, count, 1, fmt) == NVJPEG_STATUS_SUCCESS;
            s->def_batch = ok ? count : 0;
            s->def_fmt = (int)fmt;
        }
        if (ok && a->nvjpegDecodeBatched(s->handle, s->batch_state, data, len, outs.data(), ctx->compute) == NVJPEG_STATUS_SUCCESS) {
            s->last_backend = "gpu-hybrid-batched";
            return DIP_OK;
        }
        s->def_batch = 0;
    }
    // 3. one by one
    for (int i = 0; i < count; ++i)
        DIP_NVJPEG(ctx, a->nvjpegDecode(s->handle, s->state, data[i], len[i], fmt, &outs[i], ctx->compute));
    s->last_backend = "gpu-hybrid";
    return DIP_OK;
}

int dip_jpeg_encode(dip_ctx* ctx, const dip_image* src, int frame, int quality, uint8_t* out, size_t capacity, size_t* len) {
    JpegState* s = nullptr;
    int rc = jpeg_state(ctx, &s);
    if (rc != DIP_OK) return rc;
    if ((rc = check_target(ctx, src, frame, 1, "dip_jpeg_encode")) != DIP_OK) return rc;
    if (src->v.channels != 3) return fail(ctx, DIP_ERR_UNSUPPORTED, "dip_jpeg_encode: 3-channel images only");
    if (!len) return fail(ctx, DIP_ERR_ARG, "dip_jpeg_encode: null length");
    if (quality < 1 || quality > 100) return fail(ctx, DIP_ERR_ARG, "dip_jpeg_encode: quality %d outside [1, 100]", quality);
    const NvjpegApi* a = nvjpeg_api();
    if (!s->enc_state) {
        DIP_NVJPEG(ctx, a->nvjpegEncoderStateCreate(s->handle, &s->enc_state, ctx->compute));
        DIP_NVJPEG(ctx, a->nvjpegEncoderParamsCreate(s->handle, &s->enc_params, ctx->compute));
        // cv.imwrite's defaults (opencv:114): 4:2:0 chroma, standard Huffman tables
        DIP_NVJPEG(ctx, a->nvjpegEncoderParamsSetSamplingFactors(s->enc_params, NVJPEG_CSS_420, ctx->compute));
        DIP_NVJPEG(ctx, a->nvjpegEncoderParamsSetOptimizedHuffman(s->enc_params, 0, ctx->compute));
    }
    DIP_NVJPEG(ctx, a->nvjpegEncoderParamsSetQuality(s->enc_params, quality, ctx->compute));
    if ((rc = order_after_copies(ctx, src)) != DIP_OK) return rc;
    nvjpegImage_t in{};
    in.channel[0] = src->v.base + (size_t)frame * src->v.frame_stride;
    in.pitch[0] = src->v.pitch;
    DIP_NVJPEG(ctx, a->nvjpegEncodeImage(s->handle, s->enc_state, s->enc_params, &in, NVJPEG_INPUT_BGRI, src->v.width, src->v.height, ctx->compute));
    size_t need = 0;
    DIP_NVJPEG(ctx, a->nvjpegEncodeRetrieveBitstream(s->handle, s->enc_state, nullptr, &need, ctx->compute));
    *len = need;  // the size the stream needs, also when it does not fit
    if (!out || capacity < need) return fail(ctx, DIP_ERR_NOMEM, "dip_jpeg_encode: the stream needs %zu bytes, the buffer has %zu", need, out ? capacity : (size_t)0);
    DIP_NVJPEG(ctx, a->nvjpegEncodeRetrieveBitstream(s->handle, s->enc_state, out, &need, ctx->compute));
    DIP_CUDA(ctx, cudaStreamSynchronize(ctx->compute));  // the bitstream is in `out` when the call returns
    *len = need;
    return DIP_OK;
}

const char* dip_jpeg_backend(const dip_ctx* ctx) {
    if (!ctx || !ctx->jpeg) return "none";
    return static_cast<const JpegState*>(ctx->jpeg)->last_backend;
}

}  // extern "C"
