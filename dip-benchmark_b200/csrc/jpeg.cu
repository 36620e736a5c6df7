// JPEG in and out on the GPU side of upload / download.
//
// The reference's harnesses read their input with cv.imread / cv::imread (opencv/benchmark.py:14, sycl/benchmark.cpp:445)
// and save every result with imwrite (opencv:114, sycl:374): a CPU decode of 8 Mpixel, an upload of 24.5 MB of raw
// pixels, and the way back.  Here the compressed bytes are what crosses PCIe: dip_jpeg_decode lands the pixels directly
// in a frame of a device stack, dip_jpeg_encode reads a frame of a device stack and returns the bitstream.
//
// The codec itself is nvJPEG (CUDA toolkit LIBRARY code, not one of this library's kernels, and not part of the
// operation hot path).  It is bound at run time (dlopen) so that libdip_b200.so has no load-time dependency on it: a
// box without libnvjpeg gets DIP_ERR_UNSUPPORTED from these entry points and everything else works.
//
// Decoding uses nvJPEG's decoupled phases (parse, host, transfer, device) with the GPU_HYBRID decoder -- Huffman
// decoding on the GPU; what remains on the host is a scan of the bitstream -- and falls back to the HYBRID decoder (CPU
// Huffman) for streams GPU_HYBRID refuses (progressive ...).  Measured on B200 / nvJPEG 12.4 for the 3504x2336 fundus
// image (tools/nvjpeg_probe.py): the one-call API nvjpegDecode 26.6 ms whatever backend the handle was created with,
// decoupled HYBRID 26.5 ms (25.3 of them host Huffman), decoupled GPU_HYBRID 10.4 ms (9.6 host scan, 0.4 device);
// NVJPEG_BACKEND_HARDWARE is refused (ARCH_MISMATCH): this nvJPEG drives no JPEG engine on sm_100.  The host phase is
// the cost and it is per image, so a batch is decoded by several host threads, each with its own decoder state, pinned
// and device buffers and CUDA stream (the nvJPEG handle is thread-safe, states are not); the compute stream is ordered
// before and after them with events.
#include <dlfcn.h>
#include <nvjpeg.h>

#include <algorithm>
#include <atomic>
#include <mutex>
#include <thread>

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
    X(nvjpegDecoderCreate)               \
    X(nvjpegDecoderDestroy)              \
    X(nvjpegDecoderStateCreate)          \
    X(nvjpegBufferPinnedCreate)          \
    X(nvjpegBufferPinnedDestroy)         \
    X(nvjpegBufferDeviceCreate)          \
    X(nvjpegBufferDeviceDestroy)         \
    X(nvjpegStateAttachPinnedBuffer)     \
    X(nvjpegStateAttachDeviceBuffer)     \
    X(nvjpegJpegStreamCreate)            \
    X(nvjpegJpegStreamDestroy)           \
    X(nvjpegJpegStreamParse)             \
    X(nvjpegDecodeParamsCreate)          \
    X(nvjpegDecodeParamsDestroy)         \
    X(nvjpegDecodeParamsSetOutputFormat) \
    X(nvjpegDecoderJpegSupported)        \
    X(nvjpegDecodeJpegHost)              \
    X(nvjpegDecodeJpegTransferToDevice)  \
    X(nvjpegDecodeJpegDevice)            \
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

// One decoding lane: everything a host thread needs to take streams through the four phases on its own CUDA stream.
struct JpegLane {
    nvjpegJpegDecoder_t dec[2] = {nullptr, nullptr};   // [0] GPU_HYBRID, [1] HYBRID
    nvjpegJpegState_t st[2] = {nullptr, nullptr};
    nvjpegBufferPinned_t pinned[2] = {nullptr, nullptr};  // alternate between consecutive images of the lane
    nvjpegBufferDevice_t device = nullptr;
    nvjpegJpegStream_t stream[2] = {nullptr, nullptr};
    nvjpegDecodeParams_t params = nullptr;
    cudaStream_t cuda = nullptr;
    cudaEvent_t done = nullptr;
};

// per-ctx codec state, created on first use, released by dip_shutdown (jpeg_release)
struct JpegState {
    nvjpegHandle_t handle = nullptr;
    std::vector<JpegLane*> lanes;
    nvjpegEncoderState_t enc_state = nullptr;
    nvjpegEncoderParams_t enc_params = nullptr;
    cudaEvent_t ready = nullptr;              // "the compute stream reached the decode call"
    const char* last_backend = "none";
    int last_lanes = 0;
};

static void lane_destroy(const NvjpegApi* a, JpegLane* l) {
    if (!l) return;
    if (l->cuda) cudaStreamSynchronize(l->cuda);
    if (l->params) a->nvjpegDecodeParamsDestroy(l->params);
    for (int i = 0; i < 2; ++i) {
        if (l->stream[i]) a->nvjpegJpegStreamDestroy(l->stream[i]);
        if (l->st[i]) a->nvjpegJpegStateDestroy(l->st[i]);
        if (l->dec[i]) a->nvjpegDecoderDestroy(l->dec[i]);
        if (l->pinned[i]) a->nvjpegBufferPinnedDestroy(l->pinned[i]);
    }
    if (l->device) a->nvjpegBufferDeviceDestroy(l->device);
    if (l->done) cudaEventDestroy(l->done);
    if (l->cuda) cudaStreamDestroy(l->cuda);
    delete l;
}

void jpeg_release(dip_ctx* ctx) {
    JpegState* s = static_cast<JpegState*>(ctx->jpeg);
    if (!s) return;
    const NvjpegApi* a = nvjpeg_api();
    for (JpegLane* l : s->lanes) lane_destroy(a, l);
    if (s->enc_params) a->nvjpegEncoderParamsDestroy(s->enc_params);
    if (s->enc_state) a->nvjpegEncoderStateDestroy(s->enc_state);
    if (s->ready) cudaEventDestroy(s->ready);
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
        DIP_CUDA(ctx, cudaEventCreateWithFlags(&s->ready, cudaEventDisableTiming));
    }
    *out = static_cast<JpegState*>(ctx->jpeg);
    return DIP_OK;
}

int check_target(dip_ctx* ctx, const dip_image* img, int first, int count, const char* what) {
    if (!img) return fail(ctx, DIP_ERR_ARG, "%s: null image", what);
    if (img->device != ctx->device) return fail(ctx, DIP_ERR_ARG, "%s: the image lives on another device", what);
    if (img->v.channels != 3 && img->v.channels != 1) return fail(ctx, DIP_ERR_UNSUPPORTED, "%s: 1- or 3-channel images only", what);
    if (img->v.bps != 1) return fail(ctx, DIP_ERR_UNSUPPORTED, "%s: 8-bit images only", what);
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

int lane_create(dip_ctx* ctx, JpegState* s, JpegLane** out) {
    const NvjpegApi* a = nvjpeg_api();
    JpegLane* l = new JpegLane();
    auto bail = [&](int rc) { lane_destroy(a, l); return rc; };
#define DIP_LANE(expr)                                                              \
    do {                                                                            \
        nvjpegStatus_t _s = (expr);                                                 \
        if (_s != NVJPEG_STATUS_SUCCESS) return bail(nvfail(ctx, _s, #expr));       \
    } while (0)
    if (cudaStreamCreateWithFlags(&l->cuda, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&l->done, cudaEventDisableTiming) != cudaSuccess)
        return bail(fail(ctx, DIP_ERR_CUDA, "dip_jpeg_decode: cannot create a lane's stream"));
    // GPU_HYBRID is optional (DIP_JPEG_CPU_HUFFMAN=1 skips it: A/B), HYBRID is the fallback every stream can take
    const char* cpu_only = getenv("DIP_JPEG_CPU_HUFFMAN");
    if (!(cpu_only && cpu_only[0] == '1') && a->nvjpegDecoderCreate(s->handle, NVJPEG_BACKEND_GPU_HYBRID, &l->dec[0]) == NVJPEG_STATUS_SUCCESS) {
        if (a->nvjpegDecoderStateCreate(s->handle, l->dec[0], &l->st[0]) != NVJPEG_STATUS_SUCCESS) {
            a->nvjpegDecoderDestroy(l->dec[0]);
            l->dec[0] = nullptr;
        }
    } else {
        l->dec[0] = nullptr;
    }
    DIP_LANE(a->nvjpegDecoderCreate(s->handle, NVJPEG_BACKEND_HYBRID, &l->dec[1]));
    DIP_LANE(a->nvjpegDecoderStateCreate(s->handle, l->dec[1], &l->st[1]));
    for (int i = 0; i < 2; ++i) {
        DIP_LANE(a->nvjpegBufferPinnedCreate(s->handle, nullptr, &l->pinned[i]));
        DIP_LANE(a->nvjpegJpegStreamCreate(s->handle, &l->stream[i]));
    }
    DIP_LANE(a->nvjpegBufferDeviceCreate(s->handle, nullptr, &l->device));
    DIP_LANE(a->nvjpegDecodeParamsCreate(s->handle, &l->params));
#undef DIP_LANE
    *out = l;
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
    const uint8_t* datas[1] = {data};
    const size_t lens[1] = {len};
    return dip_jpeg_decode_batch(ctx, datas, lens, 1, dst, frame);
}

int dip_jpeg_decode_batch(dip_ctx* ctx, const uint8_t* const* data, const size_t* len, int count, dip_image* dst, int first_frame) {
    JpegState* s = nullptr;
    int rc = jpeg_state(ctx, &s);
    if (rc != DIP_OK) return rc;
    if (!data || !len) return fail(ctx, DIP_ERR_ARG, "dip_jpeg_decode_batch: null stream table");
    if ((rc = check_target(ctx, dst, first_frame, count, "dip_jpeg_decode_batch")) != DIP_OK) return rc;
    if (count == 0) return DIP_OK;
    for (int i = 0; i < count; ++i) {
        int w = 0, h = 0;
        if ((rc = dip_jpeg_info(ctx, data[i], len[i], &w, &h, nullptr)) != DIP_OK) return rc;
        if (w != dst->v.width || h != dst->v.height) return fail(ctx, DIP_ERR_SHAPE, "dip_jpeg_decode: stream %d is %dx%d, the image %dx%d", i, w, h, dst->v.width, dst->v.height);
    }
    // lanes = host threads of this call
    int want = (int)std::thread::hardware_concurrency();
    if (const char* e = getenv("DIP_JPEG_THREADS")) want = atoi(e);
    want = std::max(1, std::min({want, count, 32}));
    while ((int)s->lanes.size() < want) {
        JpegLane* l = nullptr;
        if ((rc = lane_create(ctx, s, &l)) != DIP_OK) return rc;
        s->lanes.push_back(l);
    }
    // the lanes' streams start after everything enqueued on the compute stream so far (kernels still reading the frames)
    if ((rc = order_after_copies(ctx, dst)) != DIP_OK) return rc;
    DIP_CUDA(ctx, cudaEventRecord(s->ready, ctx->compute));
    const nvjpegOutputFormat_t fmt = dst->v.channels == 3 ? NVJPEG_OUTPUT_BGRI : NVJPEG_OUTPUT_Y;
    std::atomic<int> next{0};
    std::atomic<int> used_cpu_huffman{0};
    std::vector<nvjpegStatus_t> status((size_t)want, NVJPEG_STATUS_SUCCESS);
    std::vector<cudaError_t> cuda_status((size_t)want, cudaSuccess);
    const NvjpegApi* a = nvjpeg_api();
    auto work = [&](int t) {
        JpegLane* l = s->lanes[(size_t)t];
        cudaSetDevice(ctx->device);
        cudaError_t ce = cudaStreamWaitEvent(l->cuda, s->ready, 0);
        nvjpegStatus_t st = a->nvjpegDecodeParamsSetOutputFormat(l->params, fmt);
        for (int n = 0; st == NVJPEG_STATUS_SUCCESS && ce == cudaSuccess; ++n) {
            const int i = next.fetch_add(1);
            if (i >= count) break;
            const int slot = n & 1;
            nvjpegImage_t out{};
            out.channel[0] = dst->v.base + (size_t)(first_frame + i) * dst->v.frame_stride;
            out.pitch[0] = dst->v.pitch;
            if ((st = a->nvjpegJpegStreamParse(s->handle, data[i], len[i], 0, 0, l->stream[slot])) != NVJPEG_STATUS_SUCCESS) break;
            int unsupported = 1;
            int d = 0;  // GPU_HYBRID when it takes the stream
            if (!l->dec[0] || a->nvjpegDecoderJpegSupported(l->dec[0], l->stream[slot], l->params, &unsupported) != NVJPEG_STATUS_SUCCESS || unsupported) {
                d = 1;
                used_cpu_huffman.fetch_add(1);
            }
            // the pinned buffer of this slot was last read by the transfer of image n - 2 of this lane; the device buffer
            // and the state by image n - 1's device phase: wait for the lane's stream (0.2-0.7 ms of device work against
            // 10-25 ms of host work per image, so this never actually waits)
            if (n > 0 && (ce = cudaStreamSynchronize(l->cuda)) != cudaSuccess) break;
            if ((st = a->nvjpegStateAttachPinnedBuffer(l->st[d], l->pinned[slot])) != NVJPEG_STATUS_SUCCESS) break;
            if ((st = a->nvjpegStateAttachDeviceBuffer(l->st[d], l->device)) != NVJPEG_STATUS_SUCCESS) break;
            if ((st = a->nvjpegDecodeJpegHost(s->handle, l->dec[d], l->st[d], l->params, l->stream[slot])) != NVJPEG_STATUS_SUCCESS) break;
            if ((st = a->nvjpegDecodeJpegTransferToDevice(s->handle, l->dec[d], l->st[d], l->stream[slot], l->cuda)) != NVJPEG_STATUS_SUCCESS) break;
            if ((st = a->nvjpegDecodeJpegDevice(s->handle, l->dec[d], l->st[d], &out, l->cuda)) != NVJPEG_STATUS_SUCCESS) break;
        }
        if (ce == cudaSuccess) ce = cudaEventRecord(l->done, l->cuda);
        if (ce == cudaSuccess) ce = cudaStreamSynchronize(l->cuda);  // the caller's bytes are released when the call returns
        status[(size_t)t] = st;
        cuda_status[(size_t)t] = ce;
    };
    if (want == 1) {
        work(0);
    } else {
        // the workers take streams from a shared counter, so any number of them finishes the batch: if the host refuses a
        // thread (no exception may cross the C boundary) the ones that started, and this thread, do its share
        std::vector<std::thread> threads;
        for (int t = 1; t < want; ++t) {
            try {
                threads.emplace_back(work, t);
            } catch (...) {
                break;
            }
        }
        work(0);
        for (auto& th : threads) th.join();
    }
    for (int t = 0; t < want; ++t) {
        if (cuda_status[(size_t)t] != cudaSuccess) return fail_cuda(ctx, cuda_status[(size_t)t], "dip_jpeg_decode (lane stream)", __FILE__, __LINE__);
        if (status[(size_t)t] != NVJPEG_STATUS_SUCCESS) return nvfail(ctx, status[(size_t)t], "dip_jpeg_decode (decoupled nvJPEG phase)");
        DIP_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, s->lanes[(size_t)t]->done, 0));  // later dip_* calls see the pixels
    }
    s->last_backend = used_cpu_huffman.load() == 0 ? "gpu-huffman" : used_cpu_huffman.load() == count ? "cpu-huffman" : "mixed";
    s->last_lanes = want;
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
