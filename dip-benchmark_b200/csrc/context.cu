// Context, image stacks and the host<->device transfer path of libdip_b200.so.
//
// Replaces, behind include/dip.h: sycl::queue + malloc_device + q.memcpy().wait()
// (/root/reference/benchmarks/sycl/benchmark.cpp:260-264, :331-333), cv.UMat(cpu_image) /
// image.get() (/root/reference/benchmarks/opencv/benchmark.py:42-44, :92-93).
//
// Transfer design (north_star "upload/download path"): three streams per ctx (compute,
// upload, download).  A pinned host buffer goes over PCIe in a single cudaMemcpy2DAsync on
// its copy stream; a pageable one is cut into slices that are memcpy'ed into two pinned
// staging slots and sent while the next slice is being staged.  Images remember an
// outstanding copy with an event, so kernels wait only when they touch such an image.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>

#include "copy_pool.h"
#include "dip_internal.h"

namespace dip {

static thread_local std::string g_last_error;

int fail(const dip_ctx* ctx, int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    if (ctx) const_cast<dip_ctx*>(ctx)->last_error = buf;
    return code;
}

int fail_cuda(const dip_ctx* ctx, cudaError_t e, const char* what, const char* file, int line) {
    const char* base = strrchr(file, '/');
    return fail(ctx, DIP_ERR_CUDA, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e),
                base ? base + 1 : file, line, what);
}

int check_pair(dip_ctx* ctx, const dip_image* src, const dip_image* dst, bool allow_inplace) {
    if (!ctx || !src || !dst) return fail(ctx, DIP_ERR_ARG, "null ctx/src/dst");
    const ImgView &a = src->v, &b = dst->v;
    if (a.width != b.width || a.height != b.height || a.channels != b.channels || a.frames != b.frames)
        return fail(ctx, DIP_ERR_SHAPE, "src %dx%dx%d(x%d) and dst %dx%dx%d(x%d) differ", a.width,
                    a.height, a.channels, a.frames, b.width, b.height, b.channels, b.frames);
    if (a.bps != b.bps) return fail(ctx, DIP_ERR_SHAPE, "src has %d-bit samples, dst %d-bit", 8 * a.bps, 8 * b.bps);
    if (!allow_inplace && a.base == b.base)
        return fail(ctx, DIP_ERR_SHAPE, "stencil operations cannot run in place (src == dst)");
    return DIP_OK;
}

int begin_op(dip_ctx* ctx, const dip_image* src, dip_image* dst) {
    DIP_CUDA(ctx, cudaSetDevice(ctx->device));
    dip_image* imgs[2] = {root_of(const_cast<dip_image*>(src)), root_of(dst)};
    for (dip_image* im : imgs) {
        if (im && im->copy_pending) {
            // dip_graph_begin synchronised the copy streams, so a pending flag here is only stale bookkeeping
            dep_touch(ctx);
            if (!ctx->capturing) DIP_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, im->copy_done, 0));
            im->copy_pending = false;
        }
    }
    return DIP_OK;
}

static bool dep_conflict_ranges(const DepRange* a, int na, const DepRange* r, int n) {
    for (int j = 0; j < na; ++j)
        for (int i = 0; i < n; ++i)
            if ((a[j].write || r[i].write) && a[j].lo < r[i].hi && r[i].lo < a[j].hi) return true;
    return false;
}

static bool dep_conflict(const std::vector<DepRange>& set, const DepRange* r, int n) {
    for (const DepRange& a : set)
        for (int i = 0; i < n; ++i)
            if ((a.write || r[i].write) && a.lo < r[i].hi && r[i].lo < a.hi) return true;
    return false;
}

// ---- recording: data-dependency edges (CaptureTracker, dip_internal.h) -----------------------------------------
// the node of the launch planned last: after a kernel launch the stream's capture dependencies are exactly that node
static void capture_collect(dip_ctx* ctx) {
    CaptureTracker& c = ctx->cap;
    if (c.pending_n < 0) return;
    cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
    const cudaGraphNode_t* deps = nullptr;
    size_t nd = 0;
    if (cudaStreamGetCaptureInfo_v3(ctx->compute, &st, nullptr, nullptr, &deps, nullptr, &nd) == cudaSuccess &&
        st == cudaStreamCaptureStatusActive && nd == 1) {
        CaptureNode cn{};
        cn.node = deps[0];
        cn.n = c.pending_n;
        for (int i = 0; i < c.pending_n; ++i) cn.r[i] = c.pending[i];
        c.nodes.push_back(cn);
    } else {
        cudaGetLastError();
        c.barrier = true;   // unknown state: the next kernel keeps whatever the stream depends on
    }
    c.pending_n = -1;
}

int capture_join(dip_ctx* ctx) {
    CaptureTracker& c = ctx->cap;
    capture_collect(ctx);
    if (!c.nodes.empty()) {
        std::vector<cudaGraphNode_t> all;
        for (const CaptureNode& cn : c.nodes) all.push_back(cn.node);
        DIP_CUDA(ctx, cudaStreamUpdateCaptureDependencies(ctx->compute, all.data(), all.size(), cudaStreamAddCaptureDependencies));
        c.nodes.clear();
    }
    c.barrier = true;
    return DIP_OK;
}

static LaunchMode plan_capture(dip_ctx* ctx, const DepRange* ranges, int n) {
    CaptureTracker& c = ctx->cap;
    capture_collect(ctx);
    ctx->dep.touched = false;
    constexpr size_t kCap = 64;   // a join now and then keeps the search small
    if (c.barrier || n <= 0 || n > 2 || c.nodes.size() >= kCap) {
        // depends on everything recorded so far (and on what the stream itself depends on: a non-kernel node)
        if (!c.nodes.empty()) {
            std::vector<cudaGraphNode_t> all;
            for (const CaptureNode& cn : c.nodes) all.push_back(cn.node);
            if (cudaStreamUpdateCaptureDependencies(ctx->compute, all.data(), all.size(), cudaStreamAddCaptureDependencies) != cudaSuccess) cudaGetLastError();
            c.nodes.clear();
        }
        c.barrier = false;
        c.pending_n = 0;          // it stands for everything before it
        return kFull;
    }
    std::vector<cudaGraphNode_t> deps;
    for (const CaptureNode& cn : c.nodes)
        if (cn.n == 0 || dep_conflict_ranges(cn.r, cn.n, ranges, n)) deps.push_back(cn.node);
    c.pending_n = n;
    for (int i = 0; i < n; ++i) c.pending[i] = ranges[i];
    const bool only_prev = deps.size() == 1 && !c.nodes.empty() && deps[0] == c.nodes.back().node;
    if (only_prev) return pdl_enabled() ? kPdlWait : kFull;   // the stream's own dependency, with its early-launch edge
    if (cudaStreamUpdateCaptureDependencies(ctx->compute, deps.empty() ? nullptr : deps.data(), deps.size(), cudaStreamSetCaptureDependencies) != cudaSuccess) {
        cudaGetLastError();       // keep the stream's order instead
        return kFull;
    }
    return kFull;
}

LaunchMode plan_launch(dip_ctx* ctx, const DepRange* ranges, int n) {
    if (ctx->capturing) return plan_capture(ctx, ranges, n);
    DepTracker& d = ctx->dep;
    static const DepRange everything{0, ~(uintptr_t)0, true};
    if (n <= 0) { ranges = &everything; n = 1; }
    const bool touched = d.touched;
    d.touched = false;
    const bool may_skip = pdl_enabled() && !touched && !d.guaranteed.empty();
    constexpr size_t kCap = 24;   // a full launch now and then keeps the sets (and the search) small
    if (d.guaranteed.size() + d.unsafe.size() + (size_t)n <= kCap && !dep_conflict(d.unsafe, ranges, n)) {
        if (may_skip && !dep_conflict(d.guaranteed, ranges, n)) {
            // independent of everything that may still run: no wait.  Waiting on THIS kernel later proves nothing about
            // the ones before it, so they are no longer implied by "the last kernel has finished".
            d.unsafe.insert(d.unsafe.end(), d.guaranteed.begin(), d.guaranteed.end());
            d.guaranteed.assign(ranges, ranges + n);
            return kNoWait;
        }
        // depends (at most) on kernels whose completion the previous kernel's completion implies: the usual
        // programmatic wait; this kernel's completion then implies theirs too
        d.guaranteed.insert(d.guaranteed.end(), ranges, ranges + n);
        return kPdlWait;
    }
    // depends on a kernel that a wait on the previous one would not cover (or the sets are full): an ordinary launch,
    // which starts after everything before it in the stream has finished
    d.unsafe.clear();
    d.guaranteed.assign(ranges, ranges + n);
    return kFull;
}

static size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static int new_image(dip_ctx* ctx, dip_image** out) {
    dip_image* im = new (std::nothrow) dip_image();
    if (!im) return fail(ctx, DIP_ERR_NOMEM, "out of host memory");
    im->ctx = ctx;
    im->device = ctx ? ctx->device : -1;
    *out = im;  // copy_done is created by the first transfer that needs it (views never do: copies are tracked on the root)
    return DIP_OK;
}

static int ensure_copy_event(dip_ctx* ctx, dip_image* im) {
    if (!im->copy_done) DIP_CUDA(ctx, cudaEventCreateWithFlags(&im->copy_done, cudaEventDisableTiming));
    return DIP_OK;
}

// ---- TMA descriptors ------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                    const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                    const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled encode_tiled_fn() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        // resolved through the runtime so the library has no link-time libcuda dependency
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

int get_tmap(dip_ctx* ctx, dip_image* img, int box_words, int box_rows, const CUtensorMap** out, bool swizzle128) {
    const ImgView& v = img->v;
    dip_image* root = root_of(img);
    for (const TmapKey& k : root->tmaps)
        if (k.base == v.base && k.pitch == v.pitch && k.frame_stride == v.frame_stride && k.rowbytes == v.rowbytes &&
            k.height == v.height && k.frames == v.frames && k.box_words == box_words && k.box_rows == box_rows && k.swizzle128 == swizzle128) {
            *out = &k.map;
            return DIP_OK;
        }
    PFN_encodeTiled enc = encode_tiled_fn();
    if (!enc) return fail(ctx, DIP_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    if ((reinterpret_cast<uintptr_t>(v.base) & 15) || (v.pitch & 15) || (v.frame_stride & 15))
        return fail(ctx, DIP_ERR_ARG, "image base/pitch/frame stride must be 16-byte aligned for TMA");
    TmapKey key{};
    key.base = v.base; key.pitch = v.pitch; key.frame_stride = v.frame_stride;
    key.rowbytes = v.rowbytes; key.height = v.height; key.frames = v.frames;
    key.box_words = box_words;
    key.box_rows = box_rows;
    key.swizzle128 = swizzle128;
    // (x in 4-byte words, y, frame): u32 elements lift the 256-element box limit to 1 KiB rows.  The x extent is
    // the row's own (16-byte padded) length, not the pitch: a wrapped region of interest must never be read past
    // its right edge (TMA zero-fills there and the border patch supplies the halo).
    const size_t row_padded = std::min<size_t>(((size_t)v.rowbytes + 15) & ~(size_t)15, v.pitch);
    cuuint64_t dims[3] = {(cuuint64_t)(row_padded / 4), (cuuint64_t)v.height, (cuuint64_t)v.frames};
    cuuint64_t strides[2] = {(cuuint64_t)v.pitch, (cuuint64_t)(v.frames > 1 ? v.frame_stride : v.pitch * (size_t)v.height)};
    if (strides[1] == 0) strides[1] = 16;
    cuuint32_t box[3] = {(cuuint32_t)box_words, (cuuint32_t)box_rows, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(&key.map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, v.base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS)
        return fail(ctx, DIP_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for %dx%d pitch %zu box %dx%d",
                    (int)r, v.width, v.height, v.pitch, box_words, box_rows);
    if (root->tmaps.size() >= 256) root->tmaps.clear();  // a caller sweeping thousands of distinct views: start over
    root->tmaps.push_back(key);
    *out = &root->tmaps.back().map;  // valid until the next get_tmap on this allocation (launches copy it by value)
    return DIP_OK;
}

}  // namespace dip

using namespace dip;

extern "C" {

const char* dip_version(void) { return "dip_b200 0.1 (sm_100a)"; }

const char* dip_last_error_string(const dip_ctx* ctx) {
    if (ctx) return ctx->last_error.c_str();
    return g_last_error.c_str();
}

int dip_device_count(int* count) {
    if (!count) return fail(nullptr, DIP_ERR_ARG, "count is null");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail(nullptr, DIP_ERR_NO_DEVICE, "no CUDA device: %s (there is no CPU fallback)",
                    cudaGetErrorString(e));
    }
    *count = n;
    return DIP_OK;
}

int dip_init_on_stream(dip_ctx** out, int device, void* cuda_stream) {
    if (!out) return fail(nullptr, DIP_ERR_ARG, "ctx out pointer is null");
    *out = nullptr;
    int n = 0;
    int rc = dip_device_count(&n);
    if (rc != DIP_OK) return rc;
    if (n == 0) return fail(nullptr, DIP_ERR_NO_DEVICE, "no CUDA device (there is no CPU fallback)");
    if (device < 0 || device >= n) return fail(nullptr, DIP_ERR_ARG, "device %d out of range [0,%d)", device, n);
    dip_ctx* ctx = new (std::nothrow) dip_ctx();
    if (!ctx) return fail(nullptr, DIP_ERR_NOMEM, "out of host memory");
    ctx->device = device;
    cudaDeviceProp prop{};
#define INIT_CUDA(expr)                                                           \
    do {                                                                          \
        cudaError_t _e = (expr);                                                  \
        if (_e != cudaSuccess) {                                                  \
            int _rc = fail_cuda(nullptr, _e, #expr, __FILE__, __LINE__);          \
            delete ctx;                                                           \
            return _rc;                                                           \
        }                                                                         \
    } while (0)
    INIT_CUDA(cudaSetDevice(device));
    INIT_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10) {
        delete ctx;
        return fail(nullptr, DIP_ERR_NO_DEVICE,
                    "device %d is sm_%d%d; this library is built for sm_100a only (no fallback)", device,
                    prop.major, prop.minor);
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->l2_bytes = (size_t)prop.l2CacheSize;
    if (cuda_stream) {
        ctx->compute = reinterpret_cast<cudaStream_t>(cuda_stream);
        ctx->own_compute = false;
    } else {
        INIT_CUDA(cudaStreamCreateWithFlags(&ctx->compute, cudaStreamNonBlocking));
    }
    INIT_CUDA(cudaStreamCreateWithFlags(&ctx->upload, cudaStreamNonBlocking));
    INIT_CUDA(cudaStreamCreateWithFlags(&ctx->download, cudaStreamNonBlocking));
    INIT_CUDA(cudaStreamCreateWithFlags(&ctx->halo, cudaStreamNonBlocking));
    INIT_CUDA(cudaEventCreateWithFlags(&ctx->ev_compute, cudaEventDisableTiming));
    INIT_CUDA(cudaEventCreateWithFlags(&ctx->ev_halo, cudaEventDisableTiming));
    INIT_CUDA(cudaEventCreate(&ctx->t0));
    INIT_CUDA(cudaEventCreate(&ctx->t1));
    ctx->stage_bytes = (size_t)8 << 20;
    for (int i = 0; i < dip_ctx::kSlots; ++i) {
        INIT_CUDA(cudaMallocHost(&ctx->stage_up[i], ctx->stage_bytes));
        INIT_CUDA(cudaMallocHost(&ctx->stage_dn[i], ctx->stage_bytes));
        INIT_CUDA(cudaEventCreateWithFlags(&ctx->stage_up_free[i], cudaEventDisableTiming));
        INIT_CUDA(cudaEventCreateWithFlags(&ctx->stage_dn_done[i], cudaEventDisableTiming));
    }
#undef INIT_CUDA
    *out = ctx;
    return DIP_OK;
}

int dip_init(dip_ctx** out, int device) { return dip_init_on_stream(out, device, nullptr); }

int dip_shutdown(dip_ctx* ctx) {
    if (!ctx) return DIP_OK;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->compute);
    cudaStreamSynchronize(ctx->upload);
    cudaStreamSynchronize(ctx->download);
    cudaStreamSynchronize(ctx->halo);
    for (dip_image*& sc : ctx->scratch) { dip_image_destroy(sc); sc = nullptr; }
    jpeg_release(ctx);
    for (int i = 0; i < dip_ctx::kSlots; ++i) {
        if (ctx->stage_up[i]) cudaFreeHost(ctx->stage_up[i]);
        if (ctx->stage_dn[i]) cudaFreeHost(ctx->stage_dn[i]);
        if (ctx->stage_up_free[i]) cudaEventDestroy(ctx->stage_up_free[i]);
        if (ctx->stage_dn_done[i]) cudaEventDestroy(ctx->stage_dn_done[i]);
    }
    for (cudaEvent_t e : ctx->marks) cudaEventDestroy(e);
    if (ctx->flush_buf) cudaFree(ctx->flush_buf);
    if (ctx->d_checksums) cudaFree(ctx->d_checksums);
    cudaEventDestroy(ctx->ev_compute);
    cudaEventDestroy(ctx->ev_halo);
    cudaEventDestroy(ctx->t0);
    cudaEventDestroy(ctx->t1);
    if (ctx->own_compute) cudaStreamDestroy(ctx->compute);
    cudaStreamDestroy(ctx->upload);
    cudaStreamDestroy(ctx->download);
    cudaStreamDestroy(ctx->halo);
    delete ctx;
    return DIP_OK;
}

int dip_sync(dip_ctx* ctx) {
    if (!ctx) return fail(nullptr, DIP_ERR_ARG, "ctx is null");
    if (ctx->capturing) return fail(ctx, DIP_ERR_ARG, "dip_sync cannot be recorded into a graph");
    DIP_CUDA(ctx, cudaSetDevice(ctx->device));
    DIP_CUDA(ctx, cudaStreamSynchronize(ctx->upload));
    DIP_CUDA(ctx, cudaStreamSynchronize(ctx->halo));
    DIP_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    DIP_CUDA(ctx, cudaStreamSynchronize(ctx->download));
    ctx->dep.unsafe.clear();
    ctx->dep.guaranteed.clear();
    ctx->dep.touched = true;
    return DIP_OK;
}

int dip_get_device_info(dip_ctx* ctx, dip_device_info* info) {
    if (!ctx || !info) return fail(ctx, DIP_ERR_ARG, "null argument");
    cudaDeviceProp prop{};
    DIP_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
    memset(info, 0, sizeof *info);
    snprintf(info->name, sizeof info->name, "%.127s", prop.name);
    info->device = ctx->device;
    info->sm_count = prop.multiProcessorCount;
    info->cc_major = prop.major;
    info->cc_minor = prop.minor;
    info->total_mem = prop.totalGlobalMem;
    info->l2_bytes = (size_t)prop.l2CacheSize;
    return DIP_OK;
}

// ---- images -----------------------------------------------------------------------------
static int create_stack(dip_ctx* ctx, int w, int h, int c, int frames, int bps, dip_image** out) {
    if (!ctx || !out) return fail(ctx, DIP_ERR_ARG, "null argument");
    *out = nullptr;
    if (w < 1 || h < 1 || frames < 1) return fail(ctx, DIP_ERR_ARG, "bad geometry %dx%d x%d frames", w, h, frames);
    if (c != 1 && c != 3 && c != 4) return fail(ctx, DIP_ERR_ARG, "channels must be 1, 3 or 4 (got %d)", c);
    if ((long long)w * c * bps > (1ll << 30)) return fail(ctx, DIP_ERR_ARG, "row too long");
    dip_image* im = nullptr;
    int rc = new_image(ctx, &im);
    if (rc != DIP_OK) return rc;
    ImgView& v = im->v;
    v.width = w; v.height = h; v.channels = c; v.frames = frames;
    v.bps = bps;
    v.rowbytes = w * c * bps;
    v.pitch = round_up((size_t)v.rowbytes, 16);
    v.frame_stride = v.pitch * (size_t)h;
    v.pad_writable = true;
    const size_t bytes = v.frame_stride * (size_t)frames;
    cudaSetDevice(ctx->device);
    cudaError_t e = cudaMalloc(&im->alloc, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        dip_image_destroy(im);
        return fail(ctx, DIP_ERR_NOMEM, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    im->owned = true;
    v.base = static_cast<uint8_t*>(im->alloc);
    *out = im;
    return DIP_OK;
}

int dip_image_create_batch(dip_ctx* ctx, int w, int h, int c, int frames, dip_image** out) { return create_stack(ctx, w, h, c, frames, 1, out); }

int dip_image_create16(dip_ctx* ctx, int w, int h, int c, int frames, dip_image** out) { return create_stack(ctx, w, h, c, frames, 2, out); }

int dip_image_depth(const dip_image* im, int* bits) {
    if (!im || !bits) return fail(nullptr, DIP_ERR_ARG, "null argument");
    *bits = 8 * im->v.bps;
    return DIP_OK;
}

int dip_image_create(dip_ctx* ctx, int w, int h, int c, dip_image** out) {
    return dip_image_create_batch(ctx, w, h, c, 1, out);
}

int dip_image_wrap(dip_ctx* ctx, void* p, int w, int h, int c, int frames, size_t pitch,
                   size_t frame_stride, dip_image** out) {
    if (!ctx || !out || !p) return fail(ctx, DIP_ERR_ARG, "null argument");
    *out = nullptr;
    if (w < 1 || h < 1 || frames < 1 || (c != 1 && c != 3 && c != 4))
        return fail(ctx, DIP_ERR_ARG, "bad geometry %dx%dx%d x%d", w, h, c, frames);
    if (pitch == 0) pitch = round_up((size_t)w * c, 16);
    if (frame_stride == 0) frame_stride = pitch * (size_t)h;
    if (pitch < (size_t)w * c || (pitch & 15) || (frame_stride & 15) || (reinterpret_cast<uintptr_t>(p) & 15))
        return fail(ctx, DIP_ERR_ARG, "wrapped memory needs a 16-byte aligned base, pitch and frame stride");
    if (frames > 1 && frame_stride < pitch * (size_t)h) return fail(ctx, DIP_ERR_ARG, "frame stride too small");
    dip_image* im = nullptr;
    int rc = new_image(ctx, &im);
    if (rc != DIP_OK) return rc;
    im->v = ImgView{static_cast<uint8_t*>(p), pitch, frame_stride, w, h, c, frames, w * c, 1, false};
    *out = im;
    return DIP_OK;
}

int dip_image_view(dip_ctx* ctx, dip_image* parent, int first, int count, dip_image** out) {
    if (!ctx || !parent || !out) return fail(ctx, DIP_ERR_ARG, "null argument");
    *out = nullptr;
    if (first < 0 || count < 1 || first + count > parent->v.frames)
        return fail(ctx, DIP_ERR_ARG, "frames [%d,%d) outside stack of %d", first, first + count, parent->v.frames);
    dip_image* im = nullptr;
    int rc = new_image(ctx, &im);
    if (rc != DIP_OK) return rc;
    im->v = parent->v;
    im->v.base = parent->v.base + parent->v.frame_stride * (size_t)first;
    im->v.frames = count;
    im->parent = parent;
    *out = im;
    return DIP_OK;
}

int dip_image_rows(dip_ctx* ctx, dip_image* parent, int row0, int rows, dip_image** out) {
    if (!ctx || !parent || !out) return fail(ctx, DIP_ERR_ARG, "null argument");
    *out = nullptr;
    if (row0 < 0 || rows < 1 || row0 + rows > parent->v.height)
        return fail(ctx, DIP_ERR_ARG, "rows [%d,%d) outside image of %d rows", row0, row0 + rows, parent->v.height);
    dip_image* im = nullptr;
    int rc = new_image(ctx, &im);
    if (rc != DIP_OK) return rc;
    im->v = parent->v;
    im->v.base = parent->v.base + parent->v.pitch * (size_t)row0;
    im->v.height = rows;
    im->parent = parent;
    *out = im;
    return DIP_OK;
}

int dip_image_destroy(dip_image* im) {
    if (!im) return DIP_OK;
    if (im->device >= 0) cudaSetDevice(im->device);
    if (im->copy_pending) cudaEventSynchronize(im->copy_done);
    if (im->copy_done) cudaEventDestroy(im->copy_done);
    if (im->ipc_mapped && im->alloc) cudaIpcCloseMemHandle(im->alloc);
    else if (im->owned && im->alloc) cudaFree(im->alloc);
    delete im;
    return DIP_OK;
}

int dip_image_get_info(const dip_image* im, dip_image_info* info) {
    if (!im || !info) return fail(nullptr, DIP_ERR_ARG, "null argument");
    info->width = im->v.width; info->height = im->v.height; info->channels = im->v.channels;
    info->frames = im->v.frames; info->pitch = im->v.pitch; info->frame_stride = im->v.frame_stride;
    info->device_ptr = im->v.base; info->device = im->device;
    return DIP_OK;
}

// ---- pinned host memory -----------------------------------------------------------------
int dip_host_alloc(void** p, size_t bytes) {
    if (!p) return fail(nullptr, DIP_ERR_ARG, "null argument");
    cudaError_t e = cudaMallocHost(p, bytes ? bytes : 1);
    if (e != cudaSuccess) { cudaGetLastError(); return fail(nullptr, DIP_ERR_NOMEM, "cudaMallocHost(%zu): %s", bytes, cudaGetErrorString(e)); }
    return DIP_OK;
}
int dip_host_free(void* p) {
    if (!p) return DIP_OK;
    cudaError_t e = cudaFreeHost(p);
    if (e != cudaSuccess) return fail_cuda(nullptr, e, "cudaFreeHost", __FILE__, __LINE__);
    return DIP_OK;
}
int dip_host_register(void* p, size_t bytes) {
    if (!p || !bytes) return fail(nullptr, DIP_ERR_ARG, "null argument");
    cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterDefault);
    if (e != cudaSuccess) { cudaGetLastError(); return fail_cuda(nullptr, e, "cudaHostRegister", __FILE__, __LINE__); }
    return DIP_OK;
}
int dip_host_unregister(void* p) {
    if (!p) return DIP_OK;
    cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) { cudaGetLastError(); return fail_cuda(nullptr, e, "cudaHostUnregister", __FILE__, __LINE__); }
    return DIP_OK;
}

// Host-side staging copy (pageable <-> pinned) split over a few threads: one core's memcpy
// (~12 GB/s) would otherwise cap pageable transfers far below the PCIe rate.  The workers are a process-wide pool
// created on first use (a transfer is a train of 8 MiB slices; creating threads per slice cost as much as a third of the
// copy); they spin for the next slice for 0.5 ms (waking sleepers one by one is what made single-frame calls slow), then
// sleep.  DIP_COPY_THREADS / DIP_COPY_SPIN_US override the pool width (default min(12, 3/4 of the cores)) and the spin time.
static void host_copy_rows(uint8_t* dst, size_t dpitch, const uint8_t* src, size_t spitch, size_t rowbytes, int rows) {
    const size_t total = rowbytes * (size_t)rows;
    if (total < ((size_t)1 << 20)) {  // small: not worth waking anybody
        if (dpitch == rowbytes && spitch == rowbytes) memcpy(dst, src, total);
        else for (int r = 0; r < rows; ++r) memcpy(dst + dpitch * r, src + spitch * r, rowbytes);
        return;
    }
    dip::CopyPool& pool = dip::CopyPool::get();
    const unsigned nthreads = (unsigned)std::min<size_t>(pool.width(), total >> 19);  // at least 512 KiB per thread
    pool.run(nthreads, [=](unsigned t) {
        if (dpitch == rowbytes && spitch == rowbytes) {
            const size_t a = (total * t / nthreads) & ~(size_t)63, b = t + 1 == nthreads ? total : (total * (t + 1) / nthreads) & ~(size_t)63;
            memcpy(dst + a, src + a, b - a);
        } else {
            const int r0 = (int)((size_t)rows * t / nthreads), r1 = (int)((size_t)rows * (t + 1) / nthreads);
            for (int r = r0; r < r1; ++r) memcpy(dst + dpitch * r, src + spitch * r, rowbytes);
        }
    });
}

static bool is_pinned(const void* p) {
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return a.type == cudaMemoryTypeHost;
}

// ---- upload / download --------------------------------------------------------------------
// One 2-D copy per frame unless everything is contiguous, in which case a single flat copy.
static int copy_2d(dip_ctx* ctx, cudaStream_t s, void* dst, size_t dpitch, size_t dframe, const void* src,
                   size_t spitch, size_t sframe, size_t rowbytes, int rows, int frames, cudaMemcpyKind kind) {
    const bool flat = dpitch == rowbytes && spitch == rowbytes && (frames == 1 || (dframe == rowbytes * (size_t)rows && sframe == dframe));
    if (flat) {
        DIP_CUDA(ctx, cudaMemcpyAsync(dst, src, rowbytes * (size_t)rows * (size_t)frames, kind, s));
        return DIP_OK;
    }
    for (int f = 0; f < frames; ++f)
        DIP_CUDA(ctx, cudaMemcpy2DAsync(static_cast<uint8_t*>(dst) + dframe * (size_t)f, dpitch,
                                        static_cast<const uint8_t*>(src) + sframe * (size_t)f, spitch, rowbytes,
                                        (size_t)rows, kind, s));
    return DIP_OK;
}

}  // extern "C"

namespace dip {

int enqueue_upload(dip_ctx* ctx, const ImgView& v, const uint8_t* host, size_t host_pitch) {
    const size_t rowb = (size_t)v.rowbytes;
    if (is_pinned(host))
        return copy_2d(ctx, ctx->upload, v.base, v.pitch, v.frame_stride, host, host_pitch, host_pitch * (size_t)v.height, rowb,
                       v.height, v.frames, cudaMemcpyHostToDevice);
    // pageable: stage slices through two pinned slots, copy of slice k overlaps staging of k+1.  A slice is a run of
    // whole rows, or -- when one row is longer than a staging slot -- a run of bytes within one row.
    const size_t cols_per_slice = std::min(rowb, ctx->stage_bytes);
    const int rows_per_slice = (int)std::max<size_t>(1, ctx->stage_bytes / rowb);
    int slot = 0;
    for (int f = 0; f < v.frames; ++f) {
        for (int y = 0; y < v.height; y += rows_per_slice) {
            const int rows = std::min(rows_per_slice, v.height - y);
            for (size_t x = 0; x < rowb; x += cols_per_slice) {
                const size_t nb = std::min(cols_per_slice, rowb - x);   // == rowb unless a row exceeds the slot
                DIP_CUDA(ctx, cudaEventSynchronize(ctx->stage_up_free[slot]));
                uint8_t* st = ctx->stage_up[slot];
                const uint8_t* hsrc = host + host_pitch * ((size_t)f * v.height + y) + x;
                host_copy_rows(st, nb, hsrc, host_pitch, nb, rows);
                DIP_CUDA(ctx, cudaMemcpy2DAsync(v.base + v.frame_stride * (size_t)f + v.pitch * (size_t)y + x, v.pitch, st, nb, nb,
                                                rows, cudaMemcpyHostToDevice, ctx->upload));
                DIP_CUDA(ctx, cudaEventRecord(ctx->stage_up_free[slot], ctx->upload));
                slot ^= 1;
            }
        }
    }
    return DIP_OK;
}

int enqueue_download(dip_ctx* ctx, const ImgView& v, uint8_t* host, size_t host_pitch, bool* completed) {
    const size_t rowb = (size_t)v.rowbytes;
    if (is_pinned(host)) {
        *completed = false;
        return copy_2d(ctx, ctx->download, host, host_pitch, host_pitch * (size_t)v.height, v.base, v.pitch, v.frame_stride, rowb,
                       v.height, v.frames, cudaMemcpyDeviceToHost);
    }
    // pageable destination: D2H into pinned slots, drain slot k while slot k+1 is in flight.
    // The host memcpy needs the data, so this path completes before returning.
    *completed = true;
    const size_t cols_per_slice = std::min(rowb, ctx->stage_bytes);   // a row longer than a slot is cut into byte runs
    const int rows_per_slice = (int)std::max<size_t>(1, ctx->stage_bytes / rowb);
    struct Pending { uint8_t* hdst; size_t nb; int rows; bool live; } pend[dip_ctx::kSlots] = {};
    int slot = 0;
    auto drain = [&](int s) -> int {
        if (!pend[s].live) return DIP_OK;
        DIP_CUDA(ctx, cudaEventSynchronize(ctx->stage_dn_done[s]));
        host_copy_rows(pend[s].hdst, host_pitch, ctx->stage_dn[s], pend[s].nb, pend[s].nb, pend[s].rows);
        pend[s].live = false;
        return DIP_OK;
    };
    for (int f = 0; f < v.frames; ++f) {
        for (int y = 0; y < v.height; y += rows_per_slice) {
            const int rows = std::min(rows_per_slice, v.height - y);
            for (size_t x = 0; x < rowb; x += cols_per_slice) {
                const size_t nb = std::min(cols_per_slice, rowb - x);
                int rc = drain(slot);
                if (rc != DIP_OK) return rc;
                DIP_CUDA(ctx, cudaMemcpy2DAsync(ctx->stage_dn[slot], nb, v.base + v.frame_stride * (size_t)f + v.pitch * (size_t)y + x,
                                                v.pitch, nb, rows, cudaMemcpyDeviceToHost, ctx->download));
                DIP_CUDA(ctx, cudaEventRecord(ctx->stage_dn_done[slot], ctx->download));
                pend[slot] = Pending{host + host_pitch * ((size_t)f * v.height + y) + x, nb, rows, true};
                slot ^= 1;
            }
        }
    }
    for (int s = 0; s < dip_ctx::kSlots; ++s) {
        int rc = drain(s);
        if (rc != DIP_OK) return rc;
    }
    return DIP_OK;
}

}  // namespace dip

extern "C" {

int dip_upload(dip_ctx* ctx, dip_image* dst_view, const uint8_t* host, size_t host_pitch) {
    if (!ctx || !dst_view || !host) return fail(ctx, DIP_ERR_ARG, "null argument");
    if (ctx->capturing) return fail(ctx, DIP_ERR_UNSUPPORTED, "transfers cannot be recorded into a graph");
    DIP_CUDA(ctx, cudaSetDevice(ctx->device));
    const ImgView& v = dst_view->v;
    dip_image* dst = root_of(dst_view);  // copy ordering is tracked on the owning allocation
    const size_t rowb = (size_t)v.rowbytes;
    if (host_pitch == 0) host_pitch = rowb;
    if (host_pitch < rowb) return fail(ctx, DIP_ERR_ARG, "host pitch %zu < row bytes %zu", host_pitch, rowb);
    // the copy must not overtake kernels still reading/writing dst
    dep_touch(ctx);
    DIP_CUDA(ctx, cudaEventRecord(ctx->ev_compute, ctx->compute));
    DIP_CUDA(ctx, cudaStreamWaitEvent(ctx->upload, ctx->ev_compute, 0));
    if (dst->copy_pending) DIP_CUDA(ctx, cudaStreamWaitEvent(ctx->upload, dst->copy_done, 0));
    int rc = enqueue_upload(ctx, v, host, host_pitch);
    if (rc != DIP_OK) return rc;
    if ((rc = ensure_copy_event(ctx, dst)) != DIP_OK) return rc;
    DIP_CUDA(ctx, cudaEventRecord(dst->copy_done, ctx->upload));
    dst->copy_pending = true;
    return DIP_OK;
}

int dip_download(dip_ctx* ctx, const dip_image* src_c, uint8_t* host, size_t host_pitch) {
    if (!ctx || !src_c || !host) return fail(ctx, DIP_ERR_ARG, "null argument");
    if (ctx->capturing) return fail(ctx, DIP_ERR_UNSUPPORTED, "transfers cannot be recorded into a graph");
    DIP_CUDA(ctx, cudaSetDevice(ctx->device));
    const ImgView& v = src_c->v;
    dip_image* src = root_of(const_cast<dip_image*>(src_c));
    const size_t rowb = (size_t)v.rowbytes;
    if (host_pitch == 0) host_pitch = rowb;
    if (host_pitch < rowb) return fail(ctx, DIP_ERR_ARG, "host pitch %zu < row bytes %zu", host_pitch, rowb);
    dep_touch(ctx);
    DIP_CUDA(ctx, cudaEventRecord(ctx->ev_compute, ctx->compute));
    DIP_CUDA(ctx, cudaStreamWaitEvent(ctx->download, ctx->ev_compute, 0));
    if (src->copy_pending) DIP_CUDA(ctx, cudaStreamWaitEvent(ctx->download, src->copy_done, 0));
    bool completed = false;
    int rc = enqueue_download(ctx, v, host, host_pitch, &completed);
    if (rc != DIP_OK) return rc;
    if (!completed) {
        if ((rc = ensure_copy_event(ctx, src)) != DIP_OK) return rc;
        DIP_CUDA(ctx, cudaEventRecord(src->copy_done, ctx->download));
        src->copy_pending = true;
    }
    return DIP_OK;
}

// ---- CUDA graphs: record a sequence of operations once, replay it with one launch ---------------------------
// The reference's harness enqueues the same few-microsecond kernel thousands of times (measure_time, opencv:33-35 /
// sycl:33-37); replayed from a graph the host cost of a call drops from ~4-11 us to the graph launch divided by the
// operations in it.  The nodes' edges are the data dependencies between the recorded kernels (plan_capture above), not
// the order of the calls.
int dip_graph_begin(dip_ctx* ctx) {
    if (!ctx) return fail(nullptr, DIP_ERR_ARG, "ctx is null");
    if (ctx->capturing) return fail(ctx, DIP_ERR_ARG, "a graph is already being recorded on this ctx");
    int rc = dip_sync(ctx);  // outstanding transfers finish now: a recording cannot wait for work outside itself
    if (rc != DIP_OK) return rc;
    dep_touch(ctx);
    DIP_CUDA(ctx, cudaStreamBeginCapture(ctx->compute, cudaStreamCaptureModeThreadLocal));
    ctx->cap = CaptureTracker{};
    ctx->capturing = true;
    ctx->capture_launches0 = ctx->launches;
    return DIP_OK;
}

int dip_graph_end(dip_ctx* ctx, dip_graph** out) {
    if (!ctx || !out) return fail(ctx, DIP_ERR_ARG, "null argument");
    *out = nullptr;
    if (!ctx->capturing) return fail(ctx, DIP_ERR_ARG, "dip_graph_begin was not called");
    capture_join(ctx);   // the end of the recording depends on every branch
    ctx->capturing = false;
    ctx->cap = CaptureTracker{};
    dep_touch(ctx);
    ctx->dep.unsafe.clear();
    ctx->dep.guaranteed.clear();   // what was recorded did not run
    cudaGraph_t graph = nullptr;
    cudaError_t e = cudaStreamEndCapture(ctx->compute, &graph);
    if (e != cudaSuccess || !graph) { cudaGetLastError(); return fail(ctx, DIP_ERR_CUDA, "recording failed: %s (transfers, dip_sync and first-time allocations cannot be recorded)", cudaGetErrorString(e)); }
    dip_graph* g = new (std::nothrow) dip_graph();
    if (!g) { cudaGraphDestroy(graph); return fail(ctx, DIP_ERR_NOMEM, "out of host memory"); }
    g->ctx = ctx;
    g->graph = graph;
    g->kernels = ctx->launches - ctx->capture_launches0;
    ctx->launches = ctx->capture_launches0;  // recorded, not run
    e = cudaGraphInstantiate(&g->exec, graph, 0);
    if (e != cudaSuccess) { cudaGetLastError(); cudaGraphDestroy(graph); delete g; return fail_cuda(ctx, e, "cudaGraphInstantiate", __FILE__, __LINE__); }
    *out = g;
    return DIP_OK;
}

int dip_graph_launch(dip_ctx* ctx, dip_graph* g) {
    if (!ctx || !g || g->ctx != ctx) return fail(ctx, DIP_ERR_ARG, "bad graph");
    if (ctx->capturing) return fail(ctx, DIP_ERR_ARG, "cannot replay a graph while recording");
    DIP_CUDA(ctx, cudaSetDevice(ctx->device));
    dep_touch(ctx);
    DIP_CUDA(ctx, cudaGraphLaunch(g->exec, ctx->compute));
    ctx->launches += g->kernels;
    return DIP_OK;
}

int dip_graph_destroy(dip_graph* g) {
    if (!g) return DIP_OK;
    if (g->exec) cudaGraphExecDestroy(g->exec);
    if (g->graph) cudaGraphDestroy(g->graph);
    delete g;
    return DIP_OK;
}

// ---- timing, counters, L2 flush ---------------------------------------------------------
int dip_timer_start(dip_ctx* ctx) {
    if (!ctx) return fail(nullptr, DIP_ERR_ARG, "ctx is null");
    dep_touch(ctx);
    DIP_CUDA(ctx, cudaEventRecord(ctx->t0, ctx->compute));
    return DIP_OK;
}

int dip_timer_stop(dip_ctx* ctx, float* ms) {
    if (!ctx || !ms) return fail(ctx, DIP_ERR_ARG, "null argument");
    dep_touch(ctx);
    DIP_CUDA(ctx, cudaEventRecord(ctx->t1, ctx->compute));
    DIP_CUDA(ctx, cudaEventSynchronize(ctx->t1));
    DIP_CUDA(ctx, cudaEventElapsedTime(ms, ctx->t0, ctx->t1));
    return DIP_OK;
}

int dip_mark(dip_ctx* ctx, int slot) {
    if (!ctx || slot < 0 || slot > (1 << 20)) return fail(ctx, DIP_ERR_ARG, "bad marker slot");
    while ((int)ctx->marks.size() <= slot) {
        cudaEvent_t e = nullptr;
        DIP_CUDA(ctx, cudaEventCreate(&e));
        ctx->marks.push_back(e);
    }
    dep_touch(ctx);
    DIP_CUDA(ctx, cudaEventRecord(ctx->marks[slot], ctx->compute));
    return DIP_OK;
}

int dip_elapsed(dip_ctx* ctx, int a, int b, float* ms) {
    if (!ctx || !ms || a < 0 || b < 0 || a >= (int)ctx->marks.size() || b >= (int)ctx->marks.size())
        return fail(ctx, DIP_ERR_ARG, "marker slot not recorded");
    DIP_CUDA(ctx, cudaEventSynchronize(ctx->marks[b]));
    DIP_CUDA(ctx, cudaEventElapsedTime(ms, ctx->marks[a], ctx->marks[b]));
    return DIP_OK;
}

int dip_launch_count(const dip_ctx* ctx, uint64_t* n) {
    if (!ctx || !n) return fail(ctx, DIP_ERR_ARG, "null argument");
    *n = ctx->launches;
    return DIP_OK;
}

int dip_flush_l2(dip_ctx* ctx) {
    if (!ctx) return fail(nullptr, DIP_ERR_ARG, "ctx is null");
    DIP_CUDA(ctx, cudaSetDevice(ctx->device));
    if (!ctx->flush_buf) {
        ctx->flush_bytes = std::max<size_t>(ctx->l2_bytes * 2, (size_t)256 << 20);
        DIP_CUDA(ctx, cudaMalloc(&ctx->flush_buf, ctx->flush_bytes));
    }
    return launch_fill(ctx, ctx->flush_buf, ctx->flush_bytes, 0);
}

int dip_checksum_rows(dip_ctx* ctx, const dip_image* img, uint64_t first_row, uint64_t* per_frame) {
    if (!ctx || !img || !per_frame) return fail(ctx, DIP_ERR_ARG, "null argument");
    int rc = begin_op(ctx, img, nullptr);
    if (rc != DIP_OK) return rc;
    const int n = img->v.frames;
    if (ctx->d_checksums_cap < n) {
        if (ctx->d_checksums) cudaFree(ctx->d_checksums);
        ctx->d_checksums = nullptr;
        DIP_CUDA(ctx, cudaMalloc(&ctx->d_checksums, sizeof(uint64_t) * (size_t)n));
        ctx->d_checksums_cap = n;
    }
    dep_touch(ctx);
    DIP_CUDA(ctx, cudaMemsetAsync(ctx->d_checksums, 0, sizeof(uint64_t) * (size_t)n, ctx->compute));
    rc = launch_checksum(ctx, img->v, first_row, ctx->d_checksums);
    if (rc != DIP_OK) return rc;
    dep_touch(ctx);
    DIP_CUDA(ctx, cudaMemcpyAsync(per_frame, ctx->d_checksums, sizeof(uint64_t) * (size_t)n, cudaMemcpyDeviceToHost, ctx->compute));
    DIP_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    return DIP_OK;
}

int dip_checksum(dip_ctx* ctx, const dip_image* img, uint64_t* per_frame) { return dip_checksum_rows(ctx, img, 0, per_frame); }

int dip_synth_fill_rows(dip_ctx* ctx, dip_image* img, uint32_t seed, uint64_t first_row) {
    if (!ctx || !img) return fail(ctx, DIP_ERR_ARG, "null argument");
    int rc = begin_op(ctx, nullptr, img);
    if (rc != DIP_OK) return rc;
    return launch_synth(ctx, img->v, seed, first_row);
}

int dip_synth_fill(dip_ctx* ctx, dip_image* img, uint32_t seed, uint64_t first_frame) {
    if (!img) return fail(ctx, DIP_ERR_ARG, "null argument");
    return dip_synth_fill_rows(ctx, img, seed, first_frame * (uint64_t)img->v.height);
}

// ---- IPC + row copies for mosaic bands ---------------------------------------------------
int dip_image_ipc_export(dip_ctx* ctx, const dip_image* img, dip_ipc_handle* h) {
    if (!ctx || !img || !h) return fail(ctx, DIP_ERR_ARG, "null argument");
    if (!img->owned || img->v.base != img->alloc) return fail(ctx, DIP_ERR_ARG, "only whole library-owned images can be exported");
    static_assert(sizeof(cudaIpcMemHandle_t) <= sizeof(dip_ipc_handle), "handle size");
    cudaIpcMemHandle_t mh;
    DIP_CUDA(ctx, cudaSetDevice(ctx->device));
    DIP_CUDA(ctx, cudaIpcGetMemHandle(&mh, img->alloc));
    memset(h, 0, sizeof *h);
    memcpy(h->bytes, &mh, sizeof mh);
    return DIP_OK;
}

int dip_image_ipc_import(dip_ctx* ctx, const dip_ipc_handle* h, int w, int hgt, int c, int frames, size_t pitch,
                         size_t frame_stride, dip_image** out) {
    if (!ctx || !h || !out) return fail(ctx, DIP_ERR_ARG, "null argument");
    *out = nullptr;
    cudaIpcMemHandle_t mh;
    memcpy(&mh, h->bytes, sizeof mh);
    void* p = nullptr;
    DIP_CUDA(ctx, cudaSetDevice(ctx->device));
    DIP_CUDA(ctx, cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
    int rc = dip_image_wrap(ctx, p, w, hgt, c, frames, pitch, frame_stride, out);
    if (rc != DIP_OK) { cudaIpcCloseMemHandle(p); return rc; }
    (*out)->alloc = p;
    (*out)->ipc_mapped = true;
    (*out)->v.pad_writable = true;  // the exporter's whole library-owned allocation: its row padding is padding here too
    return DIP_OK;
}

int dip_copy_rows(dip_ctx* ctx, const dip_image* src, int src_row, dip_image* dst, int dst_row, int rows) {
    if (!ctx || !src || !dst) return fail(ctx, DIP_ERR_ARG, "null argument");
    const ImgView &a = src->v, &b = dst->v;
    if (a.rowbytes != b.rowbytes || a.frames != b.frames) return fail(ctx, DIP_ERR_SHAPE, "row length / frames differ");
    if (rows < 1 || src_row < 0 || dst_row < 0 || src_row + rows > a.height || dst_row + rows > b.height)
        return fail(ctx, DIP_ERR_ARG, "row range out of bounds");
    DIP_CUDA(ctx, cudaSetDevice(ctx->device));
    // ordered after everything enqueued on the compute stream so far
    dep_touch(ctx);
    DIP_CUDA(ctx, cudaEventRecord(ctx->ev_compute, ctx->compute));
    DIP_CUDA(ctx, cudaStreamWaitEvent(ctx->halo, ctx->ev_compute, 0));
    for (int f = 0; f < a.frames; ++f)
        DIP_CUDA(ctx, cudaMemcpy2DAsync(b.base + b.frame_stride * (size_t)f + b.pitch * (size_t)dst_row, b.pitch,
                                        a.base + a.frame_stride * (size_t)f + a.pitch * (size_t)src_row, a.pitch,
                                        (size_t)a.rowbytes, (size_t)rows, cudaMemcpyDefault, ctx->halo));
    return DIP_OK;
}

}  // extern "C"

// ---- cross-GPU step flags (mosaic bands whose input changes from step to step) ------------------------------------
// A flag word lives in a small library-owned image (dip_flags_create) that neighbours map through CUDA IPC like a band.
// dip_flag_signal is a one-thread kernel on the compute stream: everything enqueued before it has completed and been
// flushed when it runs, then it publishes `value` with a system-scope release store -- to local memory or to the
// neighbour's over NVLink.  dip_flag_wait is a stream memory operation (cuStreamWaitValue32, >=) on LOCAL memory: the
// stream stalls, no kernel spins, so ranks time-slicing one GPU cannot starve each other.
namespace {

__global__ void flag_signal_kernel(uint32_t* word, uint32_t value) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(word), "r"(value) : "memory");
}

typedef CUresult (*PFN_streamWaitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
PFN_streamWaitValue32 stream_wait_value_fn() {
    static PFN_streamWaitValue32 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_streamWaitValue32>(p);
    });
    return fn;
}

typedef CUresult (*PFN_streamWriteValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int);
PFN_streamWriteValue32 stream_write_value_fn() {
    static PFN_streamWriteValue32 fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* e = getenv("DIP_FLAG_SIGNAL_KERNEL");   // 1: always signal with the one-thread kernel (A/B)
        if (e && e[0] == '1') return;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuStreamWriteValue32", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_streamWriteValue32>(p);
    });
    return fn;
}

int flag_word(dip_ctx* ctx, const dip_image* flags, int index, uint32_t** word) {
    if (!ctx || !flags) return fail(ctx, DIP_ERR_ARG, "null argument");
    if (index < 0 || (size_t)(index + 1) * 4 > (size_t)flags->v.rowbytes) return fail(ctx, DIP_ERR_ARG, "flag index %d out of range", index);
    *word = reinterpret_cast<uint32_t*>(flags->v.base) + index;
    return DIP_OK;
}

}  // namespace

extern "C" {

int dip_flags_create(dip_ctx* ctx, int count, dip_image** flags) {
    if (!ctx || !flags || count < 1 || count > (1 << 20)) return fail(ctx, DIP_ERR_ARG, "bad argument");
    int rc = dip_image_create(ctx, 4 * count, 1, 1, flags);
    if (rc != DIP_OK) return rc;
    cudaError_t e = cudaMemset((*flags)->v.base, 0, (*flags)->v.pitch);
    if (e != cudaSuccess) { dip_image_destroy(*flags); *flags = nullptr; return fail_cuda(ctx, e, "cudaMemset(flags)", __FILE__, __LINE__); }
    return DIP_OK;
}

int dip_flag_signal(dip_ctx* ctx, dip_image* flags, int index, uint32_t value) {
    uint32_t* word = nullptr;
    int rc = flag_word(ctx, flags, index, &word);
    if (rc != DIP_OK) return rc;
    DIP_CUDA(ctx, cudaSetDevice(ctx->device));
    dep_touch(ctx);
    // A stream memory operation where the driver takes it for this address (the default flags put a system-scope fence in
    // front of the write, and as a stream operation it runs after every earlier kernel of the stream has finished): no
    // launch at all.  Otherwise -- some drivers refuse peer / IPC-mapped addresses -- a one-thread kernel with a release store.
    if (!ctx->capturing && ctx->flag_write_ok) {
        if (PFN_streamWriteValue32 fn = stream_write_value_fn()) {
            if (fn(reinterpret_cast<CUstream>(ctx->compute), reinterpret_cast<CUdeviceptr>(word), value, 0x0 /* CU_STREAM_WRITE_VALUE_DEFAULT */) == CUDA_SUCCESS)
                return DIP_OK;
            cudaGetLastError();
        }
        ctx->flag_write_ok = false;   // not again on this ctx
    }
    flag_signal_kernel<<<1, 1, 0, ctx->compute>>>(word, value);  // plain launch: runs after every earlier kernel of the stream has finished
    DIP_CHECK_LAUNCH(ctx, "flag_signal_kernel");
    return DIP_OK;
}

int dip_flag_wait(dip_ctx* ctx, const dip_image* flags, int index, uint32_t value) {
    uint32_t* word = nullptr;
    int rc = flag_word(ctx, flags, index, &word);
    if (rc != DIP_OK) return rc;
    if (flags->ipc_mapped || flags->device != ctx->device) return fail(ctx, DIP_ERR_ARG, "wait on a flag in this GPU's own memory (neighbours signal into it)");
    PFN_streamWaitValue32 fn = stream_wait_value_fn();
    if (!fn) return fail(ctx, DIP_ERR_UNSUPPORTED, "cuStreamWaitValue32 is not available from the driver");
    DIP_CUDA(ctx, cudaSetDevice(ctx->device));
    dep_touch(ctx);
    const CUresult r = fn(reinterpret_cast<CUstream>(ctx->compute), reinterpret_cast<CUdeviceptr>(word), value, 0x0 /* CU_STREAM_WAIT_VALUE_GEQ */);
    if (r != CUDA_SUCCESS) return fail(ctx, DIP_ERR_CUDA, "cuStreamWaitValue32 failed (%d)", (int)r);
    return DIP_OK;
}

int dip_flag_read(dip_ctx* ctx, const dip_image* flags, int index, uint32_t* value) {
    uint32_t* word = nullptr;
    int rc = flag_word(ctx, flags, index, &word);
    if (rc != DIP_OK) return rc;
    if (!value) return fail(ctx, DIP_ERR_ARG, "null argument");
    DIP_CUDA(ctx, cudaSetDevice(ctx->device));
    DIP_CUDA(ctx, cudaMemcpy(value, word, 4, cudaMemcpyDefault));
    return DIP_OK;
}

int dip_halo_fence(dip_ctx* ctx) {
    if (!ctx) return fail(nullptr, DIP_ERR_ARG, "ctx is null");
    DIP_CUDA(ctx, cudaEventRecord(ctx->ev_halo, ctx->halo));
    dep_touch(ctx);
    DIP_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, ctx->ev_halo, 0));
    return DIP_OK;
}

}  // extern "C"
