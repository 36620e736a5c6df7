// The reference's operation list as data, the op dispatcher and the batched host pipeline.
//
// Mirrors the `operations` list of /root/reference/benchmarks/opencv/benchmark.py:92-105
// (sycl/benchmark.cpp:331-358) and the per-op loop around it (:109-114 / :365-375):
// every op reads the uploaded input image(s) and writes the output image(s).
#include <algorithm>

#include "dip_internal.h"

using namespace dip;

// Chunked three-stream pipeline: while chunk k's kernels run, chunk k+1 uploads and chunk k-1's
// results download.  Two input stacks and two sets of output stacks rotate; ordering between the
// streams is carried by per-slot events owned by the pipeline (not by the per-image tracking of
// dip_upload / dip_download, which would serialise an upload behind all enqueued kernels).
struct dip_pipeline {
    dip_ctx* ctx = nullptr;
    int w = 0, h = 0, c = 0, chunk = 0, n_ops = 0;
    std::vector<dip_op> ops;
    dip_image* in[2] = {nullptr, nullptr};
    std::vector<dip_image*> out;              // [slot][op]
    cudaEvent_t in_ready[2] = {nullptr, nullptr};   // upload of the slot finished
    cudaEvent_t in_free[2] = {nullptr, nullptr};    // last kernel reading the slot finished
    std::vector<cudaEvent_t> out_ready;       // [slot][op] kernel finished
    std::vector<cudaEvent_t> out_free;        // [slot][op] download finished
};

namespace {

// masks and kernels of opencv/benchmark.py:46-72
const uint8_t kCrossMask[9] = {0, 1, 0, 1, 1, 1, 0, 1, 0};
const uint8_t kSquareMask[9] = {1, 1, 1, 1, 1, 1, 1, 1, 1};
const float kBlur3[9] = {1 / 16.f, 2 / 16.f, 1 / 16.f, 2 / 16.f, 4 / 16.f, 2 / 16.f, 1 / 16.f, 2 / 16.f, 1 / 16.f};
const float kBlur3Sep[3] = {1 / 4.f, 1 / 2.f, 1 / 4.f};
const float kBlur5[25] = {1 / 256.f, 4 / 256.f,  6 / 256.f,  4 / 256.f,  1 / 256.f, 4 / 256.f, 16 / 256.f, 24 / 256.f, 16 / 256.f,
                          4 / 256.f, 6 / 256.f,  24 / 256.f, 36 / 256.f, 24 / 256.f, 6 / 256.f, 4 / 256.f,  16 / 256.f, 24 / 256.f,
                          16 / 256.f, 4 / 256.f, 1 / 256.f,  4 / 256.f,  6 / 256.f,  4 / 256.f, 1 / 256.f};
const float kBlur5Sep[5] = {1 / 16.f, 4 / 16.f, 6 / 16.f, 4 / 16.f, 1 / 16.f};

struct RefOp {
    const char* description;
    const char* prefix;
    dip_op op;
};

dip_op make_op(int kind) {
    dip_op o{};
    o.kind = kind;
    return o;
}

const RefOp* reference_ops() {
    static RefOp ops[12];
    static bool init = false;
    if (!init) {
        ops[0] = {"Copy", "copy", make_op(DIP_OP_COPY)};
        ops[1] = {"Inversion", "inversion", make_op(DIP_OP_INVERT)};
        ops[2] = {"Grayscale", "grayscale", make_op(DIP_OP_GRAYSCALE)};
        ops[3] = {"Threshold", "threshold", make_op(DIP_OP_THRESHOLD)};
        ops[3].op.thresh = 127; ops[3].op.maxval = 255;
        ops[4] = {"Erosion (3x3 Cross Kernel)", "erosion-cross", make_op(DIP_OP_ERODE)};
        ops[4].op.mask = kCrossMask; ops[4].op.mask_w = 3; ops[4].op.mask_h = 3;
        ops[5] = {"Erosion (3x3 Square Kernel)", "erosion-square", make_op(DIP_OP_ERODE)};
        ops[5].op.mask = kSquareMask; ops[5].op.mask_w = 3; ops[5].op.mask_h = 3;
        ops[6] = {"Erosion (1x3+3x1 Square Kernel)", "erosion-square-separated", make_op(DIP_OP_ERODE_SEPARABLE)};
        ops[6].op.mask_w = 3; ops[6].op.mask_h = 3;
        ops[7] = {"Convolution (3x3 Gaussian Blur Kernel)", "convolution-gaussian-blur-3x3", make_op(DIP_OP_CONVOLVE)};
        ops[7].op.kernel = kBlur3; ops[7].op.kw = 3; ops[7].op.kh = 3;
        ops[8] = {"Convolution (1x3+3x1 Gaussian Blur Kernel)", "convolution-gaussian-blur-3x3-separated", make_op(DIP_OP_CONVOLVE_SEPARABLE)};
        ops[8].op.kernel = kBlur3Sep; ops[8].op.kernel_col = kBlur3Sep; ops[8].op.kw = 3; ops[8].op.kh = 3;
        ops[9] = {"Convolution (5x5 Gaussian Blur Kernel)", "convolution-gaussian-blur-5x5", make_op(DIP_OP_CONVOLVE)};
        ops[9].op.kernel = kBlur5; ops[9].op.kw = 5; ops[9].op.kh = 5;
        ops[10] = {"Convolution (1x5+5x1 Gaussian Blur Kernel)", "convolution-gaussian-blur-5x5-separated", make_op(DIP_OP_CONVOLVE_SEPARABLE)};
        ops[10].op.kernel = kBlur5Sep; ops[10].op.kernel_col = kBlur5Sep; ops[10].op.kw = 5; ops[10].op.kh = 5;
        ops[11] = {"Gaussian Blur (3x3 Kernel)", "gaussian-blur-3x3", make_op(DIP_OP_GAUSSIAN_BLUR3X3)};
        init = true;
    }
    return ops;
}

}  // namespace

extern "C" {

int dip_reference_ops(dip_op* out, int capacity) {
    if (!out || capacity < 12) return fail(nullptr, DIP_ERR_ARG, "need room for 12 operations");
    const RefOp* r = reference_ops();
    for (int i = 0; i < 12; ++i) out[i] = r[i].op;
    return 12;
}

const char* dip_reference_op_prefix(int i) { return (i >= 0 && i < 12) ? reference_ops()[i].prefix : nullptr; }
const char* dip_reference_op_description(int i) { return (i >= 0 && i < 12) ? reference_ops()[i].description : nullptr; }

int dip_apply(dip_ctx* ctx, const dip_op* op, const dip_image* src, dip_image* dst) {
    if (!ctx || !op) return fail(ctx, DIP_ERR_ARG, "null argument");
    switch (op->kind) {
        case DIP_OP_COPY: return dip_copy(ctx, src, dst);
        case DIP_OP_INVERT: return dip_invert(ctx, src, dst);
        case DIP_OP_GRAYSCALE: return dip_grayscale(ctx, src, dst);
        case DIP_OP_THRESHOLD: return dip_threshold(ctx, src, dst, op->thresh, op->maxval);
        case DIP_OP_ERODE: return dip_erode(ctx, src, dst, op->mask, op->mask_w, op->mask_h);
        case DIP_OP_DILATE: return dip_dilate(ctx, src, dst, op->mask, op->mask_w, op->mask_h);
        case DIP_OP_ERODE_SEPARABLE: return dip_erode_separable(ctx, src, dst, op->mask_w, op->mask_h);
        case DIP_OP_CONVOLVE: return dip_convolve(ctx, src, dst, op->kernel, op->kw, op->kh);
        case DIP_OP_CONVOLVE_SEPARABLE: return dip_convolve_separable(ctx, src, dst, op->kernel, op->kw, op->kernel_col, op->kh);
        case DIP_OP_GAUSSIAN_BLUR3X3: return dip_gaussian_blur3x3(ctx, src, dst);
        case DIP_OP_GAUSS5X5_ERODE3X3: return dip_gauss5x5_erode3x3(ctx, src, dst);
    }
    return fail(ctx, DIP_ERR_ARG, "unknown operation kind %d", op->kind);
}

// One-shot form of the pipeline below.
int dip_run_ops_host(dip_ctx* ctx, const uint8_t* host_in, int w, int h, int c, int frames, const dip_op* ops, int n_ops,
                     uint8_t* const* host_out, int chunk_frames) {
    if (!ctx || !host_in || !ops || !host_out || n_ops < 1 || frames < 1) return fail(ctx, DIP_ERR_ARG, "bad argument");
    if (chunk_frames < 1) chunk_frames = 4;
    chunk_frames = std::min(chunk_frames, frames);
    dip_pipeline* p = nullptr;
    int rc = dip_pipeline_create(ctx, w, h, c, chunk_frames, ops, n_ops, &p);
    if (rc != DIP_OK) return rc;
    const int ring = (frames + chunk_frames - 1) / chunk_frames * chunk_frames;  // no wrap: plain arrays
    rc = dip_pipeline_run(p, host_in, ring, frames, host_out, ring);
    dip_pipeline_destroy(p);
    return rc;
}

}  // extern "C"

extern "C" int dip_apply_rows(dip_ctx* ctx, const dip_op* op, const dip_image* src, int valid_row0, int valid_rows,
                              int out_row0, int out_rows, dip_image* dst, int dst_row0) {
    if (!ctx || !op || !src || !dst) return fail(ctx, DIP_ERR_ARG, "null argument");
    if (out_rows < 1 || valid_rows < 1) return fail(ctx, DIP_ERR_ARG, "empty row range");
    const RowGeom g{valid_row0, valid_row0 + valid_rows, out_row0, out_row0 + out_rows, dst_row0};
    int rc = begin_op(ctx, src, dst);
    if (rc != DIP_OK) return rc;
    switch (op->kind) {
        case DIP_OP_ERODE: return stencil_morph(ctx, Morph::Erode, src, dst, op->mask, op->mask_w, op->mask_h, g);
        case DIP_OP_DILATE: return stencil_morph(ctx, Morph::Dilate, src, dst, op->mask, op->mask_w, op->mask_h, g);
        case DIP_OP_ERODE_SEPARABLE: {
            if (op->mask_w < 1 || op->mask_h < 1 || op->mask_w * op->mask_h > 225) return fail(ctx, DIP_ERR_ARG, "bad separable mask");
            std::vector<uint8_t> ones((size_t)op->mask_w * op->mask_h, 1);
            return stencil_morph(ctx, Morph::Erode, src, dst, ones.data(), op->mask_w, op->mask_h, g);
        }
        case DIP_OP_CONVOLVE: return stencil_convolve(ctx, src, dst, op->kernel, op->kw, op->kh, g);
        case DIP_OP_CONVOLVE_SEPARABLE: return stencil_convolve_sep(ctx, src, dst, op->kernel, op->kw, op->kernel_col, op->kh, g);
        case DIP_OP_GAUSSIAN_BLUR3X3: return stencil_gaussian3(ctx, src, dst, g);
        case DIP_OP_GAUSS5X5_ERODE3X3: return stencil_gauss5_erode3(ctx, src, dst, g);
        default: break;
    }
    // point operations: plain row views
    if (out_row0 < 0 || out_row0 + out_rows > src->v.height || dst_row0 < 0 || dst_row0 + out_rows > dst->v.height)
        return fail(ctx, DIP_ERR_ARG, "row range out of bounds");
    dip_image *sv = nullptr, *dv = nullptr;
    rc = dip_image_rows(ctx, const_cast<dip_image*>(src), out_row0, out_rows, &sv);
    if (rc == DIP_OK) rc = dip_image_rows(ctx, dst, dst_row0, out_rows, &dv);
    if (rc == DIP_OK) rc = dip_apply(ctx, op, sv, dv);
    dip_image_destroy(sv);
    dip_image_destroy(dv);
    return rc;
}

extern "C" {

int dip_pipeline_create(dip_ctx* ctx, int w, int h, int c, int chunk_frames, const dip_op* ops, int n_ops, dip_pipeline** out) {
    if (!ctx || !ops || !out || n_ops < 1 || chunk_frames < 1) return fail(ctx, DIP_ERR_ARG, "bad argument");
    *out = nullptr;
    dip_pipeline* p = new (std::nothrow) dip_pipeline();
    if (!p) return fail(ctx, DIP_ERR_NOMEM, "out of host memory");
    p->ctx = ctx; p->w = w; p->h = h; p->c = c; p->chunk = chunk_frames; p->n_ops = n_ops;
    p->ops.assign(ops, ops + n_ops);
    p->out.assign(2 * (size_t)n_ops, nullptr);
    p->out_ready.assign(2 * (size_t)n_ops, nullptr);
    p->out_free.assign(2 * (size_t)n_ops, nullptr);
    int rc = DIP_OK;
    cudaSetDevice(ctx->device);
    auto mk = [&](cudaEvent_t* e) { if (rc == DIP_OK && cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) rc = fail(ctx, DIP_ERR_CUDA, "cudaEventCreate failed"); };
    for (int s = 0; s < 2 && rc == DIP_OK; ++s) {
        rc = dip_image_create_batch(ctx, w, h, c, chunk_frames, &p->in[s]);
        mk(&p->in_ready[s]); mk(&p->in_free[s]);
        for (int o = 0; o < n_ops && rc == DIP_OK; ++o) {
            rc = dip_image_create_batch(ctx, w, h, c, chunk_frames, &p->out[(size_t)s * n_ops + o]);
            mk(&p->out_ready[(size_t)s * n_ops + o]); mk(&p->out_free[(size_t)s * n_ops + o]);
        }
    }
    if (rc != DIP_OK) { dip_pipeline_destroy(p); return rc; }
    *out = p;
    return DIP_OK;
}

int dip_pipeline_destroy(dip_pipeline* p) {
    if (!p) return DIP_OK;
    dip_sync(p->ctx);
    for (int s = 0; s < 2; ++s) {
        dip_image_destroy(p->in[s]);
        if (p->in_ready[s]) cudaEventDestroy(p->in_ready[s]);
        if (p->in_free[s]) cudaEventDestroy(p->in_free[s]);
    }
    for (dip_image* im : p->out) dip_image_destroy(im);
    for (cudaEvent_t e : p->out_ready) if (e) cudaEventDestroy(e);
    for (cudaEvent_t e : p->out_free) if (e) cudaEventDestroy(e);
    delete p;
    return DIP_OK;
}

int dip_pipeline_run(dip_pipeline* p, const uint8_t* host_in, int in_ring, int frames, uint8_t* const* host_out, int out_ring) {
    if (!p || !host_in || !host_out || frames < 1) return fail(p ? p->ctx : nullptr, DIP_ERR_ARG, "bad argument");
    dip_ctx* ctx = p->ctx;
    if (in_ring < p->chunk || out_ring < p->chunk || in_ring % p->chunk || out_ring % p->chunk)
        return fail(ctx, DIP_ERR_ARG, "host ring sizes must be positive multiples of chunk_frames (%d)", p->chunk);
    DIP_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t frame_bytes = (size_t)p->w * p->c * p->h;
    int rc = DIP_OK;
    int slot = 0;
    bool used[2] = {false, false};
    for (int f0 = 0; f0 < frames && rc == DIP_OK; f0 += p->chunk, slot ^= 1) {
        const int n = std::min(p->chunk, frames - f0);
        dip_image* iv = nullptr;
        if ((rc = dip_image_view(ctx, p->in[slot], 0, n, &iv)) != DIP_OK) break;
        // H2D: wait until the kernels of two chunks ago released this input slot
        if (used[slot]) DIP_CUDA(ctx, cudaStreamWaitEvent(ctx->upload, p->in_free[slot], 0));
        rc = enqueue_upload(ctx, iv->v, host_in + frame_bytes * (size_t)(f0 % in_ring), frame_bytes / p->h);
        if (rc == DIP_OK) {
            DIP_CUDA(ctx, cudaEventRecord(p->in_ready[slot], ctx->upload));
            DIP_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, p->in_ready[slot], 0));
        }
        for (int o = 0; o < p->n_ops && rc == DIP_OK; ++o) {
            const size_t idx = (size_t)slot * p->n_ops + o;
            dip_image* ov = nullptr;
            if ((rc = dip_image_view(ctx, p->out[idx], 0, n, &ov)) != DIP_OK) break;
            // the kernel may overwrite this output slot only after its previous download finished
            if (used[slot]) DIP_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, p->out_free[idx], 0));
            rc = dip_apply(ctx, &p->ops[o], iv, ov);
            if (rc == DIP_OK) {
                DIP_CUDA(ctx, cudaEventRecord(p->out_ready[idx], ctx->compute));
                DIP_CUDA(ctx, cudaStreamWaitEvent(ctx->download, p->out_ready[idx], 0));
                bool completed = false;
                rc = enqueue_download(ctx, ov->v, host_out[o] + frame_bytes * (size_t)(f0 % out_ring), frame_bytes / p->h, &completed);
                if (rc == DIP_OK) DIP_CUDA(ctx, cudaEventRecord(p->out_free[idx], ctx->download));
            }
            dip_image_destroy(ov);
        }
        if (rc == DIP_OK) DIP_CUDA(ctx, cudaEventRecord(p->in_free[slot], ctx->compute));
        used[slot] = true;
        dip_image_destroy(iv);
    }
    int rc2 = dip_sync(ctx);
    return rc != DIP_OK ? rc : rc2;
}

}  // extern "C"
