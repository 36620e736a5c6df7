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
    bool chain = false;                       // ops applied in sequence, one output (dip_pipeline_create_chain)
    int n_out = 0;                            // outputs per chunk: n_ops, or 1 for a chain
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

struct RefOps {
    RefOp ops[12];
    RefOps() {
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
    }
};

const RefOp* reference_ops() {
    static const RefOps table;  // C++11 magic static: initialised once, also when several threads (one ctx each) race here
    return table.ops;
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
        case DIP_OP_THRESHOLD:
            if (src && src->v.bps == 2 && (op->thresh16 || op->maxval16)) return dip_threshold16(ctx, src, dst, op->thresh16, op->maxval16);
            return dip_threshold(ctx, src, dst, op->thresh, op->maxval);
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

// ---- op chains ----------------------------------------------------------------------------------
}  // extern "C"

namespace {

// One pass over HBM of a planned chain.
struct ChainPass {
    enum Kind { Point, Stencil, GaussErode } kind = Point;
    // Point: dst = post(gray?(pre(src)));  Stencil: ops[op_index] with pre / post fused (tiled kernels only)
    ByteMap pre, post;
    bool gray = false;
    int op_index = -1;
    bool tiled = false;
};

bool is_byte_op(int kind) { return kind == DIP_OP_COPY || kind == DIP_OP_INVERT || kind == DIP_OP_THRESHOLD; }
bool is_stencil_op(int kind) {
    return kind == DIP_OP_ERODE || kind == DIP_OP_DILATE || kind == DIP_OP_ERODE_SEPARABLE || kind == DIP_OP_CONVOLVE ||
           kind == DIP_OP_CONVOLVE_SEPARABLE || kind == DIP_OP_GAUSSIAN_BLUR3X3 || kind == DIP_OP_GAUSS5X5_ERODE3X3;
}

// Greedy plan.  A pending point segment is (A, gray?, B): bytes through A, then grayscale, then bytes through B.
int plan_chain(dip_ctx* ctx, const dip_op* ops, int n, int channels, std::vector<ChainPass>& plan) {
    ByteMap a, b;
    bool gray = false, pending = false;
    auto flush_point = [&] {  // the pending segment becomes a pass of its own, or the epilogue of the previous tiled stencil
        if (!pending) return;
        if (!gray && !plan.empty() && plan.back().kind != ChainPass::Point && plan.back().tiled) {
            plan.back().post = plan.back().post.then(a);
        } else if (gray || !a.identity()) {
            ChainPass p;
            p.pre = a; p.gray = gray; p.post = b;
            plan.push_back(p);
        }
        a = ByteMap{}; b = ByteMap{}; gray = false; pending = false;
    };
    for (int i = 0; i < n; ++i) {
        const dip_op& op = ops[i];
        if (is_byte_op(op.kind)) {
            const ByteMap m = op.kind == DIP_OP_INVERT ? ByteMap::invert() : op.kind == DIP_OP_THRESHOLD ? ByteMap::step(op.thresh, op.maxval, 0) : ByteMap{};
            if (gray) b = b.then(m); else a = a.then(m);
            pending = true;
        } else if (op.kind == DIP_OP_GRAYSCALE) {
            if (channels == 1) { pending = true; continue; }  // GRAY2GRAY: identity
            if (gray) {
                // gray(B(Y,Y,Y)) = B(Y) on the colour channels (the weights sum to 2^15), so a second grayscale changes
                // nothing on 3 channels; with alpha it resets alpha to 255, which (A, gray, B) cannot express: new pass
                if (channels == 3) continue;
                ChainPass p;
                p.pre = a; p.gray = true; p.post = b;
                plan.push_back(p);
                a = ByteMap{}; b = ByteMap{};
            }
            gray = true; pending = true;
        } else if (is_stencil_op(op.kind)) {
            ChainPass p;
            p.kind = ChainPass::Stencil;
            p.op_index = i;
            p.tiled = stencil_takes_maps(op);
            if (op.kind == DIP_OP_GAUSS5X5_ERODE3X3) p.kind = ChainPass::GaussErode;
            // 5x5 Gaussian directly followed by the 3x3 square erosion: the fused kernel
            const bool pair = stencil_is_gauss5(op) && i + 1 < n && stencil_is_erode_sq3(ops[i + 1]);
            if (pair) p.kind = ChainPass::GaussErode;
            if (p.kind == ChainPass::GaussErode) p.tiled = true;   // the fused kernel takes the maps like a tiled stencil
            if (pending) {
                const bool prev_takes_it = !gray && !plan.empty() && plan.back().kind != ChainPass::Point && plan.back().tiled;
                if (prev_takes_it || !p.tiled) {
                    flush_point();
                } else if (gray) {  // (A, gray) is a pass; B rides in front of the stencil
                    ChainPass g;
                    g.pre = a; g.gray = true;
                    plan.push_back(g);
                    p.pre = b;
                } else {
                    p.pre = a;
                }
                a = ByteMap{}; b = ByteMap{}; gray = false; pending = false;
            }
            plan.push_back(p);
            if (pair) ++i;
        } else {
            return fail(ctx, DIP_ERR_ARG, "unknown operation kind %d at chain position %d", op.kind, i);
        }
    }
    flush_point();
    if (plan.empty()) plan.push_back(ChainPass{});  // nothing but identities: one copy
    return DIP_OK;
}

// The ctx-owned intermediate stack of dip_apply_chain: same frame geometry as `like`, at least as many frames (a
// pipeline's partial last chunk reuses the full-chunk stack instead of reallocating it on every run).  *out is a view
// of exactly like->frames frames that the caller destroys.
int ensure_scratch(dip_ctx* ctx, const dip_image* like, dip_image** out) { return dip::scratch_view(ctx, 0, like, out); }

}  // namespace

int dip::scratch_view(dip_ctx* ctx, int slot, const dip_image* like, dip_image** out) {
    dip_image* s = ctx->scratch[slot];
    if (!(s && s->v.width == like->v.width && s->v.height == like->v.height && s->v.channels == like->v.channels && s->v.bps == like->v.bps && s->v.frames >= like->v.frames)) {
        if (ctx->capturing) return fail(ctx, DIP_ERR_UNSUPPORTED, "run this chain once before recording it: its scratch stack is allocated on first use");
        if (s) {
            dep_touch(ctx);
            DIP_CUDA(ctx, cudaStreamSynchronize(ctx->compute));  // kernels may still be reading the old one
            dip_image_destroy(s);
            ctx->scratch[slot] = nullptr;
        }
        int rc = like->v.bps == 2 ? dip_image_create16(ctx, like->v.width, like->v.height, like->v.channels, like->v.frames, &ctx->scratch[slot])
                                  : dip_image_create_batch(ctx, like->v.width, like->v.height, like->v.channels, like->v.frames, &ctx->scratch[slot]);
        if (rc != DIP_OK) return rc;
    }
    return dip_image_view(ctx, ctx->scratch[slot], 0, like->v.frames, out);
}

namespace {

int run_stencil_pass(dip_ctx* ctx, const dip_op& op, const ChainPass& ps, const dip_image* src, dip_image* dst) {
    const RowGeom g = full_geom(src);
    ChainMaps maps;
    maps.pre = ps.pre; maps.post = ps.post;
    const ChainMaps* mp = (ps.pre.identity() && ps.post.identity()) ? nullptr : &maps;
    if (ps.kind == ChainPass::GaussErode) return stencil_gauss5_erode3(ctx, src, dst, g, mp);
    switch (op.kind) {
        case DIP_OP_ERODE: return stencil_morph(ctx, Morph::Erode, src, dst, op.mask, op.mask_w, op.mask_h, g, mp);
        case DIP_OP_DILATE: return stencil_morph(ctx, Morph::Dilate, src, dst, op.mask, op.mask_w, op.mask_h, g, mp);
        case DIP_OP_ERODE_SEPARABLE: {
            if (op.mask_w < 1 || op.mask_h < 1 || op.mask_w * op.mask_h > 225) return fail(ctx, DIP_ERR_ARG, "bad separable mask %dx%d", op.mask_w, op.mask_h);
            std::vector<uint8_t> ones((size_t)op.mask_w * op.mask_h, 1);
            return stencil_morph(ctx, Morph::Erode, src, dst, ones.data(), op.mask_w, op.mask_h, g, mp);
        }
        case DIP_OP_CONVOLVE: return stencil_convolve(ctx, src, dst, op.kernel, op.kw, op.kh, g, mp);
        case DIP_OP_CONVOLVE_SEPARABLE: return stencil_convolve_sep(ctx, src, dst, op.kernel, op.kw, op.kernel_col, op.kh, g, mp);
        case DIP_OP_GAUSSIAN_BLUR3X3: return stencil_gaussian3(ctx, src, dst, g, mp);
    }
    return fail(ctx, DIP_ERR_ARG, "not a stencil operation: %d", op.kind);
}

}  // namespace

extern "C" {

int dip_chain_passes(const dip_op* ops, int n, int channels, int* passes) {
    if (!ops || n < 1 || !passes || (channels != 1 && channels != 3 && channels != 4)) return fail(nullptr, DIP_ERR_ARG, "bad argument");
    std::vector<ChainPass> plan;
    const int rc = plan_chain(nullptr, ops, n, channels, plan);
    if (rc != DIP_OK) return rc;
    *passes = (int)plan.size();
    return DIP_OK;
}

int dip_chain_lut(const dip_op* ops, int n, uint8_t lut[256]) {
    if (!ops || n < 1 || !lut) return fail(nullptr, DIP_ERR_ARG, "bad argument");
    ByteMap m;
    for (int i = 0; i < n; ++i) {
        if (!is_byte_op(ops[i].kind)) return fail(nullptr, DIP_ERR_ARG, "operation %d (kind %d) is not byte-wise", i, ops[i].kind);
        m = m.then(ops[i].kind == DIP_OP_INVERT ? ByteMap::invert() : ops[i].kind == DIP_OP_THRESHOLD ? ByteMap::step(ops[i].thresh, ops[i].maxval, 0) : ByteMap{});
    }
    for (int x = 0; x < 256; ++x) lut[x] = m.eval((uint8_t)x);
    return DIP_OK;
}

int dip_apply_chain(dip_ctx* ctx, const dip_op* ops, int n, const dip_image* src, dip_image* dst) {
    if (!ctx || !ops || n < 1) return fail(ctx, DIP_ERR_ARG, "bad argument");
    int rc = check_pair(ctx, src, dst, true);
    if (rc != DIP_OK) return rc;
    if (src->v.bps == 2) {
        // 16-bit samples: the operations one after the other (the per-byte maps a chain folds are 8-bit maps), with the
        // same ping-pong between the scratch stack and dst
        dip_image* scratch16 = nullptr;
        if (n > 1 && (rc = ensure_scratch(ctx, src, &scratch16)) != DIP_OK) return rc;
        const dip_image* in = src;
        for (int j = 1; j <= n && rc == DIP_OK; ++j) {
            dip_image* out = ((n - j) & 1) ? scratch16 : dst;
            rc = dip_apply(ctx, &ops[j - 1], in, out);
            in = out;
        }
        dip_image_destroy(scratch16);
        return rc;
    }
    std::vector<ChainPass> plan;
    if ((rc = plan_chain(ctx, ops, n, src->v.channels, plan)) != DIP_OK) return rc;
    const int k = (int)plan.size();
    bool any_stencil = false;
    for (const ChainPass& p : plan) any_stencil |= p.kind != ChainPass::Point;
    if (any_stencil && (rc = check_pair(ctx, src, dst, false)) != DIP_OK) return rc;
    dip_image* scratch = nullptr;
    if (k > 1 && (rc = ensure_scratch(ctx, src, &scratch)) != DIP_OK) return rc;
    rc = begin_op(ctx, src, dst);
    // pass j (1-based) writes dst when k - j is even, the scratch stack otherwise: the last pass lands in dst and no
    // pass reads what it writes
    const dip_image* in = src;
    for (int j = 1; j <= k && rc == DIP_OK; ++j) {
        const ChainPass& p = plan[j - 1];
        dip_image* out = ((k - j) & 1) ? scratch : dst;
        if (p.kind == ChainPass::Point) rc = launch_point_chain(ctx, in->v, out->v, p.pre, p.gray, p.post);
        else rc = run_stencil_pass(ctx, ops[p.op_index], p, in, out);
        in = out;
    }
    dip_image_destroy(scratch);
    return rc;
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

namespace {

// Translates a dip_peer_rows into the kernel's view: the peer byte that corresponds to byte 0 of source row 0.
int set_peer(dip_ctx* ctx, const dip_image* src, const dip_peer_rows* pr, const uint8_t** base, int* y0, int* y1) {
    if (!pr || !pr->image || pr->rows == 0) return DIP_OK;
    const ImgView &a = src->v, &b = pr->image->v;
    if (pr->rows < 0 || pr->src_row0 < 0 || pr->src_row0 + pr->rows > a.height || pr->image_row0 < 0 || pr->image_row0 + pr->rows > b.height)
        return fail(ctx, DIP_ERR_ARG, "peer rows out of bounds");
    if (a.rowbytes != b.rowbytes || a.channels != b.channels || a.frames != b.frames || a.pitch != b.pitch ||
        (a.frames > 1 && a.frame_stride != b.frame_stride))
        return fail(ctx, DIP_ERR_SHAPE, "a peer band must have the source's row length, pitch and frame layout");
    *base = b.base + ((ptrdiff_t)pr->image_row0 - (ptrdiff_t)pr->src_row0) * (ptrdiff_t)a.pitch;
    *y0 = pr->src_row0;
    *y1 = pr->src_row0 + pr->rows;
    return DIP_OK;
}

int apply_rows_impl(dip_ctx* ctx, const dip_op* op, const dip_image* src, int valid_row0, int valid_rows, int out_row0, int out_rows,
                    dip_image* dst, int dst_row0, const dip_peer_rows* above, const dip_peer_rows* below);

}  // namespace

extern "C" int dip_apply_rows(dip_ctx* ctx, const dip_op* op, const dip_image* src, int valid_row0, int valid_rows,
                              int out_row0, int out_rows, dip_image* dst, int dst_row0) {
    return apply_rows_impl(ctx, op, src, valid_row0, valid_rows, out_row0, out_rows, dst, dst_row0, nullptr, nullptr);
}

extern "C" int dip_apply_rows_peer(dip_ctx* ctx, const dip_op* op, const dip_image* src, int valid_row0, int valid_rows,
                                   int out_row0, int out_rows, dip_image* dst, int dst_row0, const dip_peer_rows* above,
                                   const dip_peer_rows* below) {
    return apply_rows_impl(ctx, op, src, valid_row0, valid_rows, out_row0, out_rows, dst, dst_row0, above, below);
}

namespace {

int apply_rows_impl(dip_ctx* ctx, const dip_op* op, const dip_image* src, int valid_row0, int valid_rows, int out_row0, int out_rows,
                    dip_image* dst, int dst_row0, const dip_peer_rows* above, const dip_peer_rows* below) {
    if (!ctx || !op || !src || !dst) return fail(ctx, DIP_ERR_ARG, "null argument");
    if (out_rows < 1 || valid_rows < 1) return fail(ctx, DIP_ERR_ARG, "empty row range");
    if (src->v.bps != dst->v.bps) return fail(ctx, DIP_ERR_SHAPE, "src has %d-bit samples, dst %d-bit", 8 * src->v.bps, 8 * dst->v.bps);
    RowGeom g{valid_row0, valid_row0 + valid_rows, out_row0, out_row0 + out_rows, dst_row0};
    int rc = set_peer(ctx, src, above, &g.peer_top, &g.peer_top_y0, &g.peer_top_y1);
    if (rc == DIP_OK) rc = set_peer(ctx, src, below, &g.peer_bot, &g.peer_bot_y0, &g.peer_bot_y1);
    if (rc != DIP_OK) return rc;
    if ((g.peer_top || g.peer_bot) && op->kind != DIP_OP_GAUSS5X5_ERODE3X3 && stencil_tiled_op(*op) < 0)
        return fail(ctx, DIP_ERR_UNSUPPORTED, "in-kernel peer rows need a tiled stencil kernel; use dip_copy_rows for this operation");
    rc = begin_op(ctx, src, dst);
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

}  // namespace

extern "C" {

static int pipeline_create(dip_ctx* ctx, int w, int h, int c, int chunk_frames, const dip_op* ops, int n_ops, bool chain, dip_pipeline** out) {
    if (!ctx || !ops || !out || n_ops < 1 || chunk_frames < 1) return fail(ctx, DIP_ERR_ARG, "bad argument");
    *out = nullptr;
    dip_pipeline* p = new (std::nothrow) dip_pipeline();
    if (!p) return fail(ctx, DIP_ERR_NOMEM, "out of host memory");
    p->ctx = ctx; p->w = w; p->h = h; p->c = c; p->chunk = chunk_frames; p->n_ops = n_ops;
    p->chain = chain;
    p->n_out = chain ? 1 : n_ops;
    p->ops.assign(ops, ops + n_ops);
    const int n_out = p->n_out;
    p->out.assign(2 * (size_t)n_out, nullptr);
    p->out_ready.assign(2 * (size_t)n_out, nullptr);
    p->out_free.assign(2 * (size_t)n_out, nullptr);
    int rc = DIP_OK;
    cudaSetDevice(ctx->device);
    auto mk = [&](cudaEvent_t* e) { if (rc == DIP_OK && cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) rc = fail(ctx, DIP_ERR_CUDA, "cudaEventCreate failed"); };
    for (int s = 0; s < 2 && rc == DIP_OK; ++s) {
        rc = dip_image_create_batch(ctx, w, h, c, chunk_frames, &p->in[s]);
        mk(&p->in_ready[s]); mk(&p->in_free[s]);
        for (int o = 0; o < n_out && rc == DIP_OK; ++o) {
            rc = dip_image_create_batch(ctx, w, h, c, chunk_frames, &p->out[(size_t)s * n_out + o]);
            mk(&p->out_ready[(size_t)s * n_out + o]); mk(&p->out_free[(size_t)s * n_out + o]);
        }
    }
    if (rc != DIP_OK) { dip_pipeline_destroy(p); return rc; }
    *out = p;
    return DIP_OK;
}

int dip_pipeline_create(dip_ctx* ctx, int w, int h, int c, int chunk_frames, const dip_op* ops, int n_ops, dip_pipeline** out) {
    return pipeline_create(ctx, w, h, c, chunk_frames, ops, n_ops, false, out);
}

int dip_pipeline_create_chain(dip_ctx* ctx, int w, int h, int c, int chunk_frames, const dip_op* ops, int n_ops, dip_pipeline** out) {
    return pipeline_create(ctx, w, h, c, chunk_frames, ops, n_ops, true, out);
}

int dip_run_chain_host(dip_ctx* ctx, const uint8_t* host_in, int w, int h, int c, int frames, const dip_op* ops, int n_ops,
                       uint8_t* host_out, int chunk_frames) {
    if (!ctx || !host_in || !ops || !host_out || n_ops < 1 || frames < 1) return fail(ctx, DIP_ERR_ARG, "bad argument");
    if (chunk_frames < 1) chunk_frames = 4;
    chunk_frames = std::min(chunk_frames, frames);
    dip_pipeline* p = nullptr;
    int rc = dip_pipeline_create_chain(ctx, w, h, c, chunk_frames, ops, n_ops, &p);
    if (rc != DIP_OK) return rc;
    const int ring = (frames + chunk_frames - 1) / chunk_frames * chunk_frames;
    uint8_t* outs[1] = {host_out};
    rc = dip_pipeline_run(p, host_in, ring, frames, outs, ring);
    dip_pipeline_destroy(p);
    return rc;
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
    // CUDA calls inside the loop record their error in rc and fall through, so that the per-chunk views are released
    // and the trailing dip_sync still runs (host buffers must not be in flight when the call returns)
    auto cu = [&](cudaError_t e, const char* what) { if (e != cudaSuccess && rc == DIP_OK) rc = fail_cuda(ctx, e, what, __FILE__, __LINE__); };
    for (int f0 = 0; f0 < frames && rc == DIP_OK; f0 += p->chunk, slot ^= 1) {
        const int n = std::min(p->chunk, frames - f0);
        dip_image* iv = nullptr;
        if ((rc = dip_image_view(ctx, p->in[slot], 0, n, &iv)) != DIP_OK) break;
        // H2D: wait until the kernels of two chunks ago released this input slot
        if (used[slot]) cu(cudaStreamWaitEvent(ctx->upload, p->in_free[slot], 0), "cudaStreamWaitEvent(upload, in_free)");
        if (rc == DIP_OK) rc = enqueue_upload(ctx, iv->v, host_in + frame_bytes * (size_t)(f0 % in_ring), frame_bytes / p->h);
        if (rc == DIP_OK) {
            cu(cudaEventRecord(p->in_ready[slot], ctx->upload), "cudaEventRecord(in_ready)");
            dep_touch(ctx);
            cu(cudaStreamWaitEvent(ctx->compute, p->in_ready[slot], 0), "cudaStreamWaitEvent(compute, in_ready)");
        }
        for (int o = 0; o < p->n_out && rc == DIP_OK; ++o) {
            const size_t idx = (size_t)slot * p->n_out + o;
            dip_image* ov = nullptr;
            if ((rc = dip_image_view(ctx, p->out[idx], 0, n, &ov)) != DIP_OK) break;
            // the kernel may overwrite this output slot only after its previous download finished
            dep_touch(ctx);
            if (used[slot]) cu(cudaStreamWaitEvent(ctx->compute, p->out_free[idx], 0), "cudaStreamWaitEvent(compute, out_free)");
            if (rc == DIP_OK) rc = p->chain ? dip_apply_chain(ctx, p->ops.data(), p->n_ops, iv, ov) : dip_apply(ctx, &p->ops[o], iv, ov);
            if (rc == DIP_OK) {
                dep_touch(ctx);
                cu(cudaEventRecord(p->out_ready[idx], ctx->compute), "cudaEventRecord(out_ready)");
                cu(cudaStreamWaitEvent(ctx->download, p->out_ready[idx], 0), "cudaStreamWaitEvent(download, out_ready)");
                bool completed = false;
                if (rc == DIP_OK) rc = enqueue_download(ctx, ov->v, host_out[o] + frame_bytes * (size_t)(f0 % out_ring), frame_bytes / p->h, &completed);
                if (rc == DIP_OK) cu(cudaEventRecord(p->out_free[idx], ctx->download), "cudaEventRecord(out_free)");
            }
            dip_image_destroy(ov);
        }
        dep_touch(ctx);
        if (rc == DIP_OK) cu(cudaEventRecord(p->in_free[slot], ctx->compute), "cudaEventRecord(in_free)");
        used[slot] = true;
        dip_image_destroy(iv);
    }
    int rc2 = dip_sync(ctx);
    return rc != DIP_OK ? rc : rc2;
}

}  // extern "C"
