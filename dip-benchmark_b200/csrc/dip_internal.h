// Internal declarations shared by the translation units of libdip_b200.so.
// Nothing here is part of the ABI (see include/dip.h).
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdlib>
#include <utility>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "dip.h"

namespace dip {

// Device-visible description of an image stack (by value in kernel params).
struct ImgView {
    uint8_t* base;        // first byte of frame 0, row 0
    size_t pitch;         // bytes between rows (multiple of 16)
    size_t frame_stride;  // bytes between frames (multiple of 16)
    int width, height, channels, frames;
    int rowbytes;         // width * channels * bps
    int bps;              // bytes per sample: 1 (u8) or 2 (u16, little endian like the host's arrays)
    // The bytes between the end of a row and the next 16-byte boundary are padding the kernels may overwrite (library-owned
    // allocations).  False for dip_image_wrap: there they can be the caller's pixels (a region of interest of a wider image).
    bool pad_writable;
};

// A cached TMA descriptor.  The cache lives on the ROOT allocation (root_of) and is keyed by the view's geometry, so the
// short-lived frame / row views the pipelines create per chunk find the descriptor their predecessor encoded.
struct TmapKey {
    const uint8_t* base;
    size_t pitch, frame_stride;
    int rowbytes, height, frames;
    int box_words, box_rows;
    bool swizzle128;      // CU_TENSOR_MAP_SWIZZLE_128B (the tensor-core kernel's operand layout); box_words must be 32
    CUtensorMap map;
};

}  // namespace dip

struct dip_image {
    dip_ctx* ctx = nullptr;      // may be gone by the time the image is destroyed: never dereferenced there
    int device = -1;
    dip::ImgView v{};
    void* alloc = nullptr;       // cudaMalloc'ed block when owned
    bool owned = false;
    bool ipc_mapped = false;     // opened with cudaIpcOpenMemHandle
    dip_image* parent = nullptr; // views: the image whose memory is shared (must outlive the view)
    // A copy (upload/download/halo) touching this image is outstanding on a copy stream.
    cudaEvent_t copy_done = nullptr;
    bool copy_pending = false;
    std::vector<dip::TmapKey> tmaps;  // cached TMA descriptors of this allocation's views (used on the root image only)
};

namespace dip {

// Which earlier kernels of the compute stream may still be running when the next one starts, and what they touch.
// Every kernel is launched with programmatic stream serialization, so its CTAs may start while its predecessor still
// runs; a kernel that depends on earlier work executes griddepcontrol.wait first (kPdlWait).  An operation whose source
// and destination do not overlap anything an unfinished kernel touches skips that wait (kNoWait): consecutive
// independent calls -- the reference's loop over a ring of images, a batch driver walking its frames -- then overlap
// the tail of one kernel with the ramp of the next (one 3504x2336 image: 8.8 -> 7.2 us for a point op, 15.9 -> 10 us
// for the fused 1x5+5x1 Gaussian).  tools/ubench/pdl_order.cu checks what this relies on: whatever follows a no-wait
// kernel in the stream (an event record, a copy, a normally launched kernel, a synchronize) completes only after ALL
// earlier kernels have.
//   guaranteed: possibly running when the next kernel starts, but complete once the LAST launched kernel is (it, and
//               whatever it waited for, transitively)
//   unsafe:     possibly running and NOT implied by the last kernel's completion (launched before a no-wait kernel)
enum LaunchMode { kPdlWait = 0, kNoWait = 1, kFull = 2 };
struct DepRange {
    uintptr_t lo, hi;   // [lo, hi) device bytes
    bool write;
};
struct DepTracker {
    std::vector<DepRange> guaranteed, unsafe;
    bool touched = true;   // something other than a tracked kernel launch was enqueued on the compute stream since
};

// While a graph is being recorded (dip_graph_begin .. dip_graph_end) the kernels become nodes whose edges are the DATA
// dependencies between them, not the order of the calls: an operation depends on exactly the recorded kernels whose
// byte ranges it conflicts with (read-after-write, write-after-read, write-after-write).  Independent calls -- the
// reference's loop over a ring of images -- become parallel branches of the graph, so their tails and ramps overlap on
// replay just as the no-wait launches overlap them on a stream.  An operation that depends on nothing but the kernel
// recorded right before it keeps the programmatic (early-launch) edge the stream capture gives it.
struct CaptureNode {
    cudaGraphNode_t node;
    DepRange r[2];
    int n;               // 0: touches memory the tracker knows nothing about (conflicts with everything)
};
struct CaptureTracker {
    std::vector<CaptureNode> nodes;   // recorded kernels a later operation may have to depend on
    DepRange pending[2];
    int pending_n = -1;               // ranges of the last planned launch whose node has not been looked up yet (-1: none)
    bool barrier = false;             // something other than a tracked kernel was recorded: the next kernel keeps the stream's own dependencies
};

}  // namespace dip

struct dip_graph {
    dip_ctx* ctx = nullptr;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    uint64_t kernels = 0;  // kernel launches recorded (added to the ctx's launch count on every replay)
};

struct dip_ctx {
    int device = -1;
    cudaStream_t compute = nullptr;  // kernels
    bool own_compute = true;
    cudaStream_t upload = nullptr;   // H2D
    cudaStream_t download = nullptr; // D2H
    cudaStream_t halo = nullptr;     // peer row copies
    cudaEvent_t ev_compute = nullptr;  // scratch: "compute reached here"
    cudaEvent_t ev_halo = nullptr;
    cudaEvent_t t0 = nullptr, t1 = nullptr;
    int sm_count = 0;
    size_t l2_bytes = 0;
    uint64_t launches = 0;
    // pinned staging for pageable host buffers (two slots per direction)
    static constexpr int kSlots = 2;
    uint8_t* stage_up[kSlots] = {nullptr, nullptr};
    uint8_t* stage_dn[kSlots] = {nullptr, nullptr};
    cudaEvent_t stage_up_free[kSlots] = {nullptr, nullptr};
    cudaEvent_t stage_dn_done[kSlots] = {nullptr, nullptr};
    size_t stage_bytes = 0;
    // scratch
    uint8_t* flush_buf = nullptr;
    size_t flush_bytes = 0;
    uint64_t* d_checksums = nullptr;
    int d_checksums_cap = 0;
    std::vector<cudaEvent_t> marks;
    // library-owned intermediate stacks (grown on shape change): [0] dip_apply_chain's ping-pong partner, [1] the `aux`
    // image of the two-pass separable convolution (a chain may run such a convolution between its own two buffers)
    dip_image* scratch[2] = {nullptr, nullptr};
    dip::DepTracker dep;
    bool capturing = false;              // between dip_graph_begin and dip_graph_end: kernels are recorded, not run
    dip::CaptureTracker cap;
    uint64_t capture_launches0 = 0;
    void* jpeg = nullptr;                // jpeg.cu's codec state (nvJPEG handles), created on first use
    bool flag_write_ok = true;           // dip_flag_signal: cuStreamWriteValue32 has not been refused on this ctx
    std::string last_error;
};

namespace dip {

// ---- error plumbing -----------------------------------------------------------------
int fail(const dip_ctx* ctx, int code, const char* fmt, ...);
int fail_cuda(const dip_ctx* ctx, cudaError_t e, const char* what, const char* file, int line);

#define DIP_CUDA(ctx, expr)                                                       \
    do {                                                                          \
        cudaError_t _e = (expr);                                                  \
        if (_e != cudaSuccess) return dip::fail_cuda((ctx), _e, #expr, __FILE__, __LINE__); \
    } while (0)

#define DIP_CHECK_LAUNCH(ctx, name)                                               \
    do {                                                                          \
        cudaError_t _e = cudaGetLastError();                                      \
        if (_e != cudaSuccess) return dip::fail_cuda((ctx), _e, name, __FILE__, __LINE__); \
        (ctx)->launches++;                                                        \
    } while (0)

// Programmatic dependent launch: every kernel of the library is launched with the
// programmatic-stream-serialization attribute, signals `launch_dependents` at its start and
// executes `griddepcontrol.wait` before its first global-memory access, so the CTA launch and
// prologue of operation i+1 overlap the tail of operation i (the reference's loops issue
// thousands of ~10 us kernels back to back).  DIP_NO_PDL=1 disables the attribute.
inline bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("DIP_NO_PDL"); return !(e && e[0] == '1'); }();
    return on;
}

// ranges = what the kernel reads / writes (n == 0: unknown, treated as everything).  Returns how to launch it and
// updates the tracker; the caller passes `mode != kNoWait` to the kernel as its wait flag.
LaunchMode plan_launch(dip_ctx* ctx, const DepRange* ranges, int n);
// a non-kernel operation is about to be enqueued on the compute stream (event, wait, copy, memset, graph launch)
// While recording: everything recorded so far becomes a dependency of it (capture_join).
int capture_join(dip_ctx* ctx);
inline void dep_touch(dip_ctx* ctx) {
    ctx->dep.touched = true;
    if (ctx->capturing) capture_join(ctx);
}
inline DepRange dep_range(const ImgView& v, bool write) {
    const uintptr_t lo = reinterpret_cast<uintptr_t>(v.base);
    return DepRange{lo, lo + v.frame_stride * (size_t)(v.frames - 1) + v.pitch * (size_t)v.height, write};
}
inline LaunchMode plan_launch(dip_ctx* ctx, const ImgView& src, const ImgView& dst) {
    const DepRange r[2] = {dep_range(src, false), dep_range(dst, true)};
    return plan_launch(ctx, r, 2);
}

template <typename... KArgs, typename... Args>
cudaError_t launch_kernel(LaunchMode mode, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, Args&&... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl_enabled() && mode != kFull) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// ---- per-byte maps (dip_apply_chain) ---------------------------------------------------
// The byte-wise point ops (copy, inversion, threshold) are closed under composition: any sequence of them is
// x, 255 - x, or a step  x > t ? hi : lo.  A chain therefore carries ONE map where the reference would run
// one kernel per op; kernels evaluate it branch-free on four bytes at a time as
//     (x & keep) ^ (lo4 ^ (gt(x, t4) & flip4))          keep = ~0 for x / 255 - x, 0 for a step.
struct ByteMap {
    int kind = 0;            // 0: x   1: 255 - x   2: x > t ? hi : lo
    uint8_t t = 0, hi = 0, lo = 0;
    bool identity() const { return kind == 0; }
    uint8_t eval(uint8_t x) const { return kind == 0 ? x : kind == 1 ? (uint8_t)(255 - x) : (x > t ? hi : lo); }
    static ByteMap invert() { ByteMap m; m.kind = 1; return m; }
    static ByteMap step(uint8_t t, uint8_t hi, uint8_t lo) { ByteMap m; m.kind = 2; m.t = t; m.hi = hi; m.lo = lo; return m; }
    // the map "first *this, then g"
    ByteMap then(const ByteMap& g) const {
        if (kind == 2) return step(t, g.eval(hi), g.eval(lo));
        if (kind == 0) return g;
        // *this is 255 - x
        if (g.kind == 0) return *this;
        if (g.kind == 1) return ByteMap{};
        // g(255 - x) = (255 - x > g.t) ? g.hi : g.lo = (x > 254 - g.t) ? g.lo : g.hi
        if (g.t == 255) return step(255, g.lo, g.lo);  // g is the constant lo
        return step((uint8_t)(254 - g.t), g.lo, g.hi);
    }
};
struct MapParams {
    uint32_t t4 = 0, keep = 0xFFFFFFFFu, flip4 = 0, lo4 = 0;  // identity
};
inline MapParams map_params(const ByteMap& m) {
    MapParams p;
    if (m.kind == 1) { p.lo4 = 0xFFFFFFFFu; }
    else if (m.kind == 2) { p.t4 = m.t * 0x01010101u; p.keep = 0; p.flip4 = (uint32_t)(m.hi ^ m.lo) * 0x01010101u; p.lo4 = m.lo * 0x01010101u; }
    return p;
}
// Maps fused into a tiled stencil launch: `pre` rewrites the source tile in shared memory, `post` the result words.
struct ChainMaps {
    ByteMap pre, post;
};

#ifdef __CUDACC__
// 16-byte result store.  Library-owned images pad every row to 16 bytes, so the last vector of a row may spill
// into the padding; a wrapped region of interest has real pixels there, so only `valid` bytes are written.
static __device__ __noinline__ void store_partial(uint8_t* p, uint32_t x, uint32_t y, uint32_t z, uint32_t w_, int valid) {
    const uint32_t w[4] = {x, y, z, w_};
    for (int i = 0; i < valid; ++i) p[i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
}
__device__ __forceinline__ void store16(uint8_t* p, const uint4& v, int valid) {
    if (valid >= 16) __stcs(reinterpret_cast<uint4*>(p), v);          // the hot path: one predicated vector store
    else store_partial(p, v.x, v.y, v.z, v.w, valid);                 // out of line: keeps the row loops small
}
// bytes of the 16-byte chunk starting at byte `cx` of a row that may be written
__device__ __forceinline__ int chunk_valid(int cx, int rowbytes, bool exact) { return exact ? rowbytes - cx : 16; }

__device__ __forceinline__ uint32_t map_word(uint32_t x, const MapParams& m) {  // branch-free form
    return (x & m.keep) ^ (m.lo4 ^ (__vcmpgtu4(x, m.t4) & m.flip4));
}
__device__ __forceinline__ uint32_t map_step(uint32_t x, const MapParams& m) { return m.lo4 ^ (__vcmpgtu4(x, m.t4) & m.flip4); }
// N words through one map; the kind test is uniform over the launch, so x / 255 - x cost one LOP3 per word
template <int N>
__device__ __forceinline__ void map_words(uint32_t (&w)[N], const MapParams& m) {
    if (m.keep) {
#pragma unroll
        for (int i = 0; i < N; ++i) w[i] ^= m.lo4;
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) w[i] = map_step(w[i], m);
    }
}
__device__ __forceinline__ uint4 map_vec(uint4 v, const MapParams& m) {
    uint32_t w[4] = {v.x, v.y, v.z, v.w};
    map_words(w, m);
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// `wait` = false: the host proved this launch independent of every kernel that may still be running (plan_launch)
__device__ __forceinline__ void pdl_wait(bool wait) { if (wait) asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// Make the compute stream wait for outstanding copies on the operands, and set the device.
int begin_op(dip_ctx* ctx, const dip_image* src, dip_image* dst);
int check_pair(dip_ctx* ctx, const dip_image* src, const dip_image* dst, bool allow_inplace);

// Raw transfers on the ctx's copy streams without any cross-stream ordering (callers order
// them with their own events).  Pageable host memory goes through the pinned staging slots.
int enqueue_upload(dip_ctx* ctx, const ImgView& v, const uint8_t* host, size_t host_pitch);
int enqueue_download(dip_ctx* ctx, const ImgView& v, uint8_t* host, size_t host_pitch, bool* completed);

// ---- 16-bit images (depth16.cu): the point operations that are not byte-wise, and the direct stencil kernel --------
int launch_threshold16(dip_ctx* ctx, const ImgView& src, const ImgView& dst, uint16_t thresh, uint16_t maxval);
int launch_grayscale16(dip_ctx* ctx, const ImgView& src, const ImgView& dst);
// mode 0 erode, 1 dilate, 2 correlation rounded half to even (filter2D), 3 correlation rounded half up (GaussianBlur);
// taps as (dx, dy, weight) triples in pixels
int launch_direct16(dip_ctx* ctx, int mode, const dip_image* src, dip_image* dst, int valid_y0, int valid_y1, int out_y0, int out_y1,
                    int dst_y0, const int8_t* dx, const int8_t* dy, const float* wt, int ntaps);

// ---- kernels' host launchers (point_ops.cu) -----------------------------------------
enum PointOp { kCopy = 0, kInvert = 1, kThreshold = 2 };
int launch_point(dip_ctx* ctx, PointOp op, const ImgView& src, const ImgView& dst, uint8_t thresh,
                 uint8_t maxval);
int launch_grayscale(dip_ctx* ctx, const ImgView& src, const ImgView& dst);
// dst = post(gray?(pre(src))) in one pass (dip_apply_chain's point segment)
int launch_point_chain(dip_ctx* ctx, const ImgView& src, const ImgView& dst, const ByteMap& pre, bool gray, const ByteMap& post);
int launch_synth(dip_ctx* ctx, const ImgView& dst, uint32_t seed, uint64_t first_row);
int launch_checksum(dip_ctx* ctx, const ImgView& src, uint64_t first_row, uint64_t* d_out);
int launch_fill(dip_ctx* ctx, uint8_t* p, size_t bytes, uint32_t value);


// ---- stencils (stencil.cu) ------------------------------------------------------------
// Row geometry of a stencil launch.  The source's rows [valid_y0, valid_y1) are the image
// for border purposes (outside: reflect-101 / +inf); output rows [out_y0, out_y1) in source
// row coordinates are written to dst rows starting at dst_y0.
struct RowGeom {
    int valid_y0, valid_y1, out_y0, out_y1, dst_y0;
    // dip_apply_rows_peer: source rows [peer_*_y0, peer_*_y1) are read from a neighbour GPU's band (see StencilParams)
    const uint8_t* peer_top = nullptr;
    const uint8_t* peer_bot = nullptr;
    int peer_top_y0 = 0, peer_top_y1 = 0, peer_bot_y0 = 0, peer_bot_y1 = 0;
};
enum class Morph { Erode, Dilate };
// `maps` (optional): per-byte maps fused into the launch; only the 2-D tiled kernels take them (stencil_takes_maps)
int stencil_morph(dip_ctx* ctx, Morph m, const dip_image* src, dip_image* dst, const uint8_t* mask,
                  int mw, int mh, const RowGeom& g, const ChainMaps* maps = nullptr);
int stencil_convolve(dip_ctx* ctx, const dip_image* src, dip_image* dst, const float* k, int kw,
                     int kh, const RowGeom& g, const ChainMaps* maps = nullptr);
int stencil_convolve_sep(dip_ctx* ctx, const dip_image* src, dip_image* dst, const float* krow,
                         int kw, const float* kcol, int kh, const RowGeom& g, const ChainMaps* maps = nullptr);
int stencil_gaussian3(dip_ctx* ctx, const dip_image* src, dip_image* dst, const RowGeom& g, const ChainMaps* maps = nullptr);
// The tiled kernel (stencil_tiled.cuh: StencilOp) that computes `op`, or -1 (direct kernel, point op, fused chain).
int stencil_tiled_op(const dip_op& op);
bool stencil_takes_maps(const dip_op& op);    // a two-dimensional tiled kernel: instantiated with maps (stencil_mapped.cu)
bool stencil_is_gauss5(const dip_op& op);     // filter2D with the 5x5 Gaussian
bool stencil_is_erode_sq3(const dip_op& op);  // 3x3 all-ones erosion (either form)
// stencil_mapped.cu: tiled kernel `op` with the maps fused in
int launch_tiled_mapped(dip_ctx* ctx, int op, const dip_image* src, dip_image* dst, const RowGeom& g, const ChainMaps& maps);
int stencil_gauss5_erode3(dip_ctx* ctx, const dip_image* src, dip_image* dst, const RowGeom& g, const ChainMaps* maps = nullptr);

// A view of like->frames frames of the ctx-owned scratch stack `slot` (same frame geometry as `like`; allocated or grown on
// first use).  The caller destroys the view.  Used by dip_apply_chain and by the two-pass separable convolution.
int scratch_view(dip_ctx* ctx, int slot, const dip_image* like, dip_image** out);

// stencil_generic.cu: arbitrary masks / kernels of up to 7 x 7 taps on the TMA tile.  `w[dy][k]` is the tap at row dy,
// column k of the (kh x kw) mask or kernel, anchored at (kh/2, kw/2); morphology: non-zero = tap.  mode 0 erode,
// 1 dilate, 2 correlation.
struct SmallTaps {
    float w[7][7];
    int kh;
};
int launch_generic_tiled(dip_ctx* ctx, int mode, int kw, const dip_image* src, dip_image* dst, const RowGeom& g, const SmallTaps& taps);
// does the tile's 16-byte halo reach a kw x kh mask / kernel on this image (8-bit: up to 7 x 7; 16-bit: 5 x 5, 7 x 7 for one channel)?
bool generic_tiled_reach(const dip_image* src, int kw, int kh);
extern bool g_generic_float_morph;   // development knob "float_morph"

// TMA descriptor of an image view seen as a (pitch/4 words, rows, frames) u32 tensor.
int get_tmap(dip_ctx* ctx, dip_image* img, int box_words, int box_rows, const CUtensorMap** out, bool swizzle128 = false);

// jpeg.cu: releases the ctx's codec state (dip_shutdown)
void jpeg_release(dip_ctx* ctx);

// gauss_umma.cu: the fused 1x5+5x1 Gaussian on the tensor cores (tcgen05.mma.kind::i8)
bool umma_gauss5sep_ok(const dip_image* src, const dip_image* dst, const RowGeom& g);
int launch_gauss5sep_umma(dip_ctx* ctx, const dip_image* src, dip_image* dst, const RowGeom& g);

inline dip_image* root_of(dip_image* im) {
    while (im && im->parent) im = im->parent;
    return im;
}

// true when a row's last, partial 16-byte chunk must be stored byte-exactly (the bytes behind it are not ours to write)
inline bool exact_rows(const ImgView& v) { return (v.rowbytes & 15) != 0 && !v.pad_writable; }

inline RowGeom full_geom(const dip_image* src) {
    return RowGeom{0, src->v.height, 0, src->v.height, 0};
}

}  // namespace dip
