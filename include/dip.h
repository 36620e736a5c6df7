/*
 * dip.h -- C-ABI of the B200-native image-op library (libdip_b200.so).
 *
 * This is the drop-in boundary for the dip-benchmark "operation layer": the 14
 * operations every backend of the reference implements inside its
 * perform_benchmark closures.  The reference has no FFI of its own; each entry
 * point below cites the reference call it replaces:
 *   opencv:N  = /root/reference/benchmarks/opencv/benchmark.py line N (the parity oracle)
 *   sycl:N    = /root/reference/benchmarks/sycl/benchmark.cpp line N (native call shape)
 *   vgl:N     = /root/reference/benchmarks/visiongl/benchmark.cpp line N (C call shape)
 *
 * Conventions
 *  - plain C, no exceptions cross the boundary; every function returns a
 *    dip_status (0 = OK, negative = error) and records a message retrievable
 *    with dip_last_error_string().
 *  - images are opaque handles to DEVICE memory owned by the library: row-major
 *    H x W x C interleaved u8 (C in {1,3,4}, BGR(A) order as cv.imread gives),
 *    `frames` images stacked in one allocation, rows `pitch` bytes apart
 *    (pitch = W*C rounded up to 16).  A plain image is a stack of 1 frame; every
 *    operation processes all frames of its operands in ONE kernel launch.
 *  - host pointers are borrowed for the duration of the call chain: transfers
 *    and kernels are stream-ordered and asynchronous; call dip_sync() before
 *    touching a host buffer handed to dip_upload/dip_download (the reference's
 *    SYCL backend waits after every call, sycl:335; dip_sync() gives that).
 *  - one dip_ctx per (host thread, GPU); a ctx is not thread-safe.
 *  - there is no CPU fallback: without a CUDA device dip_init fails.
 *  - pixel type: 8-bit unsigned, like every backend of the reference (cv.imread's CV_8U, sycl uint8_t*).  16-bit
 *    unsigned images (last item) are a second sample type: dip_image_create16 makes a stack of
 *    them (little-endian u16 samples, as a numpy / cv::Mat CV_16U array holds them) and every operation below accepts it
 *    with the semantics OpenCV gives CV_16U (the same formulas on the wider type; src and dst must have the same
 *    depth).  They are not benchmark kernels: stencils run on the generic tiled / direct kernels, op chains apply
 *    their operations one by one, and the host pipelines, JPEG, peer rows and dip_image_wrap are 8-bit only
 *    (DIP_ERR_UNSUPPORTED / DIP_ERR_SHAPE).
 *  - masks / kernels: the benchmark's seven run on dedicated kernels; any other mask / kernel of up to 7 x 7 taps on a
 *    tiled generic kernel, up to 15 x 15 (225 taps) on a direct one; larger ones return DIP_ERR_UNSUPPORTED.
 */
#ifndef DIP_H_
#define DIP_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define DIP_API
#else
#define DIP_API __attribute__((visibility("default")))
#endif

typedef struct dip_ctx dip_ctx;
typedef struct dip_image dip_image;

typedef enum dip_status {
    DIP_OK = 0,
    DIP_ERR_ARG = -1,         /* null / out-of-range argument                         */
    DIP_ERR_SHAPE = -2,       /* src/dst geometry mismatch, or in-place stencil       */
    DIP_ERR_UNSUPPORTED = -3, /* valid request this build does not implement          */
    DIP_ERR_CUDA = -4,        /* CUDA runtime/driver error (message has the detail)   */
    DIP_ERR_NOMEM = -5,       /* host or device allocation failed                     */
    DIP_ERR_NO_DEVICE = -6    /* no usable CUDA device: there is no CPU fallback      */
} dip_status;

typedef struct dip_image_info {
    int width, height, channels, frames;
    size_t pitch;        /* bytes between rows                */
    size_t frame_stride; /* bytes between frames of the stack */
    void* device_ptr;    /* first byte of frame 0             */
    int device;
} dip_image_info;

typedef struct dip_device_info {
    char name[128];
    int device, sm_count, cc_major, cc_minor;
    size_t total_mem, l2_bytes;
} dip_device_info;

/* ---- lifecycle (sycl::queue creation sycl:428, vglClInit vgl:86) ------------------ */
DIP_API const char* dip_version(void);
DIP_API int dip_device_count(int* count);
DIP_API int dip_init(dip_ctx** ctx, int device);
/* as dip_init, but compute work is enqueued on the caller's cudaStream_t */
DIP_API int dip_init_on_stream(dip_ctx** ctx, int device, void* cuda_stream);
DIP_API int dip_shutdown(dip_ctx* ctx);
DIP_API int dip_sync(dip_ctx* ctx); /* q.wait() sycl:335 / vglClFlush vgl:24 */
DIP_API const char* dip_last_error_string(const dip_ctx* ctx); /* ctx may be NULL */
DIP_API int dip_get_device_info(dip_ctx* ctx, dip_device_info* info);

/* ---- images (sycl::malloc_device sycl:260-262, vglCreateImage vgl:92-93,
 *      cv.UMat(cpu_image) opencv:42-44) ------------------------------------------- */
DIP_API int dip_image_create(dip_ctx* ctx, int width, int height, int channels, dip_image** img);
DIP_API int dip_image_create_batch(dip_ctx* ctx, int width, int height, int channels, int frames,
                                   dip_image** img);
/* borrow caller-owned device memory (16-byte aligned base, pitch % 16 == 0) */
DIP_API int dip_image_wrap(dip_ctx* ctx, void* device_ptr, int width, int height, int channels,
                           int frames, size_t pitch, size_t frame_stride, dip_image** img);
/* a view of frames [first, first+count) of a stack; shares memory with the parent */
DIP_API int dip_image_view(dip_ctx* ctx, dip_image* parent, int first, int count, dip_image** view);
/* a view of rows [row0, row0+rows) of every frame (row bands of a mosaic) */
DIP_API int dip_image_rows(dip_ctx* ctx, dip_image* parent, int row0, int rows, dip_image** view);
DIP_API int dip_image_destroy(dip_image* img);
DIP_API int dip_image_get_info(const dip_image* img, dip_image_info* info);
/* a stack of `frames` images of 16-bit unsigned samples (pitch = W*C*2 rounded up to 16); dip_image_depth: 8 or 16 */
DIP_API int dip_image_create16(dip_ctx* ctx, int width, int height, int channels, int frames, dip_image** img);
DIP_API int dip_image_depth(const dip_image* img, int* bits_per_sample);

/* ---- pinned host buffers for the transfer path -------------------------------- */
DIP_API int dip_host_alloc(void** ptr, size_t bytes);
DIP_API int dip_host_free(void* ptr);
DIP_API int dip_host_register(void* ptr, size_t bytes);
DIP_API int dip_host_unregister(void* ptr);

/* ---- memory operations -------------------------------------------------------- */
/* "Upload": cv.UMat(cpu_image) opencv:92, q.memcpy(d_input, image.data) sycl:331,
 * vglClUpload vgl:119.  host = frames x H rows of W*C bytes, rows host_pitch apart
 * (0 = tightly packed), frames host_pitch*H apart.  Pinned host memory (dip_host_alloc / dip_host_register) is one
 * asynchronous copy at the PCIe rate (55-57 GB/s); pageable memory goes through two 8 MiB pinned staging slots filled by
 * a small pool of host threads (38-44 GB/s; env DIP_COPY_THREADS / DIP_COPY_SPIN_US tune the pool). */
DIP_API int dip_upload(dip_ctx* ctx, dip_image* dst, const uint8_t* host, size_t host_pitch);
/* "Download": image.get() opencv:93, sycl:332, vglClDownload vgl:120 */
DIP_API int dip_download(dip_ctx* ctx, const dip_image* src, uint8_t* host, size_t host_pitch);
/* "Copy": cv.copyTo opencv:94 (intended dst = src), q.memcpy(d_output, d_input) sycl:333,
 * vglClCopy vgl:121 */
DIP_API int dip_copy(dip_ctx* ctx, const dip_image* src, dip_image* dst);

/* ---- JPEG in and out on the GPU side of upload / download ----------------------------------
 * The reference reads its input with cv.imread (opencv:14, sycl:445) and writes every result with imwrite (opencv:114,
 * sycl:374).  Here only the compressed bytes cross PCIe: decoding lands the pixels directly in a frame of a device
 * stack, encoding reads one.  The codec is nvJPEG, bound at run time (DIP_ERR_UNSUPPORTED when libnvjpeg is absent; env
 * DIP_NVJPEG_LIB names another copy); it is library code around the hot path, not one of this library's kernels, and
 * JPEG being lossy, a decode agrees with cv.imread's libjpeg-turbo only to within rounding of the inverse DCT and
 * the chroma upsampling filter (tests compare by PSNR).
 *   decode: 3-channel images receive interleaved BGR like cv.imread, 1-channel images the luma plane; the stream's size
 *           must equal the image's.  Ordered on the compute stream on both sides (it runs after what was enqueued
 *           before, later dip_* calls see the pixels); the host bytes are free again when the call returns.  Huffman
 *           decoding runs on the GPU (nvJPEG's GPU_HYBRID decoder; CPU Huffman for streams it refuses: progressive).
 *   decode_batch: `count` streams into frames [first_frame, first_frame + count), decoded by up to
 *           min(count, host cores, 32) host threads (env DIP_JPEG_THREADS), each with its own decoder state and CUDA
 *           stream: the per-image host phase (a scan of the bitstream) is what limits a decode, not the GPU.
 *   encode: frame `frame` of a 3-channel image, quality 1..100, 4:2:0 chroma (imwrite's defaults are quality 95,
 *           4:2:0).  *len = the stream's size; DIP_ERR_NOMEM (with *len set) when `capacity` is too small.
 *           Synchronous: the bytes are in `out` on return.
 *   backend: where the last decode call's Huffman decoding ran: "gpu-huffman", "cpu-huffman", "mixed" or "none". */
DIP_API int dip_jpeg_info(dip_ctx* ctx, const uint8_t* data, size_t len, int* width, int* height, int* components);
DIP_API int dip_jpeg_decode(dip_ctx* ctx, const uint8_t* data, size_t len, dip_image* dst, int frame);
DIP_API int dip_jpeg_decode_batch(dip_ctx* ctx, const uint8_t* const* data, const size_t* len, int count,
                                  dip_image* dst, int first_frame);
DIP_API int dip_jpeg_encode(dip_ctx* ctx, const dip_image* src, int frame, int quality, uint8_t* out,
                            size_t capacity, size_t* len);
DIP_API const char* dip_jpeg_backend(const dip_ctx* ctx);

/* ---- point operations ------------------------------------------------------------ */
/* "Inversion": cv.bitwise_not opencv:95, InvertKernel sycl:54-69, vglClInvert vgl:122 */
DIP_API int dip_invert(dip_ctx* ctx, const dip_image* src, dip_image* dst);
/* "Grayscale": cvtColor BGR2GRAY -> GRAY2BGR opencv:96 (BT.601 15-bit fixed point, exact),
 * GrayscaleKernel sycl:71-92, vglClGrayscale vgl:123.  4 channels: alpha := 255; 1: copy */
DIP_API int dip_grayscale(dip_ctx* ctx, const dip_image* src, dip_image* dst);
/* "Threshold": cv.threshold(image,127,255,THRESH_BINARY) opencv:97,
 * ThresholdKernel(...,127,255) sycl:94-114,:337, vglClThreshold vgl:124:
 * dst = src > thresh ? maxval : 0 on every byte */
DIP_API int dip_threshold(dip_ctx* ctx, const dip_image* src, dip_image* dst, uint8_t thresh,
                          uint8_t maxval);
/* the same on 16-bit images with 16-bit arguments (dip_threshold on a 16-bit image compares with its 8-bit numbers) */
DIP_API int dip_threshold16(dip_ctx* ctx, const dip_image* src, dip_image* dst, uint16_t thresh,
                            uint16_t maxval);

/* ---- stencil operations ---------------------------------------------------------- */
/* "Erosion": cv.erode(image, mask) opencv:98-99 (and the 1x3 / 3x1 calls :77-78),
 * ErosionKernel(in,out,w,h,c,mask,mw,mh) sycl:116-157,:339-343, vglClErode(img,out,mask,w,h)
 * vgl:125-127.  mask = mask_h x mask_w row-major bytes (non-zero = tap), anchor = centre,
 * out-of-image taps ignored (OpenCV's +inf morphology border). */
DIP_API int dip_erode(dip_ctx* ctx, const dip_image* src, dip_image* dst, const uint8_t* mask,
                      int mask_w, int mask_h);
/* cv.dilate, defined but unused by the reference (opencv:80-82) */
DIP_API int dip_dilate(dip_ctx* ctx, const dip_image* src, dip_image* dst, const uint8_t* mask,
                       int mask_w, int mask_h);
/* "Erosion (1x3+3x1)": erosion_separated opencv:76-78,:100, sycl:341-344 -- the two passes
 * fused in one kernel (row minimum then column minimum through shared memory; no `aux`
 * round trip).  All-ones 1 x mask_w then mask_h x 1. */
DIP_API int dip_erode_separable(dip_ctx* ctx, const dip_image* src, dip_image* dst, int mask_w,
                                int mask_h);
/* "Convolution": cv.filter2D(image,-1,kernel) opencv:101,103 (and :85-86,:89-90),
 * ConvolutionKernel(in,out,w,h,c,kernel,kw,kh) sycl:159-202,:346-356,
 * vglClConvolution(img,out,kernel,w,h) vgl:128-131.  kernel = kh x kw row-major float,
 * correlation, anchor = centre, BORDER_REFLECT_101, result rounded half-to-even and
 * saturated.  Kernels whose weights are k/2^s (the benchmark's Gaussians) take the exact
 * packed-integer path; anything else an fp32 path (within 1 LSB of OpenCV). */
DIP_API int dip_convolve(dip_ctx* ctx, const dip_image* src, dip_image* dst, const float* kernel,
                         int kw, int kh);
/* "Convolution (1xK+Kx1)": convolution_*_separated opencv:84-90,:102,:104, sycl:347-356 --
 * row pass (1 x kw) rounded to u8 exactly as the reference's `aux` image is, then column
 * pass (kh x 1).  The benchmark's binomial pairs ([1 2 1]/4, [1 4 6 4 1]/16) are ONE fused kernel; any other pair
 * runs as the reference's two calls, with a scratch stack owned by the ctx as `aux` (allocated on first use). */
DIP_API int dip_convolve_separable(dip_ctx* ctx, const dip_image* src, dip_image* dst,
                                   const float* krow, int kw, const float* kcol, int kh);
/* "Gaussian Blur (3x3)": cv.GaussianBlur(image,(3,3),0) opencv:105,
 * GaussianBlur3x3Kernel sycl:204-252,:358, vglClBlurSq3 vgl:132 -- OpenCV's bit-exact
 * fixed-point blur: [1 2 1]x[1 2 1]/16 rounded half-up, BORDER_REFLECT_101. */
DIP_API int dip_gaussian_blur3x3(dip_ctx* ctx, const dip_image* src, dip_image* dst);

/* Fused chain (SURVEY.md 8f rank 4 / BASELINE.json config 5): dst = erode3x3_square(filter2D(src, 5x5 Gaussian))
 * in one kernel -- each pixel crosses HBM once in and once out.  Same result, bit for bit, as
 * dip_convolve(5x5 Gaussian) followed by dip_erode(3x3 ones) (opencv:103 then opencv:99). */
DIP_API int dip_gauss5x5_erode3x3(dip_ctx* ctx, const dip_image* src, dip_image* dst);

/* ---- the reference's operation list as data (opencv:92-105, sycl:331-358) ---------- */
typedef enum dip_op_kind {
    DIP_OP_COPY = 0,
    DIP_OP_INVERT = 1,
    DIP_OP_GRAYSCALE = 2,
    DIP_OP_THRESHOLD = 3,
    DIP_OP_ERODE = 4,
    DIP_OP_ERODE_SEPARABLE = 5,
    DIP_OP_CONVOLVE = 6,
    DIP_OP_CONVOLVE_SEPARABLE = 7,
    DIP_OP_GAUSSIAN_BLUR3X3 = 8,
    DIP_OP_DILATE = 9,
    DIP_OP_GAUSS5X5_ERODE3X3 = 10
} dip_op_kind;

typedef struct dip_op {
    int kind;              /* dip_op_kind */
    uint8_t thresh, maxval;
    const uint8_t* mask;   /* erode / dilate */
    int mask_w, mask_h;    /* also the all-ones extents of ERODE_SEPARABLE */
    const float* kernel;   /* convolve, or the row kernel of CONVOLVE_SEPARABLE */
    int kw, kh;
    const float* kernel_col; /* column kernel of CONVOLVE_SEPARABLE */
    uint16_t thresh16, maxval16; /* THRESHOLD on 16-bit images; both 0: use thresh / maxval */
} dip_op;

/* dst = op(src) for one entry of an operation list */
DIP_API int dip_apply(dip_ctx* ctx, const dip_op* op, const dip_image* src, dip_image* dst);

/* Op chain: dst = ops[n-1]( ... ops[1]( ops[0](src) ) ), bit for bit what n dip_apply calls
 * through intermediate images give -- the reference's separable pairs through `aux` (opencv:76-90) are its two-op
 * instance -- in as few passes over HBM as the chain allows:
 *   - runs of byte-wise ops (copy, inversion, threshold) collapse into ONE per-byte map (identity, 255 - x or a step);
 *   - that map is fused into the neighbouring tiled stencil kernel (applied to the result, or to the source words, in
 *     registers; in front of an erosion / dilation it moves behind it, min / max commuting with monotone maps) or
 *     into the grayscale kernel, so it costs no pass of its own (next to a direct kernel for an arbitrary mask, or
 *     a one-dimensional pass, it becomes one point pass instead);
 *   - 5x5 Gaussian followed directly by 3x3 square erosion runs as the fused dip_gauss5x5_erode3x3 kernel (which takes
 *     the maps on either side as well);
 *   - every other stencil, and every grayscale, is one pass.
 * e.g. grayscale -> GaussianBlur -> threshold -> erode: 3 passes (gray | blur+threshold | erode) instead of 4.
 * Intermediate results live in one scratch stack owned by the ctx (allocated on first use, same shape as src) and in
 * dst.  src and dst must not overlap unless the whole chain is byte-wise / grayscale. */
DIP_API int dip_apply_chain(dip_ctx* ctx, const dip_op* ops, int n, const dip_image* src, dip_image* dst);
/* number of kernel launches (= passes over HBM) dip_apply_chain makes for this chain on `channels`-channel images */
DIP_API int dip_chain_passes(const dip_op* ops, int n, int channels, int* passes);
/* The single per-byte map a run of byte-wise ops (copy, inversion, threshold only) folds into, tabulated:
 * lut[x] = ops[n-1](...ops[0](x)).  Host-only (no GPU needed); this is the map the kernels evaluate. */
DIP_API int dip_chain_lut(const dip_op* ops, int n, uint8_t lut[256]);

/* Fill `ops` (capacity >= 12) with the reference's 12 device operations in its order:
 * copy, inversion, grayscale, threshold(127,255), erosion cross / square / 1x3+3x1,
 * convolution 3x3 / 1x3+3x1 / 5x5 / 1x5+5x1 Gaussian, GaussianBlur 3x3.  Mask and kernel
 * pointers refer to static storage inside the library.  Returns the count (12). */
DIP_API int dip_reference_ops(dip_op* ops, int capacity);
/* the reference's file prefix of entry i ("inversion", "erosion-cross", ...) opencv:94-105 */
DIP_API const char* dip_reference_op_prefix(int i);
/* the reference's row description of entry i ("Inversion", "Erosion (3x3 Cross Kernel)", ...) */
DIP_API const char* dip_reference_op_description(int i);

/* ---- batched host pipeline (the measure_time/perform_benchmark loop, opencv:109-114,
 *      sycl:365-375, taken end to end): for each of `frames` host images upload once,
 *      run every op of the list input -> output, download every result.  Uploads,
 *      kernels and downloads of consecutive chunks overlap on three streams through
 *      pinned staging.  host_out[i] receives op i's `frames` results. --------------- */
DIP_API int dip_run_ops_host(dip_ctx* ctx, const uint8_t* host_in, int width, int height,
                             int channels, int frames, const dip_op* ops, int n_ops,
                             uint8_t* const* host_out, int chunk_frames);

/* The same pipeline with its device buffers and events kept across calls, and the host side
 * addressed as rings (a streaming source and sink): frame f is read from
 * host_in[f % in_ring_frames] and op i's result is written to host_out[i][f % out_ring_frames].
 * Both ring sizes must be multiples of chunk_frames.  Returns after the last download. */
typedef struct dip_pipeline dip_pipeline;
DIP_API int dip_pipeline_create(dip_ctx* ctx, int width, int height, int channels, int chunk_frames,
                                const dip_op* ops, int n_ops, dip_pipeline** pipe);
DIP_API int dip_pipeline_run(dip_pipeline* pipe, const uint8_t* host_in, int in_ring_frames, int frames,
                             uint8_t* const* host_out, int out_ring_frames);
DIP_API int dip_pipeline_destroy(dip_pipeline* pipe);
/* Chain form: the ops are applied IN SEQUENCE (dip_apply_chain) and only the final result comes back, so a frame
 * crosses PCIe once in each direction: host_out[0] receives the `frames` results.  Same rings, same overlap. */
DIP_API int dip_pipeline_create_chain(dip_ctx* ctx, int width, int height, int channels, int chunk_frames,
                                      const dip_op* ops, int n_ops, dip_pipeline** pipe);
DIP_API int dip_run_chain_host(dip_ctx* ctx, const uint8_t* host_in, int width, int height, int channels, int frames,
                               const dip_op* ops, int n_ops, uint8_t* host_out, int chunk_frames);

/* ---- CUDA-graph replay (SURVEY.md 8f rank 4 "graph capture"): record what the calls between dip_graph_begin and
 *      dip_graph_end enqueue on the compute stream -- any dip_* operation, chain or row-band call; not transfers, not
 *      dip_sync -- and replay it with ONE launch per dip_graph_launch.  The reference's measure_time loops
 *      (opencv:33-35, sycl:33-37) enqueue thousands of identical few-microsecond kernels; recorded in blocks, their host
 *      cost per call drops to the graph launch divided by the block size.  The graph's edges are the data dependencies
 *      between the recorded kernels (overlapping byte ranges of their operands), not the order of the calls, so
 *      independent operations overlap on replay.  The images used while recording must stay
 *      alive (and keep their contents' meaning) for every replay; a chain's scratch stack must exist (run it once). */
typedef struct dip_graph dip_graph;
DIP_API int dip_graph_begin(dip_ctx* ctx);
DIP_API int dip_graph_end(dip_ctx* ctx, dip_graph** graph);
DIP_API int dip_graph_launch(dip_ctx* ctx, dip_graph* graph);
DIP_API int dip_graph_destroy(dip_graph* graph);

/* ---- device-side timing on the ctx's compute stream (measure_time opencv:26-38,
 *      sycl:25-39 use a host clock; these are CUDA events) --------------------------- */
DIP_API int dip_timer_start(dip_ctx* ctx);
DIP_API int dip_timer_stop(dip_ctx* ctx, float* elapsed_ms); /* synchronises */
/* lap markers: dip_mark records CUDA event `slot` on the compute stream (slots grow on demand);
 * dip_elapsed waits for slot_b and returns the device time between two recorded slots */
DIP_API int dip_mark(dip_ctx* ctx, int slot);
DIP_API int dip_elapsed(dip_ctx* ctx, int slot_a, int slot_b, float* elapsed_ms);
/* number of kernels this ctx has launched since dip_init (bench.py's gpu_launches) */
DIP_API int dip_launch_count(const dip_ctx* ctx, uint64_t* launches);
/* write `bytes` of scratch device memory to evict L2 between timed iterations */
DIP_API int dip_flush_l2(dip_ctx* ctx);

/* ---- synthetic inputs and checksums (BASELINE.json configs 4 and 5) ---------------- */
/* byte i of frame f := hash32(global byte index, seed) >> 24; frame numbering starts at
 * first_frame so that shards of one batch are generated independently per GPU */
DIP_API int dip_synth_fill(dip_ctx* ctx, dip_image* img, uint32_t seed, uint64_t first_frame);
/* the same generator addressed by global row: byte x of row y of frame f := hash32(((first_row +
 * f*height + y) * row_bytes + x), seed) >> 24 -- a row band of a tall mosaic fills only its own rows */
DIP_API int dip_synth_fill_rows(dip_ctx* ctx, dip_image* img, uint32_t seed, uint64_t first_row);
/* position-weighted 64-bit checksum of every frame's valid bytes -> host array [frames] */
DIP_API int dip_checksum(dip_ctx* ctx, const dip_image* img, uint64_t* per_frame);
/* as dip_checksum with row y weighted as global row first_row + y: the checksums of the row bands
 * of a mosaic add up (mod 2^64) to the checksum of the whole mosaic */
DIP_API int dip_checksum_rows(dip_ctx* ctx, const dip_image* img, uint64_t first_row, uint64_t* per_frame);

/* ---- multi-GPU mosaic row bands (BASELINE.json config 5) ---------------------------
 * A band holds rows [row0, row0+rows) of a tall single-frame mosaic plus `halo` rows
 * above and below.  Interior bands get their halo rows from the neighbour GPUs' bands
 * over NVLink (peer copies between IPC-mapped allocations); the first and last band
 * use the operation's own border rule (reflect-101 / +inf). */
typedef struct dip_ipc_handle { unsigned char bytes[64]; } dip_ipc_handle;
DIP_API int dip_image_ipc_export(dip_ctx* ctx, const dip_image* img, dip_ipc_handle* handle);
DIP_API int dip_image_ipc_import(dip_ctx* ctx, const dip_ipc_handle* handle, int width, int height,
                                 int channels, int frames, size_t pitch, size_t frame_stride,
                                 dip_image** img);
/* copy `rows` rows starting at src_row of src into dst starting at dst_row (same width and
 * channels; src may live on a peer GPU).  Stream-ordered on the ctx's halo stream. */
DIP_API int dip_copy_rows(dip_ctx* ctx, const dip_image* src, int src_row, dip_image* dst,
                          int dst_row, int rows);
/* make the compute stream wait for all halo copies issued so far */
DIP_API int dip_halo_fence(dip_ctx* ctx);
/* dst rows [dst_row0, dst_row0+out_rows) = op(src) rows [out_row0, out_row0+out_rows), where
 * only src rows [valid_row0, valid_row0+valid_rows) exist for border purposes (rows outside
 * follow the op's border rule: reflect-101 / +inf).  A band with neighbours above and below
 * passes its halo rows as valid; the first / last band of a mosaic exclude them.  Point
 * operations ignore the valid range. */
DIP_API int dip_apply_rows(dip_ctx* ctx, const dip_op* op, const dip_image* src, int valid_row0,
                           int valid_rows, int out_row0, int out_rows, dip_image* dst,
                           int dst_row0);

/* The same as dip_apply_rows, with the halo rows read from the neighbour GPUs INSIDE the kernel (SURVEY.md 8e
 * "preferred, fused"): source rows [src_row0, src_row0+rows) are not taken from src but from rows [image_row0, ...)
 * of `image`, the neighbour's band buffer mapped with dip_image_ipc_import (same width, channels and pitch as src).
 * Only the CTAs whose tile covers such rows touch the peer (16-byte loads over NVLink into their shared-memory tile);
 * there is no exchange step, no copy into src's halo rows and no extra pass over HBM.  Tiled stencil kernels only
 * (the benchmark's masks / kernels, the 1-D passes, dip_gauss5x5_erode3x3); DIP_ERR_UNSUPPORTED otherwise. */
typedef struct dip_peer_rows {
    const dip_image* image; /* NULL: no neighbour on this side */
    int src_row0;           /* first row of src served by the neighbour */
    int image_row0;         /* the row of `image` that is src row src_row0 */
    int rows;
} dip_peer_rows;
DIP_API int dip_apply_rows_peer(dip_ctx* ctx, const dip_op* op, const dip_image* src, int valid_row0,
                                int valid_rows, int out_row0, int out_rows, dip_image* dst, int dst_row0,
                                const dip_peer_rows* above, const dip_peer_rows* below);

/* Step flags for bands whose input changes from step to step (each rank's result becomes its next input).  A flag
 * array is a tiny library-owned image (`count` 32-bit words, zeroed) that neighbours map with dip_image_ipc_export /
 * _import (width 4*count, height 1, 1 channel).
 *   dip_flag_signal: after everything enqueued so far on the compute stream has finished, store `value` into word
 *                    `index` (release, system scope); `flags` may be the neighbour's mapped array (a store over NVLink).
 *                    A stream memory operation (cuStreamWriteValue32) where the driver accepts the address, else a
 *                    one-thread kernel.
 *   dip_flag_wait:   the compute stream waits until word `index` >= value (cuStreamWaitValue32 -- a stream operation,
 *                    no kernel spins); `flags` must live on this ctx's GPU.
 *   dip_flag_read:   host read-back (synchronous), for tests.
 * Protocol of MosaicBand (mosaic.py): step k reads buffer k%2 and writes (k+1)%2; before it a rank waits for
 * flag[neighbour] >= k-1 (the neighbour finished step k-1: its rows of buffer k%2 are written, and it no longer reads
 * this rank's buffer (k+1)%2); after it, it signals k into both neighbours' arrays. */
DIP_API int dip_flags_create(dip_ctx* ctx, int count, dip_image** flags);
DIP_API int dip_flag_signal(dip_ctx* ctx, dip_image* flags, int index, uint32_t value);
DIP_API int dip_flag_wait(dip_ctx* ctx, const dip_image* flags, int index, uint32_t value);
DIP_API int dip_flag_read(dip_ctx* ctx, const dip_image* flags, int index, uint32_t* value);

/* ---- development hook: kernel-selection knobs for A/B measurements inside one process.
 * keys: "tile_rows" (stencil tile height: 0 = automatic, >= 60 the stack tiles, < 60 the single-image tiles),
 * "no_pad" (1 = do not pad shared memory to limit the resident tiles of the HBM-bound stencils), "int_math" (1 = the
 * packed-integer forms of the separable Gaussians instead of the packed-half ones), "umma" (1 = the fused 1x5+5x1
 * Gaussian on the tensor cores), "direct_generic" (1 = small arbitrary masks / kernels on the direct kernel),
 * "float_morph" (1 = arbitrary masks on the fp32 tiled kernel instead of packed minima), "band_shape" (rows x warps of
 * the 5-tap kernels on RGB stacks: 0 = shipped, 7 = 15 x 4, 1..6 = other shapes), "fused_shape" (band shape / unrolling
 * of the fused Gaussian -> erosion kernel on one-channel images: 0 = shipped, 14 = 15 x 4).  Not needed by applications. */
DIP_API int dip_debug_set(const char* key, int value);

#ifdef __cplusplus
}
#endif
#endif /* DIP_H_ */
