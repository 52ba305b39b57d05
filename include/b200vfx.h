/*
 * b200vfx.h -- C ABI of libb200vfx.so: the B200 (sm_100a) implementation of the
 * per-pixel video-filter hot path of sdroege/gst-plugin-rs.
 *
 * This is the drop-in boundary: a gstreamer-rs BaseTransform / VideoFilter /
 * VideoAggregator subclass keeps the reference's element surface (factory
 * names, caps, GObject properties) and replaces the body of its
 * transform_frame* / aggregate_frames vfunc with ONE call below.  Every entry
 * point cites the reference code it replaces (paths relative to the
 * gst-plugins-rs tree).  INTEGRATION.md shows the Rust `extern "C"` binding.
 *
 * Conventions
 *  - plain pointers and sizes only; frames are {format, width, height, data,
 *    stride} exactly as a mapped GstVideoFrame exposes them (plane 0).
 *  - `src`/`dst`/`data` may be HOST pointers (pageable or pinned) or DEVICE
 *    pointers; the library detects which (cudaPointerGetAttributes).
 *      host   : synchronous call; H2D copy, kernel and D2H copy are pipelined
 *               by row chunks on internal streams; on return `dst` is filled.
 *      device : the kernel is enqueued on the context's stream
 *               (b200vfx_ctx_set_stream) and the call returns immediately;
 *               use b200vfx_ctx_synchronize() or your own stream sync.
 *  - return 0 on success, <0 on error (maps to gst::FlowError::Error /
 *    NotNegotiated); b200vfx_last_error() gives the message.
 *  - one context per element instance; a context is not re-entrant (GStreamer
 *    serialises transform calls per element with the pad stream lock).
 *  - there is NO CPU fallback: without a CUDA device every call fails.
 */
#ifndef B200VFX_H
#define B200VFX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200VFX_ABI_VERSION 1

/* video/x-raw formats on this path (names follow GstVideoFormat) */
typedef enum {
  B200VFX_FORMAT_RGBX = 0,
  B200VFX_FORMAT_XRGB = 1,
  B200VFX_FORMAT_BGRX = 2,
  B200VFX_FORMAT_XBGR = 3,
  B200VFX_FORMAT_RGBA = 4,
  B200VFX_FORMAT_ARGB = 5,
  B200VFX_FORMAT_BGRA = 6,
  B200VFX_FORMAT_ABGR = 7,
  B200VFX_FORMAT_RGB = 8,
  B200VFX_FORMAT_BGR = 9,
  B200VFX_FORMAT_RGBA64_LE = 10,
  B200VFX_FORMAT_RGBA64_BE = 11,
  B200VFX_FORMAT_I420 = 12,
  B200VFX_FORMAT_A420 = 13
} b200vfx_format;

typedef enum {
  B200VFX_OK = 0,
  B200VFX_ERR_INVALID = -1,      /* bad argument                 -> FlowError::Error         */
  B200VFX_ERR_CUDA = -2,         /* CUDA runtime / no device     -> FlowError::Error         */
  B200VFX_ERR_NOT_NEGOTIATED = -3, /* e.g. colorlut without a LUT -> FlowError::Error (imp.rs:210-213) */
  B200VFX_ERR_UNSUPPORTED = -4,  /* format not on this element's caps                        */
  B200VFX_ERR_PARSE = -5,        /* .cube InvalidLut             -> ResourceError::Read      */
  B200VFX_ERR_IO = -6            /* .cube Io error               -> ResourceError::Read      */
} b200vfx_status;

typedef struct b200vfx_ctx b200vfx_ctx;

/* ---- library / context -------------------------------------------------- */
int b200vfx_abi_version(void);
int b200vfx_device_count(void); /* <=0: no usable CUDA device */

/* Created in BaseTransformImpl::start() (or first set_caps), destroyed in stop().
 * device < 0 selects the current CUDA device.
 * Threading: a context belongs to one streaming thread at a time, like the GstBaseTransform instance that owns it (the
 * vfuncs of one element run under its stream lock).  Different contexts may be used from different threads concurrently;
 * calls on ONE context must be serialised by the caller.  b200vfx_fence_wait / _query / _destroy may be called from any
 * thread (e.g. a downstream element mapping the buffer). */
int b200vfx_ctx_create(b200vfx_ctx **out, int device);
void b200vfx_ctx_destroy(b200vfx_ctx *ctx);
const char *b200vfx_last_error(const b200vfx_ctx *ctx); /* ctx may be NULL: last create/parse error of this thread */

/* stream used for DEVICE-pointer calls (a cudaStream_t / CUstream handle; NULL = legacy default stream).
 * Default: a non-blocking stream owned by the context. */
int b200vfx_ctx_set_stream(b200vfx_ctx *ctx, void *cuda_stream);
int b200vfx_ctx_synchronize(b200vfx_ctx *ctx);
/* Asynchronous host-frame mode.  By default a call on HOST frames returns when its output is in host memory (what
 * GstBaseTransform's transform_frame promises), so the upload of frame i+1 can never overlap the download of frame i: with
 * the default four chunks per frame the two PCIe directions idle a fifth of the time.  With the mode enabled, the
 * per-pixel entry points on host frames (colorlut, colorlut_fmt, hsvfilter, hsvdetector, convert_packed) return as soon as
 * their copies and kernels are enqueued; b200vfx_fence_create marks "everything submitted so far", and the caller must not
 * read the output frames, or reuse the input frames, before that fence has been waited on (or b200vfx_ctx_synchronize).
 * This is the contract of the reference's own GPU element: D3D12ColorLut::transform submits its command list, stamps the
 * output memories with the queue's fence value and returns (video/colorlut/src/d3d12colorlut/imp.rs:698-718; input fences
 * are waited for on the queue, :579-602, :694-696) -- whoever maps the memory waits.  Here the pinned GstMemory handed
 * downstream carries the fence and its map() waits (rust-shim/src/allocator.rs).  Frames must be pinned (b200vfx_host_alloc) to be copied asynchronously
 * at all.  Reductions and hashes (blockhash, colordetect, hash_image) stay synchronous: they return values. */
typedef struct b200vfx_fence b200vfx_fence;
int b200vfx_ctx_set_host_async(b200vfx_ctx *ctx, int enable);
int b200vfx_fence_create(b200vfx_ctx *ctx, b200vfx_fence **fence_out);
int b200vfx_fence_wait(b200vfx_fence *fence);          /* blocks the calling thread */
int b200vfx_fence_query(b200vfx_fence *fence);         /* 1 = reached, 0 = pending, < 0 = error */
void b200vfx_fence_destroy(b200vfx_fence *fence);
/* rows per H2D/kernel/D2H pipeline chunk for HOST-pointer calls (0 = auto) */
int b200vfx_ctx_set_chunk_rows(b200vfx_ctx *ctx, int rows);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t b200vfx_ctx_kernel_launches(const b200vfx_ctx *ctx);
/* kernel-variant knobs for A/B measurements (results are identical for every setting):
 * "stream_path" -1 auto | 0 | 1 (TMA-pipelined streaming kernels), "stream_cfg" 0..7 (tile/stage/thread variant),
 * "pdl" 0|1 (overlap consecutive independent frames by programmatic dependent launch; hazards are detected),
 * "stream_ctas" CTAs per SM (0 = occupancy maximum), "stream_hint" 0|1 (L2 evict_last policy on table gathers),
 * "memo_px" 4|8|16 (pixels per thread of the non-TMA kernel),
 * "hsv_memo" -1 auto | 0 never | 1 at once (settings-keyed answer tables of hsvfilter / hsvdetector; auto builds
 * the table after the same settings have processed 2^24 pixels),
 * "cd_cluster" 1|2|4|8 (colordetect: CTAs per cluster that merge their shared-memory histograms over DSMEM),
 * "cd_split" 0 auto | 1 | 2..8 (colordetect with a DEVICE histogram: fraction 1/n of the SMs per launch, so that the launches
 * of a train run beside each other),
 * "memo_ctas" 2..8 (CTAs per SM of the persistent table-lookup kernels; 4 = two consecutive frames resident together),
 * "memo_tile" 0|1 (4-byte-pixel table lookups through a per-tile shared-memory copy of the colour sub-cube; wins on
 * medium-noise content only, profiles/r01_memo_tile_experiment.jsonl),
 * "rgba64_x4" 0|1 (RGBA64 3D-LUT kernel that keeps the LUT cell of four neighbouring pixels in registers),
 * "blockhash_rows" 0..8 (videocompare block sums: 0 = one CTA per hash block, n > 0 = whole-row streaming CTAs,
 * 2 * SMs / n of them; default 2 = one per SM), "blockhash_tma" 0|1 (TMA-fed block-sum tiles),
 * "tile_gather_path" 0|1 (fused colorlut + all-gather: register stores | TMA bulk stores), "tile_gather_cfg" (path 0:
 * 1 = 32-byte stores; path 1: tile size 16/8/4/32 KB), "tile_gather_ctas" CTAs per SM, "peer_timeout_ms". */
int b200vfx_ctx_set_option(b200vfx_ctx *ctx, const char *name, int value);

/* Page-locked host memory for a GstAllocator handed out in
 * propose_allocation()/decide_allocation(): buffers allocated here are copied
 * with full-speed asynchronous DMA. */
void *b200vfx_host_alloc(size_t bytes);
void b200vfx_host_free(void *p);

/* ---- device-resident frames (chaining without PCIe round trips) ----------------------
 * What a `memory:CUDAMemory`-style GstAllocator of these elements needs: device frames that several elements
 * process back to back (every *_process call accepts device pointers and then only enqueues its kernel), one
 * upload at the head of the chain and one download at its tail.  upload/download are asynchronous on the
 * context stream (use b200vfx_ctx_synchronize, or stream order, before touching the host buffer again);
 * full PCIe rate needs page-locked host memory (b200vfx_host_alloc). */
void *b200vfx_device_alloc(b200vfx_ctx *ctx, size_t bytes);
void b200vfx_device_free(b200vfx_ctx *ctx, void *dev_ptr);
int b200vfx_upload(b200vfx_ctx *ctx, void *dev_dst, int dst_stride, const void *host_src, int src_stride,
                   size_t row_bytes, int rows);
int b200vfx_download(b200vfx_ctx *ctx, void *host_dst, int dst_stride, const void *dev_src, int src_stride,
                     size_t row_bytes, int rows);
/* plane copy between any two of {host, device} (direction inferred through UVA), asynchronous on the context stream;
 * 1 if p is device (or managed) memory, 0 otherwise */
int b200vfx_copy_plane(b200vfx_ctx *ctx, void *dst, int dst_stride, const void *src, int src_stride, size_t row_bytes,
                       int rows);
int b200vfx_pointer_is_device(const void *p);

/* ---- .cube parser -------------------------------------------------------
 * replaces CubeLut::parse / parse_file, video/colorlut/src/parser.rs:105-281
 * (same grammar, same error cases).  kind: 1 = LUT_1D, 3 = LUT_3D.
 * values: n x [r,g,b] f32 in file order, n = size (1D) or size^3 (3D, R
 * fastest); release with b200vfx_cube_free().  err may be NULL. */
int b200vfx_cube_parse(const char *text, size_t len, int *kind, int *size, float **values,
                       float domain_scale[3], float domain_offset[3], char *err, size_t errlen);
int b200vfx_cube_parse_file(const char *path, int *kind, int *size, float **values,
                            float domain_scale[3], float domain_offset[3], char *err, size_t errlen);
void b200vfx_cube_free(float *values);

/* ---- colorlut -----------------------------------------------------------
 * ColorLut::start  (video/colorlut/src/colorlut/imp.rs:168-194): parse `location` and upload. */
int b200vfx_colorlut_load_file(b200vfx_ctx *ctx, const char *location);
/* upload an already parsed LUT (State{lut}, imp.rs:50-53,191) */
int b200vfx_colorlut_set_lut(b200vfx_ctx *ctx, int kind, int size, const float *values,
                             const float domain_scale[3], const float domain_offset[3]);
/* ColorLut::stop (imp.rs:196-199) */
int b200vfx_colorlut_clear(b200vfx_ctx *ctx);
/* evaluation strategy for 8-bit RGBA: 0 = auto (memoised 2^24-entry table, L2 resident),
 * 1 = direct (trilinear evaluated per pixel).  Results are bit-identical. */
int b200vfx_colorlut_set_mode(b200vfx_ctx *ctx, int mode);
/* ColorLut::transform_frame (imp.rs:203-224) incl. transform_rgba / transform_rgba64<LE|BE>
 * 1D and 3D (imp.rs:226-397).  fmt: RGBA, RGBA64_LE, RGBA64_BE.  Never in place. */
int b200vfx_colorlut_process(b200vfx_ctx *ctx, int fmt, int width, int height, const void *src,
                             int src_stride, void *dst, int dst_stride);

/* ColorLut::transform_frame with the videoconverts that surround colorlut in real pipelines (colorlut/imp.rs:18 doc
 * pipeline) folded into the kernel's load and store: in_fmt / out_fmt are any of the eight 4-byte 8-bit RGB formats
 * (RGBx xRGB BGRx xBGR RGBA ARGB BGRA ABGR), independently.  Colour bytes go through the LUT; the 4th byte is copied
 * when both formats carry alpha and written as 255 otherwise.  Same answers as convert -> colorlut -> convert. */
int b200vfx_colorlut_process_fmt(b200vfx_ctx *ctx, int in_fmt, int out_fmt, int width, int height, const void *src,
                                 int src_stride, void *dst, int dst_stride);

/* ---- format conversion (the `videoconvert` either side of these elements, kept on the device) ----
 * b200vfx_convert_packed: any of the ten packed 8-bit RGB formats to any other -- a byte permutation, exact.  A missing
 * alpha / padding byte is written as 255.  Host or device pointers (device: asynchronous on the context stream).
 * b200vfx_convert_to_planar / _from_planar: packed RGB <-> I420 / A420 (planes Y, U, V[, A]; SURVEY App. E geometry).
 * GStreamer's converter is not part of the reference tree: the arithmetic is specified in csrc/convert.cuh (8-bit fixed
 * point, limited range, BT.601 for <= 576 lines else BT.709 when matrix == 0; 601 / 709 force one) and its parity with
 * `videoconvert` is UNPINNED.  Planes must be all host or all device memory.
 * b200vfx_a420_append: RoundedCorners::prepare_output_buffer (border/imp.rs:482-559) for device frames: I420 planes + the
 * A8 mask -> the four planes of an A420 frame (planes that are shared with the input are not copied). */
int b200vfx_convert_packed(b200vfx_ctx *ctx, int src_fmt, int dst_fmt, int width, int height, const void *src,
                           int src_stride, void *dst, int dst_stride);
int b200vfx_convert_to_planar(b200vfx_ctx *ctx, int src_fmt, int dst_fmt, int width, int height, const void *src,
                              int src_stride, void *const *planes, const int *strides, int matrix);
int b200vfx_convert_from_planar(b200vfx_ctx *ctx, int src_fmt, int dst_fmt, int width, int height,
                                const void *const *planes, const int *strides, void *dst, int dst_stride, int matrix);
int b200vfx_a420_append(b200vfx_ctx *ctx, int width, int height, const void *const *i420_planes, const int *i420_strides,
                        const void *a8, int a8_stride, void *const *out_planes, const int *out_strides);
/* ColorLut::transform_frame (colorlut/imp.rs:204-235) on an I420 / A420 frame ("planar YUV via fused
 * convert"): I420 -> RGB, ColorLut::transform_rgba (imp.rs:267-294), RGB -> I420 in ONE kernel -- 3 bytes of HBM traffic per
 * pixel instead of 19 through two RGBA intermediates.  The answer is the composition b200vfx_convert_from_planar ->
 * b200vfx_colorlut_process(RGBA) -> b200vfx_convert_to_planar, bit for bit (conversion arithmetic: csrc/convert.cuh, parity
 * with `videoconvert` unpinned); the A plane of A420 is copied.  Planes all host or all device memory; never in place. */
int b200vfx_colorlut_process_planar(b200vfx_ctx *ctx, int fmt, int width, int height, const void *const *src_planes,
                                    const int *src_strides, void *const *dst_planes, const int *dst_strides, int matrix);

/* ---- hsvfilter ----------------------------------------------------------
 * HsvFilter::transform_frame_ip + hsv_filter (video/hsv/src/hsvfilter/imp.rs:76-120,323-376).
 * In place.  Settings by value = the per-frame snapshot of imp.rs:85.
 * fmt: RGBx xRGB BGRx xBGR RGBA ARGB BGRA ABGR RGB BGR. */
int b200vfx_hsvfilter_process(b200vfx_ctx *ctx, int fmt, int width, int height, void *data,
                              int stride, float hue_shift, float saturation_mul,
                              float saturation_off, float value_mul, float value_off);

/* ---- hsvdetector --------------------------------------------------------
 * HsvDetector::transform_frame + hsv_detect (video/hsv/src/hsvdetector/imp.rs:100-160,423-707).
 * in_fmt: RGBx xRGB BGRx xBGR RGB BGR; out_fmt: RGBA ARGB BGRA ABGR. */
int b200vfx_hsvdetector_process(b200vfx_ctx *ctx, int in_fmt, int out_fmt, int width, int height,
                                const void *src, int src_stride, void *dst, int dst_stride,
                                float hue_ref, float hue_var, float saturation_ref,
                                float saturation_var, float value_ref, float value_var);

/* ---- roundedcorners -----------------------------------------------------
 * RoundedCorners::generate_alpha_mask + draw_rounded_corners
 * (video/videofx/src/border/imp.rs:57-180).  Writes stride * round_up_2(height) bytes (A8). */
int b200vfx_roundmask_generate(b200vfx_ctx *ctx, int width, int height, int stride,
                               unsigned border_radius_px, void *a8_out);

/* ---- videocompare -------------------------------------------------------
 * HasherEngine::hash_image (video/videofx/src/videocompare/hashed_image.rs:24-64,110-130 -> image_hasher 3.1.1,
 * image 0.25.10: third-party crates, restated as recalled -- bit patterns are parity-unpinned; the reference's tests pin
 * only `distance 0 for identical frames` and `> 0 for snow vs red`, tests/videocompare.rs:57-139).
 * All five values of the `hash-algo` property (GstVideoCompareHashAlgorithm, videocompare/mod.rs:57-92): */
typedef enum {
  B200VFX_HASH_MEAN = 0,
  B200VFX_HASH_GRADIENT = 1,
  B200VFX_HASH_VERTGRADIENT = 2,
  B200VFX_HASH_DOUBLEGRADIENT = 3,
  B200VFX_HASH_BLOCKHASH = 4
} b200vfx_hash_algo;
#define B200VFX_HASH_MAX_BITS 64
/* one frame -> hash bits (bytes of 0/1; 64 bits, 40 for doublegradient), any frame size larger than 8x8.
 * fmt: RGB or RGBA; src host or device.  Synchronous. */
int b200vfx_hash_image(b200vfx_ctx *ctx, int algo, int fmt, int width, int height, const void *src, int stride,
                       uint8_t *bits_out, int *n_bits);
/* HashAlg::Blockhash, integer fast path (width % hw == 0 && height % hh == 0): the hw*hh u32 block sums of
 * (A==0 ? 765 : R+G+B).  `sums` may be a host or device pointer. */
int b200vfx_blockhash_sums(b200vfx_ctx *ctx, int fmt, int width, int height, const void *src,
                           int stride, int hw, int hh, uint32_t *sums);
/* VideoCompare::aggregate_frames (videocompare/imp.rs:297-353) hashes the reference pad's frame and then every other
 * pad's frame on each tick: the same block sums for n_frames (1..8) equally sized frames of one format in ONE launch;
 * sums receives n_frames * hw*hh values, frame f at sums + f*hw*hh.  Host and device frames may be mixed. */
#define B200VFX_BLOCKHASH_MAX_FRAMES 8
int b200vfx_blockhash_sums_batch(b200vfx_ctx *ctx, int fmt, int width, int height, int n_frames,
                                 const void *const *srcs, const int *strides, int hw, int hh, uint32_t *sums);
/* HashAlg::Blockhash for every other frame size (blockhash_slow): f32 block sums, block index = floor(x / (W/hw)) with
 * the reference's f32 division, accumulated in raster order (exact integers below 2^24; a sequential kernel beyond). */
int b200vfx_blockhash_sums_f32(b200vfx_ctx *ctx, int fmt, int width, int height, const void *src, int stride, int hw,
                               int hh, float *sums);
/* median/bit rules (gen_hash!) + Hamming distance (host side, tiny): bits_out hw*hh bytes of 0/1 */
void b200vfx_blockhash_bits(const uint32_t *sums, int hw, int hh, int width, int height,
                            uint8_t *bits_out);
void b200vfx_blockhash_bits_f32(const float *sums, int hw, int hh, int width, int height, uint8_t *bits_out);
int b200vfx_hash_distance(const uint8_t *bits_a, const uint8_t *bits_b, int nbits);
/* Mean / Gradient / VertGradient / DoubleGradient: image::imageops::grayscale + resize(FilterType::Lanczos3) of the frame
 * to nw x nh luma bytes (row-major) on the GPU, then the bit rule on the host.
 * b200vfx_hash_resize_dims: HashAlg::resize_dimensions for the 8x8 hash -> (8,8) (9,8) (8,9) (5,5). */
int b200vfx_luma_resize(b200vfx_ctx *ctx, int fmt, int width, int height, const void *src, int stride, int nw, int nh,
                        uint8_t *out);
int b200vfx_hash_resize_dims(int algo, int *nw, int *nh);
int b200vfx_hash_bits_from_luma(int algo, const uint8_t *luma, int nw, int nh, uint8_t *bits);

/* ---- colordetect --------------------------------------
 * ColorDetect::detect_color (video/videofx/src/colordetect/imp.rs:57-86) = color_thief::get_palette(plane 0,
 * format, quality, max_colors) + color_name::css::Color::similar(palette[0]).  The per-pixel part of get_palette --
 * the 5-bit-per-channel histogram over every `quality`-th pixel of the FLAT plane (stride * height bytes, stride
 * padding included, exactly as frame.plane_data(0) exposes it) -- runs on the GPU; pixels with a < 125 or
 * r,g,b all > 250 are skipped.  hist: 32768 u32 counts, index (r>>3)<<10 | (g>>3)<<5 | (b>>3), host or device
 * pointer.  fmt: RGB RGBA ARGB BGR BGRA (imp.rs:214-222).  quality 1..10 (0 is accepted by the GObject property
 * but trips color-thief's range check: B200VFX_ERR_INVALID). */
#define B200VFX_COLORDETECT_BINS 32768
int b200vfx_colordetect_histogram(b200vfx_ctx *ctx, int fmt, int width, int height, const void *src, int stride,
                                  int quality, uint32_t *hist);
/* host side, once per frame on 32768 integers: modified median cut of color-thief 0.2.2 (third-party, restated
 * from its published algorithm -- parity unpinned).  Writes min(*n_colors, palette_cap) RGB triples, most
 * significant first; palette[0] is the dominant colour.  max_colors 2..255 (imp.rs:134-141). */
int b200vfx_colordetect_palette(const uint32_t *hist, int max_colors, uint8_t *palette_rgb, int palette_cap,
                                int *n_colors);
/* color_name::css::Color::similar(rgb).to_lowercase() (imp.rs:76-79): nearest CSS keyword by squared distance */
const char *b200vfx_css_color_similar(unsigned r, unsigned g, unsigned b);

/* ---- multi-GPU: colorlut on a row tile fused with the all-gather of the tiles over peer memory --------
 * The reference has no multi-device path; 8(e) shards the frame into contiguous row tiles (one process per GPU) and asks
 * for one in-place all-gather when a DEVICE-side consumer wants the whole frame.  These entry points replace
 * "tile kernel + ncclAllGather" by ONE kernel that stores its results into every rank's frame buffer over NVLink.
 *
 * Buffers that peers write into must come from b200vfx_peer_alloc (plain cudaMalloc + a CUDA IPC handle the owner sends
 * to the other processes through any channel it likes, e.g. torch.distributed.all_gather_object); a peer maps it with
 * b200vfx_peer_open.  Within one process, b200vfx_peer_enable_access + the raw pointers do the same.
 * A flag block is B200VFX_PEER_FLAG_BYTES of zero-initialised peer memory per rank (b200vfx_peer_alloc zeroes). */
#define B200VFX_IPC_HANDLE_BYTES 64
#define B200VFX_MAX_PEERS 16
#define B200VFX_PEER_FLAG_BYTES 256
int b200vfx_peer_alloc(b200vfx_ctx *ctx, size_t bytes, void **dev_ptr, unsigned char handle_out[B200VFX_IPC_HANDLE_BYTES]);
int b200vfx_peer_free(b200vfx_ctx *ctx, void *dev_ptr);
int b200vfx_peer_open(b200vfx_ctx *ctx, const unsigned char handle[B200VFX_IPC_HANDLE_BYTES], void **dev_ptr);
int b200vfx_peer_close(b200vfx_ctx *ctx, void *dev_ptr);
int b200vfx_peer_enable_access(b200vfx_ctx *ctx, int peer_device);
/* synchronises the context stream and reports the epoch of the last call whose peer wait timed out (0 = none) */
int b200vfx_peer_status(b200vfx_ctx *ctx, const void *flags, uint32_t *error_epoch);
/* ColorLut::transform_frame (video/colorlut/src/colorlut/imp.rs:204-235, RGBA arm :267-294) on rows
 * [frame_row0, frame_row0 + tile_rows) of the frame: reads the tile from src (device memory of this rank), writes the
 * result into frames[p] + frame_row0 * frame_stride for every p in [0, world) (frames[rank] is this rank's own buffer).
 * Asynchronous on the context stream; when it completes there, ALL ranks' tiles of this epoch are visible in
 * frames[rank].  epoch must be > 0, identical on all ranks for one frame and increase by one per call.
 * RGBA + memo mode only (B200VFX_ERR_UNSUPPORTED otherwise). */
int b200vfx_colorlut_process_tile_gather(b200vfx_ctx *ctx, int fmt, int width, int tile_rows, const void *src,
                                         int src_stride, int world, int rank, void *const *frames, int frame_stride,
                                         int frame_row0, void *const *flags, uint32_t epoch);
/* The same with ONE NVSwitch multicast mapping in front of the frame buffers: multicast_frame is a device address from a
 * multicast object (cuMulticastCreate / cuMulticastBindMem, or torch.distributed._symmetric_memory's multicast_ptr) to
 * which frames[p] of every rank is bound at the same offset.  Every result vector then leaves the GPU once
 * (multimem.st) and the switch writes it into all `world` buffers, this rank's own included: egress per GPU is the
 * tile instead of (world - 1) x the tile.  frames[] is still needed (hazard tracking, frames[rank] is where the frame is
 * read); NULL multicast_frame = the unicast call above.  Needs width % 4 == 0 and 16-byte aligned frames. */
int b200vfx_colorlut_process_tile_gather_mc(b200vfx_ctx *ctx, int fmt, int width, int tile_rows, const void *src,
                                            int src_stride, int world, int rank, void *const *frames, void *multicast_frame,
                                            int frame_stride, int frame_row0, void *const *flags, uint32_t epoch);

/* ---- test hooks (host logic only, no GPU needed) ---------------------------------------------------------------
 * The admission rule for overlapping consecutive frames (programmatic dependent launch): would a launch with these
 * source / destination byte ranges be allowed to start before our earlier launches on `stream_key` have completed?
 * Records the launch exactly like a real one (threads = grid x block size, lingers = it ends on griddepcontrol.wait;
 * lingers = 3: it passes griddepcontrol.wait BEFORE its first write to dst -- the reductions -- so only its early reads of
 * src can race with earlier launches).
 * b200vfx_debug_pdl_reset forgets the stream (what a stream synchronisation does). */
int b200vfx_debug_pdl_admit(void *stream_key, uintptr_t src_lo, uintptr_t src_hi, uintptr_t dst_lo, uintptr_t dst_hi,
                            int want_pdl, long long threads, int lingers);
void b200vfx_debug_pdl_reset(void *stream_key);
/* the normalised Lanczos3 tap weights image::imageops::resize uses for output sample `out` when in_len samples become
 * out_len (host computation behind b200vfx_luma_resize); returns the number of taps written to ws, < 0 on error */
int b200vfx_debug_resize_taps(int in_len, int out_len, int out, int *left, float *ws, int cap);

#ifdef __cplusplus
}
#endif
#endif /* B200VFX_H */
