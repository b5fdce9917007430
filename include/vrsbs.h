/*
 * vrsbs.h — C ABI of the B200-native SBS (side-by-side stereo) warp hot path.
 *
 * The reference (Gia-Huynh/VR-Video-Generator) has no FFI/plugin interface for this path: the
 * boundary is the Python class `SbsProcessor` (PredictAndGenerate.py:63-198) plus the depth tail
 * of the producer (depth_anything_v2/dpt.py:196-199, PredictAndGenerate.py:27-34,55).  The entry
 * points below are what a maintainer would bind (ctypes) to replace those
 * call sites; each one cites the reference lines it replaces.
 *
 * Conventions
 *   - plain C, no C++ types, no exceptions across the boundary; every call returns 0 on success or
 *     a negative VRSBS_E_* code and records a message retrievable with vrsbs_last_error();
 *   - "dev" pointers are CUDA device pointers on the context's device, "host" pointers are host
 *     memory (pinned or pageable); the caller owns every buffer it passes in;
 *   - calls taking a stream are asynchronous and stream-ordered (stream = a cudaStream_t cast to
 *     void*, NULL = the legacy default stream); nothing synchronises the host unless stated;
 *   - a context is NOT thread-safe; use one per (process, device, clip range);
 *   - frames are uint8 RGB, [H,W,3] row-major (what `left_side_sbs` receives); full-resolution depth is
 *     IEEE fp16 or fp32 (vrsbs_params.depth_dtype; the DPT-resolution map is always fp16); an SBS frame is
 *     [H,2W,3] = [warped view | input frame].
 *   - there is no CPU fallback: without a CUDA device vrsbs_create fails.
 */
#ifndef VRSBS_H
#define VRSBS_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VRSBS_ABI_VERSION 2

enum {
    VRSBS_OK            = 0,
    VRSBS_E_INVALID     = -1,  /* bad argument (NULL, size out of the context's limits, ...)        */
    VRSBS_E_CUDA        = -2,  /* a CUDA runtime call failed; message has cudaGetErrorString        */
    VRSBS_E_NOMEM       = -3,
    VRSBS_E_FRAME       = -4,  /* a frame of the batch was rejected on the device: NaN depth (the    */
                               /* reference raises in math.ceil) or more layers than max_layers.   */
                               /* Like the reference's worker, which dies on that exception, the   */
                               /* clip-range state is not defined afterwards: call vrsbs_reset.    */
                               /* The device-pointer calls are asynchronous and report it through  */
                               /* vrsbs_get_frame_info only; the host calls return it themselves.  */
    VRSBS_E_STATE       = -5   /* call order violated (e.g. warp before prepare)                    */
};

/* per-frame status bits written by the device (vrsbs_frame_info.status) */
#define VRSBS_FRAME_NAN        1u   /* depth.max() is NaN                                   */
#define VRSBS_FRAME_OVERFLOW   2u   /* layer count > max_layers                             */
#define VRSBS_FRAME_GENERIC    4u   /* non-monotone bounds: brute-force membership was used */

typedef struct vrsbs_ctx vrsbs_ctx;

/* Parameters of one clip range.  Mirrors what SbsProcessor.__init__ copies out of `args_god`
 * (PredictAndGenerate.py:63-100) and the constants it derives. */
typedef struct vrsbs_params {
    double offset_fg;          /* --offset_fg        (PredictAndGenerate.py:340, default 0.025)  */
    double offset_bg;          /* --offset_bg        (PredictAndGenerate.py:342, default -0.01)  */
    int    offset_step_size;   /* --offset_step_size (PredictAndGenerate.py:344, default 1)      */
    int    blur;               /* 1 = full path; 0 = stop after hole fill (pre-blur view, strip  */
                               /*     not restored) — parity tiers T3/T4 only                    */
    int    depth_dtype;        /* VRSBS_DEPTH_F16 (default) or VRSBS_DEPTH_F32: element type of    */
                               /*     every full-resolution depth buffer of this clip range.  The  */
                               /*     reference is dtype-agnostic: smoothing (:139-142), the max    */
                               /*     (:102) and the bin comparison (:173) run in the tensor's      */
                               /*     dtype.  fp16 is what autocast produced with the reference's   */
                               /*     pinned torch; torch >= 2.4 on CUDA returns fp32 from the     */
                               /*     bicubic tail (upsample_bicubic2d is on autocast's fp32 list). */
                               /*     fp32 runs the general row kernel (slower, same contract).    */
} vrsbs_params;

enum { VRSBS_DEPTH_F16 = 0, VRSBS_DEPTH_F32 = 1 };

/* What the device decided for one frame (the python lists `get_cutoff` returns, flattened). */
typedef struct vrsbs_frame_info {
    uint32_t status;           /* VRSBS_FRAME_* bits                                             */
    int32_t  layers;           /* L = len(step_list)                                             */
    int32_t  limit_step;       /* math.ceil(depth.max())                                         */
    int32_t  fill_layer;       /* int(L*3/5)                                                     */
    int32_t  strip;            /* columns [0,strip) restored from the input                      */
    int32_t  reserved;
    float    depth_max;        /* depth.max() of the smoothed frame                              */
    float    reserved2;
    double   offset_range[2];  /* the averaged [bg,fg] pixel range (self.last_offset_range)      */
    uint64_t holes;            /* number of unpainted destination pixels (before fill)           */
} vrsbs_frame_info;

int  vrsbs_abi_version(void);

/* Life cycle.  Replaces SbsProcessor.__init__ (PredictAndGenerate.py:63-100): allocates scratch
 * for frames up to max_h x max_w, batches up to max_batch, tables up to max_layers. */
int  vrsbs_create(vrsbs_ctx **out, int device, int max_h, int max_w, int max_batch, int max_layers);
int  vrsbs_destroy(vrsbs_ctx *ctx);
const char *vrsbs_last_error(const vrsbs_ctx *ctx);   /* ctx may be NULL: last create() failure */

/* Sets offsets/step and forgets depth history + range EMA — what constructing a fresh
 * SbsProcessor per clip range does (PredictAndGenerate.py:209). */
int  vrsbs_reset(vrsbs_ctx *ctx, const vrsbs_params *params);

/* The range-EMA half of the clip state (SbsProcessor.last_offset_range, PredictAndGenerate.py:105-108),
 * readable and writable so a host-side get_cutoff() and the device tables share one state.
 * has_last = 0 means `last_offset_range is None`.  Both synchronise the device. */
int  vrsbs_get_range_state(vrsbs_ctx *ctx, int *has_last, double range[2]);
int  vrsbs_set_range_state(vrsbs_ctx *ctx, int has_last, const double range[2]);

/* Gaussian weights for the hole blur, [ky,kx] row-major fp32, exactly as the caller's torchvision
 * computes them (torchvision _misc.py:86-98 via PredictAndGenerate.py:191-193).  kx = 2k+3 along W,
 * ky = 2k+1 along H, k = round(0.0036*H) (PredictAndGenerate.py:165).  Host pointer; copied. */
int  vrsbs_set_blur_weights(vrsbs_ctx *ctx, const float *weights_host, int kx, int ky);

/* ---- stage 1: depth tail + temporal smoothing + per-frame max --------------------------------
 * Replaces dpt.py:196 (bicubic, A=-0.75, align_corners=True, fp32 accumulate -> fp16),
 * PredictAndGenerate.py:55 (`* scaler` in fp16) and SbsProcessor.get_depth's smoothing
 * (PredictAndGenerate.py:134-144: 0.58*d_t + 0.30*d_{t-1} + 0.12*d_{t-2}, fp16 rounding after every
 * op, history holds RAW depths), plus the `depth.max()` of get_cutoff (PredictAndGenerate.py:102).
 * depth_lo_dev  [B,h,w] fp16 DPT output;  depth_out_dev [B,H,W] fp16 smoothed full-res depth.
 * Frames are consecutive frames of the clip range; history carries over between calls. */
int  vrsbs_depth_from_lowres(vrsbs_ctx *ctx, const void *depth_lo_dev, int B, int h, int w,
                             float scaler, int H, int W, void *depth_out_dev, void *stream);

/* Same, but the raw depth is already full resolution ([B,H,W] fp16) — the tensor the reference's
 * result_queue delivers to get_depth (PredictAndGenerate.py:133). */
int  vrsbs_depth_from_full(vrsbs_ctx *ctx, const void *depth_raw_dev, int B, int H, int W,
                           void *depth_out_dev, void *stream);

/* ---- stage 2: layer tables on the device --------------------------------------------------------
 * Replaces SbsProcessor.get_cutoff (PredictAndGenerate.py:101-126) for the B frames whose maxima
 * the previous depth call left on the device: ceil(max), EMA of the offset range across frames,
 * thresholds, sorted, steps, re-rounded offsets, bounds narrowed to fp16, fill layer, strip width.
 * IEEE double arithmetic in the reference's operation order; no host round trip. */
int  vrsbs_build_tables(vrsbs_ctx *ctx, int B, int H, int W, void *stream);

/* ---- stage 3: layered warp + hole fill + hole blur + strip + SBS pack -----------------------------
 * Replaces gpu_roll_with_offset and the body of left_side_sbs (PredictAndGenerate.py:150-155,
 * 161-197) for B frames.  frames_dev [B,H,W,3] u8, depth_dev [B,H,W] fp16 (output of stage 1),
 * sbs_dev [B,H,2W,3] u8.  Uses the tables stage 2 left in the context. */
int  vrsbs_warp_batch(vrsbs_ctx *ctx, const uint8_t *frames_dev, const void *depth_dev,
                      int B, int H, int W, uint8_t *sbs_dev, void *stream);

/* Stage 1(full-res)+2+3 in one call on device buffers (the "warp stage" bench.py times).  The
 * smoothed depth is written to depth_scratch_dev ([B,H,W] fp16, required) and consumed by the warp
 * kernel.  With the option "smooth_in_warp" = 1 (and a frame width that is a multiple of 16, 16-byte
 * aligned pointers) the depth pass computes the maxima only and the warp kernel recomputes the
 * smoothing from the raw rows; depth_scratch_dev is then not touched and may be NULL. */
int  vrsbs_process_batch(vrsbs_ctx *ctx, const uint8_t *frames_dev, const void *depth_raw_dev,
                         int B, int H, int W, void *depth_scratch_dev, uint8_t *sbs_dev, void *stream);

/* ---- host-buffer entry: what left_side_sbs does end to end ----------------------------------------
 * Replaces the H2D copies (PredictAndGenerate.py:133,158), the whole warp and the blocking D2H
 * (`.cpu().numpy()`, PredictAndGenerate.py:197) for B frames with HOST buffers.  depth_host is
 * either full-res raw depth [B,H,W] fp16 (lowres_h = lowres_w = 0) or DPT low-res [B,h,w] fp16.
 * Internally pipelined: three streams (H2D / kernels / D2H) and three pinned slots; frames are
 * processed in chunks ("host_chunk", default 8) so copy-in, kernels and copy-out of neighbouring
 * chunks overlap.  Only the synthesised (left) half of every SBS row crosses PCIe on the way back;
 * the right half is the caller's own frame and is copied host-to-host by the library's copy threads
 * (option "host_right_half", default 1).  Returns after sbs_host is complete (like the reference's
 * blocking D2H).  depth_host may also be a DEVICE pointer (a depth producer in the same process that keeps
 * its output on the GPU, PredictAndGenerate.py:23-61 without the `.to('cpu')` at :55): it is then read where
 * it is, and the caller guarantees that the work which produced it has completed. */
int  vrsbs_process_host(vrsbs_ctx *ctx, const uint8_t *frames_host, const void *depth_host,
                        int B, int H, int W, int lowres_h, int lowres_w, float scaler,
                        uint8_t *sbs_host);

/* ---- asynchronous host-buffer entry: decode and encode overlap the kernels -----------------------------------
 * vrsbs_submit_host enqueues one batch (the copies in, the kernels and the copy out of its chunks, on the library's
 * three streams) and returns; vrsbs_collect blocks until that batch's SBS frames are complete in sbs_host and reports
 * its per-frame status like vrsbs_process_host.  Batches run in submission order and share the clip-range state, so
 * sub-clip k+1 uploads and warps while the caller encodes sub-clip k and decodes sub-clip k+2 - what the serial loop of
 * nibba_woka (PredictAndGenerate.py:216-250) cannot do.  All host buffers must be page-locked and stay valid until
 * collected; at most three chunks are in flight, so submit itself blocks while the pipeline is full.
 * frames_host may be pitched: row y of frame t starts at frames_host + t * frame_pitch + y * frame_row_pitch
 * (0 = packed).  With VRSBS_HOST_RIGHT_IN_PLACE the caller has decoded its frames straight into the right halves of
 * sbs_host (frames_host = sbs_host + 3W, frame_row_pitch = 6W, frame_pitch = 6WH): the input then never gets copied
 * on the host and only the synthesised left halves come back across PCIe.  depth_host as in vrsbs_process_host.
 * Mixing the two entry styles on one context: the device-pointer calls run on the caller's stream, the host calls on the
 * library's; a host call orders itself behind the device-pointer calls made before it (event), a device-pointer call is
 * refused (VRSBS_E_STATE) while submitted batches have not been collected. */
#define VRSBS_HOST_RIGHT_IN_PLACE 1u
int  vrsbs_submit_host(vrsbs_ctx *ctx, const uint8_t *frames_host, size_t frame_row_pitch, size_t frame_pitch,
                       const void *depth_host, int B, int H, int W, int lowres_h, int lowres_w, float scaler,
                       uint8_t *sbs_host, unsigned flags, uint64_t *ticket);
int  vrsbs_collect(vrsbs_ctx *ctx, uint64_t ticket);
/* Orders the host pipeline's kernels behind everything queued so far on `producer_stream` (a depth producer that left
 * its output on the device): an event wait on the device, the host is not blocked. */
int  vrsbs_host_depends_on(vrsbs_ctx *ctx, void *producer_stream);

/* ---- introspection (parity tiers T1..T3, error reporting) ------------------------------------------
 * All of these synchronise `stream` first. */
int  vrsbs_get_frame_info(vrsbs_ctx *ctx, int B, vrsbs_frame_info *info_host, void *stream);
/* Tables of batch frame `frame`: cutoffs [L+1] (double), offsets [L] (int32), bounds lo/hi [L] as
 * fp16 bit patterns.  Any output pointer may be NULL.  cap = capacity in elements of each array. */
int  vrsbs_get_tables(vrsbs_ctx *ctx, int frame, int cap, double *cutoffs, int32_t *offsets,
                      uint16_t *lo_f16, uint16_t *hi_f16, void *stream);
/* The [lo, hi) bounds the device compares the depth against, as floats (exact in both depth dtypes: fp16-narrowed
 * values for fp16 depth, fp32-narrowed ones for fp32 depth).  Returns L. */
int  vrsbs_get_bounds(vrsbs_ctx *ctx, int frame, int cap, float *lo, float *hi, void *stream);
/* Hole bitmask of the last vrsbs_warp_batch: [B,H,ceil(W/32)] uint32, bit x%32 of word x/32. */
int  vrsbs_get_hole_mask(vrsbs_ctx *ctx, int B, int H, int W, uint32_t *mask_host, void *stream);
/* Number of kernels this library has launched on this context (bench.py's gpu_launches). */
uint64_t vrsbs_launch_count(const vrsbs_ctx *ctx);

/* Per-stage device time (CUDA events recorded on the launching stream around every kernel while the
 * option "stage_timing" is 1).  Synchronises, then ADDS the elapsed milliseconds of all launches
 * recorded since the previous call to ms[0..4] = {depth, tables, warp, blur, commit+strip} and the number
 * of launches to count[0..4]; the caller zeroes the arrays.  bench.py's roofline uses ms[2]. */
#define VRSBS_NUM_STAGES 5
int  vrsbs_get_stage_times(vrsbs_ctx *ctx, double ms[VRSBS_NUM_STAGES], uint64_t count[VRSBS_NUM_STAGES]);

/* Tuning knobs (defaults are the measured best; the others stay for tests and experiments):
 *   "fused" (1 = TMA warp kernels when the shape allows, 0 = general row kernel), "warp_ws" (1 = warp-specialised
 *   k_warp_ws, 0 = barrier-synchronised k_warp_fused), "ws_scatter_warps" (scatter/destination split of k_warp_ws),
 *   "smooth_in_warp" (see vrsbs_process_batch), "fast_tables" (0 forces the slow membership path),
 *   "scatter_mode" of the general row kernel (2 = atomicMax for every key, 1 = plain store + verify),
 *   "blur_sep" (1 = separable screening kernel k_blur_sep when the weights are near rank 1, 0 = k_blur_holes_fixed,
 *   2 = k_blur_sep with every value sent to its exact fallback),
 *   "blur_band" (1 = band-driven k_blur_band where it is built - the 1080p and 720p footprints -, 0 = per-word k_blur_sep),
 *   "ws_no_list" (1 = k_warp_ws leaves the per-word hole list to k_word_list), "f32_fast" (1 = the vectorised fp32 depth pass
 *   and the warp-specialised kernel's fp32 instantiation, 0 = the general fp32 kernels),
 *   "blur_screen" (1 = screening sum before the exact integer blur, 0 = exact sum for every hole),
 *   "commit_mode", "lowres_tiled", "bicubic_contract", "blocks_per_sm", "host_chunk", "copy_threads",
 *   "pageable_direct", "host_right_half", "host_async" (0 = page-locked callers of vrsbs_process_host use the blocking
 *   chunk loop instead of submit + collect), "pdl" (bit mask of programmatic-dependent-launch edges), "stage_timing". */
int  vrsbs_set_option(vrsbs_ctx *ctx, const char *name, int value);

#ifdef __cplusplus
}
#endif
#endif /* VRSBS_H */
