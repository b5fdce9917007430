// C ABI of the SBS warp hot path (see include/vrsbs.h for the contract and reference citations).
#include <cstdarg>
#include <cstdio>
#include <cmath>
#include <cstring>
#if defined(__x86_64__) || defined(_M_X64)
#include <emmintrin.h>
#define VRSBS_HAVE_SSE2 1
#endif
#include <new>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <utility>
#include <vector>

#include "blur_holes.cuh"
#include "common.cuh"
#include "depth_kernels.cuh"
#include "table_kernel.cuh"
#include "warp_fused.cuh"
#include "warp_ws.cuh"
#include "warp_kernel.cuh"

using namespace vrsbs;

namespace {

thread_local char g_create_error[256] = "";

struct Scratch {                 // per-batch device scratch; one per pipeline slot
    uint32_t *frame_max = nullptr, *frame_nan = nullptr;   // one allocation: [B] max | [B] nan | [1] hole_count | [1] blur ticket | [1] band count
    FrameTab *tabs = nullptr;
    float2 *bounds = nullptr;
    int *offm = nullptr;
    double *cutoffs = nullptr;
    int *offsets = nullptr;
    uint16_t *lo16 = nullptr, *hi16 = nullptr;
    uint32_t *hole_mask = nullptr;
    uint32_t *hole_list = nullptr;   // mask words that contain holes (blur work list)
    uint32_t *band_map = nullptr;    // bit per band column (8 rows x one mask word) with a hole; all zero between batches (set by k_warp_ws or
                                     // k_band_map, compacted into band_list and cleared by k_band_list)
    uint32_t *band_list = nullptr;   // the marked band columns: work list of k_blur_band and k_blur_commit
    bool band_marked = false;        // the last warp launch marked the band columns itself (otherwise k_band_map does)
    bool list_missing = false;       // the last warp launch left the per-word hole list to k_word_list
    uint32_t *hole_count = nullptr;  // points into frame_max's allocation: one memset clears all three
    uint8_t *blobs = nullptr;        // [B][kBlobMax] fast-path tables
    uint8_t *plane = nullptr;        // [B,H,W,3] blurred hole values (allocated on first blur)
    bool holes_dirty = false;        // a warp kernel has appended to hole_list since hole_count was last cleared
    int cap_batch = 0;               // frames this scratch set was allocated for
};

constexpr int kEntCapMax = 255, kLutCapMax = 8192;

constexpr int kSlots = 3;         // chunks in flight in vrsbs_process_host

struct HostSlot {                // one chunk of host<->device staging for vrsbs_process_host
    cudaEvent_t in_done = nullptr, k_done = nullptr, out_done = nullptr;
    FrameTab *pin_tabs = nullptr;    // [max_batch] pinned copy of the chunk's frame records
    uint8_t *pin_frames = nullptr, *pin_depth = nullptr, *pin_sbs = nullptr;
    uint8_t *dev_frames = nullptr, *dev_depth_in = nullptr, *dev_depth = nullptr, *dev_sbs = nullptr;
    size_t cap_frames = 0, cap_depth_in = 0, cap_depth = 0, cap_sbs = 0;
    size_t cap_pin_frames = 0, cap_pin_depth = 0, cap_pin_sbs = 0;
};

// Row copy with non-temporal stores: the destinations are large, written once and read much later (by ffmpeg or
// the DMA engine), so bypassing the cache avoids the read-for-ownership and leaves host DRAM bandwidth to PCIe.
inline void stream_copy(char *dst, const char *src, size_t n) {
#ifdef VRSBS_HAVE_SSE2
    if (n >= 256 && ((uintptr_t)dst & 15) == 0) {
        size_t i = 0;
        for (; i + 64 <= n; i += 64) {
            const __m128i a = _mm_loadu_si128((const __m128i *)(src + i)), b = _mm_loadu_si128((const __m128i *)(src + i + 16));
            const __m128i c = _mm_loadu_si128((const __m128i *)(src + i + 32)), d = _mm_loadu_si128((const __m128i *)(src + i + 48));
            _mm_stream_si128((__m128i *)(dst + i), a);      _mm_stream_si128((__m128i *)(dst + i + 16), b);
            _mm_stream_si128((__m128i *)(dst + i + 32), c); _mm_stream_si128((__m128i *)(dst + i + 48), d);
        }
        if (i < n) memcpy(dst + i, src + i, n - i);
        return;
    }
#endif
    memcpy(dst, src, n);
}

// Persistent host threads for the row copies of the host pipeline (staging of pageable buffers, and the right half
// of the SBS frame, which is the caller's own input and never needs to cross PCIe).
class CopyPool {
public:
    explicit CopyPool(int nthreads) {
        for (int i = 0; i < nthreads; ++i) workers_.emplace_back([this] { run(); });
    }
    ~CopyPool() {
        { std::lock_guard<std::mutex> g(m_); stop_ = true; }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    // rows x width bytes from src (pitch spitch) to dst (pitch dpitch); returns a ticket for wait()
    int submit(void *dst, size_t dpitch, const void *src, size_t spitch, size_t width, size_t rows) {
        if (rows == 0 || width == 0) return -1;
        size_t tail;
        if (dpitch == width && spitch == width) {               // contiguous: re-shape into 256 KiB pieces
            const size_t total = width * rows, piece = 256u << 10;
            width = total < piece ? total : piece;
            rows = (total + width - 1) / width;
            tail = total - (rows - 1) * width;
            dpitch = spitch = width;
        } else {
            tail = width;
        }
        const size_t parts = rows < (size_t)workers_.size() * 2 ? rows : (size_t)workers_.size() * 2;
        std::lock_guard<std::mutex> g(m_);
        const int ticket = next_ticket_++;
        pending_.push_back({ticket, (int)parts});
        for (size_t p = 0; p < parts; ++p) {
            const size_t r0 = rows * p / parts, r1 = rows * (p + 1) / parts;
            jobs_.push_back({ticket, (char *)dst + r0 * dpitch, (const char *)src + r0 * spitch, dpitch, spitch, width, r1 - r0,
                             (r1 == rows) ? tail : width});
        }
        cv_.notify_all();
        return ticket;
    }
    void wait(int ticket) {
        if (ticket < 0) return;
        std::unique_lock<std::mutex> g(m_);
        done_cv_.wait(g, [&] {
            for (auto &p : pending_) if (p.ticket == ticket) return false;
            return true;
        });
    }
private:
    struct Job { int ticket; char *dst; const char *src; size_t dpitch, spitch, width, rows, last_width; };
    struct Pending { int ticket, left; };
    void run() {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> g(m_);
                cv_.wait(g, [&] { return stop_ || !jobs_.empty(); });
                if (jobs_.empty()) return;
                j = jobs_.front();
                jobs_.pop_front();
            }
            for (size_t r = 0; r < j.rows; ++r)
                stream_copy(j.dst + r * j.dpitch, j.src + r * j.spitch, r + 1 == j.rows ? j.last_width : j.width);
#ifdef VRSBS_HAVE_SSE2
            _mm_sfence();
#endif
            {
                std::lock_guard<std::mutex> g(m_);
                for (size_t i = 0; i < pending_.size(); ++i)
                    if (pending_[i].ticket == j.ticket && --pending_[i].left == 0) { pending_.erase(pending_.begin() + i); break; }
            }
            done_cv_.notify_all();
        }
    }
    std::vector<std::thread> workers_;
    std::deque<Job> jobs_;
    std::vector<Pending> pending_;
    std::mutex m_;
    std::condition_variable cv_, done_cv_;
    int next_ticket_ = 0;
    bool stop_ = false;
};

}  // namespace

struct vrsbs_ctx {
    int device = 0, max_h = 0, max_w = 0, max_batch = 0, max_layers = 0;
    int sm_count = 0;
    vrsbs_params params{0.025, -0.01, 1, 1, VRSBS_DEPTH_F16};
    int f32 = 0;                         // params.depth_dtype == VRSBS_DEPTH_F32: every depth pointer is float, general route
    SmoothWeights sw{};
    // clip-range state
    __half *hist[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};   // [set][0: raw t-1, 1: raw t-2]; set hist_idx is current
    int hist_idx = 0;
    RangeState *state = nullptr;         // [2], ping-pong
    int state_idx = 0;
    long long depth_frames = 0;          // frames pushed through stage 1 since reset
    int state_h = 0, state_w = 0;
    // scratch
    Scratch scratch[kSlots];
    cudaStream_t st_in = nullptr, st_k = nullptr, st_out = nullptr;   // H2D / kernels / D2H of the host pipeline
    float *weights = nullptr;
    uint32_t *wq = nullptr;              // integer blur weights [parts][(ky/2+1)*(kx/2+1)]
    std::vector<uint32_t> wq_host, wh_host;   // exact parts / screening weights floor(w * 2^s1) of the integer blur
    uint32_t ws1 = 1, wrmax = 2;
    int blur_screen = 1;
    std::vector<uint32_t> sep_hy, sep_hx;     // separable screening kernel of k_blur_sep (scale 2^sep_s), empty if the weights are not near rank 1
    uint32_t sep_s = 0, sep_eps32 = 0;
    int blur_sep = 1;
    int ws_no_list = 1;                       // option: 1 = k_warp_ws leaves the per-word hole list to k_word_list (list-driven blur kernels)
    int blur_band = 1;                        // option: 1 = band-driven k_blur_band for the footprints it is built for (1080p, 720p)
    int pdl = 0;                              // option (bit mask): programmatic dependent launch of 1 tables, 2 warp, 4 blur, 8 commit.
                                              // Measured slower than plain stream order with all four edges on (0.532 vs 0.512 ms per
                                              // 1080p step), so off by default                         // option: 1 = k_blur_sep when the weights allow it, 0 = k_blur_holes_fixed
    int kx = 0, ky = 0, wparts = 0, wshift = 0;
    int ent_cap = 0, lut_cap = 0;        // fast-path table capacities of the last vrsbs_build_tables
    int key_pad = 0;                     // fast path: bound on |signed layer offset| in pixels (multiple of 32)
    int fused = 1;                       // option: use the fused route in vrsbs_process_batch when possible
    int fast_tables = 1;                 // option (tests): 0 forces the slow membership path of k_warp_fused
    int ws_scatter_warps = 4;            // option: scatter warps of k_warp_ws (3, 4 or 5 of 8 warps); 4 measured best (9- and 10-warp CTAs were slower
                                         // and are no longer built)
    int commit_mode = 1;                 // experiments: 1 = commit + strip, 0 = strip only, 3 = commit only, 2 = neither
    int f32_fast = 1;                    // option: 1 = vectorised fp32 depth pass and the warp-specialised warp kernel for fp32 depth when they fit
    int warp_ws = 1;                     // option: 1 = warp-specialised warp kernel (k_warp_ws) when it fits, 0 = k_warp_fused
    int lowres_tiled = 1;                // option: 0 = one-pixel-per-thread bicubic kernel (tests)
    int smooth_in_warp = 0;              // option: 1 = smoothing recomputed inside the warp kernel (no smoothed depth in HBM);
                                         // measured slower than materialising it, so off by default
    // host pipeline
    HostSlot slot[kSlots];
    int host_chunk = 8;                  // frames per chunk of the host pipeline (measured: 2 -> 4 196, 4 -> 4 381, 8 -> 4 429, 16 -> 4 466 frames/s end to end)
    int copy_threads = 8;
    int pageable_direct = 0;             // option: 1 = hand pageable host pointers to cudaMemcpyAsync instead of staging them
    int host_right_half = 1;             // option: 1 = the right half of the SBS frame (= the caller's input) is copied on the host
    CopyPool *pool = nullptr;            // created on first use by vrsbs_process_host
    int host_async = 1;                  // option: page-locked callers of vrsbs_process_host go through submit + collect
    int skip_right = 0;                  // set around the host pipeline's warp launches: the right half stays on the host
    // asynchronous host pipeline (vrsbs_submit_host / vrsbs_collect)
    struct Inflight { uint64_t ticket; int rc; int chunks; char err[256]; };
    std::vector<Inflight> inflight;      // submitted, not yet collected
    uint64_t next_ticket = 1;
    uint64_t slot_ticket[kSlots] = {0, 0, 0};   // 0 = slot free
    int slot_n[kSlots] = {0, 0, 0}, slot_first[kSlots] = {0, 0, 0};
    long long chunk_seq = 0;             // chunks enqueued so far: slot = chunk_seq % kSlots
    cudaEvent_t dep_event = nullptr;     // vrsbs_host_depends_on
    cudaStream_t last_dev_stream = nullptr;   // stream of the last device-pointer call that touched the clip state ...
    bool dev_state_dirty = false;        // ... since the host pipeline last ordered itself behind it (mixed_entry_sync)
    // options / accounting
    int scatter_mode = 2;
    int bicubic_contract = 1;
    int blocks_per_sm = 0;               // 0 = occupancy API
    uint64_t launches = 0;
    int stage_timing = 0;
    struct Stamp { int stage; cudaEvent_t a, b; };
    std::vector<Stamp> stamps;           // recorded, not yet resolved
    std::vector<cudaEvent_t> event_pool;
    char err[256] = "";
};

namespace {

int fail(vrsbs_ctx *ctx, int code, const char *fmt, ...) {
    char *dst = ctx ? ctx->err : g_create_error;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 256, fmt, ap);
    va_end(ap);
    return code;
}

#define CU_TRY(ctx, call)                                                                              \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(ctx, VRSBS_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__); \
    } while (0)

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

template <typename T>
cudaError_t dmalloc(T **p, size_t count) { return cudaMalloc(reinterpret_cast<void **>(p), count * sizeof(T)); }

void free_scratch(Scratch &s) {
    cudaFree(s.frame_max); cudaFree(s.tabs); cudaFree(s.bounds); cudaFree(s.offm);
    cudaFree(s.cutoffs); cudaFree(s.offsets); cudaFree(s.lo16); cudaFree(s.hi16); cudaFree(s.hole_mask);
    cudaFree(s.hole_list); cudaFree(s.band_map); cudaFree(s.band_list); cudaFree(s.blobs); cudaFree(s.plane);
    s = Scratch{};
}

int alloc_scratch(vrsbs_ctx *c, Scratch &s, int cap_batch) {
    const size_t B = cap_batch, L = c->max_layers;
    s.cap_batch = cap_batch;
    const size_t mask_words = B * c->max_h * ((c->max_w + 31) / 32);
    CU_TRY(c, dmalloc(&s.frame_max, 2 * B + 4));
    s.frame_nan = s.frame_max + B;
    s.hole_count = s.frame_max + 2 * B;
    CU_TRY(c, dmalloc(&s.tabs, B));
    CU_TRY(c, dmalloc(&s.bounds, B * L));
    CU_TRY(c, dmalloc(&s.offm, B * (L + 1)));
    CU_TRY(c, dmalloc(&s.cutoffs, B * (L + 1)));
    CU_TRY(c, dmalloc(&s.offsets, B * L));
    CU_TRY(c, dmalloc(&s.lo16, B * L));
    CU_TRY(c, dmalloc(&s.hi16, B * L));
    CU_TRY(c, dmalloc(&s.hole_mask, mask_words));
    CU_TRY(c, dmalloc(&s.hole_list, mask_words));
    {
        const size_t words = B * (size_t)((c->max_h + kBandRows - 1) / kBandRows) * (((c->max_w + 31) / 32 + 31) / 32);
        CU_TRY(c, dmalloc(&s.band_map, words));
        CU_TRY(c, dmalloc(&s.band_list, words * 32));
        CU_TRY(c, cudaMemset(s.band_map, 0, words * sizeof(uint32_t)));
    }
    CU_TRY(c, dmalloc(&s.blobs, B * (size_t)blob_bytes(kEntCapMax, kLutCapMax, true)));
    return VRSBS_OK;
}

int check_dims(vrsbs_ctx *c, int B, int H, int W) {
    if (!c) return VRSBS_E_INVALID;
    if (B < 1 || B > c->max_batch) return fail(c, VRSBS_E_INVALID, "batch %d outside [1,%d]", B, c->max_batch);
    if ((long long)B * H >= (1 << 24)) return fail(c, VRSBS_E_INVALID, "batch of %d frames x %d rows exceeds 2^24 rows", B, H);
    if (H < 2 || H > c->max_h || W < 2 || W > c->max_w)
        return fail(c, VRSBS_E_INVALID, "frame %dx%d outside the context limit %dx%d", H, W, c->max_h, c->max_w);
    return VRSBS_OK;
}

// Records an event pair around one kernel launch when stage timing is on.
struct StageTimer {
    vrsbs_ctx *c; cudaStream_t st; int stage; cudaEvent_t a = nullptr, b = nullptr;
    static cudaEvent_t get(vrsbs_ctx *c) {
        if (!c->event_pool.empty()) { cudaEvent_t e = c->event_pool.back(); c->event_pool.pop_back(); return e; }
        cudaEvent_t e = nullptr;
        cudaEventCreate(&e);
        return e;
    }
    StageTimer(vrsbs_ctx *c_, cudaStream_t st_, int stage_) : c(c_), st(st_), stage(stage_) {
        if (!c->stage_timing) return;
        a = get(c); b = get(c);
        if (a) cudaEventRecord(a, st);
    }
    ~StageTimer() {
        if (!a || !b) return;
        cudaEventRecord(b, st);
        c->stamps.push_back({stage, a, b});
    }
};

// Kernel launch with the programmatic-dependent-launch attribute (see common.cuh): only for kernels that execute
// pdl_wait() before their first global read, and only directly behind another kernel of this library (events recorded
// for stage timing, or a memset, in between simply turn the edge back into an ordinary one).
template <typename... KArgs, typename... Args>
cudaError_t launch_pdl(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// ---- stage launchers (stream-ordered, no host sync) ---------------------------------------------------
int clear_counters(vrsbs_ctx *c, Scratch &s, int B, cudaStream_t st) {
    (void)B;
    CU_TRY(c, cudaMemsetAsync(s.frame_max, 0, sizeof(uint32_t) * (2 * s.cap_batch + 4), st));
    s.holes_dirty = false;
    return VRSBS_OK;
}

// the hole work list must start empty for every warp launch (a second vrsbs_warp_batch without a depth call in between)
int fresh_hole_list(vrsbs_ctx *c, Scratch &s, cudaStream_t st) {
    if (s.holes_dirty) CU_TRY(c, cudaMemsetAsync(s.hole_count, 0, 4 * sizeof(uint32_t), st));   // hole count, blur ticket, band count
    s.holes_dirty = true;
    return VRSBS_OK;
}

int check_state_dims(vrsbs_ctx *c, int H, int W) {
    if (c->depth_frames > 0 && (c->state_h != H || c->state_w != W))
        return fail(c, VRSBS_E_STATE, "frame size changed from %dx%d to %dx%d without vrsbs_reset", c->state_h,
                    c->state_w, H, W);
    return VRSBS_OK;
}

// staged route: smoothed depth is materialised (vrsbs_depth_from_full / vrsbs_depth_from_lowres)
int launch_depth(vrsbs_ctx *c, Scratch &s, const void *raw_v, const __half *lowres, int B, int H, int W, int h, int w,
                 float scaler, void *out_v, cudaStream_t st) {
    int rc = check_state_dims(c, H, W);
    if (rc) return rc;
    if ((rc = clear_counters(c, s, B, st))) return rc;
    if (c->f32) {
        DepthArgs32 a{};
        a.raw = static_cast<const float *>(raw_v); a.lowres = lowres; a.out = static_cast<float *>(out_v);
        a.hist1 = reinterpret_cast<float *>(c->hist[c->hist_idx][0]); a.hist2 = reinterpret_cast<float *>(c->hist[c->hist_idx][1]);
        a.frame_max = s.frame_max; a.frame_nan = s.frame_nan; a.sw = c->sw;
        a.B = B; a.H = H; a.W = W; a.h = h; a.w = w; a.first = c->depth_frames == 0; a.scaler = scaler;
        a.scale_y = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
        a.scale_x = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
        const size_t smem32 = sizeof(uint32_t) * 2 * B;
        dim3 grid((W + 31) / 32, (H + 7) / 8);
        StageTimer timer(c, st, 0);
        const size_t n32 = (size_t)H * W;
        if (raw_v && n32 % 4 == 0 && ((uintptr_t)raw_v % 16 == 0) && ((uintptr_t)out_v % 16 == 0) && c->f32_fast)
            k_depth_pass_f32<<<(unsigned)((n32 / 4 + 255) / 256), 256, smem32, st>>>(a);
        else if (raw_v) k_depth_f32<false, true><<<grid, 256, smem32, st>>>(a);
        else if (c->bicubic_contract) k_depth_f32<true, true><<<grid, 256, smem32, st>>>(a);
        else k_depth_f32<true, false><<<grid, 256, smem32, st>>>(a);
        CU_TRY(c, cudaGetLastError());
        c->launches++;
        c->depth_frames += B;
        c->state_h = H; c->state_w = W;
        return VRSBS_OK;
    }
    const __half *raw = static_cast<const __half *>(raw_v);
    __half *out = static_cast<__half *>(out_v);
    DepthArgs a{};
    a.raw = raw; a.lowres = lowres; a.out = out; a.hist1 = c->hist[c->hist_idx][0]; a.hist2 = c->hist[c->hist_idx][1];
    a.frame_max = s.frame_max; a.frame_nan = s.frame_nan; a.sw = c->sw;
    a.B = B; a.H = H; a.W = W; a.h = h; a.w = w; a.first = c->depth_frames == 0; a.scaler = scaler;
    a.scale_y = H > 1 ? (float)(h - 1) / (float)(H - 1) : 0.f;
    a.scale_x = W > 1 ? (float)(w - 1) / (float)(W - 1) : 0.f;
    const size_t smem = sizeof(uint32_t) * 2 * B;
    const size_t n = (size_t)H * W;
    StageTimer timer(c, st, 0);
    if (raw) {
        const bool vec = (n % 8 == 0) && ((uintptr_t)raw % 16 == 0) && ((uintptr_t)out % 16 == 0);
        if (vec) {
            // same kernel as the fused route's max pass, additionally storing the smoothed depth; history in place
            DepthMaxArgs m{};
            m.out = out; m.raw = raw; m.hist1 = a.hist1; m.hist2 = a.hist2; m.hist1_out = a.hist1; m.hist2_out = a.hist2;
            m.frame_max = s.frame_max; m.frame_nan = s.frame_nan; m.sw = c->sw; m.B = B; m.first = a.first; m.n = n;
            k_depth_pass<true><<<(unsigned)((n / 8 + 255) / 256), 256, smem, st>>>(m);
        } else {
            k_depth_full<1><<<(unsigned)((n + 255) / 256), 256, smem, st>>>(a);
        }
    } else {
        // tiled kernel: input rows a 16-row output tile can touch; very strong downscaling keeps the simple kernel
        const int rmax = (int)ceilf(a.scale_y * (kLrTileH - 1)) + 5, cmax = (int)ceilf(a.scale_x * (kLrTileW - 1)) + 5;
        const bool tiled = c->lowres_tiled && rmax * cmax <= 512 && (W % 2 == 0) && ((uintptr_t)out % 4 == 0);
        if (tiled) {
            dim3 grid((W + kLrTileW - 1) / kLrTileW, (H + kLrTileH - 1) / kLrTileH);
            const size_t tsmem = sizeof(float) * ((size_t)rmax * kLrTileW + 2 * (size_t)rmax * cmax) + smem;
            if (c->bicubic_contract) k_depth_lowres_tiled<true><<<grid, 256, tsmem, st>>>(a, rmax, cmax);
            else k_depth_lowres_tiled<false><<<grid, 256, tsmem, st>>>(a, rmax, cmax);
        } else {
            dim3 grid((W + 31) / 32, (H + 7) / 8);
            if (c->bicubic_contract) k_depth_lowres<true><<<grid, 256, smem, st>>>(a);
            else k_depth_lowres<false><<<grid, 256, smem, st>>>(a);
        }
    }
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    c->depth_frames += B;
    c->state_h = H; c->state_w = W;
    return VRSBS_OK;
}

// fused route, pass 1: per-frame max of the smoothed depth only; history goes to the other ping-pong set
int launch_depth_max(vrsbs_ctx *c, Scratch &s, const __half *raw, int B, int H, int W, cudaStream_t st) {
    int rc = check_state_dims(c, H, W);
    if (rc) return rc;
    if ((rc = clear_counters(c, s, B, st))) return rc;
    DepthMaxArgs a{};
    a.raw = raw; a.hist1 = c->hist[c->hist_idx][0]; a.hist2 = c->hist[c->hist_idx][1];
    a.hist1_out = c->hist[c->hist_idx ^ 1][0]; a.hist2_out = c->hist[c->hist_idx ^ 1][1];
    a.frame_max = s.frame_max; a.frame_nan = s.frame_nan; a.sw = c->sw;
    a.B = B; a.first = c->depth_frames == 0; a.n = (size_t)H * W;
    StageTimer timer(c, st, 0);
    k_depth_pass<false><<<(unsigned)((a.n / 8 + 255) / 256), 256, sizeof(uint32_t) * 2 * B, st>>>(a);
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return VRSBS_OK;
}

// Capacities of the fast-path tables, from the parameters alone (no device round trip): enough layers and
// LUT cells for limit_step <= 32; frames that need more take the slow path inside k_warp_fused.
void fast_caps(vrsbs_ctx *c, int H, int W, int *ent_cap, int *lut_cap) {
    const double span_per_limit = fabs(c->params.offset_fg - c->params.offset_bg) * H / 14.0;   // pixels of offset per unit of limit
    const int step = c->params.offset_step_size;
    double layers = span_per_limit * 32.0 / step + 6.0;
    int ec = layers > kEntCapMax ? kEntCapMax : (int)layers;
    if (ec < 8) ec = 8;
    // layer width in depth units ~ step / span_per_limit; a LUT cell must be narrower than ~0.9 of it
    const double width = span_per_limit > 0 ? 0.85 * 0.9 * step / span_per_limit : 1e9;
    int cells = 64;
    for (int e = -14; e <= 4; ++e) {                       // binade [2^e, 2^(e+1)) of the frame maximum, up to 32
        const double ulp = ldexp(1.0, e - 10);
        int sh = 0;
        while (sh < 9 && ulp * (2 << sh) <= width) ++sh;
        const int top_bits = (e + 15 + 1) << 10;           // fp16 bits of 2^(e+1)
        const int n = (top_bits >> sh) + 2;
        if (n > cells) cells = n;
    }
    if (cells > kLutCapMax) cells = kLutCapMax;
    if (cells > 5120 && ec > 128) cells = 5120;            // keep two 4K CTAs per SM
    // bound on the signed layer offsets for limit_step <= 32, in whole 32-pixel segments
    const double far = fabs(c->params.offset_fg) > fabs(c->params.offset_bg) ? fabs(c->params.offset_fg) : fabs(c->params.offset_bg);
    long long pad = ((long long)ceil(far * H * 32.0 / 14.0) + 2 + 31) / 32 * 32;
    // (the warp kernel relies on pad <= 32 * warps per CTA: 256 threads up to W = 2048, 512 above)
    if (pad * 4 > W || pad * 8 > 65535 || pad > 32 * (W <= 2048 ? 8 : 16)) { pad = 0; ec = 0; cells = 16; }   // slow path
    c->key_pad = (int)pad;
    if (!c->fast_tables) { ec = 0; cells = 16; }
    *ent_cap = ec;
    *lut_cap = (cells + 15) / 16 * 16;
}

int launch_tables(vrsbs_ctx *c, Scratch &s, int B, int H, int W, cudaStream_t st) {
    TableArgs a{};
    a.frame_max = s.frame_max; a.frame_nan = s.frame_nan;
    a.state_in = c->state + c->state_idx; a.state_out = c->state + (c->state_idx ^ 1);
    a.tabs = s.tabs; a.bounds = s.bounds; a.offm = s.offm; a.cutoffs = s.cutoffs; a.offsets = s.offsets;
    a.lo16 = s.lo16; a.hi16 = s.hi16;
    a.offset_fg = c->params.offset_fg; a.offset_bg = c->params.offset_bg; a.step = c->params.offset_step_size;
    a.B = B; a.H = H; a.W = W; a.Lcap = c->max_layers; a.f32 = c->f32;
    fast_caps(c, H, W, &c->ent_cap, &c->lut_cap);
    {
        const double span_per_limit = fabs(c->params.offset_fg - c->params.offset_bg) * H / 14.0;
        a.cell_width = span_per_limit > 0 ? (float)(0.85 * 0.9 * c->params.offset_step_size / span_per_limit) : 0.f;
    }
    a.blobs = s.blobs; a.ent_cap = c->ent_cap; a.lut_cap = c->lut_cap; a.key_pad = c->key_pad;
    const size_t smem = sizeof(double) * 2 * (size_t)(c->max_layers + 2 > B ? c->max_layers + 2 : B) +
                        sizeof(int) * (size_t)(c->max_layers + 2);
    StageTimer timer(c, st, 1);
    CU_TRY(c, launch_pdl((c->pdl & 1) != 0, k_build_tables, dim3(B), dim3(256), smem, st, a));
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    c->state_idx ^= 1;
    return VRSBS_OK;
}

template <int MODE, bool TMA, int NT, bool F32 = false>
int launch_warp_inst(vrsbs_ctx *c, const WarpArgs &a, cudaStream_t st) {
    auto kern = k_warp_rows<MODE, TMA, NT, F32>;
    const size_t smem = warp_smem_layout(a.W, a.Lcap, F32 ? 4 : 2).total;
    CU_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = c->blocks_per_sm;
    if (occ <= 0) CU_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    if (occ < 1) return fail(c, VRSBS_E_INVALID, "warp kernel does not fit: %zu B shared memory", smem);
    long long rows = (long long)a.B * a.H;
    long long grid = (long long)c->sm_count * occ;
    if (grid > rows) grid = rows;
    StageTimer timer(c, st, 2);
    kern<<<(unsigned)grid, NT, smem, st>>>(a);
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return VRSBS_OK;
}

template <int MODE, bool TMA, bool F32 = false>
int launch_warp_nt(vrsbs_ctx *c, const WarpArgs &a, cudaStream_t st) {
    if (a.W <= 2048) return launch_warp_inst<MODE, TMA, 256, F32>(c, a, st);
    if (a.W <= 4096) return launch_warp_inst<MODE, TMA, 512, F32>(c, a, st);
    return launch_warp_inst<MODE, TMA, 1024, F32>(c, a, st);
}

template <bool SMOOTH, int NT>
int launch_fused_inst(vrsbs_ctx *c, const FusedArgs &a, cudaStream_t st) {
    auto kern = k_warp_fused<SMOOTH, NT>;
    const size_t smem = fused_smem_layout(a.W, a.blob_bytes).total;
    CU_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    CU_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, smem));
    if (c->blocks_per_sm > 0 && c->blocks_per_sm < occ) occ = c->blocks_per_sm;
    if (occ < 1) return fail(c, VRSBS_E_INVALID, "fused warp kernel does not fit: %zu B shared memory", smem);
    long long iters = (long long)a.B * a.H;
    long long grid = (long long)c->sm_count * occ;
    if (grid > iters) grid = iters;
    StageTimer timer(c, st, 2);
    kern<<<(unsigned)grid, NT, smem, st>>>(a);
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return VRSBS_OK;
}

// warp-specialised variant: 5 scatter + 3 destination warps; used when it reaches the same residency as k_warp_fused
// (4 CTAs per SM), the mask row fits the destination warps and the key-row bound fits the scatter warps
template <int NT, int NS, bool F32 = false>
int launch_ws_inst(vrsbs_ctx *c, const FusedArgs &a, cudaStream_t st, bool *launched) {
    *launched = false;
    const int wwords32 = ((a.W + 31) / 32 + 31) / 32 * 32;
    if (a.W % 32 != 0 || a.key_pad > 32 * NS || wwords32 > (NT / 32 - NS) * 32) return VRSBS_OK;
    WsArgs w{};
    w.f = a;
    w.lay = ws_smem_layout(a.W, a.blob_bytes, F32 ? 4 : 2);
    auto kern = k_warp_ws<NT, NS, F32>;
    if (w.lay.total > 200 * 1024) return VRSBS_OK;
    CU_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)w.lay.total));
    int occ = 0;
    CU_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, NT, w.lay.total));
    if (occ * NT < (F32 ? 512 : 768)) return VRSBS_OK;    // too few resident warps per SM: k_warp_fused / k_warp_rows do better (three
                                                          // 8-warp CTAs still beat k_warp_fused: 0.315 vs 0.338 ms per 64 frames of 1080p)
    if (c->blocks_per_sm > 0 && c->blocks_per_sm < occ) occ = c->blocks_per_sm;
    long long iters = (long long)a.B * a.H, grid = (long long)c->sm_count * occ;
    if (grid > iters) grid = iters;
    StageTimer timer(c, st, 2);
    CU_TRY(c, launch_pdl((c->pdl & 2) != 0, kern, dim3((unsigned)grid), dim3(NT), w.lay.total, st, w));
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    *launched = true;
    return VRSBS_OK;
}
int launch_ws(vrsbs_ctx *c, const FusedArgs &a, cudaStream_t st, bool *launched) {
    if (c->f32) return a.W > 2048 ? launch_ws_inst<512, 8, true>(c, a, st, launched) : launch_ws_inst<256, 4, true>(c, a, st, launched);
    if (a.W > 2048) {                                  // 4K rows: 2 CTAs of 16 warps per SM (shared memory bound)
        switch (c->ws_scatter_warps) {
            case 3: return launch_ws_inst<512, 6>(c, a, st, launched);
            case 5: return launch_ws_inst<512, 10>(c, a, st, launched);
            default: return launch_ws_inst<512, 8>(c, a, st, launched);
        }
    }
    switch (c->ws_scatter_warps) {
        case 3: return launch_ws_inst<256, 3>(c, a, st, launched);
        case 4: return launch_ws_inst<256, 4>(c, a, st, launched);
        case 5: return launch_ws_inst<256, 5>(c, a, st, launched);
        default: return launch_ws_inst<256, 4>(c, a, st, launched);
    }
}

bool fused_capable(const vrsbs_ctx *c, const void *frames, const void *depth, const void *sbs, int W) {
    const size_t smem = fused_smem_layout(W, blob_bytes(c->ent_cap, c->lut_cap)).total;
    return (W % 16 == 0) && W * 4 <= 65535 && ((uintptr_t)frames % 16 == 0) && ((uintptr_t)depth % 16 == 0) &&
           ((uintptr_t)sbs % 16 == 0) && smem <= 200 * 1024;
}

// hole values -> SBS frame (when blur is on) and strip restore, one kernel
int launch_commit(vrsbs_ctx *c, const BlurArgs &b, int do_commit, cudaStream_t st) {
    StageTimer timer(c, st, 4);
    CU_TRY(c, launch_pdl((c->pdl & 8) != 0, k_blur_commit, dim3((unsigned)(c->sm_count * 32)), dim3(256), 0, st, b, do_commit));   // short latency-bound tasks: one or two per warp
    CU_TRY(c, cudaGetLastError());
    c->launches++;
    return VRSBS_OK;
}

template <int PARTS, int CX, int CY>
int launch_blur_fixed(vrsbs_ctx *c, const BlurArgs &b, cudaStream_t st) {
    BlurWeights<PARTS, CX, CY> wts;
    memcpy(wts.q, c->wq_host.data(), sizeof(wts.q));
    if (c->blur_screen) {
        memcpy(wts.h, c->wh_host.data(), sizeof(wts.h));
        wts.s1 = c->ws1; wts.rmax = c->wrmax;
    } else {                                           // screening sum 0 is never decisive: every hole takes the exact path
        memset(wts.h, 0, sizeof(wts.h));
        wts.s1 = 1; wts.rmax = 2;
    }
    auto kern = k_blur_holes_fixed<PARTS, CX, CY>;
    const size_t per_warp = blur_fixed_warp_smem<CX, CY>();
    const int warps = 8;
    const size_t bsmem = per_warp * warps;
    CU_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
    int occ = 0;
    CU_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, bsmem));
    if (occ < 1) occ = 1;
    CU_TRY(c, launch_pdl((c->pdl & 4) != 0, kern, dim3((unsigned)(c->sm_count * occ)), dim3(warps * 32), bsmem, st, b, wts));
    return VRSBS_OK;
}

template <int PARTS, int CX, int CY>
int launch_blur_sep(vrsbs_ctx *c, const BlurArgs &b, cudaStream_t st) {
    BlurSepWeights<CX, CY> wts;
    memcpy(wts.hy, c->sep_hy.data(), sizeof(wts.hy));
    memcpy(wts.hx, c->sep_hx.data(), sizeof(wts.hx));
    wts.s = c->sep_s; wts.eps32 = c->blur_sep == 2 ? 0x7fffffffu : c->sep_eps32;   // blur_sep = 2 (tests): every value takes the exact path
    auto kern = k_blur_sep<PARTS, CX, CY>;
    const int warps = 8;
    const size_t bsmem = blur_sep_warp_smem<CX>() * warps;
    CU_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
    int occ = 0;
    CU_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, bsmem));
    if (occ < 1) occ = 1;
    CU_TRY(c, launch_pdl((c->pdl & 4) != 0, kern, dim3((unsigned)(c->sm_count * occ)), dim3(warps * 32), bsmem, st, b, wts));
    return VRSBS_OK;
}

// true when launch_blur will run the band-driven kernel for this frame size (the warp kernel then lists the band columns itself)
bool band_route(const vrsbs_ctx *c, int B, int H, int W, const uint8_t *sbs) {
    const int cx = c->kx / 2, cy = c->ky / 2;
    const bool sep = c->blur_sep && c->blur_screen && !c->sep_hy.empty();
    const bool aligned = (W % 2 == 0) && ((uintptr_t)sbs % 4 == 0);
    const bool built = (cx == 5 && cy == 4) || (cx == 4 && cy == 3);
    return c->params.blur && c->blur_band && sep && aligned && built && (c->wparts == 2 || c->wparts == 3) &&
           (long long)B * ((H + kBandRows - 1) / kBandRows) < (1 << 24) && (W + 31) / 32 <= 256;
}

// band-driven blur: the list of band columns that hold a hole (unless the warp kernel made it), then one warp per band column
template <int PARTS, int CX, int CY>
int launch_blur_band(vrsbs_ctx *c, const BlurArgs &b, cudaStream_t st) {
    BlurSepWeights<CX, CY> wts;
    memcpy(wts.hy, c->sep_hy.data(), sizeof(wts.hy));
    memcpy(wts.hx, c->sep_hx.data(), sizeof(wts.hx));
    wts.s = c->sep_s; wts.eps32 = c->blur_sep == 2 ? 0x7fffffffu : c->sep_eps32;
    if (b.band_prepass) {
        const long long cells = (long long)b.B * b.Hb * b.Wwords;
        CU_TRY(c, launch_pdl((c->pdl & 4) != 0, k_band_map, dim3((unsigned)((cells + 255) / 256)), dim3(256), 0, st, b));
        c->launches++;
    }
    {
        const long long words = (long long)b.B * b.Hb * b.band_groups;
        CU_TRY(c, launch_pdl((c->pdl & 4) != 0, k_band_list, dim3((unsigned)((words + 255) / 256)), dim3(256), 0, st, b));
        c->launches++;
    }
    auto kern = k_blur_band<PARTS, CX, CY>;
    const int warps = 8;
    const size_t bsmem = blur_sep_warp_smem<CX>() * warps;
    CU_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
    int occ = 0;
    CU_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, bsmem));
    if (occ < 1) occ = 1;
    CU_TRY(c, launch_pdl((c->pdl & 4) != 0, kern, dim3((unsigned)(c->sm_count * occ)), dim3(warps * 32), bsmem, st, b, wts));
    return VRSBS_OK;
}

int launch_blur(vrsbs_ctx *c, Scratch &s, const uint8_t *frames, int B, int H, int W, uint8_t *sbs, cudaStream_t st) {
    if (!s.plane) CU_TRY(c, dmalloc(&s.plane, (size_t)s.cap_batch * c->max_h * c->max_w * 3));
    BlurArgs b{};
    b.frames = frames; b.sbs = sbs; b.tabs = s.tabs; b.hole_mask = s.hole_mask; b.hole_list = s.hole_list;
    b.hole_count = s.hole_count; b.ticket = s.hole_count + 1; b.plane = s.plane; b.wq = c->wq; b.weights = c->weights;
    b.band_list = s.band_list; b.band_count = s.hole_count + 2; b.Hb = (H + kBandRows - 1) / kBandRows;
    b.magic_hb = ((1ull << 40) + (unsigned long long)b.Hb - 1) / (unsigned long long)b.Hb;
    b.band_groups = ((W + 31) / 32 + 31) / 32;
    const bool band = band_route(c, B, H, W, sbs);
    b.band_map = band ? s.band_map : nullptr;
    b.band_prepass = (band && !s.band_marked) ? 1 : 0;
    s.band_marked = false;
    b.B = B; b.H = H; b.W = W; b.Wwords = (W + 31) / 32; b.kx = c->kx; b.ky = c->ky; b.wshift = c->wshift;
    b.magic_h = ((1ull << 40) + (unsigned long long)H - 1) / (unsigned long long)H;
    const int cx = c->kx / 2, cy = c->ky / 2;
    const bool aligned = (W % 2 == 0) && ((uintptr_t)sbs % 4 == 0);
    bool band_done = false;
    {
        StageTimer timer(c, st, 3);
        bool done = false;
        const bool sep = c->blur_sep && c->blur_screen && !c->sep_hy.empty();
#define VRSBS_BLUR_BAND(P, CXv, CYv)                                                    \
        if (!done && band && c->wparts == P && cx == CXv && cy == CYv) {                \
            int rc = launch_blur_band<P, CXv, CYv>(c, b, st);                           \
            if (rc) return rc;                                                          \
            done = true; band_done = true;                                              \
        }
        VRSBS_BLUR_BAND(2, 5, 4)       // 1080p: 11 x 9
        VRSBS_BLUR_BAND(3, 5, 4)
        VRSBS_BLUR_BAND(2, 4, 3)       // 720p: 9 x 7
        VRSBS_BLUR_BAND(3, 4, 3)
#undef VRSBS_BLUR_BAND
        if (!done && s.list_missing) {                   // the list-driven kernels below need the per-word list
            const long long nw = (long long)B * H * b.Wwords;
            CU_TRY(c, launch_pdl((c->pdl & 4) != 0, k_word_list, dim3((unsigned)((nw + 4095) / 4096)), dim3(1024), 0, st,
                                 (const uint32_t *)s.hole_mask, s.hole_list, s.hole_count, nw, b.Wwords));
            c->launches++;
        }
        s.list_missing = false;
#define VRSBS_BLUR_SEP(P, CXv, CYv)                                                     \
        if (!done && sep && aligned && c->wparts == P && cx == CXv && cy == CYv) {      \
            int rc = launch_blur_sep<P, CXv, CYv>(c, b, st);                            \
            if (rc) return rc;                                                          \
            done = true;                                                                \
        }
        VRSBS_BLUR_SEP(2, 5, 4)        // 1080p: 11 x 9
        VRSBS_BLUR_SEP(3, 5, 4)
        VRSBS_BLUR_SEP(2, 9, 8)        // 4K: 19 x 17
        VRSBS_BLUR_SEP(3, 9, 8)
        VRSBS_BLUR_SEP(2, 4, 3)        // 720p: 9 x 7
        VRSBS_BLUR_SEP(3, 4, 3)
        VRSBS_BLUR_SEP(2, 6, 5)        // 1440p: 13 x 11
        VRSBS_BLUR_SEP(3, 6, 5)
#undef VRSBS_BLUR_SEP
#define VRSBS_BLUR_FIXED(P, CXv, CYv)                                                   \
        if (!done && aligned && c->wparts == P && cx == CXv && cy == CYv) {             \
            int rc = launch_blur_fixed<P, CXv, CYv>(c, b, st);                          \
            if (rc) return rc;                                                          \
            done = true;                                                                \
        }
        VRSBS_BLUR_FIXED(2, 5, 4)      // 1080p: 11 x 9
        VRSBS_BLUR_FIXED(3, 5, 4)
        VRSBS_BLUR_FIXED(2, 9, 8)      // 4K: 19 x 17
        VRSBS_BLUR_FIXED(3, 9, 8)
        VRSBS_BLUR_FIXED(2, 4, 3)      // 720p: 9 x 7
        VRSBS_BLUR_FIXED(3, 4, 3)
        VRSBS_BLUR_FIXED(2, 6, 5)      // 1440p: 13 x 11
        VRSBS_BLUR_FIXED(3, 6, 5)
#undef VRSBS_BLUR_FIXED
        if (!done) {
            const size_t per_warp = blur_warp_smem(c->kx, c->ky, c->wparts > 0);
            int warps = 8;
            while (warps > 1 && per_warp * warps > 96 * 1024) warps >>= 1;
            const size_t bsmem = per_warp * warps;
            if (bsmem > 200 * 1024) return fail(c, VRSBS_E_INVALID, "blur kernel %dx%d needs %zu B shared memory per warp", c->kx, c->ky, per_warp);
            auto kern = c->wparts == 2 ? k_blur_holes<2> : (c->wparts == 3 ? k_blur_holes<3> : k_blur_holes<0>);
            CU_TRY(c, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bsmem));
            int occ = 0;
            CU_TRY(c, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, warps * 32, bsmem));
            if (occ < 1) occ = 1;
            kern<<<(unsigned)(c->sm_count * occ), warps * 32, bsmem, st>>>(b);
        }
        CU_TRY(c, cudaGetLastError());
        c->launches++;
    }
    return launch_commit(c, b, band_done ? (1 | 4) : c->commit_mode, st);   // (the band route has no per-word list to commit from)
}

int check_blur_ready(vrsbs_ctx *c, int H, int W) {
    if (c->params.blur && (c->kx <= 0 || c->ky <= 0))
        return fail(c, VRSBS_E_STATE, "vrsbs_set_blur_weights must be called before the warp");
    if (c->params.blur && (c->kx / 2 >= W || c->ky / 2 >= H))
        return fail(c, VRSBS_E_INVALID, "blur kernel %dx%d too large for a %dx%d frame (reflect padding)", c->kx, c->ky, W, H);
    return VRSBS_OK;
}

FusedArgs make_fused_args(vrsbs_ctx *c, Scratch &s, const uint8_t *frames, const __half *depth, int B, int H, int W, uint8_t *sbs) {
    FusedArgs a{};
    a.frames = frames; a.depth = depth; a.hist1 = c->hist[c->hist_idx][0]; a.hist2 = c->hist[c->hist_idx][1];
    a.sbs = sbs; a.blobs = s.blobs; a.tabs = s.tabs; a.bounds = s.bounds; a.offm = s.offm;
    a.hole_mask = s.hole_mask; a.hole_list = s.hole_list; a.hole_count = s.hole_count;
    a.B = B; a.H = H; a.W = W; a.Lcap = c->max_layers; a.Wwords = (W + 31) / 32;
    a.band_map = nullptr;
    a.Hb = (H + kBandRows - 1) / kBandRows; a.band_groups = (a.Wwords + 31) / 32;
    a.first = 0;
    a.skip_right = c->skip_right;
    a.blob_bytes = blob_bytes(c->ent_cap, c->lut_cap, c->f32 != 0); a.ent_bytes = blob_ent_bytes(c->ent_cap, c->f32 != 0); a.key_pad = c->key_pad;
    a.w0 = c->sw.w_now; a.w1 = c->sw.w_prev1; a.w2 = c->sw.w_prev2;
    const FusedSmem L = fused_smem_layout(W, a.blob_bytes);
    a.lay.img = (uint32_t)L.img; a.lay.img_stride = (uint32_t)L.img_stride; a.lay.dep = (uint32_t)L.dep; a.lay.dep_stride = (uint32_t)L.dep_stride;
    a.lay.out = (uint32_t)L.out; a.lay.keys = (uint32_t)L.keys; a.lay.blob = (uint32_t)L.blob; a.lay.blob_stride = (uint32_t)L.blob_stride;
    a.lay.mask = (uint32_t)L.mask; a.lay.bars = (uint32_t)L.bars;
    return a;
}

// staged route, stage 3: depth is the SMOOTHED depth left by launch_depth
int launch_warp(vrsbs_ctx *c, Scratch &s, const uint8_t *frames, const void *depth_v, int B, int H, int W, uint8_t *sbs,
                cudaStream_t st) {
    int rc = check_blur_ready(c, H, W);
    if (rc) return rc;
    if ((rc = fresh_hole_list(c, s, st))) return rc;
    const __half *depth = static_cast<const __half *>(depth_v);
    if (c->f32) {                                           // fp32 depth, comparison in fp32
        const bool al = (W % 32 == 0) && ((uintptr_t)frames % 16 == 0) && ((uintptr_t)depth_v % 16 == 0) && ((uintptr_t)sbs % 16 == 0);
        if (c->f32_fast && c->fused && c->warp_ws && al && c->ent_cap > 0) {      // the warp-specialised kernel's fp32 instantiation
            FusedArgs fa = make_fused_args(c, s, frames, depth, B, H, W, sbs);
            if (band_route(c, B, H, W, sbs)) fa.band_map = s.band_map;
            if (c->params.blur && c->ws_no_list) fa.hole_list = nullptr;
            bool done = false;
            if ((rc = launch_ws(c, fa, st, &done))) return rc;
            s.band_marked = done && fa.band_map;
            s.list_missing = done && !fa.band_map && !fa.hole_list;
            if (done) {
                if (!c->params.blur) return VRSBS_OK;
                return launch_blur(c, s, frames, B, H, W, sbs, st);
            }
        }
        WarpArgs a{};                                       // the general row kernel
        a.frames = frames; a.depth = depth_v; a.sbs = sbs; a.tabs = s.tabs; a.bounds = s.bounds; a.offm = s.offm;
        a.hole_mask = s.hole_mask; a.B = B; a.H = H; a.W = W; a.Lcap = c->max_layers; a.Wwords = (W + 31) / 32;
        a.hole_list = s.hole_list; a.hole_count = s.hole_count;
        const bool tma = (W % 16 == 0) && ((uintptr_t)frames % 16 == 0) && ((uintptr_t)depth_v % 16 == 0) && ((uintptr_t)sbs % 16 == 0);
        rc = tma ? launch_warp_nt<2, true, true>(c, a, st) : launch_warp_nt<2, false, true>(c, a, st);
        if (rc) return rc;
        if (!c->params.blur) return VRSBS_OK;
        return launch_blur(c, s, frames, B, H, W, sbs, st);
    }
    if (c->fused && fused_capable(c, frames, depth, sbs, W)) {
        FusedArgs a = make_fused_args(c, s, frames, depth, B, H, W, sbs);
        if (c->warp_ws && band_route(c, B, H, W, sbs)) a.band_map = s.band_map;
        if (c->warp_ws && c->params.blur && c->ws_no_list) a.hole_list = nullptr;
        bool done = false;
        if (c->warp_ws) rc = launch_ws(c, a, st, &done);
        s.band_marked = !rc && done && a.band_map;
        s.list_missing = !rc && done && !a.band_map && !a.hole_list;
        if (!done) { a.band_map = nullptr; a.hole_list = s.hole_list; }    // the kernels below keep the per-word list
        if (!rc && !done) rc = W <= 2048 ? launch_fused_inst<false, 256>(c, a, st) : launch_fused_inst<false, 512>(c, a, st);
    } else {
        WarpArgs a{};
        a.frames = frames; a.depth = depth; a.sbs = sbs; a.tabs = s.tabs; a.bounds = s.bounds; a.offm = s.offm;
        a.hole_mask = s.hole_mask; a.B = B; a.H = H; a.W = W; a.Lcap = c->max_layers; a.Wwords = (W + 31) / 32;
        a.hole_list = s.hole_list; a.hole_count = s.hole_count;
        const bool tma = (W % 16 == 0) && ((uintptr_t)frames % 16 == 0) && ((uintptr_t)depth % 16 == 0) &&
                         ((uintptr_t)sbs % 16 == 0);
        if (c->scatter_mode == 2) rc = tma ? launch_warp_nt<2, true>(c, a, st) : launch_warp_nt<2, false>(c, a, st);
        else rc = tma ? launch_warp_nt<1, true>(c, a, st) : launch_warp_nt<1, false>(c, a, st);
    }
    if (rc) return rc;
    if (!c->params.blur) return VRSBS_OK;
    return launch_blur(c, s, frames, B, H, W, sbs, st);
}

// fused route: raw depth in, smoothing recomputed inside the warp kernel (no smoothed depth in HBM)
int launch_process_fused(vrsbs_ctx *c, Scratch &s, const uint8_t *frames, const __half *raw, int B, int H, int W,
                         uint8_t *sbs, cudaStream_t st) {
    int rc = check_blur_ready(c, H, W);
    if (rc) return rc;
    const bool first = c->depth_frames == 0;
    if ((rc = launch_depth_max(c, s, raw, B, H, W, st))) return rc;
    if ((rc = launch_tables(c, s, B, H, W, st))) return rc;
    if ((rc = fresh_hole_list(c, s, st))) return rc;
    FusedArgs a = make_fused_args(c, s, frames, raw, B, H, W, sbs);
    a.first = first ? 1 : 0;
    rc = W <= 2048 ? launch_fused_inst<true, 256>(c, a, st) : launch_fused_inst<true, 512>(c, a, st);
    if (rc) return rc;
    c->hist_idx ^= 1;                        // k_depth_pass wrote the next batch's history into the other set
    c->depth_frames += B;
    c->state_h = H; c->state_w = W;
    if (!c->params.blur) return VRSBS_OK;
    return launch_blur(c, s, frames, B, H, W, sbs, st);
}

bool is_pinned(const void *p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

bool is_device(const void *p) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeDevice;
}

int ensure_slot(vrsbs_ctx *c, HostSlot &s, size_t frames_b, size_t depth_in_b, size_t depth_b, size_t sbs_b, bool need_pin_in,
                bool need_pin_out) {
    if (!c->st_in) {
        CU_TRY(c, cudaStreamCreateWithFlags(&c->st_in, cudaStreamNonBlocking));
        CU_TRY(c, cudaStreamCreateWithFlags(&c->st_k, cudaStreamNonBlocking));
        CU_TRY(c, cudaStreamCreateWithFlags(&c->st_out, cudaStreamNonBlocking));
    }
    if (!s.in_done) {
        CU_TRY(c, cudaEventCreateWithFlags(&s.in_done, cudaEventDisableTiming));
        CU_TRY(c, cudaEventCreateWithFlags(&s.k_done, cudaEventDisableTiming));
        CU_TRY(c, cudaEventCreateWithFlags(&s.out_done, cudaEventDisableTiming));
        CU_TRY(c, cudaHostAlloc(reinterpret_cast<void **>(&s.pin_tabs), sizeof(FrameTab) * c->max_batch, cudaHostAllocDefault));
    }
    auto grow_dev = [&](uint8_t *&dev, size_t &cap, size_t need) -> cudaError_t {
        if (need <= cap) return cudaSuccess;
        cudaFree(dev); dev = nullptr; cap = 0;
        cudaError_t e = cudaMalloc(reinterpret_cast<void **>(&dev), need);
        if (e == cudaSuccess) cap = need;
        return e;
    };
    auto grow_pin = [&](uint8_t *&pin, size_t &cap, size_t need) -> cudaError_t {
        if (need <= cap) return cudaSuccess;
        cudaFreeHost(pin); pin = nullptr; cap = 0;
        cudaError_t e = cudaHostAlloc(reinterpret_cast<void **>(&pin), need, cudaHostAllocDefault);
        if (e == cudaSuccess) cap = need;
        return e;
    };
    CU_TRY(c, grow_dev(s.dev_frames, s.cap_frames, frames_b));
    CU_TRY(c, grow_dev(s.dev_depth_in, s.cap_depth_in, depth_in_b));
    CU_TRY(c, grow_dev(s.dev_depth, s.cap_depth, depth_b));
    CU_TRY(c, grow_dev(s.dev_sbs, s.cap_sbs, sbs_b));
    if (need_pin_in) {
        CU_TRY(c, grow_pin(s.pin_frames, s.cap_pin_frames, frames_b));
        CU_TRY(c, grow_pin(s.pin_depth, s.cap_pin_depth, depth_in_b));
    }
    if (need_pin_out) CU_TRY(c, grow_pin(s.pin_sbs, s.cap_pin_sbs, sbs_b));
    return VRSBS_OK;
}

void free_slot(HostSlot &s) {
    if (s.in_done) cudaEventDestroy(s.in_done);
    if (s.k_done) cudaEventDestroy(s.k_done);
    if (s.out_done) cudaEventDestroy(s.out_done);
    cudaFree(s.dev_frames); cudaFree(s.dev_depth_in); cudaFree(s.dev_depth); cudaFree(s.dev_sbs);
    cudaFreeHost(s.pin_frames); cudaFreeHost(s.pin_depth); cudaFreeHost(s.pin_sbs); cudaFreeHost(s.pin_tabs);
    s = HostSlot{};
}

int frame_status_error(vrsbs_ctx *c, const FrameTab *tabs, int B, int first_index) {
    for (int i = 0; i < B; ++i) {
        if (tabs[i].status & VRSBS_FRAME_NAN)
            return fail(c, VRSBS_E_FRAME, "frame %d: depth.max() is NaN (the reference raises in math.ceil)", first_index + i);
        if (tabs[i].status & VRSBS_FRAME_OVERFLOW)
            return fail(c, VRSBS_E_FRAME, "frame %d: layer count exceeds max_layers=%d (limit_step=%d)", first_index + i,
                        c->max_layers, tabs[i].limit_step);
    }
    return VRSBS_OK;
}

// The device-pointer calls run on the caller's stream, the host pipeline on the context's own streams; both share the depth
// history and the range state.  A host call that follows device-pointer calls orders its kernel stream behind them.
int mixed_entry_sync(vrsbs_ctx *c) {
    if (!c->dev_state_dirty || !c->st_k) { c->dev_state_dirty = c->dev_state_dirty && !c->st_k; return VRSBS_OK; }
    if (!c->dep_event) CU_TRY(c, cudaEventCreateWithFlags(&c->dep_event, cudaEventDisableTiming));
    CU_TRY(c, cudaEventRecord(c->dep_event, c->last_dev_stream));
    CU_TRY(c, cudaStreamWaitEvent(c->st_k, c->dep_event, 0));
    c->dev_state_dirty = false;
    return VRSBS_OK;
}

}  // namespace

// =========================================================================================================
extern "C" {

int vrsbs_abi_version(void) { return VRSBS_ABI_VERSION; }

const char *vrsbs_last_error(const vrsbs_ctx *ctx) { return ctx ? ctx->err : g_create_error; }

int vrsbs_create(vrsbs_ctx **out, int device, int max_h, int max_w, int max_batch, int max_layers) {
    if (!out) return fail(nullptr, VRSBS_E_INVALID, "out is NULL");
    *out = nullptr;
    if (max_h < 2 || max_w < 2 || max_w > 8192 || max_batch < 1 || max_layers < 1 || max_layers > 4096)
        return fail(nullptr, VRSBS_E_INVALID, "limits out of range (max_w <= 8192, max_layers <= 4096)");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, VRSBS_E_CUDA, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, VRSBS_E_INVALID, "device %d of %d", device, ndev);
    DeviceGuard g(device);
    if (!g.ok) return fail(nullptr, VRSBS_E_CUDA, "cudaSetDevice(%d) failed", device);
    cudaDeviceProp prop{};
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess)
        return fail(nullptr, VRSBS_E_CUDA, "cudaGetDeviceProperties: %s", cudaGetErrorString(e));
    if (prop.major != 10)
        return fail(nullptr, VRSBS_E_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                    prop.minor);
    vrsbs_ctx *c = new (std::nothrow) vrsbs_ctx();
    if (!c) return fail(nullptr, VRSBS_E_NOMEM, "out of host memory");
    c->device = device; c->max_h = max_h; c->max_w = max_w; c->max_batch = max_batch; c->max_layers = max_layers;
    c->sm_count = prop.multiProcessorCount;
    auto init = [&]() -> int {
        const size_t n = (size_t)max_h * max_w;
        for (int i = 0; i < 2; ++i)
            for (int j = 0; j < 2; ++j) CU_TRY(c, dmalloc(&c->hist[i][j], 2 * n));   // fp16 or fp32 raw depth
        CU_TRY(c, dmalloc(&c->state, 2));
        int rc = alloc_scratch(c, c->scratch[0], c->max_batch);
        if (rc) return rc;
        vrsbs_params p = c->params;
        return vrsbs_reset(c, &p);
    };
    int rc = init();
    if (rc) {
        snprintf(g_create_error, sizeof g_create_error, "%s", c->err);
        vrsbs_destroy(c);
        return rc;
    }
    *out = c;
    return VRSBS_OK;
}

int vrsbs_destroy(vrsbs_ctx *c) {
    if (!c) return VRSBS_OK;
    DeviceGuard g(c->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j) cudaFree(c->hist[i][j]);
    cudaFree(c->state); cudaFree(c->weights); cudaFree(c->wq);
    for (int i = 0; i < kSlots; ++i) { free_scratch(c->scratch[i]); free_slot(c->slot[i]); }
    if (c->st_in) { cudaStreamDestroy(c->st_in); cudaStreamDestroy(c->st_k); cudaStreamDestroy(c->st_out); }
    if (c->dep_event) cudaEventDestroy(c->dep_event);
    delete c->pool;
    for (auto &s : c->stamps) { cudaEventDestroy(s.a); cudaEventDestroy(s.b); }
    for (auto e : c->event_pool) cudaEventDestroy(e);
    delete c;
    return VRSBS_OK;
}

int vrsbs_reset(vrsbs_ctx *c, const vrsbs_params *p) {
    if (!c) return VRSBS_E_INVALID;
    DeviceGuard g(c->device);
    if (p) {
        if (p->offset_step_size < 1) return fail(c, VRSBS_E_INVALID, "offset_step_size must be >= 1");
        if (p->depth_dtype != VRSBS_DEPTH_F16 && p->depth_dtype != VRSBS_DEPTH_F32) return fail(c, VRSBS_E_INVALID, "depth_dtype must be VRSBS_DEPTH_F16 or VRSBS_DEPTH_F32");
        if (!c->inflight.empty()) return fail(c, VRSBS_E_STATE, "vrsbs_reset with submitted batches not yet collected");
        c->params = *p;
        c->f32 = p->depth_dtype == VRSBS_DEPTH_F32;
    }
    // SbsProcessor.__init__: t = 0.3; sum += t; t *= 0.4 (twice); weight of the current frame = 1 - sum
    double t = 0.3, acc = 0.0, taps[2];
    for (int i = 0; i < 2; ++i) { acc = acc + t; taps[i] = t; t = t * 0.4; }
    c->sw.w_now = (float)(1 - acc);
    c->sw.w_prev1 = (float)taps[0];
    c->sw.w_prev2 = (float)taps[1];
    CU_TRY(c, cudaDeviceSynchronize());
    CU_TRY(c, cudaMemset(c->state, 0, sizeof(RangeState) * 2));
    c->state_idx = 0;
    c->depth_frames = 0;
    c->state_h = c->state_w = 0;
    return VRSBS_OK;
}

int vrsbs_get_range_state(vrsbs_ctx *c, int *has_last, double range[2]) {
    if (!c || !has_last || !range) return fail(c, VRSBS_E_INVALID, "NULL argument");
    DeviceGuard g(c->device);
    CU_TRY(c, cudaDeviceSynchronize());
    RangeState st;
    CU_TRY(c, cudaMemcpy(&st, c->state + c->state_idx, sizeof st, cudaMemcpyDeviceToHost));
    *has_last = st.has_last; range[0] = st.range[0]; range[1] = st.range[1];
    return VRSBS_OK;
}

int vrsbs_set_range_state(vrsbs_ctx *c, int has_last, const double range[2]) {
    if (!c || (has_last && !range)) return fail(c, VRSBS_E_INVALID, "NULL argument");
    DeviceGuard g(c->device);
    CU_TRY(c, cudaDeviceSynchronize());
    RangeState st{};
    st.has_last = has_last ? 1 : 0;
    if (has_last) { st.range[0] = range[0]; st.range[1] = range[1]; }
    CU_TRY(c, cudaMemcpy(c->state + c->state_idx, &st, sizeof st, cudaMemcpyHostToDevice));
    return VRSBS_OK;
}

int vrsbs_set_blur_weights(vrsbs_ctx *c, const float *w, int kx, int ky) {
    if (!c || !w || kx < 1 || ky < 1 || kx > 255 || ky > 255 || !(kx & 1) || !(ky & 1))
        return fail(c, VRSBS_E_INVALID, "bad blur kernel %dx%d", kx, ky);
    DeviceGuard g(c->device);
    CU_TRY(c, cudaDeviceSynchronize());
    cudaFree(c->weights); c->weights = nullptr;
    cudaFree(c->wq); c->wq = nullptr;
    CU_TRY(c, dmalloc(&c->weights, (size_t)kx * ky));
    CU_TRY(c, cudaMemcpy(c->weights, w, sizeof(float) * kx * ky, cudaMemcpyHostToDevice));
    c->kx = kx; c->ky = ky; c->wparts = 0; c->wshift = 0;
    // Integer path (blur_holes.cuh): 4-fold symmetric, positive weights whose common scale 2^S makes them
    // integers that split into 2 x 15 or 3 x 13 bits without overflowing 32-bit accumulators.
    const int cx = kx / 2, cy = ky / 2, nu = (cx + 1) * (cy + 1);
    bool sym = true;
    for (int i = 0; i < ky && sym; ++i)
        for (int j = 0; j < kx; ++j) {
            const float v = w[i * kx + j];
            if (!(v >= 0.f) || !std::isfinite(v) || v != w[(ky - 1 - i) * kx + j] || v != w[i * kx + (kx - 1 - j)]) { sym = false; break; }
        }
    if (sym) {
        int S = -1;
        for (int s = 1; s <= 62 && S < 0; ++s) {
            bool all = true;
            for (int i = 0; i < kx * ky && all; ++i) { const double v = ldexp((double)w[i], s); all = v == floor(v); }
            if (all) S = s;
        }
        double wmax = 0;
        for (int i = 0; i < kx * ky; ++i) wmax = w[i] > wmax ? w[i] : wmax;
        int parts = 0;
        if (S > 0) {
            const double qmax = ldexp(wmax, S);
            if (qmax < ldexp(1.0, 30) && nu <= 100) parts = 2;
            else if (qmax < ldexp(1.0, 39) && nu <= 400) parts = 3;
        }
        if (parts) {
            const int pbits = parts == 2 ? 15 : 13;
            std::vector<uint32_t> q((size_t)parts * nu);
            for (int i = 0; i <= cy; ++i)
                for (int j = 0; j <= cx; ++j) {
                    unsigned long long v = (unsigned long long)ldexp((double)w[(cy - i) * kx + (cx - j)], S);
                    for (int p = 0; p < parts; ++p) { q[(size_t)p * nu + i * (cx + 1) + j] = (uint32_t)(v & ((1ull << pbits) - 1ull)); v >>= pbits; }
                }
            c->wq_host = q;
            // screening weights (blur_holes.cuh): h = floor(w * 2^s1) with 255 * sum(h) < 2^32, and the bound
            // rmax >= (exact total - screening total) in units of 2^-s1
            std::vector<uint32_t> hq(nu, 0u);
            uint32_t s1 = 1, rmax = 2;
            for (int t = S < 24 ? S : 24; t >= 2; --t) {
                unsigned long long sum_h = 0, sum_r = 0;
                const int drop = S - t;
                for (int i = 0; i <= cy; ++i)
                    for (int j = 0; j <= cx; ++j) {
                        const unsigned long long v = (unsigned long long)ldexp((double)w[(cy - i) * kx + (cx - j)], S);
                        const unsigned long long mult = (i ? 2ull : 1ull) * (j ? 2ull : 1ull);
                        hq[i * (cx + 1) + j] = (uint32_t)(v >> drop);
                        sum_h += mult * (v >> drop);
                        sum_r += mult * (v & ((1ull << drop) - 1ull));
                    }
                const unsigned long long rm = ((255ull * sum_r) >> drop) + 1ull;
                if (255ull * sum_h < (1ull << 32) && rm < (1ull << (t - 2))) { s1 = (uint32_t)t; rmax = (uint32_t)rm; break; }
            }
            if (s1 == 1) std::fill(hq.begin(), hq.end(), 0u);
            c->wh_host = hq; c->ws1 = s1; c->wrmax = rmax;
            CU_TRY(c, dmalloc(&c->wq, q.size()));
            CU_TRY(c, cudaMemcpy(c->wq, q.data(), sizeof(uint32_t) * q.size(), cudaMemcpyHostToDevice));
            c->wparts = parts; c->wshift = S;
            // separable screening kernel (k_blur_sep): A[i][j] = hy[i] * hx[j] ~ w[i][j] * 2^s, taken from the centre row and
            // column; eps = 255 * max(sum of positive, sum of negative) errors, in exact integer arithmetic
            c->sep_hy.clear(); c->sep_hx.clear();
            const double wcc = (double)w[cy * kx + cx];
            if (wcc > 0.0 && S <= 100) {
                const int a_bits = 30;
                int b_bits = 22;
                std::vector<uint32_t> hy(cy + 1), hx(cx + 1);
                bool ok = false;
                for (; b_bits >= 12 && !ok; --b_bits) {
                    unsigned long long vsum = 0;
                    for (int i = 0; i <= cy; ++i) {
                        const double f = (double)w[(cy - i) * kx + cx] / sqrt(wcc);
                        hy[i] = (uint32_t)llround(ldexp(f, b_bits));
                        vsum += (i ? 2ull : 1ull) * hy[i];
                    }
                    ok = 2ull * 255ull * vsum < (1ull << 32);
                    if (ok) break;
                }
                if (ok) {
                    unsigned __int128 total = 0;
                    bool fits = true;
                    for (int j = 0; j <= cx; ++j) {
                        const double f = (double)w[cy * kx + (cx - j)] / sqrt(wcc);
                        const double v = ldexp(f, a_bits);
                        if (!(v < 4294967295.0)) { fits = false; break; }
                        hx[j] = (uint32_t)llround(v);
                    }
                    const int s_bits = a_bits + b_bits;
                    if (fits && s_bits >= 40 && s_bits <= 60) {
                        // common scale 2^m, m = max(s, S): errors e = A * 2^(m-s) - Wint * 2^(m-S)
                        const int m = s_bits > S ? s_bits : S;
                        unsigned __int128 pos = 0, neg = 0;
                        for (int i = 0; i < ky; ++i)
                            for (int j = 0; j < kx; ++j) {
                                const int di = i < cy ? cy - i : i - cy, dj = j < cx ? cx - j : j - cx;
                                const unsigned __int128 A = ((unsigned __int128)hy[di] * hx[dj]) << (m - s_bits);
                                const unsigned __int128 Wi = ((unsigned __int128)(unsigned long long)ldexp((double)w[i * kx + j], S)) << (m - S);
                                total += (unsigned __int128)hy[di] * hx[dj];
                                if (A >= Wi) pos += A - Wi; else neg += Wi - A;
                            }
                        unsigned __int128 e = (pos > neg ? pos : neg) * 255u;
                        e = (e >> (m - s_bits)) + 2;                                   // units of 2^-s, rounded up
                        const unsigned __int128 e32 = (e >> (s_bits - 32)) + 2;        // units of 2^-32 of one output step
                        if (total * 255u < ((unsigned __int128)1 << 63) && e32 < (1u << 22)) {   // < 0.1 % of the values undecided
                            c->sep_hy = hy; c->sep_hx = hx; c->sep_s = (uint32_t)s_bits; c->sep_eps32 = (uint32_t)e32;
                        }
                    }
                }
            }
        }
    }
    return VRSBS_OK;
}

int vrsbs_depth_from_lowres(vrsbs_ctx *c, const void *lo, int B, int h, int w, float scaler, int H, int W, void *out,
                            void *stream) {
    int rc = check_dims(c, B, H, W);
    if (rc) return rc;
    if (!lo || !out || h < 1 || w < 1) return fail(c, VRSBS_E_INVALID, "bad low-res depth arguments");
    DeviceGuard g(c->device);
    if (!c->inflight.empty()) return fail(c, VRSBS_E_STATE, "device-pointer call with submitted host batches not yet collected");
    c->last_dev_stream = (cudaStream_t)stream; c->dev_state_dirty = true;
    return launch_depth(c, c->scratch[0], nullptr, (const __half *)lo, B, H, W, h, w, scaler, out, (cudaStream_t)stream);
}

int vrsbs_depth_from_full(vrsbs_ctx *c, const void *raw, int B, int H, int W, void *out, void *stream) {
    int rc = check_dims(c, B, H, W);
    if (rc) return rc;
    if (!raw || !out) return fail(c, VRSBS_E_INVALID, "NULL depth pointer");
    DeviceGuard g(c->device);
    if (!c->inflight.empty()) return fail(c, VRSBS_E_STATE, "device-pointer call with submitted host batches not yet collected");
    c->last_dev_stream = (cudaStream_t)stream; c->dev_state_dirty = true;
    return launch_depth(c, c->scratch[0], raw, nullptr, B, H, W, 0, 0, 1.f, out, (cudaStream_t)stream);
}

int vrsbs_build_tables(vrsbs_ctx *c, int B, int H, int W, void *stream) {
    int rc = check_dims(c, B, H, W);
    if (rc) return rc;
    DeviceGuard g(c->device);
    if (!c->inflight.empty()) return fail(c, VRSBS_E_STATE, "device-pointer call with submitted host batches not yet collected");
    c->last_dev_stream = (cudaStream_t)stream; c->dev_state_dirty = true;
    return launch_tables(c, c->scratch[0], B, H, W, (cudaStream_t)stream);
}

int vrsbs_warp_batch(vrsbs_ctx *c, const uint8_t *frames, const void *depth, int B, int H, int W, uint8_t *sbs, void *stream) {
    int rc = check_dims(c, B, H, W);
    if (rc) return rc;
    if (!frames || !depth || !sbs) return fail(c, VRSBS_E_INVALID, "NULL buffer");
    DeviceGuard g(c->device);
    if (!c->inflight.empty()) return fail(c, VRSBS_E_STATE, "device-pointer call with submitted host batches not yet collected");
    c->last_dev_stream = (cudaStream_t)stream; c->dev_state_dirty = true;
    return launch_warp(c, c->scratch[0], frames, depth, B, H, W, sbs, (cudaStream_t)stream);
}

int vrsbs_process_batch(vrsbs_ctx *c, const uint8_t *frames, const void *raw, int B, int H, int W, void *depth_scratch,
                        uint8_t *sbs, void *stream) {
    int rc = check_dims(c, B, H, W);
    if (rc) return rc;
    if (!frames || !raw || !sbs) return fail(c, VRSBS_E_INVALID, "NULL buffer");
    DeviceGuard g(c->device);
    if (!c->inflight.empty()) return fail(c, VRSBS_E_STATE, "device-pointer call with submitted host batches not yet collected");
    c->last_dev_stream = (cudaStream_t)stream; c->dev_state_dirty = true;
    cudaStream_t st = (cudaStream_t)stream;
    Scratch &s = c->scratch[0];
    fast_caps(c, H, W, &c->ent_cap, &c->lut_cap);
    if (!c->f32 && c->fused && c->smooth_in_warp && ((size_t)H * W) % 8 == 0 && fused_capable(c, frames, raw, sbs, W))
        return launch_process_fused(c, s, frames, (const __half *)raw, B, H, W, sbs, st);
    if (!depth_scratch) return fail(c, VRSBS_E_INVALID, "depth_scratch_dev is required for this frame size / alignment");
    if ((rc = launch_depth(c, s, raw, nullptr, B, H, W, 0, 0, 1.f, depth_scratch, st))) return rc;
    if ((rc = launch_tables(c, s, B, H, W, st))) return rc;
    return launch_warp(c, s, frames, depth_scratch, B, H, W, sbs, st);
}

int vrsbs_process_host(vrsbs_ctx *c, const uint8_t *frames, const void *depth, int B, int H, int W, int lh, int lw,
                       float scaler, uint8_t *sbs) {
    if (!c) return VRSBS_E_INVALID;
    if (B < 1) return fail(c, VRSBS_E_INVALID, "batch %d", B);
    int rc = check_dims(c, 1, H, W);
    if (rc) return rc;
    if (!frames || !depth || !sbs) return fail(c, VRSBS_E_INVALID, "NULL buffer");
    if (!c->inflight.empty()) return fail(c, VRSBS_E_STATE, "vrsbs_process_host with %zu submitted batches not yet collected", c->inflight.size());
    const bool lowres = lh > 0 && lw > 0;
    DeviceGuard g(c->device);
    if (c->host_async && !c->pageable_direct && c->host_right_half != 2 && !(c->fused && c->smooth_in_warp) && is_pinned(frames) && is_pinned(sbs) &&
        (is_device(depth) || is_pinned(depth))) {
        // page-locked buffers: the asynchronous pipeline, collected right away (same result, one code path to maintain)
        uint64_t t = 0;
        rc = vrsbs_submit_host(c, frames, 0, 0, depth, B, H, W, lh, lw, scaler, sbs, 0u, &t);
        if (rc) return rc;
        return vrsbs_collect(c, t);
    }
    int chunk = c->host_chunk < c->max_batch ? c->host_chunk : c->max_batch;
    if (chunk < 1) chunk = 1;
    const size_t fb = (size_t)H * W * 3, sb = fb * 2, db = (size_t)H * W * (c->f32 ? 4 : 2);
    const size_t dib = lowres ? (size_t)lh * lw * 2 : db;
    const size_t row3 = (size_t)W * 3;
    const bool direct = c->pageable_direct != 0;
    // the depth may already live on the device (a depth producer in the same process): it is then used
    // where it is - no staging, no H2D - and the caller has made sure the producer's stream is done with it
    const bool dev_d = is_device(depth);
    const bool pin_f = direct || is_pinned(frames), pin_d = direct || dev_d || is_pinned(depth), pin_s = direct || is_pinned(sbs);
    // the right half of every SBS row is the caller's own input row: copy it host to host and move only the
    // synthesised half across PCIe (halves the D2H traffic, which is the longer leg)
    const bool host_right = c->host_right_half != 0;
    for (int i = 0; i < kSlots; ++i) {
        if (!c->scratch[i].tabs && (rc = alloc_scratch(c, c->scratch[i], chunk))) return rc;
        if (c->scratch[i].cap_batch < chunk) {            // host_chunk was raised after the first call
            CU_TRY(c, cudaDeviceSynchronize());
            free_scratch(c->scratch[i]);
            if ((rc = alloc_scratch(c, c->scratch[i], chunk))) return rc;
        }
        if (c->params.blur && !c->scratch[i].plane)       // not inside the pipeline: cudaMalloc synchronises the device
            CU_TRY(c, dmalloc(&c->scratch[i].plane, (size_t)c->scratch[i].cap_batch * c->max_h * c->max_w * 3));
        if ((rc = ensure_slot(c, c->slot[i], fb * chunk, dib * chunk, db * chunk, sb * chunk, !pin_f || !pin_d, !pin_s))) return rc;
    }
    const bool need_pool = !pin_f || !pin_d || !pin_s || host_right;
    if (!c->pool && need_pool) c->pool = new (std::nothrow) CopyPool(c->copy_threads);
    if (need_pool && !c->pool) return fail(c, VRSBS_E_NOMEM, "cannot start the host copy threads");
    CopyPool *pool = c->pool;
    fast_caps(c, H, W, &c->ent_cap, &c->lut_cap);
    if ((rc = check_blur_ready(c, H, W))) return rc;
    if ((rc = mixed_entry_sync(c))) return rc;
    const bool use_fused = !c->f32 && !dev_d && !lowres && c->fused && c->smooth_in_warp && ((size_t)H * W) % 8 == 0 &&
                           fused_capable(c, c->slot[0].dev_frames, c->slot[0].dev_depth_in, c->slot[0].dev_sbs, W);

    // Three streams: st_in copies chunk i+1 in while st_k runs the kernels of chunk i and st_out copies chunk i-1
    // out.  All kernels run on st_k in clip order, so the clip-range state needs no extra synchronisation.  Host
    // copies (staging of pageable buffers, right halves) run on the copy pool and are waited for by ticket.
    const int nchunks = (B + chunk - 1) / chunk;
    int pending_n[kSlots] = {0}, pending_first[kSlots] = {0};
    int tk_in[kSlots], tk_in2[kSlots], tk_out[kSlots], tk_right[kSlots];
    for (int i = 0; i < kSlots; ++i) tk_in[i] = tk_in2[i] = tk_out[i] = tk_right[i] = -1;
    auto count_of = [&](int ci) { const int first = ci * chunk; return (B - first < chunk) ? B - first : chunk; };
    auto stage_in = [&](int ci) {                      // pageable inputs of chunk ci -> the slot's pinned buffers (async)
        const int si = ci % kSlots, first = ci * chunk, n = count_of(ci);
        HostSlot &s = c->slot[si];
        if (pin_f && pin_d) return;
        cudaEventSynchronize(s.in_done);               // the slot's previous H2D has left the pinned buffers (no-op if never recorded)
        if (!pin_f) tk_in[si] = pool->submit(s.pin_frames, fb * n, frames + (size_t)first * fb, fb * n, fb * n, 1);
        if (!pin_d) tk_in2[si] = pool->submit(s.pin_depth, dib * n, (const uint8_t *)depth + (size_t)first * dib, dib * n, dib * n, 1);
    };
    auto drain = [&](int si) -> int {     // wait for slot si's chunk, deliver its output, check its status
        HostSlot &s = c->slot[si];
        if (!pending_n[si]) return VRSBS_OK;
        const cudaError_t ee = cudaEventSynchronize(s.out_done);
        if (ee != cudaSuccess) { pending_n[si] = 0; return fail(c, VRSBS_E_CUDA, "cudaEventSynchronize failed: %s", cudaGetErrorString(ee)); }
        const int n = pending_n[si], first = pending_first[si];
        if (!pin_s) {
            uint8_t *dst = sbs + (size_t)first * sb;
            if (host_right) tk_out[si] = pool->submit(dst, 2 * row3, s.pin_sbs, row3, row3, (size_t)n * H);   // compact left halves
            else tk_out[si] = pool->submit(dst, sb * n, s.pin_sbs, sb * n, sb * n, 1);
        }
        pending_n[si] = 0;
        return frame_status_error(c, s.pin_tabs, n, first);
    };
    auto drain_all = [&](int rc_in) -> int {          // never leave work in flight behind an error return
        if (rc_in) { cudaStreamSynchronize(c->st_in); cudaStreamSynchronize(c->st_k); cudaStreamSynchronize(c->st_out); }
        c->skip_right = 0;
        for (int i = 0; i < kSlots; ++i) { int r = drain(i); if (!rc_in) rc_in = r; }
        if (pool) for (int i = 0; i < kSlots; ++i) { pool->wait(tk_in[i]); pool->wait(tk_in2[i]); pool->wait(tk_out[i]); pool->wait(tk_right[i]); }
        return rc_in;
    };
#define CU_TRY_DRAIN(call)                                                                             \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return drain_all(fail(c, VRSBS_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); \
    } while (0)
    stage_in(0);
    for (int ci = 0; ci < nchunks; ++ci) {
        const int si = ci % kSlots, first = ci * chunk, n = count_of(ci);
        HostSlot &s = c->slot[si];
        if ((rc = drain(si))) return drain_all(rc);            // frees the slot's device + pinned output buffers
        if (ci + 1 < nchunks) {
            // pageable output: hand finished chunks to the copy pool as early as possible
            if (!pin_s && (rc = drain((ci + 1) % kSlots))) return drain_all(rc);
            stage_in(ci + 1);                                   // overlaps with this chunk's enqueue and the GPU
        }
        const uint8_t *hf = frames + (size_t)first * fb;
        const uint8_t *hd = (const uint8_t *)depth + (size_t)first * dib;
        if (pool) { pool->wait(tk_in[si]); pool->wait(tk_in2[si]); }
        if (!pin_f) hf = s.pin_frames;
        if (!pin_d) hd = s.pin_depth;
        CU_TRY_DRAIN(cudaMemcpyAsync(s.dev_frames, hf, fb * n, cudaMemcpyHostToDevice, c->st_in));
        if (!dev_d) CU_TRY_DRAIN(cudaMemcpyAsync(s.dev_depth_in, hd, dib * n, cudaMemcpyHostToDevice, c->st_in));
        const __half *din = dev_d ? (const __half *)hd : (const __half *)s.dev_depth_in;
        CU_TRY_DRAIN(cudaEventRecord(s.in_done, c->st_in));
        if (host_right && c->host_right_half != 2)               // right halves: caller's frames -> caller's output, on the pool
            { pool->wait(tk_right[si]); tk_right[si] = pool->submit(sbs + (size_t)first * sb + row3, 2 * row3, frames + (size_t)first * fb, row3, row3, (size_t)n * H); }
        CU_TRY_DRAIN(cudaStreamWaitEvent(c->st_k, s.in_done, 0));
        Scratch &sc = c->scratch[si];
        c->skip_right = host_right ? 1 : 0;                      // nobody reads the right half of dev_sbs then
        if (use_fused) {
            rc = launch_process_fused(c, sc, s.dev_frames, din, n, H, W, s.dev_sbs, c->st_k);
        } else {
            if (lowres) rc = launch_depth(c, sc, nullptr, din, n, H, W, lh, lw, scaler, s.dev_depth, c->st_k);
            else rc = launch_depth(c, sc, din, nullptr, n, H, W, 0, 0, 1.f, s.dev_depth, c->st_k);
            if (!rc) rc = launch_tables(c, sc, n, H, W, c->st_k);
            if (!rc) rc = launch_warp(c, sc, s.dev_frames, s.dev_depth, n, H, W, s.dev_sbs, c->st_k);
        }
        c->skip_right = 0;
        if (rc) return drain_all(rc);
        CU_TRY_DRAIN(cudaMemcpyAsync(s.pin_tabs, sc.tabs, sizeof(FrameTab) * n, cudaMemcpyDeviceToHost, c->st_k));
        CU_TRY_DRAIN(cudaEventRecord(s.k_done, c->st_k));
        CU_TRY_DRAIN(cudaStreamWaitEvent(c->st_out, s.k_done, 0));
        if (pool) pool->wait(tk_out[si]);                        // the slot's pinned output has been delivered
        if (host_right) {
            // left halves only: [n*H rows] x W*3 bytes out of rows of 2*W*3
            if (pin_s) CU_TRY_DRAIN(cudaMemcpy2DAsync(sbs + (size_t)first * sb, 2 * row3, s.dev_sbs, 2 * row3, row3, (size_t)n * H, cudaMemcpyDeviceToHost, c->st_out));
            else CU_TRY_DRAIN(cudaMemcpy2DAsync(s.pin_sbs, row3, s.dev_sbs, 2 * row3, row3, (size_t)n * H, cudaMemcpyDeviceToHost, c->st_out));
        } else {
            uint8_t *ho = pin_s ? sbs + (size_t)first * sb : s.pin_sbs;
            CU_TRY_DRAIN(cudaMemcpyAsync(ho, s.dev_sbs, sb * n, cudaMemcpyDeviceToHost, c->st_out));
        }
        CU_TRY_DRAIN(cudaEventRecord(s.out_done, c->st_out));
        pending_n[si] = n; pending_first[si] = first;
    }
    return drain_all(VRSBS_OK);
#undef CU_TRY_DRAIN
}

// ---- asynchronous host pipeline ----------------------------------------------------------------------------------
namespace {

vrsbs_ctx::Inflight *find_inflight(vrsbs_ctx *c, uint64_t ticket) {
    for (auto &f : c->inflight) if (f.ticket == ticket) return &f;
    return nullptr;
}

// Waits for the chunk that occupies slot si, records its per-frame status on the batch it belongs to, frees the slot.
void retire_slot(vrsbs_ctx *c, int si) {
    if (!c->slot_ticket[si]) return;
    HostSlot &s = c->slot[si];
    vrsbs_ctx::Inflight *f = find_inflight(c, c->slot_ticket[si]);
    const cudaError_t e = cudaEventSynchronize(s.out_done);
    if (f) {
        if (e != cudaSuccess && !f->rc) {
            f->rc = VRSBS_E_CUDA;
            snprintf(f->err, sizeof f->err, "cudaEventSynchronize failed: %s", cudaGetErrorString(e));
        } else if (!f->rc) {
            char keep[256];
            memcpy(keep, c->err, sizeof keep);
            const int rc = frame_status_error(c, s.pin_tabs, c->slot_n[si], c->slot_first[si]);
            if (rc) { f->rc = rc; memcpy(f->err, c->err, sizeof f->err); }
            memcpy(c->err, keep, sizeof keep);
        }
        f->chunks--;
    }
    c->slot_ticket[si] = 0;
}

// An enqueue failed half way: let everything already queued finish, then fail every batch in flight.
int abort_inflight(vrsbs_ctx *c, int rc) {
    if (c->st_in) { cudaStreamSynchronize(c->st_in); cudaStreamSynchronize(c->st_k); cudaStreamSynchronize(c->st_out); }
    c->skip_right = 0;
    for (int i = 0; i < kSlots; ++i) c->slot_ticket[i] = 0;
    for (auto &f : c->inflight) if (!f.rc) { f.rc = rc; memcpy(f.err, c->err, sizeof f.err); f.chunks = 0; }
    return rc;
}

}  // namespace

int vrsbs_host_depends_on(vrsbs_ctx *c, void *producer_stream) {
    if (!c) return VRSBS_E_INVALID;
    DeviceGuard g(c->device);
    if (!c->st_in) {
        CU_TRY(c, cudaStreamCreateWithFlags(&c->st_in, cudaStreamNonBlocking));
        CU_TRY(c, cudaStreamCreateWithFlags(&c->st_k, cudaStreamNonBlocking));
        CU_TRY(c, cudaStreamCreateWithFlags(&c->st_out, cudaStreamNonBlocking));
    }
    if (!c->dep_event) CU_TRY(c, cudaEventCreateWithFlags(&c->dep_event, cudaEventDisableTiming));
    CU_TRY(c, cudaEventRecord(c->dep_event, (cudaStream_t)producer_stream));
    CU_TRY(c, cudaStreamWaitEvent(c->st_k, c->dep_event, 0));      // every kernel of the host pipeline runs on st_k, in order
    return VRSBS_OK;
}

int vrsbs_submit_host(vrsbs_ctx *c, const uint8_t *frames, size_t frame_row_pitch, size_t frame_pitch, const void *depth,
                      int B, int H, int W, int lh, int lw, float scaler, uint8_t *sbs, unsigned flags, uint64_t *ticket) {
    if (!c) return VRSBS_E_INVALID;
    if (!ticket) return fail(c, VRSBS_E_INVALID, "ticket is NULL");
    *ticket = 0;
    if (B < 1) return fail(c, VRSBS_E_INVALID, "batch %d", B);
    int rc = check_dims(c, 1, H, W);
    if (rc) return rc;
    if (!frames || !depth || !sbs) return fail(c, VRSBS_E_INVALID, "NULL buffer");
    const size_t row3 = (size_t)W * 3, fb = (size_t)H * row3, sb = fb * 2, db = (size_t)H * W * (c->f32 ? 4 : 2);
    if (!frame_row_pitch) frame_row_pitch = row3;
    if (!frame_pitch) frame_pitch = frame_row_pitch * H;
    if (frame_row_pitch < row3 || frame_pitch < frame_row_pitch * (size_t)(H - 1) + row3)
        return fail(c, VRSBS_E_INVALID, "frame pitches %zu / %zu too small for %dx%d", frame_row_pitch, frame_pitch, W, H);
    const bool in_place = (flags & VRSBS_HOST_RIGHT_IN_PLACE) != 0;
    if (in_place && (frames != sbs + row3 || frame_row_pitch != 2 * row3 || frame_pitch != sb))
        return fail(c, VRSBS_E_INVALID, "VRSBS_HOST_RIGHT_IN_PLACE: frames must be the right halves of sbs_host (frames = sbs + 3W, row pitch 6W)");
    const bool lowres = lh > 0 && lw > 0;
    const size_t dib = lowres ? (size_t)lh * lw * 2 : db;
    DeviceGuard g(c->device);
    const bool dev_d = is_device(depth);
    if (!is_pinned(frames) || !is_pinned(sbs) || !(dev_d || is_pinned(depth)))
        return fail(c, VRSBS_E_INVALID, "vrsbs_submit_host needs page-locked host buffers (cudaHostAlloc / cudaHostRegister); "
                                          "vrsbs_process_host stages pageable ones");
    int chunk = c->host_chunk < c->max_batch ? c->host_chunk : c->max_batch;
    if (chunk < 1) chunk = 1;
    bool grow = false;
    for (int i = 0; i < kSlots; ++i)
        grow = grow || c->slot[i].cap_frames < fb * chunk || c->slot[i].cap_depth_in < dib * chunk || c->slot[i].cap_depth < db * chunk ||
               c->slot[i].cap_sbs < sb * chunk || !c->scratch[i].tabs || c->scratch[i].cap_batch < chunk ||
               (c->params.blur && !c->scratch[i].plane);
    if (grow) {                                           // (re)allocation: nothing may be in flight on the old buffers
        for (int i = 0; i < kSlots; ++i) retire_slot(c, i);
        if (c->st_in) CU_TRY(c, cudaDeviceSynchronize());
        for (int i = 0; i < kSlots; ++i) {
            if (c->scratch[i].tabs && c->scratch[i].cap_batch < chunk) free_scratch(c->scratch[i]);
            if (!c->scratch[i].tabs && (rc = alloc_scratch(c, c->scratch[i], i == 0 ? c->max_batch : chunk))) return rc;
            if (c->params.blur && !c->scratch[i].plane)
                CU_TRY(c, dmalloc(&c->scratch[i].plane, (size_t)c->scratch[i].cap_batch * c->max_h * c->max_w * 3));
            if ((rc = ensure_slot(c, c->slot[i], fb * chunk, dib * chunk, db * chunk, sb * chunk, false, false))) return rc;
        }
    }
    fast_caps(c, H, W, &c->ent_cap, &c->lut_cap);
    if ((rc = check_blur_ready(c, H, W))) return rc;
    if ((rc = mixed_entry_sync(c))) return rc;
    const bool plain = frame_row_pitch == row3 && frame_pitch == fb;
    // without the in-place layout the right halves are the caller's own frames: copied host to host by the pool
    const bool host_right = !in_place && c->host_right_half != 0;
    if (host_right && !c->pool) {
        c->pool = new (std::nothrow) CopyPool(c->copy_threads);
        if (!c->pool) return fail(c, VRSBS_E_NOMEM, "cannot start the host copy threads");
    }
    vrsbs_ctx::Inflight job{};
    job.ticket = c->next_ticket++;
    const int nchunks = (B + chunk - 1) / chunk;
    job.chunks = nchunks;
    c->inflight.push_back(job);
#define CU_TRY_ABORT(call)                                                                             \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return abort_inflight(c, fail(c, VRSBS_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e_), __FILE__, __LINE__)); \
    } while (0)
    std::vector<int> right_tickets;
    for (int ci = 0; ci < nchunks; ++ci) {
        const int si = (int)(c->chunk_seq++ % kSlots), first = ci * chunk, n = (B - first < chunk) ? B - first : chunk;
        HostSlot &s = c->slot[si];
        retire_slot(c, si);                                // blocks while the pipeline is full (three chunks in flight)
        const uint8_t *hf = frames + (size_t)first * frame_pitch;
        const uint8_t *hd = (const uint8_t *)depth + (size_t)first * dib;
        if (plain) CU_TRY_ABORT(cudaMemcpyAsync(s.dev_frames, hf, fb * n, cudaMemcpyHostToDevice, c->st_in));
        else if (frame_pitch == frame_row_pitch * (size_t)H)   // one pitched copy for the whole chunk
            CU_TRY_ABORT(cudaMemcpy2DAsync(s.dev_frames, row3, hf, frame_row_pitch, row3, (size_t)n * H, cudaMemcpyHostToDevice, c->st_in));
        else
            for (int t = 0; t < n; ++t)
                CU_TRY_ABORT(cudaMemcpy2DAsync(s.dev_frames + (size_t)t * fb, row3, hf + (size_t)t * frame_pitch, frame_row_pitch, row3, H, cudaMemcpyHostToDevice, c->st_in));
        if (!dev_d) CU_TRY_ABORT(cudaMemcpyAsync(s.dev_depth_in, hd, dib * n, cudaMemcpyHostToDevice, c->st_in));
        const __half *din = dev_d ? (const __half *)hd : (const __half *)s.dev_depth_in;
        CU_TRY_ABORT(cudaEventRecord(s.in_done, c->st_in));
        if (host_right)
            right_tickets.push_back(c->pool->submit(sbs + (size_t)first * sb + row3, 2 * row3, hf, frame_row_pitch, row3, (size_t)n * H));
        CU_TRY_ABORT(cudaStreamWaitEvent(c->st_k, s.in_done, 0));
        Scratch &sc = c->scratch[si];
        c->skip_right = 1;                                 // the right half never leaves the host
        if (lowres) rc = launch_depth(c, sc, nullptr, din, n, H, W, lh, lw, scaler, s.dev_depth, c->st_k);
        else rc = launch_depth(c, sc, din, nullptr, n, H, W, 0, 0, 1.f, s.dev_depth, c->st_k);
        if (!rc) rc = launch_tables(c, sc, n, H, W, c->st_k);
        if (!rc) rc = launch_warp(c, sc, s.dev_frames, s.dev_depth, n, H, W, s.dev_sbs, c->st_k);
        c->skip_right = 0;
        if (rc) return abort_inflight(c, rc);
        CU_TRY_ABORT(cudaMemcpyAsync(s.pin_tabs, sc.tabs, sizeof(FrameTab) * n, cudaMemcpyDeviceToHost, c->st_k));
        CU_TRY_ABORT(cudaEventRecord(s.k_done, c->st_k));
        CU_TRY_ABORT(cudaStreamWaitEvent(c->st_out, s.k_done, 0));
        CU_TRY_ABORT(cudaMemcpy2DAsync(sbs + (size_t)first * sb, 2 * row3, s.dev_sbs, 2 * row3, row3, (size_t)n * H, cudaMemcpyDeviceToHost, c->st_out));
        CU_TRY_ABORT(cudaEventRecord(s.out_done, c->st_out));
        c->slot_ticket[si] = job.ticket; c->slot_n[si] = n; c->slot_first[si] = first;
    }
#undef CU_TRY_ABORT
    for (int t : right_tickets) c->pool->wait(t);          // host-to-host right halves (only without the in-place layout)
    *ticket = job.ticket;
    return VRSBS_OK;
}

int vrsbs_collect(vrsbs_ctx *c, uint64_t ticket) {
    if (!c) return VRSBS_E_INVALID;
    DeviceGuard g(c->device);
    vrsbs_ctx::Inflight *f = find_inflight(c, ticket);
    if (!f) return fail(c, VRSBS_E_INVALID, "unknown or already collected ticket %llu", (unsigned long long)ticket);
    for (int i = 0; i < kSlots; ++i)
        if (c->slot_ticket[i] == ticket) retire_slot(c, i);
    f = find_inflight(c, ticket);
    const int rc = f->rc;
    if (rc) memcpy(c->err, f->err, sizeof c->err);
    c->inflight.erase(c->inflight.begin() + (f - c->inflight.data()));
    return rc;
}

int vrsbs_get_frame_info(vrsbs_ctx *c, int B, vrsbs_frame_info *info, void *stream) {
    if (!c || !info || B < 1 || B > c->max_batch) return fail(c, VRSBS_E_INVALID, "bad arguments");
    DeviceGuard g(c->device);
    CU_TRY(c, cudaStreamSynchronize((cudaStream_t)stream));
    std::vector<FrameTab> tabs(B);
    CU_TRY(c, cudaMemcpy(tabs.data(), c->scratch[0].tabs, sizeof(FrameTab) * B, cudaMemcpyDeviceToHost));
    for (int i = 0; i < B; ++i) {
        vrsbs_frame_info &o = info[i];
        memset(&o, 0, sizeof o);
        o.status = tabs[i].status; o.layers = tabs[i].layers; o.limit_step = tabs[i].limit_step;
        o.fill_layer = tabs[i].fill_layer; o.strip = tabs[i].strip; o.depth_max = tabs[i].depth_max;
        o.offset_range[0] = tabs[i].range[0]; o.offset_range[1] = tabs[i].range[1]; o.holes = tabs[i].holes;
    }
    return frame_status_error(c, tabs.data(), B, 0);
}

int vrsbs_get_tables(vrsbs_ctx *c, int frame, int cap, double *cutoffs, int32_t *offsets, uint16_t *lo, uint16_t *hi,
                     void *stream) {
    if (!c || frame < 0 || frame >= c->max_batch) return fail(c, VRSBS_E_INVALID, "bad frame index");
    DeviceGuard g(c->device);
    CU_TRY(c, cudaStreamSynchronize((cudaStream_t)stream));
    FrameTab t;
    Scratch &s = c->scratch[0];
    CU_TRY(c, cudaMemcpy(&t, s.tabs + frame, sizeof t, cudaMemcpyDeviceToHost));
    const int L = t.layers, Lc = c->max_layers;
    if (cap < L + 1) return fail(c, VRSBS_E_INVALID, "capacity %d < L+1 = %d", cap, L + 1);
    if (cutoffs) CU_TRY(c, cudaMemcpy(cutoffs, s.cutoffs + (size_t)frame * (Lc + 1), sizeof(double) * (L + 1), cudaMemcpyDeviceToHost));
    if (offsets) CU_TRY(c, cudaMemcpy(offsets, s.offsets + (size_t)frame * Lc, sizeof(int32_t) * L, cudaMemcpyDeviceToHost));
    if (lo) CU_TRY(c, cudaMemcpy(lo, s.lo16 + (size_t)frame * Lc, sizeof(uint16_t) * L, cudaMemcpyDeviceToHost));
    if (hi) CU_TRY(c, cudaMemcpy(hi, s.hi16 + (size_t)frame * Lc, sizeof(uint16_t) * L, cudaMemcpyDeviceToHost));
    return L;
}

int vrsbs_get_bounds(vrsbs_ctx *c, int frame, int cap, float *lo, float *hi, void *stream) {
    if (!c || frame < 0 || frame >= c->max_batch || !lo || !hi) return fail(c, VRSBS_E_INVALID, "bad arguments");
    DeviceGuard g(c->device);
    CU_TRY(c, cudaStreamSynchronize((cudaStream_t)stream));
    FrameTab t;
    Scratch &s = c->scratch[0];
    CU_TRY(c, cudaMemcpy(&t, s.tabs + frame, sizeof t, cudaMemcpyDeviceToHost));
    const int L = t.layers;
    if (cap < L) return fail(c, VRSBS_E_INVALID, "capacity %d < L = %d", cap, L);
    std::vector<float2> b(L);
    CU_TRY(c, cudaMemcpy(b.data(), s.bounds + (size_t)frame * c->max_layers, sizeof(float2) * L, cudaMemcpyDeviceToHost));
    for (int k = 0; k < L; ++k) { lo[k] = b[k].x; hi[k] = b[k].y; }
    return L;
}

int vrsbs_get_hole_mask(vrsbs_ctx *c, int B, int H, int W, uint32_t *mask, void *stream) {
    int rc = check_dims(c, B, H, W);
    if (rc) return rc;
    if (!mask) return fail(c, VRSBS_E_INVALID, "NULL mask");
    DeviceGuard g(c->device);
    CU_TRY(c, cudaStreamSynchronize((cudaStream_t)stream));
    CU_TRY(c, cudaMemcpy(mask, c->scratch[0].hole_mask, sizeof(uint32_t) * B * H * ((W + 31) / 32), cudaMemcpyDeviceToHost));
    return VRSBS_OK;
}

int vrsbs_get_stage_times(vrsbs_ctx *c, double ms[VRSBS_NUM_STAGES], uint64_t count[VRSBS_NUM_STAGES]) {
    if (!c || !ms || !count) return fail(c, VRSBS_E_INVALID, "NULL argument");
    DeviceGuard g(c->device);
    CU_TRY(c, cudaDeviceSynchronize());
    for (auto &s : c->stamps) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, s.a, s.b) == cudaSuccess && s.stage >= 0 && s.stage < VRSBS_NUM_STAGES) {
            ms[s.stage] += t;
            count[s.stage] += 1;
        }
        c->event_pool.push_back(s.a);
        c->event_pool.push_back(s.b);
    }
    c->stamps.clear();
    return VRSBS_OK;
}

uint64_t vrsbs_launch_count(const vrsbs_ctx *c) { return c ? c->launches : 0; }

int vrsbs_set_option(vrsbs_ctx *c, const char *name, int value) {
    if (!c || !name) return VRSBS_E_INVALID;
    if (!strcmp(name, "scatter_mode")) { if (value != 1 && value != 2) return fail(c, VRSBS_E_INVALID, "scatter_mode 1|2"); c->scatter_mode = value; }
    else if (!strcmp(name, "bicubic_contract")) c->bicubic_contract = value != 0;
    else if (!strcmp(name, "blocks_per_sm")) c->blocks_per_sm = value < 0 ? 0 : value;
    else if (!strcmp(name, "host_chunk")) { if (value < 1) return fail(c, VRSBS_E_INVALID, "host_chunk >= 1"); c->host_chunk = value; }
    else if (!strcmp(name, "stage_timing")) c->stage_timing = value != 0;
    else if (!strcmp(name, "copy_threads")) c->copy_threads = value < 1 ? 1 : value;
    else if (!strcmp(name, "pageable_direct")) c->pageable_direct = value != 0;
    else if (!strcmp(name, "host_right_half")) c->host_right_half = value;   // 2 (experiments): left half only, right half not delivered
    else if (!strcmp(name, "fused")) c->fused = value != 0;
    else if (!strcmp(name, "fast_tables")) c->fast_tables = value != 0;
    else if (!strcmp(name, "smooth_in_warp")) c->smooth_in_warp = value != 0;
    else if (!strcmp(name, "lowres_tiled")) c->lowres_tiled = value != 0;
    else if (!strcmp(name, "warp_ws")) c->warp_ws = value != 0;
    else if (!strcmp(name, "commit_mode")) c->commit_mode = value;
    else if (!strcmp(name, "f32_fast")) c->f32_fast = value ? 1 : 0;
    else if (!strcmp(name, "blur_screen")) c->blur_screen = value ? 1 : 0;
    else if (!strcmp(name, "blur_sep")) c->blur_sep = value;
    else if (!strcmp(name, "blur_band")) c->blur_band = value ? 1 : 0;
    else if (!strcmp(name, "ws_no_list")) c->ws_no_list = value ? 1 : 0;
    else if (!strcmp(name, "pdl")) c->pdl = value & 15;
    else if (!strcmp(name, "host_async")) c->host_async = value ? 1 : 0;
    else if (!strcmp(name, "ws_scatter_warps")) c->ws_scatter_warps = value;
    else return fail(c, VRSBS_E_INVALID, "unknown option %s", name);
    return VRSBS_OK;
}

}  // extern "C"
