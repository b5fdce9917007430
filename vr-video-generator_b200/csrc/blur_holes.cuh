// Stage 3b: hole blur, commit and strip restore (PredictAndGenerate.py:191-196).
//
// The reference blurs the whole filled frame with a (2k+1)x(2k+3) fp32 gaussian (torchvision
// gaussian_blur: reflect padding, one depthwise conv, round_) and keeps the result only at hole
// pixels.  Here only hole pixels are evaluated.  The warp kernel leaves (a) the FILLED pre-blur view in
// the SBS frame, (b) a hole bitmask and (c) a work index: a bitmap of the band columns (8 rows x one mask word) that hold a hole
// (k_warp_ws -> k_band_list -> k_blur_band, the default at 1080p / 720p) or a list of the mask words that contain holes
// (from the warp kernel or from k_word_list; k_blur_sep, k_blur_holes_fixed, k_blur_holes).
//
//   k_blur_holes  : one warp per listed mask word (32 pixels of one row).  The warp stages the word's
//                   footprint (ky rows x (32 + kx - 1) pixels, reflect padded) once, as per-byte-column
//                   vertical pair sums T_i = row(y-i) + row(y+i), then each lane evaluates one
//                   (hole, channel): horizontal pair sums x unique weights.  Results go to a scratch
//                   plane, NOT to the SBS frame, because neighbouring holes still need the pre-blur values.
//   k_blur_commit : copies the scratch values of the listed holes into the SBS frame, then restores the strip:
//                   result_img[:, 0:strip] = img[:, 0:strip] (PredictAndGenerate.py:196).
//
// Arithmetic.  The oracle defines the blurred value as the EXACT sum of fp32-weight x u8-pixel products,
// rounded half-to-even (DESIGN.md "blur parity").  Integer path (PARTS = 2 or 3): when the weights are 4-fold
// symmetric and every w * 2^S is an integer (true for every torchvision gaussian; checked on the host in
// vrsbs_set_blur_weights), the sum is evaluated in integers: pixel quads are added first (<= 1020), the
// integer weight is split into PARTS chunks of 15 (PARTS=2) or 13 (PARTS=3) bits so that 32-bit
// accumulators never overflow, and the final value is rounded half-to-even from the 64-bit total.
// Generic path (PARTS = 0): all ky*kx taps in fp64 FMA (exact for the same reason the oracle's float64
// accumulation is).
#pragma once
#include "common.cuh"

namespace vrsbs {

struct BlurArgs {
    const uint8_t *frames;       // [B,H,W,3]
    uint8_t *sbs;                // [B,H,2W,3]
    const FrameTab *tabs;        // [B]
    const uint32_t *hole_mask;   // [B][H][Wwords]
    const uint32_t *hole_list;   // (global row << 8 | word index) of the mask words that contain holes
    const uint32_t *hole_count;
    uint32_t *ticket;            // pre-zeroed work counter of k_blur_sep / k_blur_band (next unit of work)
    uint32_t *band_map;          // [B*Hb][band_groups] bit per band column (kBandRows rows x one mask word) that holds a hole: set by
                                 // k_warp_ws (or k_band_map), turned into band_list - and zeroed again - by k_band_list
    uint32_t *band_list;         // (frame * Hb + row band) << 8 | mask word column of every marked band column: the work list of
                                 // k_blur_band and k_blur_commit
    uint32_t *band_count;        // pre-zeroed
    int band_groups;             // 32-word groups per row of band_map
    int band_prepass;            // 1: the warp kernel did not mark the band columns, k_band_map does
    int Hb;                      // row bands per frame, kBandRows rows each
    unsigned long long magic_hb; // ceil(2^40 / Hb): frame = (band * magic_hb) >> 40
    uint8_t *plane;              // [B,H,W,3] scratch: blurred values of hole pixels
    const uint32_t *wq;          // integer path: [PARTS][(cy+1)][(cx+1)] parts of w * 2^S, low part first (i = |dy|, j = |dx|)
    const float *weights;        // generic: [ky][kx]
    int B, H, W, Wwords, kx, ky, wshift;   // wshift = S
    unsigned long long magic_h;  // ceil(2^40 / H): frame = (row * magic_h) >> 40, exact for row < 2^24
};

// per-warp shared memory: T[(rows)][cols] u16, cols = 3*(32 + kx - 1) rounded up to a multiple of 32
__host__ __device__ inline int blur_cols(int kx) { return (3 * (32 + kx - 1) + 31) / 32 * 32; }
__host__ __device__ inline size_t blur_warp_smem(int kx, int ky, bool sym) {
    return (size_t)(sym ? ky / 2 + 1 : ky) * blur_cols(kx) * sizeof(uint16_t);
}

template <int PARTS>
__global__ void __launch_bounds__(256) k_blur_holes(BlurArgs a) {
    extern __shared__ __align__(16) uint8_t blur_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int kx = a.kx, ky = a.ky, cx = kx / 2, cy = ky / 2, W = a.W, H = a.H;
    const int cols = blur_cols(kx), ncol = 3 * (32 + kx - 1);
    constexpr bool SYM = PARTS > 0;
    constexpr int PBITS = PARTS == 2 ? 15 : 13;
    const int trows = SYM ? cy + 1 : ky;
    uint16_t *T = reinterpret_cast<uint16_t *>(blur_smem) + (size_t)warp * trows * cols;
    const uint32_t count = *a.hole_count;

    for (uint32_t ei = blockIdx.x * nwarps + warp; ei < count; ei += gridDim.x * nwarps) {
        const uint32_t ent = a.hole_list[ei];
        const uint32_t row = ent >> 8, w = ent & 0xffu;
        const int b = (int)(((unsigned long long)row * a.magic_h) >> 40), y = (int)row - b * H;
        const int strip = a.tabs[b].strip;
        uint32_t m = a.hole_mask[(size_t)row * a.Wwords + w];
        const int xw = (int)w * 32;
        if (strip > xw) m = (strip - xw >= 32) ? 0u : (m & ~((1u << (strip - xw)) - 1u));
        if (m == 0u) continue;
        __syncwarp();
        // ---- stage: byte column c <-> pixel X = xw - cx + c/3, channel c%3 (reflect padded) ----
        const uint8_t *left = a.sbs + (size_t)b * H * W * 6;
        for (int c = lane; c < ncol; c += 32) {
            const int px = c / 3, ch = c - px * 3;
            const int X = min(max(reflect_idx(xw - cx + px, W), 0), W - 1);
            const uint8_t *colp = left + (size_t)X * 3 + ch;
            if (SYM) {
                T[c] = colp[(size_t)y * W * 6];
                for (int i = 1; i <= cy; ++i) {
                    const int ya = reflect_idx(y - i, H), yb = reflect_idx(y + i, H);
                    T[i * cols + c] = (uint16_t)colp[(size_t)ya * W * 6] + (uint16_t)colp[(size_t)yb * W * 6];
                }
            } else {
                for (int i = 0; i < ky; ++i) T[i * cols + c] = colp[(size_t)reflect_idx(y + i - cy, H) * W * 6];
            }
        }
        __syncwarp();
        // ---- evaluate: lane = 3*slot + channel, 10 holes per pass ----
        const int nh = __popc(m);
        const int slot = lane / 3, ch = lane - slot * 3;
        for (int h0 = 0; h0 < nh; h0 += 10) {
            const int hi = h0 + slot;
            if (lane < 30 && hi < nh) {
                const int xo = __fns(m, 0, hi + 1);                  // bit position of the hi-th hole
                const int cc = 3 * (xo + cx) + ch;                   // its byte column in the footprint
                uint32_t result;
                if (SYM) {
                    uint32_t acc[3] = {0u, 0u, 0u};
                    const int nw = (cy + 1) * (cx + 1);
                    for (int i = 0; i <= cy; ++i) {
                        const uint16_t *Ti = T + i * cols + cc;
                        const uint32_t *wr = a.wq + i * (cx + 1);
                        for (int j = 0; j <= cx; ++j) {
                            const uint32_t v = j ? (uint32_t)Ti[-3 * j] + (uint32_t)Ti[3 * j] : (uint32_t)Ti[0];
#pragma unroll
                            for (int p = 0; p < PARTS; ++p) acc[p] += v * __ldg(wr + p * nw + j);
                        }
                    }
                    unsigned long long total = 0ull;
#pragma unroll
                    for (int p = PARTS - 1; p >= 0; --p) total = (total << PBITS) + acc[p];
                    const int S = a.wshift;
                    unsigned long long q = total >> S;
                    const unsigned long long r = total & ((1ull << S) - 1ull), half = 1ull << (S - 1);
                    q += (r > half || (r == half && (q & 1ull))) ? 1ull : 0ull;
                    result = (uint32_t)q;
                } else {
                    double acc = 0.0;
                    for (int i = 0; i < ky; ++i) {
                        const uint16_t *Ti = T + i * cols + cc - 3 * cx;
                        const float *wr = a.weights + i * kx;
                        for (int j = 0; j < kx; ++j) acc = fma((double)__ldg(wr + j), (double)Ti[3 * j], acc);
                    }
                    result = (uint32_t)__double2int_rn(acc);
                }
                a.plane[((size_t)row * W + xw + xo) * 3 + ch] = (uint8_t)result;
            }
        }
    }
}

// ---- specialised hole blur: kernel size known at compile time, integer weights in the kernel parameters ----
// Same algorithm as k_blur_holes, three things tightened (the generic kernel was issue bound at ~980
// warp-instructions per listed word in an early capture of this round):
//   * weights live in the parameter constant bank and every (i,j) is unrolled, so a tap costs two LDS.U16, one
//     add and PARTS IMADs with a constant operand - no weight loads, no loop counters;
//   * interior words stage their footprint with aligned 32-bit loads (one or two per lane per row) and build the
//     vertical pair sums on packed bytes (2 x 16-bit lanes per register);
//   * hole positions come from a rank table instead of __fns.
//   * screening: every hole is first evaluated with ONE multiply per tap, using h = floor(w * 2^S1) (S1 <= 24 so the
//     sum fits 32 bits).  The exact total is that sum plus a remainder in [0, rmax] (rmax = 255 * sum of the dropped
//     weight bits, computed on the host), so unless the fractional part lies within rmax below one half - or
//     exactly on it, a possible tie - the rounded value is already decided.  Only the undecided lanes (0.1-0.3 %
//     of the holes) run the exact PARTS-way sum.
template <int PARTS, int CX, int CY>
struct BlurWeights { uint32_t q[PARTS][(CY + 1) * (CX + 1)]; uint32_t h[(CY + 1) * (CX + 1)]; uint32_t s1, rmax; };

// list entries staged per warp before their holes are evaluated together (fewer for the large 4K footprint, whose
// staging buffers would otherwise cost occupancy)
template <int CX> __host__ __device__ constexpr int blur_group() { return CX <= 6 ? 4 : 2; }

template <int PARTS, int CX, int CY>
__global__ void __launch_bounds__(256, 4) k_blur_holes_fixed(BlurArgs a, const __grid_constant__ BlurWeights<PARTS, CX, CY> wts) {
    constexpr int PBITS = PARTS == 2 ? 15 : 13;
    constexpr int KX = 2 * CX + 1, NPX = 32 + KX - 1;
    constexpr int PHASE = ((-3 * CX) % 4 + 4) % 4;                    // (96 w - 3 CX) mod 4: the same for every word
    constexpr int NWORDS = (PHASE + 3 * NPX + 3) / 4;                 // aligned 32-bit words that cover the footprint
    constexpr int COLS = (NWORDS * 4 + 7) / 8 * 8;                    // u16 columns per T row (16-byte multiple)
    constexpr int TSZ = (CY + 1) * COLS + 32;                         // u16 per staged entry: T rows + rank->bit table
    constexpr int G = blur_group<CX>();
    constexpr int NP = CX > 6 ? 2 : 1;                                // pixels per evaluation lane (see "pair slots" below)
    constexpr bool PAIR = CY <= 5;                                    // row loads of two entries in flight (registers allow it)
    static_assert(NWORDS <= 64, "footprint wider than two words per lane");
    static_assert(blur_group<CX>() % 2 == 0, "entries are staged in pairs");
    extern __shared__ __align__(16) uint8_t blur_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const int W = a.W, H = a.H;
    uint16_t *Tw = reinterpret_cast<uint16_t *>(blur_smem) + (size_t)warp * (G * TSZ);
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t count = *a.hole_count;
    const size_t pitch = (size_t)W * 6;

    // entry metadata, one group ahead: lane g (< G) holds (list entry, mask with the strip columns removed) of entry
    // e0 + g, so the list -> mask / strip load chain of the next group overlaps this group's staging and evaluation
    auto fetch_meta = [&](uint32_t e0, uint32_t &ent, uint32_t &m) {
        ent = 0u; m = 0u;
        if (lane < G && e0 < count && e0 + lane < count) {
            ent = a.hole_list[e0 + lane];
            const uint32_t row = ent >> 8;
            const int xw = (int)(ent & 0xffu) * 32;
            const int strip = a.tabs[(int)(((unsigned long long)row * a.magic_h) >> 40)].strip;
            m = a.hole_mask[(size_t)row * a.Wwords + (ent & 0xffu)];
            if (strip > xw) m = (strip - xw >= 32) ? 0u : (m & ~((1u << (strip - xw)) - 1u));
        }
    };
    const uint32_t stride = gridDim.x * nwarps * G;
    uint32_t ent_n, m_n;
    fetch_meta((blockIdx.x * nwarps + warp) * G, ent_n, m_n);
    for (uint32_t e0 = (blockIdx.x * nwarps + warp) * G; e0 < count; e0 += stride) {
        const uint32_t ent_c = ent_n, m_c = m_n;
        fetch_meta(e0 + stride, ent_n, m_n);
        __syncwarp();
        // ---- stage up to G entries; remember (row, first pixel, phase) and the running task count of each ----
        uint32_t g_row[G], g_xw[G], g_phase[G], g_end[G];
        uint32_t tasks = 0;
        // interior entries, two at a time: all row loads of both in flight before the pair sums are built
        auto stage_interior_load = [&](const uint8_t *left, int y, int s0, uint32_t (&v)[2 * CY + 1]) {
            if (NWORDS < 32 && lane >= NWORDS) return;                   // narrow footprints (9 x 7): fewer words than lanes
            const uint8_t *colp = left + (s0 - PHASE) + 4 * lane;
            if ((y - CY >= 0) && (y + CY < H)) {
                const uint8_t *p0 = colp + (size_t)(y - CY) * pitch;
#pragma unroll
                for (int i = 0; i <= 2 * CY; ++i) v[i] = __ldg(reinterpret_cast<const uint32_t *>(p0 + (size_t)i * pitch));
            } else {
#pragma unroll
                for (int i = 0; i <= 2 * CY; ++i)
                    v[i] = __ldg(reinterpret_cast<const uint32_t *>(colp + (size_t)reflect_idx(y + i - CY, H) * pitch));
            }
        };
        auto stage_interior_store = [&](uint16_t *T, int k, const uint32_t (&v)[2 * CY + 1]) {
            if (k >= NWORDS) return;
#pragma unroll
            for (int i = 0; i <= CY; ++i) {                              // bytes -> u16 lanes, pair sums on two packed u16 each
                const uint32_t p = v[CY - i], q = v[CY + i];
                if (PAIR) {                                              // measured: PRMT expansion wins at 1080p, mask/shift at 4K
                    uint32_t lo = __byte_perm(p, 0u, 0x4140), hi = __byte_perm(p, 0u, 0x4342);
                    if (i) { lo += __byte_perm(q, 0u, 0x4140); hi += __byte_perm(q, 0u, 0x4342); }
                    *reinterpret_cast<uint2 *>(T + i * COLS + 4 * k) = make_uint2(lo, hi);
                } else {
                    uint32_t lo = p & 0x00ff00ffu, hi = (p >> 8) & 0x00ff00ffu;
                    if (i) { lo += q & 0x00ff00ffu; hi += (q >> 8) & 0x00ff00ffu; }
                    *reinterpret_cast<uint2 *>(T + i * COLS + 4 * k) = make_uint2(__byte_perm(lo, hi, 0x5410), __byte_perm(lo, hi, 0x7632));
                }
            }
        };
        // the few words past the 32nd (wide footprints): (word, row pair) tasks spread over the lanes
        auto stage_interior_tail = [&](uint16_t *T, const uint8_t *left, int y, int s0) {
            constexpr int TAILN = (NWORDS > 32 ? NWORDS - 32 : 0) * (CY + 1);
#pragma unroll
            for (int tt = 0; tt < (TAILN + 31) / 32; ++tt) {
                const int t = lane + 32 * tt;
                if (t >= TAILN) continue;
                const int wd = t / (CY + 1), i = t - wd * (CY + 1);
                const uint8_t *colp = left + (s0 - PHASE) + 4 * (32 + wd);
                const uint32_t p = __ldg(reinterpret_cast<const uint32_t *>(colp + (size_t)reflect_idx(y - i, H) * pitch));
                uint32_t lo = __byte_perm(p, 0u, 0x4140), hi = __byte_perm(p, 0u, 0x4342);
                if (i) {
                    const uint32_t q = __ldg(reinterpret_cast<const uint32_t *>(colp + (size_t)reflect_idx(y + i, H) * pitch));
                    lo += __byte_perm(q, 0u, 0x4140); hi += __byte_perm(q, 0u, 0x4342);
                }
                *reinterpret_cast<uint2 *>(T + i * COLS + 4 * (32 + wd)) = make_uint2(lo, hi);
            }
        };
#pragma unroll
        for (int g2 = 0; g2 < G; g2 += 2) {
            uint32_t v0[2 * CY + 1], v1[2 * CY + 1];
            bool inner[2] = {false, false};
            int yy[2] = {0, 0}, ss[2] = {0, 0};
            const uint8_t *lf[2] = {nullptr, nullptr};
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int g = g2 + h;
                g_row[g] = 0; g_xw[g] = 0; g_phase[g] = 0; g_end[g] = tasks;
                const uint32_t ent = __shfl_sync(0xffffffffu, ent_c, g), m = __shfl_sync(0xffffffffu, m_c, g);
                if (m == 0u) continue;
                const uint32_t row = ent >> 8, w = ent & 0xffu;
                const int b = (int)(((unsigned long long)row * a.magic_h) >> 40), y = (int)row - b * H;
                const int xw = (int)w * 32;
                uint16_t *T = Tw + g * TSZ;
                uint16_t *pos = T + (CY + 1) * COLS;
                // NP == 2, pair slots: pixels (2s, 2s+1) of the word are evaluated by one lane per channel (they share all
                // but two of their taps); a slot is listed when either pixel is a hole.  NP == 1: one hole per lane.
                const uint32_t pm = NP == 2 ? ((m | (m >> 1)) & 0x55555555u) : m;
                if ((pm >> lane) & 1u) pos[__popc(pm & ((1u << lane) - 1u))] = (uint16_t)lane;
                if (NP == 2 && lane == 0) { pos[16] = (uint16_t)(m & 0xffffu); pos[17] = (uint16_t)(m >> 16); }
                const uint8_t *left = a.sbs + (size_t)b * H * pitch;
                const int s0 = 3 * (xw - CX);                             // byte offset of the footprint in its row
                if (xw - CX >= 0 && xw + 31 + CX < W) {
                    g_phase[g] = PHASE;
                    inner[h] = true; yy[h] = y; ss[h] = s0; lf[h] = left;
                    if (h == 0) stage_interior_load(left, y, s0, v0); else stage_interior_load(left, y, s0, v1);
                    if (!PAIR) {                                          // large footprint: one entry's rows in registers at a time
                        stage_interior_store(T, lane, h == 0 ? v0 : v1);
                        if (NWORDS > 32 && lane < NWORDS - 32) {
                            uint32_t v2[2 * CY + 1];
                            stage_interior_load(left + 128, y, s0, v2);
                            stage_interior_store(T, lane + 32, v2);
                        }
                        inner[h] = false;
                    }
                } else {
                    // border words: byte by byte with reflect padding
                    for (int c = lane; c < 3 * NPX; c += 32) {
                        const int px = c / 3, ch = c - px * 3;
                        const int X = min(max(reflect_idx(xw - CX + px, W), 0), W - 1);
                        const uint8_t *colp = left + (size_t)X * 3 + ch;
                        T[c] = colp[(size_t)y * pitch];
#pragma unroll
                        for (int i = 1; i <= CY; ++i)
                            T[i * COLS + c] = (uint16_t)colp[(size_t)reflect_idx(y - i, H) * pitch] + (uint16_t)colp[(size_t)reflect_idx(y + i, H) * pitch];
                    }
                }
                g_row[g] = row; g_xw[g] = (uint32_t)xw;
                tasks += 3u * (uint32_t)__popc(pm);
                g_end[g] = tasks;
            }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (!inner[h]) continue;
                uint16_t *T = Tw + (g2 + h) * TSZ;
                if (h == 0) stage_interior_store(T, lane, v0); else stage_interior_store(T, lane, v1);
                if (NWORDS > 32) stage_interior_tail(T, lf[h], yy[h], ss[h]);
            }
        }
        __syncwarp();
        // ---- evaluate: task = (entry, hole rank, channel), 32 tasks per pass ----
        for (uint32_t t0 = 0; t0 < tasks; t0 += 32) {
            const uint32_t task = t0 + lane;
            if (task < tasks) {
                int g = 0;
                uint32_t start = 0;
#pragma unroll
                for (int k = 0; k < G - 1; ++k)
                    if (task >= g_end[k]) { g = k + 1; start = g_end[k]; }
                uint32_t row = g_row[0], xw = g_xw[0], phase = g_phase[0];
#pragma unroll
                for (int k = 1; k < G; ++k)
                    if (g == k) { row = g_row[k]; xw = g_xw[k]; phase = g_phase[k]; }
                const uint32_t rel = task - start, rank = rel / 3u, ch = rel - rank * 3u;
                const uint16_t *T = Tw + g * TSZ;
                const uint16_t *pos = T + (CY + 1) * COLS;
                const int xo = pos[rank];                                     // NP == 2: even, first pixel of the pair
                const uint32_t mk = NP == 2 ? ((uint32_t)pos[16] | ((uint32_t)pos[17] << 16)) >> xo : 1u;   // bit o: pixel xo+o is a hole
                const uint16_t *Tc = T + phase + 3 * (xo + CX) + ch;
                // screening sums of the NP pixels: columns xo-CX .. xo+NP-1+CX of row i are read once
                uint32_t a1[2] = {0u, 0u};
#pragma unroll
                for (int i = 0; i <= CY; ++i) {
                    uint32_t u[2 * CX + NP];
#pragma unroll
                    for (int k = 0; k < 2 * CX + NP; ++k) u[k] = Tc[i * COLS + 3 * (k - CX)];
#pragma unroll
                    for (int j = 0; j <= CX; ++j) {
                        const uint32_t hw = wts.h[i * (CX + 1) + j];
#pragma unroll
                        for (int o = 0; o < NP; ++o) a1[o] += (j ? u[CX + o - j] + u[CX + o + j] : u[CX + o]) * hw;
                    }
                }
                const uint32_t half1 = 1u << (wts.s1 - 1u);
                uint32_t q[2] = {0u, 0u};
                bool open[2] = {false, false};
#pragma unroll
                for (int o = 0; o < NP; ++o) {
                    const uint32_t r1 = a1[o] & (2u * half1 - 1u);
                    q[o] = (a1[o] >> wts.s1) + (r1 > half1 ? 1u : 0u);
                    open[o] = ((mk >> o) & 1u) && !(r1 > half1 || r1 + wts.rmax < half1);
                }
                // undecided by the screening sum: exact total in PARTS 32-bit accumulators
                auto exact = [&](const uint16_t *To) {
                    uint32_t acc[PARTS];
#pragma unroll
                    for (int p = 0; p < PARTS; ++p) acc[p] = 0u;
#pragma unroll 1
                    for (int i = 0; i <= CY; ++i) {
#pragma unroll
                        for (int j = 0; j <= CX; ++j) {
                            const uint32_t v = j ? (uint32_t)To[i * COLS - 3 * j] + (uint32_t)To[i * COLS + 3 * j] : (uint32_t)To[i * COLS];
#pragma unroll
                            for (int p = 0; p < PARTS; ++p) acc[p] += v * wts.q[p][i * (CX + 1) + j];
                        }
                    }
                    unsigned long long total = 0ull;
#pragma unroll
                    for (int p = PARTS - 1; p >= 0; --p) total = (total << PBITS) + acc[p];
                    const int S = a.wshift;
                    unsigned long long qq = total >> S;
                    const unsigned long long r = total & ((1ull << S) - 1ull), half = 1ull << (S - 1);
                    qq += (r > half || (r == half && (qq & 1ull))) ? 1ull : 0ull;
                    return (uint32_t)qq;
                };
                if (NP == 2) {
                    // (written out rather than calling `exact`: ptxas keeps everything in registers this way)
#pragma unroll 1
                    for (int o = 0; o < 2; ++o) {
                        if (!(o ? open[1] : open[0])) continue;
                        const uint16_t *To = Tc + 3 * o;
                        uint32_t acc[PARTS];
#pragma unroll
                        for (int p = 0; p < PARTS; ++p) acc[p] = 0u;
#pragma unroll 1
                        for (int i = 0; i <= CY; ++i) {
#pragma unroll
                            for (int j = 0; j <= CX; ++j) {
                                const uint32_t v = j ? (uint32_t)To[i * COLS - 3 * j] + (uint32_t)To[i * COLS + 3 * j] : (uint32_t)To[i * COLS];
#pragma unroll
                                for (int p = 0; p < PARTS; ++p) acc[p] += v * wts.q[p][i * (CX + 1) + j];
                            }
                        }
                        unsigned long long total = 0ull;
#pragma unroll
                        for (int p = PARTS - 1; p >= 0; --p) total = (total << PBITS) + acc[p];
                        const int S = a.wshift;
                        unsigned long long qq = total >> S;
                        const unsigned long long r = total & ((1ull << S) - 1ull), half = 1ull << (S - 1);
                        qq += (r > half || (r == half && (qq & 1ull))) ? 1ull : 0ull;
                        if (o) q[1] = (uint32_t)qq; else q[0] = (uint32_t)qq;
                    }
                } else if (open[0]) {
                    q[0] = exact(Tc);
                }
                uint8_t *dst = a.plane + ((size_t)row * W + xw + xo) * 3 + ch;
                if (mk & 1u) dst[0] = (uint8_t)q[0];
                if (NP == 2 && (mk & 2u)) dst[3] = (uint8_t)q[1];
            }
        }
    }
}
template <int CX, int CY>
__host__ __device__ constexpr size_t blur_fixed_warp_smem() {
    constexpr int PHASE = ((-3 * CX) % 4 + 4) % 4;
    constexpr int NWORDS = (PHASE + 3 * (32 + 2 * CX) + 3) / 4, COLS = (NWORDS * 4 + 7) / 8 * 8;
    return (size_t)blur_group<CX>() * ((size_t)(CY + 1) * COLS + 32) * sizeof(uint16_t);
}

// ---- separable screening: the default hole blur for torchvision's gaussians ------------------------------------
// The reference's 2-D kernel is w[i][j] = fl32(ky[i] * kx[j]): not separable bit for bit, but within 2^-24 relative
// of a rank-1 kernel.  The blurred value is defined by the EXACT sum of w[i][j] * pixel (see above); what decides the
// output byte is on which side of a half-integer that sum falls.  So the sum is first evaluated with a separable
// integer kernel A[i][j] = hy[|i-cy|] * hx[|j-cx|] (scale 2^s, s = 52): one vertical pass per staged byte column
// (V[c] = sum_i hy[i] * (row(y-i)[c] + row(y+i)[c]), 32-bit, done once per footprint column while staging) and one
// horizontal pass per hole and channel (sum_j hx[j] * (V[c-3j] + V[c+3j]), CX+1 64-bit multiply-adds instead of
// (CX+1)(CY+1) taps).  The host computes, in exact integer arithmetic over the caller's actual weights, a bound eps on
// |exact - separable| * 2^s for any pixel content (vrsbs_set_blur_weights); a value whose separable sum lies further
// than eps from every half-integer is decided.  The others - about 2 * eps = 0.02-0.04 % of the values, which includes
// every true tie - are recomputed EXACTLY (PARTS-way integer sum over the 2-D weights, straight from global memory).
// The result is therefore bit-identical to k_blur_holes / k_blur_holes_fixed and to the oracle for every input;
// tests run every pixel of all-hole frames through all three.
template <int CX, int CY>
struct BlurSepWeights { uint32_t hy[CY + 1]; uint32_t hx[CX + 1]; uint32_t s, eps32; };

constexpr int kBlurSepGroup = 8;                   // list entries staged per warp before their holes are evaluated together

template <int CX>
__host__ __device__ constexpr int blur_sep_vcols() {
    constexpr int PHASE = ((-3 * CX) % 4 + 4) % 4;
    return (PHASE + 3 * (32 + 2 * CX) + 3) / 4 * 4;
}
template <int CX>
__host__ __device__ constexpr size_t blur_sep_warp_smem() { return (size_t)kBlurSepGroup * (blur_sep_vcols<CX>() + 16) * sizeof(uint32_t); }

// exact blurred value of one (pixel, channel) from global memory: the rare path behind the separable screening.
// The WHOLE WARP evaluates one undecided value: the ky * kx taps are dealt out over the lanes and the PARTS partial sums
// are added across the warp (their totals fit 32 bits for the same reason the single-thread accumulators of
// k_blur_holes do).  One undecided lane used to walk all taps alone while 31 lanes waited: 1.2 % of the task passes of
// a 4K frame contain such a lane, and those passes were 27 % of the kernel's executed instructions (profiles r02a_4k).
template <int PARTS, int CX, int CY>
__device__ __noinline__ uint32_t blur_exact_warp(const uint8_t *left, size_t pitch, int H, int W, int y, int x, int ch,
                                                 const uint32_t *wq, int S) {
    constexpr int PBITS = PARTS == 2 ? 15 : 13;
    constexpr int KX = 2 * CX + 1, KY = 2 * CY + 1, NTAP = KX * KY, NU = (CY + 1) * (CX + 1);
    const int lane = threadIdx.x & 31;
    uint32_t acc[PARTS];
#pragma unroll
    for (int p = 0; p < PARTS; ++p) acc[p] = 0u;
    for (int t = lane; t < NTAP; t += 32) {
        const int i = t / KX, j = t - i * KX;
        const int di = i - CY, dj = j - CX;
        const int yy = reflect_idx(y + di, H), xx = min(max(reflect_idx(x + dj, W), 0), W - 1);
        const uint32_t v = left[(size_t)yy * pitch + (size_t)xx * 3 + ch];
        const int wi = (di < 0 ? -di : di) * (CX + 1) + (dj < 0 ? -dj : dj);
#pragma unroll
        for (int p = 0; p < PARTS; ++p) acc[p] += v * __ldg(wq + p * NU + wi);
    }
    unsigned long long total = 0ull;
#pragma unroll
    for (int p = PARTS - 1; p >= 0; --p) total = (total << PBITS) + __reduce_add_sync(0xffffffffu, acc[p]);
    unsigned long long q = total >> S;
    const unsigned long long r = total & ((1ull << S) - 1ull), half = 1ull << (S - 1);
    q += (r > half || (r == half && (q & 1ull))) ? 1ull : 0ull;
    return (uint32_t)q;
}

template <int PARTS, int CX, int CY>
__global__ void __launch_bounds__(256, 4) k_blur_sep(BlurArgs a, const __grid_constant__ BlurSepWeights<CX, CY> wts) {
    constexpr int KX = 2 * CX + 1, NPX = 32 + KX - 1;
    constexpr int PHASE = ((-3 * CX) % 4 + 4) % 4;                    // (96 w - 3 CX) mod 4: the same for every word
    constexpr int NWORDS = (PHASE + 3 * NPX + 3) / 4;                 // aligned 32-bit words that cover the footprint
    constexpr int VCOLS = NWORDS * 4;
    constexpr int G = kBlurSepGroup;
    static_assert(NWORDS <= 64 && VCOLS == blur_sep_vcols<CX>(), "footprint layout");
    extern __shared__ __align__(16) uint8_t blur_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = a.W, H = a.H;
    uint32_t *Vw = reinterpret_cast<uint32_t *>(blur_smem) + (size_t)warp * (G * (VCOLS + 16));
    uint16_t *tasktab = reinterpret_cast<uint16_t *>(Vw + G * VCOLS);  // [G * 32]: (entry << 8 | pixel) of every hole of the group
    pdl_launch_dependents();
    pdl_wait();                                  // the warp kernel's frame, hole mask and hole list
    const uint32_t count = *a.hole_count;
    const size_t pitch = (size_t)W * 6;

    // entry metadata, one group ahead: lane g (< G) holds (list entry, mask with the strip columns removed) of entry e0 + g
    auto fetch_meta = [&](uint32_t e0, uint32_t &ent, uint32_t &m) {
        ent = 0u; m = 0u;
        if (lane < G && e0 < count && e0 + lane < count) {
            ent = a.hole_list[e0 + lane];
            const uint32_t row = ent >> 8;
            const int xw = (int)(ent & 0xffu) * 32;
            const int strip = a.tabs[(int)(((unsigned long long)row * a.magic_h) >> 40)].strip;
            m = a.hole_mask[(size_t)row * a.Wwords + (ent & 0xffu)];
            if (strip > xw) m = (strip - xw >= 32) ? 0u : (m & ~((1u << (strip - xw)) - 1u));
        }
    };
    auto decode = [&](uint32_t ent, int &b, int &y, int &xw) {
        const uint32_t row = ent >> 8;
        b = (int)(((unsigned long long)row * a.magic_h) >> 40);
        y = (int)row - b * H;
        xw = (int)(ent & 0xffu) * 32;
    };
    // vertical pass on one aligned word (4 byte columns) of the footprint: V = sum_i hy[i] * (row(y-i) + row(y+i)), on packed
    // u16 pair sums; the row pairs are loaded CH at a time (2 * CH loads in flight per lane)
    auto stage_word = [&](uint32_t *Vdst, const uint8_t *colp, int y) {
        constexpr int CH = 4;
        const bool inside = (y - CY >= 0) && (y + CY < H);
        const uint8_t *pc = colp + (size_t)y * pitch;
        const uint32_t c0 = __ldg(reinterpret_cast<const uint32_t *>(pc));
        uint32_t V0, V1, V2, V3;
#pragma unroll
        for (int i0 = 1; i0 <= CY; i0 += CH) {
            uint32_t p[CH], q[CH];
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const int i = i0 + k;
                if (i > CY) continue;
                if (inside) {
                    p[k] = __ldg(reinterpret_cast<const uint32_t *>(pc - (size_t)i * pitch));
                    q[k] = __ldg(reinterpret_cast<const uint32_t *>(pc + (size_t)i * pitch));
                } else {
                    p[k] = __ldg(reinterpret_cast<const uint32_t *>(colp + (size_t)reflect_idx(y - i, H) * pitch));
                    q[k] = __ldg(reinterpret_cast<const uint32_t *>(colp + (size_t)reflect_idx(y + i, H) * pitch));
                }
            }
            if (i0 == 1) {
                const uint32_t lo = __byte_perm(c0, 0u, 0x4140), hi = __byte_perm(c0, 0u, 0x4342);
                V0 = (lo & 0xffffu) * wts.hy[0]; V1 = (lo >> 16) * wts.hy[0]; V2 = (hi & 0xffffu) * wts.hy[0]; V3 = (hi >> 16) * wts.hy[0];
            }
#pragma unroll
            for (int k = 0; k < CH; ++k) {
                const int i = i0 + k;
                if (i > CY) continue;
                const uint32_t lo = __byte_perm(p[k], 0u, 0x4140) + __byte_perm(q[k], 0u, 0x4140);
                const uint32_t hi = __byte_perm(p[k], 0u, 0x4342) + __byte_perm(q[k], 0u, 0x4342);
                V0 += (lo & 0xffffu) * wts.hy[i]; V1 += (lo >> 16) * wts.hy[i];
                V2 += (hi & 0xffffu) * wts.hy[i]; V3 += (hi >> 16) * wts.hy[i];
            }
        }
        if (CY == 0) {
            const uint32_t lo = __byte_perm(c0, 0u, 0x4140), hi = __byte_perm(c0, 0u, 0x4342);
            V0 = (lo & 0xffffu) * wts.hy[0]; V1 = (lo >> 16) * wts.hy[0]; V2 = (hi & 0xffffu) * wts.hy[0]; V3 = (hi >> 16) * wts.hy[0];
        }
        *reinterpret_cast<uint4 *>(Vdst) = make_uint4(V0, V1, V2, V3);
    };

    // Groups of G list entries are handed out by a global ticket counter (the work per entry varies with its hole count,
    // and a static split leaves a quarter of the warp slots idle at the end).  Tickets run two groups ahead: the atomic's
    // round trip and the list -> mask / strip load chain of the next group both overlap the current group's work.
    auto grab = [&]() -> uint32_t { return lane == 0 ? atomicAdd(a.ticket, 1u) : 0u; };
    uint32_t t_a = __shfl_sync(0xffffffffu, grab(), 0);
    uint32_t t_b_raw = grab();
    uint32_t ent_n, m_n;
    fetch_meta(t_a * G, ent_n, m_n);
    for (;;) {
        const uint32_t e0 = t_a * G;
        if (e0 >= count) break;
        const uint32_t ent_c = ent_n, m_c = m_n;
        t_a = __shfl_sync(0xffffffffu, t_b_raw, 0);
        fetch_meta(t_a * G, ent_n, m_n);
        t_b_raw = grab();
        __syncwarp();
        // ---- stage: hole table of up to G entries; border entries byte by byte ----
        uint32_t nholes = 0, border = 0;
#pragma unroll 2
        for (int g = 0; g < G; ++g) {
            const uint32_t ent = __shfl_sync(0xffffffffu, ent_c, g), m = __shfl_sync(0xffffffffu, m_c, g);
            if (m == 0u) continue;
            if ((m >> lane) & 1u) tasktab[nholes + __popc(m & ((1u << lane) - 1u))] = (uint16_t)((g << 8) | lane);
            nholes += (uint32_t)__popc(m);
            const int xw = (int)(ent & 0xffu) * 32;
            if (xw - CX >= 0 && xw + 31 + CX < W) continue;
            // border words: byte by byte with reflect padding (phase 0)
            border |= 1u << g;
            int b, y, xw_;
            decode(ent, b, y, xw_);
            const uint8_t *left = a.sbs + (size_t)b * H * pitch;
            uint32_t *V = Vw + g * VCOLS;
            for (int c = lane; c < 3 * NPX; c += 32) {
                const int px = c / 3, ch = c - px * 3;
                const int X = min(max(reflect_idx(xw - CX + px, W), 0), W - 1);
                const uint8_t *colp = left + (size_t)X * 3 + ch;
                uint32_t acc = (uint32_t)colp[(size_t)y * pitch] * wts.hy[0];
#pragma unroll
                for (int i = 1; i <= CY; ++i)
                    acc += ((uint32_t)colp[(size_t)reflect_idx(y - i, H) * pitch] + (uint32_t)colp[(size_t)reflect_idx(y + i, H) * pitch]) * wts.hy[i];
                V[c] = acc;
            }
        }
        // ---- stage: V columns of the interior entries.  Only the aligned words a hole's taps can reach are built: the
        // holes of a word sit in bits lo..hi, their taps in footprint bytes PHASE + 3 lo .. PHASE + 3 (hi + 2 CX) + 2.  A
        // listed word holds 5 holes on average at 1080p (a third of the footprint is needed); the (entry, word) tasks of
        // the whole group are dealt out 32 at a time, so sparse entries share a pass.
        {
            uint32_t wlo = 0, cnt = 0;
            if (lane < G && m_c != 0u && !((border >> lane) & 1u)) {
                const int lo = __ffs((int)m_c) - 1, hi = 31 - __clz((int)m_c);
                wlo = (uint32_t)(PHASE + 3 * lo) >> 2;
                cnt = ((uint32_t)(PHASE + 3 * (hi + 2 * CX) + 2) >> 2) - wlo + 1u;
            }
            uint32_t incl = cnt;                                         // inclusive prefix sum over lanes 0 .. G-1
#pragma unroll
            for (int d = 1; d < G; d <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
                if (lane >= d) incl += t;
            }
            const uint32_t total = __shfl_sync(0xffffffffu, incl, G - 1);
            const uint32_t excl = incl - cnt;
            for (uint32_t base = 0; base < total; base += 32) {
                const uint32_t task = base + lane;
                int g = 0;                                               // number of entries that end at or before this task
#pragma unroll
                for (int step = G / 2; step >= 1; step >>= 1) {
                    const uint32_t probe = __shfl_sync(0xffffffffu, incl, g + step - 1);
                    if (probe <= task) g += step;
                }
                g = min(g, G - 1);
                const uint32_t ent = __shfl_sync(0xffffffffu, ent_c, g);
                const uint32_t k = task - __shfl_sync(0xffffffffu, excl, g) + __shfl_sync(0xffffffffu, wlo, g);
                if (task < total) {
                    int b, y, xw;
                    decode(ent, b, y, xw);
                    stage_word(Vw + g * VCOLS + 4 * k, a.sbs + (size_t)b * H * pitch + (3 * (xw - CX) - PHASE) + 4 * (int)k, y);
                }
            }
        }
        __syncwarp();
        // ---- evaluate: task = (hole of the group, channel), 32 tasks per pass ----
        const uint32_t tasks = 3u * nholes;
        for (uint32_t t0 = 0; t0 < tasks; t0 += 32) {
            const uint32_t task = t0 + lane;
            const bool on = task < tasks;
            const uint32_t hidx = on ? task / 3u : 0u, ch = task - hidx * 3u;
            const uint32_t code = tasktab[hidx];
            const uint32_t g = code >> 8, xo = code & 0xffu;
            const uint32_t ent = __shfl_sync(0xffffffffu, ent_c, g);
            uint32_t q = 0u;
            bool open = false;
            if (on) {
                const uint32_t phase = ((border >> g) & 1u) ? 0u : (uint32_t)PHASE;
                const uint32_t *Vc = Vw + g * VCOLS + phase + 3u * (xo + CX) + ch;
                unsigned long long acc = (unsigned long long)Vc[0] * wts.hx[0];
#pragma unroll
                for (int j = 1; j <= CX; ++j) acc += (unsigned long long)(Vc[-3 * j] + Vc[3 * j]) * wts.hx[j];
                const uint32_t r32 = (uint32_t)(acc >> (wts.s - 32u));             // top 32 bits of the fractional part
                q = (uint32_t)(acc >> wts.s) + (r32 >> 31);
                open = r32 - 0x80000000u + wts.eps32 <= 2u * wts.eps32;            // within eps of a half-integer: exact sum decides
            }
            // undecided values, one at a time, by the whole warp
            for (unsigned need = __ballot_sync(0xffffffffu, open); need; need &= need - 1u) {
                const int src = __ffs(need) - 1;
                const uint32_t e_s = __shfl_sync(0xffffffffu, ent, src), x_s = __shfl_sync(0xffffffffu, xo, src), c_s = __shfl_sync(0xffffffffu, ch, src);
                int b, y, xw_;
                decode(e_s, b, y, xw_);
                const uint32_t qx = blur_exact_warp<PARTS, CX, CY>(a.sbs + (size_t)b * H * pitch, pitch, H, W, y, xw_ + (int)x_s, (int)c_s, a.wq, a.wshift);
                if (lane == src) q = qx;
            }
            if (!on) continue;
            const uint32_t row = ent >> 8, xw = (ent & 0xffu) * 32u;
            a.plane[((size_t)row * W + xw + xo) * 3 + ch] = (uint8_t)q;
        }
    }
}

// hole mask -> per-word hole list (row << 8 | word) for the list-driven blur kernels, when the warp kernel did not append it
// itself: k_warp_ws is 0.02 ms per 64 frames faster without the list slots and the scattered list stores, this pass over the
// 16 MB mask costs a quarter of that
__global__ void __launch_bounds__(1024) k_word_list(const uint32_t *hole_mask, uint32_t *hole_list, uint32_t *hole_count, long long nwords, int Wwords) {
    __shared__ uint32_t s_cnt[32], s_base;
    pdl_launch_dependents();
    pdl_wait();
    constexpr int V = 4;                              // mask words per thread: a quarter of the CTAs, each one atomic round trip
    const long long i0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t v[V];
#pragma unroll
    for (int k = 0; k < V; ++k) v[k] = i0 + k < nwords ? hole_mask[i0 + k] : 0u;
    uint32_t cnt = 0;
#pragma unroll
    for (int k = 0; k < V; ++k) cnt += v[k] != 0u;
    uint32_t incl = cnt;                              // inclusive prefix sum over the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) s_cnt[warp] = incl;
    __syncthreads();
    if (warp == 0) {                                  // one list-slot request per CTA
        const uint32_t c = lane < (int)(blockDim.x >> 5) ? s_cnt[lane] : 0u;
        uint32_t ws = c;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t t = __shfl_up_sync(0xffffffffu, ws, d);
            if (lane >= d) ws += t;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, ws, 31);
        if (lane == 0) s_base = total ? atomicAdd(hole_count, total) : 0u;
        s_cnt[lane] = ws - c;                         // exclusive offsets of the warps
    }
    __syncthreads();
    uint32_t at = s_base + s_cnt[warp] + incl - cnt;
#pragma unroll
    for (int k = 0; k < V; ++k)
        if (v[k]) {
            const long long i = i0 + k, row = i / Wwords;
            hole_list[at++] = ((uint32_t)row << 8) | (uint32_t)(i - row * Wwords);
        }
}

// ---- band-driven hole blur for the small footprints (the footprint of a mask word fits 32 aligned words) ------------
// k_blur_sep stages every listed mask word (32 pixels of ONE row) on its own: nine row loads, nine byte expansions and the
// weighted sum per aligned word, although the word of the row below needs eight of the same nine rows.  Here the unit of
// work is a BAND COLUMN - one mask word column over kBandRows = 8 consecutive rows of one frame: the warp walks the 8 + 2 CY
// input rows once (lane = aligned word of the footprint), keeps them expanded in a register window, and emits the V columns
// of every row of the band that has holes.  Per staged row that is CY + 1 packed additions and 4 (CY + 1) multiply-adds
// instead of nine loads, eighteen expansions and the same arithmetic.  The evaluation (one horizontal pass per hole and
// channel, undecided values exactly, by the whole warp) is k_blur_sep's.  The band columns that hold a hole are marked in a
// bitmap - by k_warp_ws itself, one fire-and-forget OR per row (it then needs no slot from a global counter for a per-word
// list: 0.260 instead of 0.278 ms), or by k_band_map from the hole mask behind the other warp kernels; k_band_list compacts the
// bitmap into the work list of k_blur_band and k_blur_commit.  Results are bit-identical to k_blur_sep (same integer sums).
constexpr int kBandRows = 8;
static_assert(kBandRows == kBlurSepGroup, "a band's rows use the V buffers of a k_blur_sep group");

__global__ void __launch_bounds__(256) k_band_map(BlurArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const long long total = (long long)a.B * a.Hb * a.Wwords;
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int w = (int)(idx % a.Wwords);
    const long long q = idx / a.Wwords;                           // frame * Hb + band
    const int t = (int)(q / a.Hb), yb = (int)(q - (long long)t * a.Hb);
    const uint32_t *mp = a.hole_mask + ((size_t)t * a.H + (size_t)yb * kBandRows) * a.Wwords + w;
    uint32_t any = 0u;
#pragma unroll
    for (int r = 0; r < kBandRows; ++r)
        if (yb * kBandRows + r < a.H) any |= mp[(size_t)r * a.Wwords];
    if (any) atomicOr(a.band_map + (size_t)q * a.band_groups + (w >> 5), 1u << (w & 31));
}

// band map -> band list (and the map is zero again): one thread per map word, list slots by one atomic per warp
__global__ void __launch_bounds__(256) k_band_list(BlurArgs a) {
    pdl_launch_dependents();
    pdl_wait();
    const uint32_t nwords = (uint32_t)a.B * (uint32_t)a.Hb * (uint32_t)a.band_groups;
    const uint32_t wi = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    uint32_t bits = wi < nwords ? a.band_map[wi] : 0u;
    if (bits) a.band_map[wi] = 0u;
    const uint32_t cnt = (uint32_t)__popc(bits);
    uint32_t incl = cnt;                                          // inclusive prefix sum over the warp
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (total == 0u) return;
    uint32_t base = 0;
    if (lane == 31) base = atomicAdd(a.band_count, total);
    base = __shfl_sync(0xffffffffu, base, 31) + incl - cnt;
    const uint32_t q = wi / (uint32_t)a.band_groups, g = wi - q * (uint32_t)a.band_groups;
    for (; bits; bits &= bits - 1u) a.band_list[base++] = (q << 8) | (32u * g + (uint32_t)__ffs((int)bits) - 1u);
}

template <int PARTS, int CX, int CY>
__global__ void __launch_bounds__(256, 4) k_blur_band(BlurArgs a, const __grid_constant__ BlurSepWeights<CX, CY> wts) {
    constexpr int KX = 2 * CX + 1, NPX = 32 + KX - 1;
    constexpr int PHASE = ((-3 * CX) % 4 + 4) % 4;
    constexpr int NWORDS = (PHASE + 3 * NPX + 3) / 4;
    constexpr int VCOLS = NWORDS * 4;
    constexpr int RB = kBandRows, NR = RB + 2 * CY;                   // rows of a band, input rows it reads
    static_assert(NWORDS <= 32 && VCOLS == blur_sep_vcols<CX>(), "one aligned word of the footprint per lane");
    extern __shared__ __align__(16) uint8_t blur_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int W = a.W, H = a.H;
    uint32_t *Vw = reinterpret_cast<uint32_t *>(blur_smem) + (size_t)warp * (RB * (VCOLS + 16));
    uint16_t *tasktab = reinterpret_cast<uint16_t *>(Vw + RB * VCOLS);   // [RB * 32]: (row of the band << 8 | pixel) of every hole
    pdl_launch_dependents();
    pdl_wait();                                  // the band list
    const uint32_t count = *a.band_count;
    const size_t pitch = (size_t)W * 6;

    // unit metadata, one unit ahead: lane r (< RB) holds the mask word of row r of the band column (strip columns removed)
    auto fetch_meta = [&](uint32_t u, uint32_t &code, uint32_t &m) {
        code = 0u; m = 0u;
        if (u >= count) return;
        code = a.band_list[u];
        const uint32_t q = code >> 8, w = code & 0xffu;
        const int t = (int)(((unsigned long long)q * a.magic_hb) >> 40), y = ((int)q - t * a.Hb) * RB + lane;
        if (lane < RB && y < H) {
            const int strip = a.tabs[t].strip, xw = (int)w * 32;
            m = a.hole_mask[((size_t)t * H + y) * a.Wwords + w];
            if (strip > xw) m = (strip - xw >= 32) ? 0u : (m & ~((1u << (strip - xw)) - 1u));
        }
    };
    auto grab = [&]() -> uint32_t { return lane == 0 ? atomicAdd(a.ticket, 1u) : 0u; };
    uint32_t t_a = __shfl_sync(0xffffffffu, grab(), 0);
    uint32_t t_b_raw = grab();
    uint32_t code_n, m_n;
    fetch_meta(t_a, code_n, m_n);
    for (;;) {
        if (t_a >= count) break;
        const uint32_t code = code_n, m_c = m_n;
        t_a = __shfl_sync(0xffffffffu, t_b_raw, 0);
        fetch_meta(t_a, code_n, m_n);
        t_b_raw = grab();
        const uint32_t q = code >> 8;
        const int t = (int)(((unsigned long long)q * a.magic_hb) >> 40), y0 = ((int)q - t * a.Hb) * RB, xw = (int)(code & 0xffu) * 32;
        const uint8_t *left = a.sbs + (size_t)t * H * pitch;
        const uint32_t rmask = __ballot_sync(0xffffffffu, m_c != 0u) & ((1u << RB) - 1u);   // rows of the band that have holes
        __syncwarp();
        // ---- hole table of the band column ----
        uint32_t nholes = 0;
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const uint32_t m = __shfl_sync(0xffffffffu, m_c, r);
            if (m == 0u) continue;
            if ((m >> lane) & 1u) tasktab[nholes + __popc(m & ((1u << lane) - 1u))] = (uint16_t)((r << 8) | lane);
            nholes += (uint32_t)__popc(m);
        }
        const bool border = !(xw - CX >= 0 && xw + 31 + CX < W);
        if (!border) {
            // ---- vertical pass over the band: input rows y0 - CY .. y0 + RB - 1 + CY, expanded once, window in registers ----
            uint32_t need = 0u;                                       // input rows some hole row reads
#pragma unroll
            for (int i = 0; i <= 2 * CY; ++i) need |= rmask << i;
            const uint8_t *colp = left + (3 * (xw - CX) - PHASE) + 4 * lane;
            const bool wl = NWORDS >= 32 || lane < NWORDS;
            uint32_t raw[NR];
            if (y0 - CY >= 0 && y0 + RB - 1 + CY < H) {                // no reflected row: walk down the column
                const uint8_t *rp = colp + (size_t)(y0 - CY) * pitch;
#pragma unroll
                for (int j = 0; j < NR; ++j) {
                    raw[j] = (wl && ((need >> j) & 1u)) ? __ldg(reinterpret_cast<const uint32_t *>(rp)) : 0u;
                    rp += pitch;
                }
            } else {
#pragma unroll
                for (int j = 0; j < NR; ++j)
                    raw[j] = (wl && ((need >> j) & 1u)) ? __ldg(reinterpret_cast<const uint32_t *>(colp + (size_t)reflect_idx(y0 - CY + j, H) * pitch)) : 0u;
            }
            uint32_t lo[NR], hi[NR];
#pragma unroll
            for (int j = 0; j < NR; ++j) {
                lo[j] = __byte_perm(raw[j], 0u, 0x4140);
                hi[j] = __byte_perm(raw[j], 0u, 0x4342);
                if (j >= 2 * CY) {
                    const int r = j - 2 * CY, c = j - CY;
                    if ((rmask >> r) & 1u) {
                        uint32_t V0 = (lo[c] & 0xffffu) * wts.hy[0], V1 = (lo[c] >> 16) * wts.hy[0];
                        uint32_t V2 = (hi[c] & 0xffffu) * wts.hy[0], V3 = (hi[c] >> 16) * wts.hy[0];
#pragma unroll
                        for (int i = 1; i <= CY; ++i) {
                            const uint32_t l = lo[c - i] + lo[c + i], h = hi[c - i] + hi[c + i];
                            V0 += (l & 0xffffu) * wts.hy[i]; V1 += (l >> 16) * wts.hy[i];
                            V2 += (h & 0xffffu) * wts.hy[i]; V3 += (h >> 16) * wts.hy[i];
                        }
                        if (wl) *reinterpret_cast<uint4 *>(Vw + r * VCOLS + 4 * lane) = make_uint4(V0, V1, V2, V3);
                    }
                }
            }
        } else {
            // border band columns: byte by byte with reflect padding (phase 0), row by row
#pragma unroll 1
            for (int r = 0; r < RB; ++r) {
                if (!((rmask >> r) & 1u)) continue;
                const int y = y0 + r;
                uint32_t *V = Vw + r * VCOLS;
                for (int c = lane; c < 3 * NPX; c += 32) {
                    const int px = c / 3, ch = c - px * 3;
                    const int X = min(max(reflect_idx(xw - CX + px, W), 0), W - 1);
                    const uint8_t *cp = left + (size_t)X * 3 + ch;
                    uint32_t acc = (uint32_t)cp[(size_t)y * pitch] * wts.hy[0];
#pragma unroll
                    for (int i = 1; i <= CY; ++i)
                        acc += ((uint32_t)cp[(size_t)reflect_idx(y - i, H) * pitch] + (uint32_t)cp[(size_t)reflect_idx(y + i, H) * pitch]) * wts.hy[i];
                    V[c] = acc;
                }
            }
        }
        __syncwarp();
        // ---- evaluate: task = (hole of the band column, channel), 32 tasks per pass ----
        const uint32_t tasks = 3u * nholes, phase = border ? 0u : (uint32_t)PHASE;
        for (uint32_t t0 = 0; t0 < tasks; t0 += 32) {
            const uint32_t task = t0 + lane;
            const bool on = task < tasks;
            const uint32_t hidx = on ? task / 3u : 0u, ch = task - hidx * 3u;
            const uint32_t tc = tasktab[hidx];
            const uint32_t g = tc >> 8, xo = tc & 0xffu;
            uint32_t qv = 0u;
            bool open = false;
            if (on) {
                const uint32_t *Vc = Vw + g * VCOLS + phase + 3u * (xo + CX) + ch;
                unsigned long long acc = (unsigned long long)Vc[0] * wts.hx[0];
#pragma unroll
                for (int j = 1; j <= CX; ++j) acc += (unsigned long long)(Vc[-3 * j] + Vc[3 * j]) * wts.hx[j];
                const uint32_t r32 = (uint32_t)(acc >> (wts.s - 32u));             // top 32 bits of the fractional part
                qv = (uint32_t)(acc >> wts.s) + (r32 >> 31);
                open = r32 - 0x80000000u + wts.eps32 <= 2u * wts.eps32;            // within eps of a half-integer: exact sum decides
            }
            for (unsigned todo = __ballot_sync(0xffffffffu, open); todo; todo &= todo - 1u) {
                const int src = __ffs(todo) - 1;
                const uint32_t g_s = __shfl_sync(0xffffffffu, g, src), x_s = __shfl_sync(0xffffffffu, xo, src), c_s = __shfl_sync(0xffffffffu, ch, src);
                const uint32_t qx = blur_exact_warp<PARTS, CX, CY>(left, pitch, H, W, y0 + (int)g_s, xw + (int)x_s, (int)c_s, a.wq, a.wshift);
                if (lane == src) qv = qx;
            }
            if (!on) continue;
            a.plane[(((size_t)t * H + y0 + g) * W + xw + xo) * 3 + ch] = (uint8_t)qv;
        }
        __syncwarp();                                 // the next unit rewrites the hole table and the V rows
    }
}

// plane -> SBS frame for the listed holes right of the strip, then result_img[:, 0:strip] = img[:, 0:strip]
// (PredictAndGenerate.py:196).  A warp takes 32 list entries at a time: lane l fetches entry l's word index, mask
// and strip (one round of dependent loads for 32 entries), then the warp walks the entries, lane = pixel.
__global__ void __launch_bounds__(256, 8) k_blur_commit(BlurArgs a, int do_commit) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    pdl_wait();                                  // every blurred value is in the plane
    if ((do_commit & 1) && (do_commit & 4)) {
        // band-driven (do_commit bit 2: k_blur_band ran, band_list is valid): a warp per band column - the eight mask words in
        // one round of loads, then lane = pixel for each row; eight independent load chains per lane instead of four
        const uint32_t count = *a.band_count;
        for (uint32_t u = blockIdx.x * nwarps + warp; u < count; u += gridDim.x * nwarps) {
            const uint32_t code = a.band_list[u];
            const uint32_t q = code >> 8, w = code & 0xffu;
            const int t = (int)(((unsigned long long)q * a.magic_hb) >> 40), y0 = ((int)q - t * a.Hb) * kBandRows;
            uint32_t m = 0u;
            if (lane < kBandRows && y0 + lane < a.H) {
                const int strip = a.tabs[t].strip, xw = (int)w * 32;
                m = a.hole_mask[((size_t)t * a.H + y0 + lane) * a.Wwords + w];
                if (strip > xw) m = (strip - xw >= 32) ? 0u : (m & ~((1u << (strip - xw)) - 1u));
            }
            const size_t px0 = ((size_t)t * a.H + y0) * a.W + w * 32u + lane;
#pragma unroll
            for (int r0 = 0; r0 < kBandRows; r0 += 4) {          // four rows at a time: four load chains per lane, no spills
                uint32_t val[4];
                bool on[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    on[k] = (__shfl_sync(0xffffffffu, m, r0 + k) >> lane) & 1u;
                    val[k] = 0u;
                    if (on[k]) {
                        const uint8_t *src = a.plane + (px0 + (size_t)(r0 + k) * a.W) * 3;
                        val[k] = (uint32_t)src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16);
                    }
                }
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (on[k]) {
                        uint8_t *dst = a.sbs + (((size_t)t * a.H + y0 + r0 + k) * 2 * a.W + w * 32u + lane) * 3;
                        dst[0] = (uint8_t)val[k]; dst[1] = (uint8_t)(val[k] >> 8); dst[2] = (uint8_t)(val[k] >> 16);
                    }
            }
        }
    }
    const uint32_t count = ((do_commit & 1) && !(do_commit & 4)) ? *a.hole_count : 0u;   // do_commit: bit 0 = hole values, bit 1 = skip the strip (experiments)
    constexpr int E = 4;                                   // entries per warp step: four independent load chains in flight
    for (uint32_t e0 = (blockIdx.x * nwarps + warp) * E; e0 < count; e0 += gridDim.x * nwarps * E) {
        uint32_t gw = 0, m = 0, strip = 0;
        if (lane < E && e0 + lane < count) {
            gw = a.hole_list[e0 + lane];
            m = a.hole_mask[(size_t)(gw >> 8) * a.Wwords + (gw & 0xffu)];
            strip = (uint32_t)a.tabs[(int)(((unsigned long long)(gw >> 8) * a.magic_h) >> 40)].strip;
        }
        uint32_t val[E];
        uint8_t *dst[E];
        bool on[E];
#pragma unroll
        for (int k = 0; k < E; ++k) {
            const uint32_t g = __shfl_sync(0xffffffffu, gw, k), mk = __shfl_sync(0xffffffffu, m, k), st = __shfl_sync(0xffffffffu, strip, k);
            const uint32_t row = g >> 8, w = g & 0xffu;
            const uint32_t x = w * 32u + lane;
            on[k] = ((mk >> lane) & 1u) && x >= st;
            dst[k] = a.sbs + ((size_t)row * 2 * a.W + x) * 3;
            val[k] = 0;
            if (on[k]) {
                const uint8_t *src = a.plane + ((size_t)row * a.W + x) * 3;
                val[k] = (uint32_t)src[0] | ((uint32_t)src[1] << 8) | ((uint32_t)src[2] << 16);
            }
        }
#pragma unroll
        for (int k = 0; k < E; ++k)
            if (on[k]) { dst[k][0] = (uint8_t)val[k]; dst[k][1] = (uint8_t)(val[k] >> 8); dst[k][2] = (uint8_t)(val[k] >> 16); }
    }
    // strip restore: 4 threads per image row, 64 bytes each per round (16-byte copies when the rows are aligned)
    const long long nthreads = (long long)gridDim.x * blockDim.x;
    const long long rows = (do_commit & 2) ? 0 : (long long)a.B * a.H;
    const bool vec = (a.W % 16 == 0) && ((reinterpret_cast<uintptr_t>(a.frames) | reinterpret_cast<uintptr_t>(a.sbs)) % 16 == 0);
    for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < rows * 4; idx += nthreads) {
        const long long r = idx >> 2;
        const int part = (int)(idx & 3);
        const int nbytes = a.tabs[(int)(((unsigned long long)r * a.magic_h) >> 40)].strip * 3;
        const uint8_t *src = a.frames + r * (size_t)a.W * 3;
        uint8_t *dst = a.sbs + r * (size_t)a.W * 6;
        for (int c0 = part * 64; c0 < nbytes; c0 += 256) {
            const int c1 = min(c0 + 64, nbytes);
            int c = c0;
            if (vec)
                for (; c + 16 <= c1; c += 16) *reinterpret_cast<uint4 *>(dst + c) = __ldg(reinterpret_cast<const uint4 *>(src + c));
            for (; c < c1; ++c) dst[c] = src[c];
        }
    }
}

}  // namespace vrsbs
