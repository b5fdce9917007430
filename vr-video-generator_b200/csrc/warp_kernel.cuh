// Stage 3a (general route: any width / alignment / layer count; the fast route is warp_fused.cuh):
// the layered painter's-order warp as a per-row scatter/gather in shared memory.
//
// Replaces gpu_roll_with_offset + the layer loop + hole fill + SBS pack of left_side_sbs
// (PredictAndGenerate.py:150-155,169-190,197).  Restatement being implemented (SURVEY.md section 0,
// checked bit-for-bit against the reference by the oracle tests):
//     every source pixel (y,xs) that belongs to layer k lands at xd = (xs + off_k) mod W;
//     the highest k wins; unhit destinations are holes and take img[y, (xd - off_f) mod W].
//
// One persistent CTA walks image rows (rows are independent).  Per row:
//   1. TMA bulk copies (cp.async.bulk, mbarrier completion) stage the RGB row and the fp16 depth
//      row in shared memory, one row ahead of the compute (double buffered);
//   2. scatter: each thread finds the layers of its source pixels (affine guess + exact walk over
//      the fp16-exact bounds) and writes key = k+1 to keys[xd].  Colliding writes are resolved
//      deterministically: MODE 1 = plain store, barrier, re-read, and atomicMax only where the
//      stored key is smaller than mine (collisions are rare: occlusion edges only);
//      MODE 2 = atomicMax for every write;
//   3. gather: keys[xd] -> source x (via the offset table; key 0 = hole -> fill offset) -> 3 bytes
//      from the staged row -> packed 32-bit words into the staged output row; hole bits are
//      ballot-packed and written as the bitmask the blur kernel consumes;
//   4. two TMA bulk stores write the row of the SBS frame: [warped view | input row], the right half
//      straight from the staged input (never touched by a thread).
// Lanes own interleaved pixels (x = 32*segment + lane) so that key traffic is bank-conflict free.
#pragma once
#include <type_traits>

#include "common.cuh"

namespace vrsbs {

struct WarpArgs {
    const uint8_t *frames;     // [B,H,W,3]
    const void *depth;         // [B,H,W] smoothed, fp16 (F32 = false) or fp32
    uint8_t *sbs;              // [B,H,2W,3]
    FrameTab *tabs;            // [B]
    const float2 *bounds;      // [B][Lcap]
    const int *offm;           // [B][Lcap+1]
    uint32_t *hole_mask;       // [B][H][Wwords]
    uint32_t *hole_list;       // (global row << 8 | word) of every mask word that has a hole (blur work list, any order)
    uint32_t *hole_count;      // pre-zeroed
    int B, H, W, Lcap, Wwords;
};

constexpr int kMaxSeg = 8;     // 32-pixel segments per warp per row (bounds register state)

struct WarpSmem {
    size_t img[2], dep[2], out[2], keys, bounds, offm, bars, total;
};
__host__ __device__ inline WarpSmem warp_smem_layout(int W, int Lcap, int depth_bytes = 2) {
    WarpSmem s;
    size_t o = 0;
    for (int i = 0; i < 2; ++i) { s.img[i] = o; o += align_up((size_t)W * 3 + 16, 128); }
    for (int i = 0; i < 2; ++i) { s.dep[i] = o; o += align_up((size_t)W * depth_bytes, 128); }
    for (int i = 0; i < 2; ++i) { s.out[i] = o; o += align_up((size_t)W * 3 + 16, 128); }
    s.keys = o;   o += align_up((size_t)W * 4, 128);
    s.bounds = o; o += align_up((size_t)Lcap * 8, 16);
    s.offm = o;   o += align_up((size_t)(Lcap + 1) * 4, 16);
    s.bars = o;   o += 16;
    s.total = o;
    return s;
}

template <int MODE, bool TMA, int NT, bool F32 = false>
__global__ void __launch_bounds__(NT) k_warp_rows(WarpArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    using DT = typename std::conditional<F32, float, __half>::type;
    const WarpSmem lay = warp_smem_layout(a.W, a.Lcap, (int)sizeof(DT));
    const DT *a_depth = static_cast<const DT *>(a.depth);
    // stage s of a double-buffered array lives at base + s * stride (no pointer arrays: keeps them out of local memory)
    const size_t img_stride = lay.img[1] - lay.img[0], dep_stride = lay.dep[1] - lay.dep[0], out_stride = lay.out[1] - lay.out[0];
    auto img_at = [&](int s) { return smem + lay.img[0] + s * img_stride; };
    auto dep_at = [&](int s) { return reinterpret_cast<DT *>(smem + lay.dep[0] + s * dep_stride); };
    auto out_at = [&](int s) { return smem + lay.out[0] + s * out_stride; };
    uint32_t *keys = reinterpret_cast<uint32_t *>(smem + lay.keys);
    float2 *bnd = reinterpret_cast<float2 *>(smem + lay.bounds);
    int *offm = reinterpret_cast<int *>(smem + lay.offm);
    uint64_t *full = reinterpret_cast<uint64_t *>(smem + lay.bars);

    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = a.W, H = a.H;
    const long long total_rows = (long long)a.B * H;
    const uint32_t img_bytes = (uint32_t)W * 3, dep_bytes = (uint32_t)W * (uint32_t)sizeof(DT);
    const int nseg = (W + 31) >> 5;

    if (TMA && tid == 0) {
        mbar_init(&full[0], 1);
        mbar_init(&full[1], 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto issue_load = [&](long long row, int s) {           // thread 0 only
        mbar_expect_tx(&full[s], img_bytes + dep_bytes);
        bulk_g2s(img_at(s), a.frames + (size_t)row * img_bytes, img_bytes, &full[s]);
        bulk_g2s(dep_at(s), a_depth + (size_t)row * W, dep_bytes, &full[s]);
    };

    long long row = blockIdx.x;
    if (TMA && tid == 0 && row < total_rows) issue_load(row, 0);

    int cur_frame = -1, L = 0, generic = 0;
    float gscale = 0.f, gbias = 0.f;
    unsigned hole_acc = 0;                                   // per-warp (lane 0) hole count of cur_frame

    for (int it = 0; row < total_rows; row += gridDim.x, ++it) {
        const int s = it & 1;
        const int frame = (int)(row / H);
        const long long next = row + gridDim.x;

        if (TMA) {
            if (tid == 0 && next < total_rows) {
                bulk_wait_read0();                           // stage s^1 buffers are no longer being read by stores
                issue_load(next, s ^ 1);
            }
        } else {
            // generic path (W*3 or pointers not 16-byte aligned): cooperative byte copies
            const uint8_t *gi = a.frames + (size_t)row * img_bytes;
            const DT *gd = a_depth + (size_t)row * W;
            uint8_t *si = img_at(s);
            DT *sd = dep_at(s);
            for (int i = tid; i < (int)img_bytes; i += NT) si[i] = gi[i];
            for (int i = tid; i < W; i += NT) sd[i] = gd[i];
        }

        if (frame != cur_frame) {                            // (re)load this frame's tables
            if (lane == 0 && hole_acc && cur_frame >= 0) atomicAdd(&a.tabs[cur_frame].holes, (unsigned long long)hole_acc);
            hole_acc = 0;
            cur_frame = frame;
            const FrameTab *t = a.tabs + frame;
            L = t->layers;
            generic = (t->status & VRSBS_FRAME_GENERIC) != 0;
            gscale = t->guess_scale;
            gbias = t->guess_bias;
            const float2 *gb = a.bounds + (size_t)frame * a.Lcap;
            const int *go = a.offm + (size_t)frame * (a.Lcap + 1);
            for (int i = tid; i < L; i += NT) bnd[i] = gb[i];
            for (int i = tid; i <= L; i += NT) offm[i] = go[i];
        }
        {   // clear keys (previous row's gather finished at the trailing barrier)
            uint4 z = make_uint4(0, 0, 0, 0);
            uint4 *k4 = reinterpret_cast<uint4 *>(keys);
            for (int i = tid; i < (W + 3) / 4; i += NT) k4[i] = z;
        }
        if (TMA) mbar_wait(&full[s], (it >> 1) & 1);
        __syncthreads();

        const DT *dep = dep_at(s);
        uint32_t memb[kMaxSeg];                              // (ktop+1) | count<<16 per owned source pixel

        // ---- scatter --------------------------------------------------------------------------------
#pragma unroll
        for (int i = 0; i < kMaxSeg; ++i) {
            const int x = ((warp + NW * i) << 5) + lane;
            memb[i] = 0;
            if (x < W) {
                float d;
                if constexpr (F32) d = dep[x]; else d = h2f(dep[x]);
                if (!generic) {
                    int k = min(max((int)floorf(fmaf(d, gscale, gbias)), 0), L - 1);
                    while (k + 1 < L && bnd[k + 1].x <= d) ++k;
                    while (k >= 0 && !(bnd[k].x <= d)) --k;
                    const int ktop = k;
                    int cnt = 0;
                    for (; k >= 0 && d < bnd[k].y; --k, ++cnt) {
                        int xd = x + offm[k + 1];
                        xd -= (xd >= W) ? W : 0;
                        if (MODE == 1) keys[xd] = (uint32_t)(k + 1);
                        else atomicMax(&keys[xd], (uint32_t)(k + 1));
                    }
                    memb[i] = (uint32_t)(ktop + 1) | ((uint32_t)cnt << 16);
                } else {
                    for (int k = 0; k < L; ++k) {
                        const float2 b = bnd[k];
                        if (b.x <= d && d < b.y) {
                            int xd = x + offm[k + 1];
                            xd -= (xd >= W) ? W : 0;
                            atomicMax(&keys[xd], (uint32_t)(k + 1));
                        }
                    }
                }
            }
        }
        __syncthreads();
        // ---- verify: whoever lost a plain-store race re-asserts with atomicMax -------------------------
        if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < kMaxSeg; ++i) {
                const int x = ((warp + NW * i) << 5) + lane;
                int k = (int)(memb[i] & 0xffffu) - 1;
                for (int c = (int)(memb[i] >> 16); c > 0; --c, --k) {
                    int xd = x + offm[k + 1];
                    xd -= (xd >= W) ? W : 0;
                    if (keys[xd] < (uint32_t)(k + 1)) atomicMax(&keys[xd], (uint32_t)(k + 1));
                }
            }
            __syncthreads();
        }
        // ---- gather -----------------------------------------------------------------------------------
        uint8_t *img_row = img_at(s), *out_row = out_at(s);
        const uint32_t *img32 = reinterpret_cast<const uint32_t *>(img_row);
        uint32_t *out32 = reinterpret_cast<uint32_t *>(out_row);
        uint32_t *mask_row = a.hole_mask + (size_t)row * a.Wwords;
        for (int sg = warp; sg < nseg; sg += NW) {
            const int x = (sg << 5) + lane;
            const bool in = x < W;
            const uint32_t key = in ? keys[x] : 1u;
            int xs = x - offm[key];
            xs += (xs < 0) ? W : 0;
            const int ab = in ? 3 * xs : 0;
            const uint32_t w0 = img32[ab >> 2], w1 = img32[(ab >> 2) + 1];
            const uint32_t px = __funnelshift_r(w0, w1, (ab & 3) * 8) & 0x00ffffffu;
            const unsigned hm = __ballot_sync(0xffffffffu, in && key == 0u);
            if (lane == 0) {
                mask_row[sg] = hm;
                if (hm) {
                    hole_acc += __popc(hm);
                    a.hole_list[atomicAdd(a.hole_count, 1u)] = ((uint32_t)row << 8) | (uint32_t)sg;
                }
            }
            if (((sg + 1) << 5) <= W) {
                // 32 pixels = 24 words: lane 4q+r (r<3) emits word 3q+r from pixels 4q+r, 4q+r+1
                const uint32_t nx = __shfl_down_sync(0xffffffffu, px, 1);
                const int r = lane & 3;
                if (r != 3) out32[sg * 24 + 3 * (lane >> 2) + r] = (px >> (8 * r)) | (nx << (24 - 8 * r));
            } else if (in) {
                uint8_t *o = out_row + 3 * x;
                o[0] = (uint8_t)px; o[1] = (uint8_t)(px >> 8); o[2] = (uint8_t)(px >> 16);
            }
        }
        if (TMA) fence_async_smem();
        __syncthreads();
        uint8_t *go = a.sbs + (size_t)row * img_bytes * 2;
        if (TMA) {
            if (tid == 0) {
                bulk_s2g(go, out_row, img_bytes);
                bulk_s2g(go + img_bytes, img_row, img_bytes);
                bulk_commit();
            }
        } else {
            for (int i = tid; i < (int)img_bytes; i += NT) { go[i] = out_row[i]; go[img_bytes + i] = img_row[i]; }
            __syncthreads();
        }
    }
    if (lane == 0 && hole_acc && cur_frame >= 0) atomicAdd(&a.tabs[cur_frame].holes, (unsigned long long)hole_acc);
    if (TMA && tid == 0) bulk_wait0();
}

}  // namespace vrsbs
