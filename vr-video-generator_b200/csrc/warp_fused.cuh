// Stage 3a (fast route): temporal smoothing + layered painter's-order warp + hole fill + SBS pack in ONE pass.
//
// Replaces get_depth's smoothing (PredictAndGenerate.py:134-144), gpu_roll_with_offset, the layer loop,
// the hole fill and the SBS pack of left_side_sbs (PredictAndGenerate.py:150-155,169-190,197).
// Restatement implemented (pinned bit-for-bit by the oracle tests):
//     every source pixel (y,xs) of layer k lands at xd = (xs + off_k) mod W; the highest k wins;
//     unhit destinations are holes and take img[y, (xd - off_f) mod W].
//
// Design (all of it chosen to cut instructions per pixel: the previous kernel was issue bound at 141
// thread-instructions per pixel, see profiles/r01a_ncu_summary.md):
//   * key = (k+1) << 24 | RGB.  One shared-memory atomicMax per (pixel, layer) moves the colour together
//     with the priority, so the painter's order is resolved by the atomic itself and the destination pass
//     is a plain read of the key row: no gather, no offset lookup.  (Two sources with the same k can never
//     meet: same k = same offset.)  Needs L <= 255; larger tables take the slow path below.
//   * layer membership = cell LUT + one packed fp16 compare.  cell = fp16 bits >> shift indexes a byte LUT
//     built (and validated cell by cell) by k_build_tables; its value e names the only two layers the
//     value can belong to, and LayerEnt[e] holds the two thresholds that decide (setp.lt.f16x2).
//   * time-major order: a CTA owns an image row y and walks the batch in time (flattened index
//     f = y*B + t, split evenly over the grid), so the two previous RAW depth rows needed by the
//     smoothing are already in its shared-memory ring and the smoothed depth never touches HBM.
//   * TMA bulk copies (mbarrier completion) stage rows two iterations ahead; the SBS row leaves through two
//     bulk stores: [warped view | input row], the right half straight from the staged input.
//   * scatter phase: lane <-> pixel interleaved (conflict-free atomics); destination phase: 4 pixels per
//     thread (128-bit key reads, keys zeroed in the same pass, 3 packed 32-bit stores).
#pragma once
#include <type_traits>

#include "common.cuh"

namespace vrsbs {

// shared-memory layout as 32-bit kernel parameters (constant bank: no per-row re-derivation); see fused_smem_layout
struct FusedLay {
    uint32_t img, img_stride, dep, dep_stride, out, keys, blob, blob_stride, mask, bars;
};

struct FusedArgs {
    const uint8_t *frames;     // [B,H,W,3]
    const __half *depth;       // [B,H,W] RAW depth (SMOOTH) or already smoothed depth (!SMOOTH)
    const __half *hist1;       // [H,W] raw t-1 of the previous batch (SMOOTH)
    const __half *hist2;       // [H,W] raw t-2
    uint8_t *sbs;              // [B,H,2W,3]
    const uint8_t *blobs;      // [B][blob_bytes]
    FrameTab *tabs;            // [B]
    const float2 *bounds;      // [B][Lcap]    (slow path)
    const int *offm;           // [B][Lcap+1]  (slow path)
    uint32_t *hole_mask;       // [B][H][Wwords]
    uint32_t *hole_list;       // (global row << 8 | word) of every mask word that has a hole (any order)
    uint32_t *hole_count;      // pre-zeroed
    uint32_t *band_map;        // k_warp_ws, optional: [B*Hb][band_groups] bit per band column (8 rows x one mask word), set when a row of it has
                               //   a hole; replaces hole_list / hole_count as the work index of k_blur_band and k_blur_commit
    int Hb, band_groups;       //   row bands per frame (8 rows each), 32-word groups per row
    int B, H, W, Lcap, Wwords;
    int first;                 // frame 0 of the batch is the first frame of the clip range
    int skip_right;            // 1: do not store the right half of the SBS row (the host pipeline's caller already holds it)
    uint32_t blob_bytes, ent_bytes;
    int key_pad;               // fast path: |signed offset| <= key_pad pixels (multiple of 32); only segments closer than
                               // that to a row end can wrap
    float w0, w1, w2;          // smoothing weights (fp32 narrowing of the python doubles)
    FusedLay lay;              // shared-memory layout, filled on the host from fused_smem_layout()
};

struct FusedSmem {
    size_t img, img_stride, dep, dep_stride, out, keys, blob, blob_stride, mask, bars, total;
};

constexpr int kImgSlots = 3, kDepSlots = 4, kBlobSlots = 2, kBars = 3;

__host__ __device__ inline FusedSmem fused_smem_layout(int W, uint32_t blob_b) {
    FusedSmem s;
    size_t o = 0;
    s.img = o;  s.img_stride = align_up((size_t)W * 3 + 16, 128);  o += kImgSlots * s.img_stride;
    s.dep = o;  s.dep_stride = align_up((size_t)W * 2, 128);       o += kDepSlots * s.dep_stride;
    s.out = o;  o += align_up((size_t)W * 3 + 16, 128);
    s.keys = o; o += align_up((size_t)W * 4, 128);
    s.blob = o; s.blob_stride = align_up((size_t)blob_b, 128);     o += kBlobSlots * s.blob_stride;
    s.mask = o; o += align_up((size_t)((W + 31) / 32) * 4, 16);
    s.bars = o; o += 8 * kBars;
    s.total = align_up(o, 16);
    return s;
}

// 3 bytes at byte offset 3*x of a 4-byte aligned row
__device__ __forceinline__ uint32_t fetch_rgb(const uint8_t *row, int x) {
    const int ab = 3 * x;
    const uint32_t *p = reinterpret_cast<const uint32_t *>(row) + (ab >> 2);
    return __funnelshift_r(p[0], p[1], (ab & 3) * 8) & 0x00ffffffu;
}

template <bool SMOOTH, int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) k_warp_fused(FusedArgs a) {
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sb = smem_u32(smem);                              // every shared access below is sb + constant-bank offset
    const uint32_t sa_keys = sb + a.lay.keys, sa_out = sb + a.lay.out, sa_mask = sb + a.lay.mask, sa_bars = sb + a.lay.bars;
    uint32_t *keys = reinterpret_cast<uint32_t *>(smem + a.lay.keys);   // generic pointer: slow path only

    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = a.W, H = a.H, B = a.B;
    const uint32_t img_bytes = (uint32_t)W * 3, dep_bytes = (uint32_t)W * 2, W4 = (uint32_t)W * 4;
    const int nseg = (W + 31) >> 5, nquad = W >> 2, Wwords = a.Wwords;

    const long long F = (long long)B * H;
    const long long f_lo = F * blockIdx.x / gridDim.x, f_hi = F * (blockIdx.x + 1) / gridDim.x;
    const int N = (int)(f_hi - f_lo);
    if (N <= 0) return;

    {   // zero the key row and the mask row once; later rows are re-zeroed by the destination pass
        for (int i = tid; i < nquad; i += NT) sts_zero128(sa_keys + 16u * i);
        for (int i = tid; i < Wwords; i += NT) sts_u32(sa_mask + 4u * i, 0u);
    }
    if (tid == 0) {
        for (int i = 0; i < kBars; ++i) mbar_init(reinterpret_cast<uint64_t *>(smem + a.lay.bars) + i, 1);
        fence_mbar_init();
    }
    __syncthreads();

    auto depth_row = [&](int y, int tt) -> const __half * {
        if (tt >= 0) return a.depth + ((size_t)tt * H + y) * W;
        if (a.first) return a.depth + (size_t)y * W;             // clip start: history = the first raw frame
        return (tt == -1 ? a.hist1 : a.hist2) + (size_t)y * W;
    };
    // thread 0: stage image row + current depth row + table blob of iteration k
    auto issue_main = [&](int k, int y, int t, int dslot, bool bnd) {
        const uint32_t bar = sa_bars + 8u * (uint32_t)(k % kBars);
        mbar_expect_tx_a(bar, img_bytes + dep_bytes + a.blob_bytes + ((SMOOTH && bnd) ? 2 * dep_bytes : 0u));
        bulk_g2s_a(sb + a.lay.img + (uint32_t)(k % kImgSlots) * a.lay.img_stride, a.frames + ((size_t)t * H + y) * img_bytes, img_bytes, bar);
        bulk_g2s_a(sb + a.lay.dep + (uint32_t)dslot * a.lay.dep_stride, depth_row(y, t), dep_bytes, bar);
        bulk_g2s_a(sb + a.lay.blob + (uint32_t)(k & 1) * a.lay.blob_stride, a.blobs + (size_t)t * a.blob_bytes, a.blob_bytes, bar);
    };
    auto issue_hist = [&](int k, int y, int t, int s_h1, int s_h2) {
        const uint32_t bar = sa_bars + 8u * (uint32_t)(k % kBars);
        bulk_g2s_a(sb + a.lay.dep + (uint32_t)s_h2 * a.lay.dep_stride, depth_row(y, t - 2), dep_bytes, bar);
        bulk_g2s_a(sb + a.lay.dep + (uint32_t)s_h1 * a.lay.dep_stride, depth_row(y, t - 1), dep_bytes, bar);
    };

    // (y,t) of iterations n, n+1, n+2 and the depth-ring slots of n and n+1 (uniform state)
    int y0 = (int)(f_lo / B), t0 = (int)(f_lo - (long long)y0 * B);
    auto next_yt = [&](int &y, int &t) { if (++t == B) { t = 0; ++y; } };
    int y1 = y0, t1 = t0; next_yt(y1, t1);
    int y2 = y1, t2 = t1; next_yt(y2, t2);
    int c0 = 0, h1_0 = 2, h2_0 = 1, c1 = 3;                      // slots: cur(0), hist(0), cur(1)
    if (tid == 0) {
        issue_main(0, y0, t0, c0, true);
        if (SMOOTH) issue_hist(0, y0, t0, h1_0, h2_0);
        if (N > 1) issue_main(1, y1, t1, c1, t1 == 0);
    }

    const int wofs = (3 * lane) >> 2, wsh = ((3 * lane) & 3) * 8;       // lane-constant part of the pixel fetch

    // per-warp scatter schedule (row independent): this warp owns segments warp, warp + NW, ... plus, for the LAST
    // nseg % NW warps, one left-over segment; a batch needs the wrap fix-up only if one of its segments lies within
    // key_pad pixels of a row end
    const int sch_full = nseg / NW, sch_rem = nseg - sch_full * NW;
    const bool sch_extra = NW - 1 - warp < sch_rem;
    const int sch_nw = sch_full + (sch_extra ? 1 : 0);
    const int sch_w_last = sch_extra ? NW - 1 - warp : warp;
    const int sch_eseg = a.key_pad >> 5;                          // segments [0, eseg) and [nseg - eseg, nseg) are edge segments
    auto batch_edge = [&](int r0, int cnt, int wl) {
        const int first = r0 * NW + warp, last = (r0 + cnt - 1) * NW + wl;
        return first < sch_eseg || last >= nseg - sch_eseg;
    };
    const int sch_r_last = sch_nw > 4 ? 4 * ((sch_nw - 1) / 4) : 0;
    const bool sch_wrap_first = batch_edge(0, 4, warp);
    const bool sch_wrap_last = batch_edge(sch_r_last, sch_nw - sch_r_last, sch_w_last);

    int i3 = 0;                 // n % 3: image slot and mbarrier of iteration n
    uint32_t par = 0;           // (n / 3) & 1: phase parity of that mbarrier
    for (int n = 0; n < N; ++n) {
        // slots of iteration n+1's history and of cur(n+2)
        const bool bnd1 = SMOOTH && (t1 == 0);
        int h1_1, h2_1, c2;
        if (!SMOOTH) { h1_1 = c0; h2_1 = h1_0; c2 = (c1 + 1) & 3; }
        else if (bnd1) { h2_1 = (c1 + 1) & 3; h1_1 = (c1 + 2) & 3; c2 = (c1 + 3) & 3; }
        else      { h1_1 = c0; h2_1 = h1_0; c2 = c1 ^ c0 ^ h1_0; }

        const uint32_t sa_imgrow = sb + a.lay.img + (uint32_t)i3 * a.lay.img_stride;
        const uint32_t sa_blob = sb + a.lay.blob + (uint32_t)(n & 1) * a.lay.blob_stride;
        const uint32_t sa_cur = sb + a.lay.dep + (uint32_t)c0 * a.lay.dep_stride;
        const uint32_t sa_p1 = sb + a.lay.dep + (uint32_t)h1_0 * a.lay.dep_stride, sa_p2 = sb + a.lay.dep + (uint32_t)h2_0 * a.lay.dep_stride;

        mbar_wait_a(sa_bars + 8u * (uint32_t)i3, par);

        const uint4 hdrw = lds_u128(sa_blob);                       // BlobHdr: fill_off, shift, ncells, flags
        const bool fast = hdrw.w & 1u;
        const int fill = (int)hdrw.x;
        const uint32_t sa_img = sa_imgrow + 4u * (uint32_t)wofs;

        // smoothed depth of pixel x (both halves of the result hold it), from raw halves
        auto smooth_px = [&](uint32_t c, uint32_t p1, uint32_t p2) -> uint32_t {
            if (SMOOTH) {
                const float m0 = __fmul_rn(__half2float(__ushort_as_half((unsigned short)c)), a.w0);
                const float m1 = __fmul_rn(__half2float(__ushort_as_half((unsigned short)p1)), a.w1);
                const float m2 = __fmul_rn(__half2float(__ushort_as_half((unsigned short)p2)), a.w2);
                __half2 d = __hadd2(__floats2half2_rn(m0, m0), __floats2half2_rn(m1, m1));
                d = __hadd2(d, __floats2half2_rn(m2, m2));
                return *reinterpret_cast<uint32_t *>(&d);
            } else {
                return c | (c << 16);
            }
        };

        // ---- scatter ------------------------------------------------------------------------------------
        if (fast) {
            const uint32_t sa_ent = sa_blob + 16u, sa_lut = sa_ent + a.ent_bytes;
            // LUT address = sa_lut + min(bits >> shift, ncells); dd holds the fp16 bits twice, so dd >> (16 + shift)
            // = umulhi(dd, 2^(16 - shift)) folds the shift and the base add into one IMAD.HI
            const uint32_t lut_mul = 1u << (16u - hdrw.y), lut_last = sa_lut + hdrw.z;
            const uint32_t sa_keys_lo = sa_keys - 4u * (uint32_t)a.key_pad;
            // U segments of 32 pixels per step, all loads of a step issued before their uses
            // w_last: warp index used for the LAST segment of the batch (differs from `warp` only for the left-over segment)
            auto batch = [&](auto Uc, auto Wc, int round0, bool tail, int w_last) {
                constexpr int U = decltype(Uc)::value;
                constexpr bool WRAP = decltype(Wc)::value;        // some segment of the batch is within key_pad of a row end
                uint32_t c[U], p1[U], p2[U], w0[U], w1[U], x4[U], ok[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int seg = (round0 + u) * NW + (u == U - 1 ? w_last : warp);
                    const int x = (seg << 5) + lane;
                    ok[u] = (!tail || x < W) ? 1u : 0u;
                    const int xc = tail ? min(x, W - 1) : x;
                    x4[u] = (uint32_t)xc * 4u;
                    c[u] = lds_u16(sa_cur + 2u * xc);
                    if (SMOOTH) { p1[u] = lds_u16(sa_p1 + 2u * xc); p2[u] = lds_u16(sa_p2 + 2u * xc); }
                    else { p1[u] = p2[u] = 0u; }
                    const uint32_t ia = sa_img + 96u * (uint32_t)(tail ? min(seg, nseg - 1) : seg);
                    w0[u] = lds_u32(ia);
                    w1[u] = lds_u32(ia + 4u);
                }
                uint32_t dd[U], e[U];
                if (SMOOTH) {
                    // d = rn16(rn16(rn16(c*w0) + rn16(p1*w1)) + rn16(p2*w2)); the fp32->fp16 roundings are packed two
                    // per F2FP (c/p1 products of one pixel, p2 products of two pixels), the adds are fp16 adds whose
                    // operand selectors replicate the result into both halves
                    __half2 d01[U];
                    float m2[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const float m0 = __fmul_rn(__half2float(__ushort_as_half((unsigned short)c[u])), a.w0);
                        const float m1 = __fmul_rn(__half2float(__ushort_as_half((unsigned short)p1[u])), a.w1);
                        m2[u] = __fmul_rn(__half2float(__ushort_as_half((unsigned short)p2[u])), a.w2);
                        const __half2 v = __floats2half2_rn(m0, m1);
                        d01[u] = __hadd2(__low2half2(v), __high2half2(v));
                    }
#pragma unroll
                    for (int u = 0; u < U; u += 2) {
                        const __half2 v = __floats2half2_rn(m2[u], m2[u + 1 < U ? u + 1 : u]);
                        __half2 r0 = __hadd2(d01[u], __low2half2(v));
                        dd[u] = *reinterpret_cast<uint32_t *>(&r0);
                        if (u + 1 < U) {
                            __half2 r1 = __hadd2(d01[u + 1], __high2half2(v));
                            dd[u + 1] = *reinterpret_cast<uint32_t *>(&r1);
                        }
                    }
                } else {
#pragma unroll
                    for (int u = 0; u < U; ++u) dd[u] = c[u] | (c[u] << 16);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) e[u] = lds_u8(min(__umulhi(dd[u], lut_mul) + sa_lut, lut_last));
                uint2 en[U];
#pragma unroll
                for (int u = 0; u < U; ++u) en[u] = lds_u64(sa_ent + 8u * e[u]);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const uint32_t px = __funnelshift_r(w0[u], w1[u], wsh) & 0x00ffffffu;
                    const uint32_t key0 = px | (e[u] << 24);
                    // offsets are signed (shortest way round the row) and biased by key_pad: interior segments cannot wrap
                    const uint32_t kb = sa_keys_lo + x4[u];
                    uint32_t a0 = kb + (en[u].y & 0xffffu), a1 = kb + (en[u].y >> 16);
                    if (WRAP) {
                        // relative byte offsets in [-4 pad, 4 W + 4 pad).  Only the first segment of a batch can lie at the
                        // left end of the row (negative = huge unsigned: min(t, t + 4W) brings it back) and only the last
                        // two at the right end (min(t, t - 4W)); key_pad <= 32 * NW guarantees that (host check).
                        uint32_t t0 = a0 - sa_keys, t1 = a1 - sa_keys;
                        if (u == 0) { t0 = min(t0, t0 + W4); t1 = min(t1, t1 + W4); }
                        if (u >= U - 2) { t0 = min(t0, t0 - W4); t1 = min(t1, t1 - W4); }
                        a0 = sa_keys + t0;
                        a1 = sa_keys + t1;
                    }
                    // paint layer e-1 iff d < hi(e-1); paint layer e iff !(d < lo(e)).  Non-members still issue the
                    // atomic, with key 0 (a no-op for max): cheaper than the branch ptxas wraps a predicated ATOMS in.
                    uint32_t k0, k1;
                    asm("{\n\t.reg .pred p, q;\n\t"
                        "setp.lt.f16x2 p|q, %2, %3;\n\t"
                        "selp.u32 %0, %4, 0, p;\n\t"
                        "selp.u32 %1, 0, %5, q;\n\t}"
                        : "=r"(k0), "=r"(k1) : "r"(dd[u]), "r"(en[u].x), "r"(key0), "r"(key0 + 0x01000000u));
                    if (tail) { k0 = ok[u] ? k0 : 0u; k1 = ok[u] ? k1 : 0u; }
                    asm volatile("red.shared.max.u32 [%0], %1;" :: "r"(a0), "r"(k0) : "memory");
                    asm volatile("red.shared.max.u32 [%0], %1;" :: "r"(a1), "r"(k1) : "memory");
                }
            };
            if ((W & 31) == 0) {
                // whole segments only: this warp owns segments warp, warp + NW, ... 4 at a time.  The nseg % NW left-over
                // segments go to the LAST warps: warp 0 also issues the TMA traffic and warps 0.. flush the mask row.
                using T = std::true_type; using F = std::false_type;
                int r = 0;
                for (; r + 4 < sch_nw; r += 4) {
                    const bool ew = r == 0 ? sch_wrap_first : batch_edge(r, 4, warp);
                    if (ew) batch(std::integral_constant<int, 4>{}, T{}, r, false, warp);
                    else batch(std::integral_constant<int, 4>{}, F{}, r, false, warp);
                }
                switch (sch_nw - r) {                         // warp-uniform; the final batch ends with the left-over segment
                    case 4: if (sch_wrap_last) batch(std::integral_constant<int, 4>{}, T{}, r, false, sch_w_last); else batch(std::integral_constant<int, 4>{}, F{}, r, false, sch_w_last); break;
                    case 3: if (sch_wrap_last) batch(std::integral_constant<int, 3>{}, T{}, r, false, sch_w_last); else batch(std::integral_constant<int, 3>{}, F{}, r, false, sch_w_last); break;
                    case 2: batch(std::integral_constant<int, 2>{}, T{}, r, false, sch_w_last); break;
                    case 1: batch(std::integral_constant<int, 1>{}, T{}, r, false, sch_w_last); break;
                    default: break;
                }
            } else {
                for (int r = 0; r * NW + warp < nseg; ++r) batch(std::integral_constant<int, 1>{}, std::true_type{}, r, true, warp);
            }
        } else {
            // slow path (L > 255, non-monotone bounds, LUT too coarse): brute-force membership, layer-only keys
            const int L = (int)(hdrw.w >> 8);
            const int frame = t0;
            const float2 *gb = a.bounds + (size_t)frame * a.Lcap;
            const int *go = a.offm + (size_t)frame * (a.Lcap + 1);
            for (int seg = warp; seg < nseg; seg += NW) {
                const int x = (seg << 5) + lane;
                if (x < W) {
                    const uint32_t du = smooth_px(lds_u16(sa_cur + 2u * x), SMOOTH ? lds_u16(sa_p1 + 2u * x) : 0u, SMOOTH ? lds_u16(sa_p2 + 2u * x) : 0u);
                    const float d = __half2float(__ushort_as_half((unsigned short)(du & 0xffffu)));
                    for (int k = 0; k < L; ++k) {
                        const float2 b = __ldg(gb + k);
                        if (b.x <= d && d < b.y) {
                            int xd = x + __ldg(go + k + 1);
                            xd -= (xd >= W) ? W : 0;
                            atomicMax(&keys[xd], (uint32_t)(k + 1));
                        }
                    }
                }
            }
        }
        if (tid == 0) bulk_wait_read0();        // the previous row's bulk stores have finished reading out_row / img slots
        __syncthreads();

        // ---- destination pass: 4 pixels per thread --------------------------------------------------------
        {
            auto fetch = [&](int xs) {                          // 3 bytes at byte offset 3*xs of the staged image row
                const uint32_t ab = 3u * (uint32_t)xs, wa = sa_imgrow + (ab & ~3u);
                return __funnelshift_r(lds_u32(wa), lds_u32(wa + 4u), (ab & 3u) * 8u) & 0x00ffffffu;
            };
            for (int j = tid; j < nquad; j += NT) {
                const uint4 k = lds_u128(sa_keys + 16u * j);
                sts_zero128(sa_keys + 16u * j);
                uint32_t kk[4] = {k.x, k.y, k.z, k.w};
                if (fast) {
                    const uint32_t mn = min(min(k.x, k.y), min(k.z, k.w));
                    if (mn < 0x01000000u) {
                        uint32_t hm = 0;
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            if (kk[i] < 0x01000000u) {
                                int xs = 4 * j + i - fill;
                                xs += (xs < 0) ? W : 0;
                                kk[i] = fetch(xs);
                                hm |= 1u << i;
                            }
                        }
                        reds_or(sa_mask + 4u * (uint32_t)(j >> 3), hm << ((j & 7) * 4));
                    }
                } else {
                    const int *go = a.offm + (size_t)t0 * (a.Lcap + 1);
                    uint32_t hm = 0;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        int xs = 4 * j + i - (kk[i] ? __ldg(go + kk[i]) : fill);
                        xs += (xs < 0) ? W : 0;
                        hm |= (kk[i] ? 0u : 1u) << i;
                        kk[i] = fetch(xs);
                    }
                    if (hm) reds_or(sa_mask + 4u * (uint32_t)(j >> 3), hm << ((j & 7) * 4));
                }
                const uint32_t oa = sa_out + 12u * j;
                sts_u32(oa, __byte_perm(kk[0], kk[1], 0x4210));
                sts_u32(oa + 4u, __byte_perm(kk[1], kk[2], 0x5421));
                sts_u32(oa + 8u, __byte_perm(kk[2], kk[3], 0x6542));
            }
        }
        fence_async_smem();
        __syncthreads();

        const uint32_t row = (uint32_t)t0 * (uint32_t)H + (uint32_t)y0;        // global row index of this iteration (< 2^24)
        if (tid == 0) {
            uint8_t *go = a.sbs + (size_t)row * img_bytes * 2;
            bulk_s2g_a(go, sa_out, img_bytes);
            if (!a.skip_right) bulk_s2g_a(go + img_bytes, sa_imgrow, img_bytes);
            bulk_commit();
            if (n + 2 < N) issue_main(n + 2, y2, t2, c2, t2 == 0);
            if (SMOOTH && bnd1 && n + 1 < N) issue_hist(n + 1, y1, t1, h1_1, h2_1);
        }
        // ---- hole mask row -> global bitmask + work list for the blur -----------------------------------------
        if (warp < ((Wwords + 31) >> 5)) {
            const int w = tid;
            uint32_t v = 0;
            if (w < Wwords) {
                v = lds_u32(sa_mask + 4u * w);
                sts_u32(sa_mask + 4u * w, 0u);
                a.hole_mask[(size_t)row * Wwords + w] = v;
            }
            const unsigned nz = __ballot_sync(0xffffffffu, v != 0u);
            if (nz) {
                const unsigned holes = __reduce_add_sync(0xffffffffu, (unsigned)__popc(v));
                uint32_t base = 0;
                if (lane == 0) {
                    base = atomicAdd(a.hole_count, (uint32_t)__popc(nz));
                    atomicAdd(&a.tabs[t0].holes, (unsigned long long)holes);
                }
                base = __shfl_sync(0xffffffffu, base, 0);
                if (v) a.hole_list[base + __popc(nz & ((1u << lane) - 1u))] = (row << 8) | (uint32_t)w;
            }
        }
        // advance the uniform state
        if (++i3 == 3) { i3 = 0; par ^= 1u; }
        y0 = y1; t0 = t1; y1 = y2; t1 = t2; next_yt(y2, t2);
        h2_0 = h2_1; h1_0 = h1_1; c0 = c1; c1 = c2;
    }
    if (tid == 0) bulk_wait0();
}

}  // namespace vrsbs
