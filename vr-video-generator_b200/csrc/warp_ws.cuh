// Stage 3a, warp-specialised variant of k_warp_fused<false> (same arithmetic, same tables, same output).
//
// k_warp_fused runs scatter -> CTA barrier -> destination pass -> CTA barrier per row; ~40 % of its stall samples
// are warps parked at those two barriers (measured mid-round, see profiles/README.md).  Here the CTA is split into NS scatter warps and ND
// destination warps that meet only through split arrive / sync barriers around a DOUBLE-BUFFERED key row:
//
//   scatter warps, row n:  wait in_full[n%3] (TMA data, mbarrier) and k_empty[n&1]  ->  atomicMax keys into keys[n&1]
//                          ->  arrive k_full[n&1]                     (and go straight on to row n+1)
//   destination warps:     wait in_full[n%3] and k_full[n&1]  ->  read + re-zero keys[n&1], fill holes, pack the
//                          row  ->  bulk stores  ->  arrive k_empty[n&1]  ->  mask flush (hole mask row, band-column bitmap or
//                          per-word list for the blur), TMA loads of row n+2
//
// k_full / k_empty are hardware named barriers (bar.arrive by the producer group, bar.sync by the consumer group):
// a parked consumer issues nothing, whereas mbarrier try_wait loops spent 23 % of the kernel's issue slots on
// polling (profiles/README.md).
//
// so a slow warp delays only its own group, and the scatter of row n+1 overlaps the destination pass of row n.
// Smoothed depth in (the depth pass materialises it); frames whose tables did not validate take the same
// brute-force membership path as in k_warp_fused.  Reference lines replaced: PredictAndGenerate.py:150-155,
// 169-190,197.
#pragma once
#include <type_traits>

#include "common.cuh"
#include "warp_fused.cuh"

namespace vrsbs {

struct WsLay {               // shared-memory layout of k_warp_ws as kernel parameters
    uint32_t img, img_stride, dep, dep_stride, out, keys, keys_stride, blob, blob_stride, mask, bars, total;
};
constexpr int kWsImgSlots = 3, kWsDepSlots = 3;
constexpr int kBarFull = 2, kBarEmpty = 4;      // named barriers 2,3: key row complete; 4,5: key row free again (1: destination group)

__host__ inline WsLay ws_smem_layout(int W, uint32_t blob_b, int depth_bytes = 2) {
    WsLay s{};
    size_t o = 0;
    // (rows only need the 16-byte alignment of the bulk copies: 128-byte strides cost 448 B at 1080p, the difference between three
    // and four CTAs per SM for offsets like fg .025 / bg -.015 whose LUT is twice the default's)
    s.img = (uint32_t)o;  s.img_stride = (uint32_t)align_up((size_t)W * 3 + 16, 16);   o += kWsImgSlots * s.img_stride;
    s.dep = (uint32_t)o;  s.dep_stride = (uint32_t)align_up((size_t)W * depth_bytes, 128);   o += kWsDepSlots * s.dep_stride;
    s.out = (uint32_t)o;  o += align_up((size_t)W * 3 + 16, 16);
    s.keys = (uint32_t)o; s.keys_stride = (uint32_t)align_up((size_t)W * 4, 128);      o += 2 * s.keys_stride;
    s.blob = (uint32_t)o; s.blob_stride = (uint32_t)align_up((size_t)blob_b, 128);     o += 2 * s.blob_stride;
    s.mask = (uint32_t)o; o += align_up((size_t)((W + 31) / 32) * 4, 16);
    s.bars = (uint32_t)o; o += 8 * 4;                    // in_full[3]
    s.total = (uint32_t)align_up(o, 16);
    return s;
}

struct WsArgs {
    FusedArgs f;             // pointers, sizes, tables (f.lay unused)
    WsLay lay;
};

__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// mbarrier wait with a suspend-time hint, used for the TMA arrival barriers (in_full).  The key-row hand-off used
// the same loop at first; a group parked for most of a row time then re-polled about every 50 cycles - hint or
// nanosleep made no difference - and 19-23 % of all executed warp-instructions were polls, which is why that hand-off
// now goes through named barriers.
#ifndef VRSBS_WS_WAIT_HINT_NS
#define VRSBS_WS_WAIT_HINT_NS 4000
#endif
__device__ __forceinline__ void mbar_wait_sleep(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n\t"
        "@!p bra WAIT_%=;\n\t}" ::"r"(bar), "r"(parity), "r"((uint32_t)VRSBS_WS_WAIT_HINT_NS) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// F32: the smoothed depth is fp32 (what torch >= 2.4's autocast hands the reference's warp), compared in fp32 against the
// double -> float narrowed bounds.  Same kernel; the differences are the depth row (4 bytes per pixel: three CTAs per SM
// instead of four), the cell of the LUT (fp32 bits >> shift, counted from the cell of `base`; everything below is one cell,
// every negative value the last one), 16-byte layer entries with the two bounds as floats, and two fp32 compares.
template <int NT, int NS, bool F32 = false>
__global__ void __launch_bounds__(NT, F32 ? 768 / NT : 1024 / NT) k_warp_ws(WsArgs wa) {
    const FusedArgs &a = wa.f;
    extern __shared__ __align__(128) uint8_t smem[];
    constexpr int NW = NT / 32, ND = NW - NS, NDT = ND * 32;
    const uint32_t sb = smem_u32(smem);
    const uint32_t sa_out = sb + wa.lay.out, sa_mask = sb + wa.lay.mask, sa_bars = sb + wa.lay.bars;
    const uint32_t bar_in = sa_bars;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int W = a.W, H = a.H, B = a.B;
    const uint32_t img_bytes = (uint32_t)W * 3, dep_bytes = (uint32_t)W * (F32 ? 4u : 2u), W4 = (uint32_t)W * 4;
    const int nseg = W >> 5, nquad = W >> 2, Wwords = a.Wwords;

    const long long F = (long long)B * H;
    const long long f_lo = F * blockIdx.x / gridDim.x, f_hi = F * (blockIdx.x + 1) / gridDim.x;
    const int N = (int)(f_hi - f_lo);
    pdl_launch_dependents();
    if (N <= 0) return;

    for (int i = tid; i < 2 * (int)(wa.lay.keys_stride / 16); i += NT) sts_zero128(sb + wa.lay.keys + 16u * i);
    for (int i = tid; i < Wwords; i += NT) sts_u32(sa_mask + 4u * i, 0u);
    if (tid == 0) {
        uint64_t *bars = reinterpret_cast<uint64_t *>(smem + wa.lay.bars);
        for (int i = 0; i < 3; ++i) mbar_init(bars + i, 1);              // in_full: one expect_tx arrival + bytes
        fence_mbar_init();
    }
    __syncthreads();
    pdl_wait();                                  // tables (and, through them, the smoothed depth) are complete; the set-up above overlapped their tail

    int y0 = (int)(f_lo / B), t0 = (int)(f_lo - (long long)y0 * B);
    auto next_yt = [&](int &y, int &t) { if (++t == B) { t = 0; ++y; } };

    if (warp < NS) {
        // =============================== scatter warps ===============================================
        const int wofs = (3 * lane) >> 2, wsh = ((3 * lane) & 3) * 8;
        const int sch_nw = nseg / NS + ((warp < nseg % NS) ? 1 : 0);           // segments warp, warp + NS, ...
        const int sch_eseg = a.key_pad >> 5;
        auto batch_edge = [&](int r0, int cnt) {
            const int first = r0 * NS + warp, last = (r0 + cnt - 1) * NS + warp;
            return first < sch_eseg || last >= nseg - sch_eseg;
        };
        int i3 = 0;
        uint32_t par3 = 0;
        for (int n = 0; n < N; ++n) {
            const uint32_t b = (uint32_t)n & 1u;
            const uint32_t sa_keys = sb + wa.lay.keys + b * wa.lay.keys_stride;
            const uint32_t sa_blob = sb + wa.lay.blob + b * wa.lay.blob_stride;
            const uint32_t sa_cur = sb + wa.lay.dep + (uint32_t)i3 * wa.lay.dep_stride;
            const uint32_t sa_img = sb + wa.lay.img + (uint32_t)i3 * wa.lay.img_stride + 4u * (uint32_t)wofs;
            mbar_wait_sleep(bar_in + 8u * (uint32_t)i3, par3);
            if (n >= 2) named_bar_sync(kBarEmpty + (int)b, NS * 32 + 32);   // keys[b] re-zeroed by the destination pass of row n-2
            const uint4 hdrw = lds_u128(sa_blob);
            if (hdrw.w & 1u) {
                const uint32_t sa_ent = sa_blob + 16u, sa_lut = sa_ent + a.ent_bytes;
                const uint32_t lut_sh = hdrw.y & 0xffu, lut_last = sa_lut + hdrw.z;   // cell = depth bits >> shift; the last cell = every negative value
                const uint32_t lut_base = hdrw.y >> 8;                            // F32: cells are counted from this one (everything below is cell 0)
                const uint32_t sa_lut0 = sa_lut - lut_base;
                const uint32_t sa_keys_lo = sa_keys - 4u * (uint32_t)a.key_pad;
                auto batch = [&](auto Uc, auto Wc, int round0) {
                    constexpr int U = decltype(Uc)::value;
                    constexpr bool WRAP = decltype(Wc)::value;
                    uint32_t c[U], w0[U], w1[U], x4[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int seg = (round0 + u) * NS + warp;
                        const int x = (seg << 5) + lane;
                        x4[u] = (uint32_t)x * 4u;
                        c[u] = F32 ? lds_u32(sa_cur + 4u * x) : lds_u16(sa_cur + 2u * x);
                        const uint32_t ia = sa_img + 96u * (uint32_t)seg;
                        w0[u] = lds_u32(ia);
                        w1[u] = lds_u32(ia + 4u);
                    }
                    uint32_t dd[U], e[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (F32) {
                            dd[u] = c[u];
                            e[u] = lds_u8(min(max(c[u] >> lut_sh, lut_base) + sa_lut0, lut_last));
                        } else {
                            asm("mov.b32 %0, {%1, %1};" : "=r"(dd[u]) : "h"((unsigned short)c[u]));
                            e[u] = lds_u8(min((c[u] >> lut_sh) + sa_lut, lut_last));
                        }
                    }
                    uint4 en[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (F32) en[u] = lds_u128(sa_ent + 16u * e[u]);
                        else { const uint2 t = lds_u64(sa_ent + 8u * e[u]); en[u] = make_uint4(t.x, 0u, t.y, 0u); }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        // key = layer byte | RGB: one PRMT on the funnel-shifted pixel word; a0 = kb + low half of the offset pair (one dp2a)
                        const uint32_t key0 = __byte_perm(__funnelshift_r(w0[u], w1[u], wsh), e[u], 0x4210);
                        const uint32_t kb = sa_keys_lo + x4[u];
                        uint32_t a0 = __dp2a_lo(en[u].z, 1u, kb), a1 = kb + (en[u].z >> 16);
                        if (WRAP) {                              // see k_warp_fused: first segment left end, last two right end
                            uint32_t q0 = a0 - sa_keys, q1 = a1 - sa_keys;
                            if (u == 0) { q0 = min(q0, q0 + W4); q1 = min(q1, q1 + W4); }
                            if (u >= U - 2) { q0 = min(q0, q0 - W4); q1 = min(q1, q1 - W4); }
                            a0 = sa_keys + q0;
                            a1 = sa_keys + q1;
                        }
                        uint32_t k0, k1;
                        if (F32) {
                            const float d = __uint_as_float(dd[u]);
                            k0 = (d < __uint_as_float(en[u].x)) ? key0 : 0u;                 // member of layer e-1 iff d < hi(e-1)
                            k1 = (d < __uint_as_float(en[u].y)) ? 0u : key0 + 0x01000000u;   // member of layer e iff !(d < lo(e))
                        } else {
                            asm("{\n\t.reg .pred p, q;\n\t"
                                "setp.lt.f16x2 p|q, %2, %3;\n\t"
                                "selp.u32 %0, %4, 0, p;\n\t"
                                "selp.u32 %1, 0, %5, q;\n\t}"
                                : "=r"(k0), "=r"(k1) : "r"(dd[u]), "r"(en[u].x), "r"(key0), "r"(key0 + 0x01000000u));
                        }
                        asm volatile("red.shared.max.u32 [%0], %1;" :: "r"(a0), "r"(k0) : "memory");
                        asm volatile("red.shared.max.u32 [%0], %1;" :: "r"(a1), "r"(k1) : "memory");
                    }
                };
                using T = std::true_type; using Fa = std::false_type;
                int r = 0;
                for (; r + 4 < sch_nw; r += 4) {
                    if (batch_edge(r, 4)) batch(std::integral_constant<int, 4>{}, T{}, r);
                    else batch(std::integral_constant<int, 4>{}, Fa{}, r);
                }
                const bool ew = batch_edge(r, sch_nw - r);
                switch (sch_nw - r) {
                    case 4: if (ew) batch(std::integral_constant<int, 4>{}, T{}, r); else batch(std::integral_constant<int, 4>{}, Fa{}, r); break;
                    case 3: if (ew) batch(std::integral_constant<int, 3>{}, T{}, r); else batch(std::integral_constant<int, 3>{}, Fa{}, r); break;
                    case 2: batch(std::integral_constant<int, 2>{}, T{}, r); break;
                    case 1: batch(std::integral_constant<int, 1>{}, T{}, r); break;
                    default: break;
                }
            } else {
                // slow path: brute-force membership, layer-only keys
                const int L = (int)(hdrw.w >> 8);
                const float2 *gb = a.bounds + (size_t)t0 * a.Lcap;
                const int *go = a.offm + (size_t)t0 * (a.Lcap + 1);
                for (int seg = warp; seg < nseg; seg += NS) {
                    const int x = (seg << 5) + lane;
                    const float d = F32 ? __uint_as_float(lds_u32(sa_cur + 4u * x)) : __half2float(__ushort_as_half((unsigned short)lds_u16(sa_cur + 2u * x)));
                    for (int k = 0; k < L; ++k) {
                        const float2 bd = __ldg(gb + k);
                        if (bd.x <= d && d < bd.y) {
                            int xd = x + __ldg(go + k + 1);
                            xd -= (xd >= W) ? W : 0;
                            asm volatile("red.shared.max.u32 [%0], %1;" :: "r"(sa_keys + 4u * (uint32_t)xd), "r"((uint32_t)(k + 1)) : "memory");
                        }
                    }
                }
            }
            __syncwarp();
            named_bar_arrive(kBarFull + (int)b, NT);                       // keys[b] complete (orders this warp's atomics)
            if (++i3 == 3) { i3 = 0; par3 ^= 1u; }
            next_yt(y0, t0);
        }
    } else {
        // =============================== destination warps ===========================================
        const int dt = tid - NS * 32;                                     // 0 .. NDT-1
        const bool elect = dt == 0;                                       // issues the bulk stores, arrives on k_empty
        const bool loader = dt == NDT - 32;                               // lane 0 of the last destination warp: TMA loads
        int y1 = y0, t1 = t0; next_yt(y1, t1);
        int y2 = y1, t2 = t1; next_yt(y2, t2);
        auto issue = [&](int k, int y, int t) {                           // elected thread: TMA loads of iteration k
            const uint32_t s3 = (uint32_t)(k % 3), bar = bar_in + 8u * s3;
            mbar_expect_tx_a(bar, img_bytes + dep_bytes + a.blob_bytes);
            bulk_g2s_a(sb + wa.lay.img + s3 * wa.lay.img_stride, a.frames + ((size_t)t * H + y) * img_bytes, img_bytes, bar);
            bulk_g2s_a(sb + wa.lay.dep + s3 * wa.lay.dep_stride, reinterpret_cast<const uint8_t *>(a.depth) + ((size_t)t * H + y) * dep_bytes, dep_bytes, bar);
            bulk_g2s_a(sb + wa.lay.blob + (uint32_t)(k & 1) * wa.lay.blob_stride, a.blobs + (size_t)t * a.blob_bytes, a.blob_bytes, bar);
        };
        if (loader) {
            issue(0, y0, t0);
            if (N > 1) issue(1, y1, t1);
        }
        int i3 = 0;
        uint32_t par3 = 0;
        for (int n = 0; n < N; ++n) {
            const uint32_t b = (uint32_t)n & 1u;
            const uint32_t sa_keys = sb + wa.lay.keys + b * wa.lay.keys_stride;
            const uint32_t sa_blob = sb + wa.lay.blob + b * wa.lay.blob_stride;
            const uint32_t sa_imgrow = sb + wa.lay.img + (uint32_t)i3 * wa.lay.img_stride;
            mbar_wait_sleep(bar_in + 8u * (uint32_t)i3, par3);
            const uint4 hdrw = lds_u128(sa_blob);
            const bool fast = hdrw.w & 1u;
            const int fill = (int)hdrw.x;
            named_bar_sync(kBarFull + (int)b, NT);                         // parked in hardware until every scatter warp has arrived
            if (elect) bulk_wait_read0();                 // the previous row's bulk stores have finished reading out / img slots
            named_bar_sync(1, NDT);
            // row n+2 goes into the slots of row n-1 (scattered, packed, and - the barrier above - read by its stores) and into
            // blob slot b, whose header every destination thread holds in registers and whose tables only scatter(n) -
            // complete, k_full seen - needed.  A different thread than the storer issues it, off the storer's critical path.
            if (loader && n + 2 < N) issue(n + 2, y2, t2);
            auto fetch = [&](int xs) {
                const uint32_t ab = 3u * (uint32_t)xs, wadr = sa_imgrow + (ab & ~3u);
                return __funnelshift_r(lds_u32(wadr), lds_u32(wadr + 4u), (ab & 3u) * 8u) & 0x00ffffffu;
            };
            for (int j = dt; j < nquad; j += NDT) {
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
            fence_async_smem();
            named_bar_sync(1, NDT);
            const uint32_t row = (uint32_t)t0 * (uint32_t)H + (uint32_t)y0;
            if (elect) {
                uint8_t *go = a.sbs + (size_t)row * img_bytes * 2;
                bulk_s2g_a(go, sa_out, img_bytes);
                if (!a.skip_right) bulk_s2g_a(go + img_bytes, sa_imgrow, img_bytes);
                bulk_commit();
            }
            __syncwarp();
            if (dt < 32 && n + 2 < N) named_bar_arrive(kBarEmpty + (int)b, NS * 32 + 32);   // keys[b] is zero again (scatter of row n+2 waits for it)
            // hole mask row -> global bitmask + blur work list (first ceil(Wwords/32) destination warps)
            if (dt < ((Wwords + 31) & ~31)) {
                const int w = dt;
                uint32_t v = 0;
                if (w < Wwords) {
                    v = lds_u32(sa_mask + 4u * w);
                    sts_u32(sa_mask + 4u * w, 0u);
                    a.hole_mask[(size_t)row * Wwords + w] = v;
                }
                const unsigned nz = __ballot_sync(0xffffffffu, v != 0u);
                if (nz) {
                    const unsigned holes = __reduce_add_sync(0xffffffffu, (unsigned)__popc(v));
                    if (lane == 0) atomicAdd(&a.tabs[t0].holes, (unsigned long long)holes);
                    if (a.band_map) {
                        // k_blur_band / k_blur_commit walk a bitmap of band columns (8 rows x one mask word) that hold a hole: one
                        // fire-and-forget OR per row and 32-word group, no list slot to wait for
                        if (lane == 0) atomicOr(a.band_map + ((size_t)t0 * a.Hb + ((uint32_t)y0 >> 3)) * a.band_groups + (dt >> 5), nz);
                    } else if (a.hole_list) {                 // (nullptr: k_word_list makes the per-word list from the hole mask)
                        uint32_t base = 0;
                        if (lane == 0) base = atomicAdd(a.hole_count, (uint32_t)__popc(nz));
                        base = __shfl_sync(0xffffffffu, base, 0);
                        if (v) a.hole_list[base + __popc(nz & ((1u << lane) - 1u))] = (row << 8) | (uint32_t)w;
                    }
                }
            }
            if (++i3 == 3) { i3 = 0; par3 ^= 1u; }
            y0 = y1; t0 = t1; y1 = y2; t1 = t2; next_yt(y2, t2);
        }
        if (elect) bulk_wait0();
    }
}

}  // namespace vrsbs
