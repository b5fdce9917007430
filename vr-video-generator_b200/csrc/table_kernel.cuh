// Stage 2: SbsProcessor.get_cutoff on the device (PredictAndGenerate.py:101-126) plus the derived
// per-frame constants of left_side_sbs (fill layer :190, strip width :196, fp16-narrowed bounds :173).
//
// One CTA per frame of the batch.  The only cross-frame dependency is the EMA of the offset range
// (`self.last_offset_range`), a chain of two double adds per frame; every CTA re-walks that chain
// from the persisted state up to its own frame (B <= a few hundred iterations of trivial work), so
// no grid-wide ordering is needed.  All arithmetic is IEEE double with explicit-rounding
// intrinsics (never FMA-contracted), in the reference's python operation order, so the tables are
// bit-identical to the python lists (parity tier T1).
#pragma once
#include "common.cuh"

namespace vrsbs {

struct TableArgs {
    const uint32_t *frame_max;   // [B] order-encoded max of the smoothed depth
    const uint32_t *frame_nan;   // [B]
    const RangeState *state_in;  // EMA state before this batch
    RangeState *state_out;       // EMA state after this batch (written by the last frame's CTA)
    FrameTab *tabs;              // [B]
    float2 *bounds;              // [B][Lcap]   (lo, hi) as floats holding exact fp16 values
    int *offm;                   // [B][Lcap+1] offsets mod W; slot 0 = fill layer's offset
    double *cutoffs;             // [B][Lcap+1] cutoff_list (introspection / T1)
    int *offsets;                // [B][Lcap]   offset_x_list (signed)
    uint16_t *lo16, *hi16;       // [B][Lcap]   fp16 bit patterns of the bounds
    uint8_t *blobs;              // [B][blob_stride] fast-path tables (see common.cuh), or nullptr
    double offset_fg, offset_bg;
    int step, B, H, W, Lcap;
    int ent_cap, lut_cap;        // capacity of the blob's entry table (layers) and cell LUT (bytes)
    int key_pad;                 // fast path: every signed offset (shortest way round the row) must satisfy |off| <= key_pad
    int f32;                     // 1: the depth is fp32 - bounds narrowed double -> float, fp32 cell LUT
    float cell_width;            // host estimate of the widest valid LUT cell in depth units (0.85 * 0.9 * layer width): where the
                                 // search for the cell shift starts (any shift that validates is correct)
};

// LUT value for the cell of depth values [vmin, vmax] (monotone bounds required): e such that every
// value of the cell can only belong to layer e-1 (iff v < hi[e-1]) or layer e (iff v >= lo[e]);
// -1 if the cell is too coarse for that to hold.
__device__ __forceinline__ int cell_entry(const float2 *bounds, int L, float vmin, float vmax) {
    int lo_i = 0, hi_i = L;                       // a = #{k : hi_k <= vmin}
    while (lo_i < hi_i) {
        int mid = (lo_i + hi_i) >> 1;
        if (bounds[mid].y <= vmin) lo_i = mid + 1; else hi_i = mid;
    }
    const int a = lo_i;
    if (a >= L) return L;                         // above every layer: nothing is painted
    if (vmin >= bounds[a].x && (a + 2 > L - 1 || vmax < bounds[a + 2].x) && (a + 1 > L - 1 || vmax < bounds[a + 1].y))
        return a + 1;
    if (vmax < bounds[a].y && (a + 1 > L - 1 || vmax < bounds[a + 1].x)) return a;
    return -1;
}

__device__ __forceinline__ double py_round(double v) { return rint(v); }   // round-half-even
__device__ __forceinline__ int clamp_int(double v) {
    return v > 2.0e9 ? 2000000000 : (v < -2.0e9 ? -2000000000 : (int)v);
}

__global__ void __launch_bounds__(256) k_build_tables(TableArgs a) {
    extern __shared__ double s_val[];          // [max(Lcap+2,B)] unsorted marks, then as many sorted
    const int sstride = max(a.Lcap + 2, a.B);
    double *s_sorted = s_val + sstride;
    int *s_off = reinterpret_cast<int *>(s_val + 2 * sstride);       // [Lcap + 1] offset_x_list, kept for the later phases
    float2 *s_bounds = reinterpret_cast<float2 *>(s_val);           // [Lcap] fp16 bounds as floats; s_val is free after the sort
    __shared__ double s_r0, s_r1, s_span, s_top;
    __shared__ int s_limit, s_start, s_nneg, s_E, s_bad;
    __shared__ float s_max;

    const int b = blockIdx.x;
    FrameTab *tab = a.tabs + b;
    pdl_launch_dependents();
    pdl_wait();                                  // the frame maxima of the depth pass

    // this frame's own range for every frame up to b, in parallel (the divisions are the expensive part) ...
    double *s_cur0 = s_val, *s_cur1 = s_sorted;                      // both reused as mark buffers afterwards
    for (int t = threadIdx.x; t <= b; t += blockDim.x) {
        const float fm = ord2f(a.frame_max[t]);
        const double c = ceil((double)fm);
        const int lim = (a.frame_nan[t] || !(c == c)) ? 0 : clamp_int(c);
        // bg * H * limit / 14, left to right
        s_cur0[t] = __ddiv_rn(__dmul_rn(__dmul_rn(a.offset_bg, (double)a.H), (double)lim), 14.0);
        s_cur1[t] = __ddiv_rn(__dmul_rn(__dmul_rn(a.offset_fg, (double)a.H), (double)lim), 14.0);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double l0 = a.state_in->range[0], l1 = a.state_in->range[1];
        int has = a.state_in->has_last;
        double r0 = 0, r1 = 0;
        // ... then the EMA chain: (last + cur) / 2 is exact as a multiplication by 0.5
        for (int t = 0; t <= b; ++t) {
            r0 = s_cur0[t]; r1 = s_cur1[t];
            if (has) {
                r0 = __dmul_rn(__dadd_rn(l0, r0), 0.5);
                r1 = __dmul_rn(__dadd_rn(l1, r1), 0.5);
            }
            l0 = r0; l1 = r1; has = 1;
        }
        const float fmax = ord2f(a.frame_max[b]);
        const double cb = ceil((double)fmax);
        const int limit = (a.frame_nan[b] || !(cb == cb)) ? 0 : clamp_int(cb);
        if (b == a.B - 1) {
            a.state_out->range[0] = r0;
            a.state_out->range[1] = r1;
            a.state_out->has_last = 1;
        }
        s_r0 = r0; s_r1 = r1; s_limit = limit; s_max = fmax;
        s_span = __dsub_rn(__dadd_rn(0.00001, r1), r0);       // 0.00001 + r1 - r0
        s_top = __dadd_rn(0.00001, (double)limit);            // 0.00001 + limit_step
        int start = clamp_int(py_round(r0)), stop = clamp_int(py_round(r1));
        long long nneg = start < 0 ? ((long long)(-(long long)start) + a.step - 1) / a.step : 0;   // range(start,0,step)
        long long npos = stop > 1 ? ((long long)stop - 1 + a.step - 1) / a.step : 0;              // range(1,stop,step)
        long long E = nneg + 1 + npos + 1;
        s_bad = (E - 1 > a.Lcap) ? 1 : 0;
        s_start = start; s_nneg = (int)min(nneg, (long long)a.Lcap);
        s_E = s_bad ? 2 : (int)E;
    }
    __syncthreads();
    const double r0 = s_r0, span = s_span, top = s_top;
    const int E = s_E, L = E - 1, nneg = s_bad ? 0 : s_nneg;

    // marks before sorting, in the order the reference appends them
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        double v;
        if (e == E - 1) {
            v = (double)s_limit;
        } else {
            int px = (e < nneg) ? s_start + e * a.step : (e == nneg ? 0 : 1 + (e - nneg - 1) * a.step);
            v = __dmul_rn(__ddiv_rn(__dsub_rn((double)px, r0), span), top);
        }
        s_val[e] = v;
    }
    __syncthreads();
    // sorted(): stable rank sort (ties keep append order, as timsort does)
    for (int e = threadIdx.x; e < E; e += blockDim.x) {
        double v = s_val[e];
        int rank = 0;
        for (int j = 0; j < E; ++j) {
            double u = s_val[j];
            rank += (u < v) || (u == v && j < e);
        }
        s_sorted[rank] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) s_sorted[0] = 0.0;                  // cutoff_list[0] = 0
    __syncthreads();

    float2 *bounds = a.bounds + (size_t)b * a.Lcap;
    int *offm = a.offm + (size_t)b * (a.Lcap + 1);
    int *offs = a.offsets + (size_t)b * a.Lcap;
    double *cuts = a.cutoffs + (size_t)b * (a.Lcap + 1);
    for (int e = threadIdx.x; e < E; e += blockDim.x) cuts[e] = s_sorted[e];
    for (int k = threadIdx.x; k < L; k += blockDim.x) {
        double c = s_sorted[k], s = __dsub_rn(s_sorted[k + 1], c);
        // round(c / top * span + r0)
        int off = clamp_int(py_round(__dadd_rn(__dmul_rn(__ddiv_rn(c, top), span), r0)));
        // python double -> float -> half (c10::Half has only a float constructor)
        const float lo32 = __double2float_rn(__dsub_rn(c, __dmul_rn(0.05, s))), hi32 = __double2float_rn(__dadd_rn(c, __dmul_rn(1.05, s)));
        __half lo = __float2half_rn(lo32);
        __half hi = __float2half_rn(hi32);
        // the comparison runs in the depth dtype (:173): fp16 depth sees the fp16-narrowed bounds, fp32 depth the fp32 ones
        const float2 bd = a.f32 ? make_float2(lo32, hi32) : make_float2(__half2float(lo), __half2float(hi));
        bounds[k] = bd;
        s_bounds[k] = bd;                                      // (s_val region: every read of the marks is behind two barriers)
        a.lo16[(size_t)b * a.Lcap + k] = __half_as_ushort(lo);
        a.hi16[(size_t)b * a.Lcap + k] = __half_as_ushort(hi);
        offs[k] = off;
        s_off[k] = off;
        offm[k + 1] = wrap_mod(off, a.W);
    }
    __syncthreads();
    int mono = 1;
    for (int k = threadIdx.x; k + 1 < L; k += blockDim.x) {
        float2 p = s_bounds[k], q = s_bounds[k + 1];
        if (!(p.x <= q.x) || !(p.y <= q.y)) mono = 0;
    }
    mono = __syncthreads_and(mono);
    const int fill_off = wrap_mod(s_off[(int)((double)(L * 3) / 5.0)], a.W);
    if (threadIdx.x == 0) {
        int fill = (int)((double)(L * 3) / 5.0);               // int(len(offset_img)*3/5)
        offm[0] = fill_off;
        double sn = py_round(__dmul_rn(__ddiv_rn((double)s_off[L - 1], 3.0), 2.0));   // round(offset_x/3*2)
        int n = clamp_int(sn);
        int strip = n >= 0 ? min(n, a.W) : max(0, a.W + n);    // python slice 0:n
        float scale = 0.f, bias = 0.f;
        if (L >= 3) {
            float lo1 = s_bounds[1].x, loN = s_bounds[L - 1].x;
            if (loN > lo1) { scale = (float)(L - 2) / (loN - lo1); bias = 1.f - lo1 * scale; }
        }
        tab->layers = L;
        tab->fill_off = fill_off;
        tab->strip = strip;
        tab->status = (a.frame_nan[b] ? VRSBS_FRAME_NAN : 0u) | (s_bad ? VRSBS_FRAME_OVERFLOW : 0u) |
                      (mono ? 0u : VRSBS_FRAME_GENERIC);
        tab->guess_scale = scale;
        tab->guess_bias = bias;
        tab->limit_step = s_limit;
        tab->fill_layer = fill;
        tab->depth_max = s_max;
        tab->range[0] = s_r0;
        tab->range[1] = s_r1;
        tab->holes = 0ull;
    }

    // ---- fast-path blob: entry table + cell LUT (validated exactly, cell by cell) ----------------------
    if (!a.blobs) { if (threadIdx.x == 0) { tab->fast = 0; tab->lut_shift = 0; tab->lut_cells = 0; } return; }
    const bool f32 = a.f32 != 0;
    const uint32_t ent_bytes = blob_ent_bytes(a.ent_cap, f32);
    uint8_t *blob = a.blobs + (size_t)b * blob_bytes(a.ent_cap, a.lut_cap, f32);
    BlobHdr *hdr = reinterpret_cast<BlobHdr *>(blob);
    LayerEnt *ent = reinterpret_cast<LayerEnt *>(blob + 16);
    LayerEnt32 *ent32 = reinterpret_cast<LayerEnt32 *>(blob + 16);
    uint8_t *lut = blob + 16 + ent_bytes;
    const float fmax = s_max;
    bool ok = mono && !s_bad && !a.frame_nan[b] && L <= a.ent_cap && L <= 255 && a.W * 4 <= 65535 && fmax < 60000.f;
    // (the LUT search binary-searches the bounds thousands of times: they are in shared memory already)
    const uint32_t maxbits = (fmax > 0.f) ? (f32 ? __float_as_uint(fmax) : (uint32_t)__half_as_ushort(__float2half_rn(fmax))) : 0u;
    {
        int fits = 1;
        for (int k = threadIdx.x; k < L; k += blockDim.x) {
            const int o = wrap_mod(s_off[k], a.W), so = o * 2 < a.W ? o : o - a.W;
            if (so > a.key_pad || so < -a.key_pad) fits = 0;
        }
        ok = __syncthreads_and(fits) && ok;
    }
    int shift = -1;
    uint32_t ncells = 0, lbase = 0;
    // first shift to try: one coarser than the estimate for the binade of the frame maximum (the search walks towards finer
    // cells until one validates; starting at the coarsest shift cost four or five failed passes per frame)
    int sh_est = 0;
    if (fmax > 0.f && a.cell_width > 0.f) {
        int e2;
        frexpf(fmax, &e2);                                     // fmax = m * 2^e2, m in [0.5, 1): binade exponent e2 - 1
        const float ulp = ldexpf(1.f, e2 - 1 - (f32 ? 23 : 10));
        while (sh_est < 30 && ulp * (float)(2u << sh_est) <= a.cell_width) ++sh_est;
    }
    if (ok && !f32) {
        for (int sh = min(9, sh_est + 1); sh >= 0; --sh) {
            const uint32_t nc = (maxbits >> sh) + 1;
            if (nc + 1 > (uint32_t)a.lut_cap) break;
            int valid = 1;
            for (uint32_t q = threadIdx.x; q <= nc; q += blockDim.x) {
                float vmin, vmax;
                if (q < nc) {
                    uint32_t b0 = q << sh, b1 = min(((q + 1) << sh) - 1, maxbits);
                    vmin = __half2float(__ushort_as_half((unsigned short)b0));
                    vmax = __half2float(__ushort_as_half((unsigned short)b1));
                } else {                           // every negative value (and -0.0)
                    vmin = -INFINITY; vmax = 0.f;
                }
                int e = cell_entry(s_bounds, L, vmin, vmax);
                if (e < 0) valid = 0; else lut[q] = (uint8_t)e;
            }
            if (__syncthreads_and(valid)) { shift = sh; ncells = nc; break; }
        }
    } else if (ok) {
        // fp32 depth: cell = max(bits >> sh, base) - base, base = the cell of 2^-6; cell 0 holds every value in [0, end of that
        // cell], the last cell every negative value.  From 8 cells per binade (sh = 20) down to 1024 (sh = 13).
        for (int sh = max(13, min(20, sh_est + 1)); sh >= 13; --sh) {
            const uint32_t cb = kF32LutFloorBits >> sh;
            const uint32_t top = max(maxbits >> sh, cb);
            const uint32_t nc = top - cb + 1;
            if (nc + 1 > (uint32_t)a.lut_cap) break;
            int valid = 1;
            for (uint32_t q = threadIdx.x; q <= nc; q += blockDim.x) {
                float vmin, vmax;
                if (q < nc) {
                    const uint32_t b0 = q ? (cb + q) << sh : 0u, b1 = max(min((((cb + q) + 1) << sh) - 1, maxbits), b0);
                    vmin = __uint_as_float(b0);
                    vmax = __uint_as_float(b1);
                } else {
                    vmin = -INFINITY; vmax = 0.f;
                }
                int e = cell_entry(s_bounds, L, vmin, vmax);
                if (e < 0) valid = 0; else lut[q] = (uint8_t)e;
            }
            if (__syncthreads_and(valid)) { shift = sh; ncells = nc; lbase = cb; break; }
        }
    }
    ok = ok && shift >= 0;
    for (int e = threadIdx.x; e <= L && e <= a.ent_cap; e += blockDim.x) {
        auto biased4 = [&](int o) { return (uint32_t)(((o * 2 < a.W ? o : o - a.W) + a.key_pad) * 4) & 0xffffu; };
        const uint32_t o0 = e >= 1 ? biased4(wrap_mod(s_off[e - 1], a.W)) : 0u;
        const uint32_t o1 = e <= L - 1 ? biased4(wrap_mod(s_off[e], a.W)) : 0u;
        if (f32) {
            LayerEnt32 le;
            le.hi_prev = e >= 1 ? s_bounds[e - 1].y : -INFINITY;
            le.lo_cur = e <= L - 1 ? s_bounds[e].x : INFINITY;
            le.off4 = (o0 & 0xffffu) | (o1 << 16);
            le.pad = 0u;
            ent32[e] = le;
        } else {
            LayerEnt le;
            const uint32_t hi = e >= 1 ? (uint32_t)__half_as_ushort(__float2half_rn(s_bounds[e - 1].y)) : 0xFC00u;   // -inf
            const uint32_t lo = e <= L - 1 ? (uint32_t)__half_as_ushort(__float2half_rn(s_bounds[e].x)) : 0x7C00u;    // +inf
            le.hi_lo = hi | (lo << 16);
            le.off4 = (o0 & 0xffffu) | (o1 << 16);
            ent[e] = le;
        }
    }
    if (threadIdx.x == 0) {
        hdr->fill_off = fill_off;
        hdr->shift = ok ? ((uint32_t)shift | (lbase << 8)) : 0u;
        hdr->ncells = ok ? ncells : 0u;
        hdr->flags = (ok ? 1u : 0u) | ((uint32_t)L << 8);
        tab->fast = ok ? 1u : 0u;
        tab->lut_shift = hdr->shift;
        tab->lut_cells = hdr->ncells;
    }
}

}  // namespace vrsbs
