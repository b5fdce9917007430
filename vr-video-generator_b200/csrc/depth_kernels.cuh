// Stage 1: depth tail (bicubic + scaler), temporal smoothing, per-frame max.
//
// Replaces dpt.py:196 (F.interpolate bicubic, align_corners=True), PredictAndGenerate.py:55
// (`* scaler`), PredictAndGenerate.py:134-144 (smoothing over RAW history) and the depth.max() of
// PredictAndGenerate.py:102.
//
// Layout trick: one thread owns a pixel (or 8 consecutive pixels) for ALL frames of the batch and
// walks the batch in time order with the two previous raw depths in registers.  The history is
// therefore read/written once per batch instead of twice per frame, and the only per-frame HBM
// traffic is "read raw (or low-res taps), write smoothed".
#pragma once
#include "common.cuh"

namespace vrsbs {

struct SmoothWeights {
    float w_now, w_prev1, w_prev2;   // 1-(0.3+0.12), 0.3, 0.3*0.4 as python doubles narrowed to fp32
};

struct DepthArgs {
    const __half *raw;        // [B, n] full-res raw depth            (full variant)
    const __half *lowres;     // [B, h, w] DPT output                 (lowres variant)
    __half *out;              // [B, n] smoothed
    __half *hist1;            // [n] raw depth of frame t-1 (state)
    __half *hist2;            // [n] raw depth of frame t-2 (state)
    uint32_t *frame_max;      // [B] order-encoded float max, pre-zeroed
    uint32_t *frame_nan;      // [B] !=0 if the smoothed frame contains a NaN, pre-zeroed
    SmoothWeights sw;
    int B, H, W, h, w;
    int first;                // 1: frame 0 of this call is the first frame of the clip range
    float scaler;
    float scale_y, scale_x;   // (h-1)/(H-1), (w-1)/(W-1) in fp32 (area_pixel_compute_scale)
};

__device__ __forceinline__ void frame_max_commit(uint32_t enc, bool nan, int t, uint32_t *s_max, uint32_t *s_nan) {
    enc = __reduce_max_sync(0xffffffffu, enc);
    unsigned any_nan = __ballot_sync(0xffffffffu, nan);
    if ((threadIdx.x & 31) == 0) {
        atomicMax(&s_max[t], enc);
        if (any_nan) s_nan[t] = 1;
    }
}

// ---- full-resolution raw depth in, 8 pixels per thread -------------------------------------------
template <int VEC>
__global__ void __launch_bounds__(256) k_depth_full(DepthArgs a) {
    extern __shared__ uint32_t s_red[];          // [B] max | [B] nan
    uint32_t *s_max = s_red, *s_nan = s_red + a.B;
    for (int i = threadIdx.x; i < 2 * a.B; i += blockDim.x) s_red[i] = 0;
    __syncthreads();

    const size_t n = (size_t)a.H * a.W;
    const size_t nvec = (n + VEC - 1) / VEC;
    const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = v < nvec;
    const size_t base = v * VEC;

    __half p1[VEC], p2[VEC], cur[VEC], res[VEC];
    auto load = [&](const __half *src, __half *dst) {
        if (VEC == 8 && base + 8 <= n) {
            *reinterpret_cast<uint4 *>(dst) = *reinterpret_cast<const uint4 *>(src + base);
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e) dst[e] = (base + e < n) ? src[base + e] : __float2half(0.f);
        }
    };
    auto store = [&](__half *dst, const __half *src) {
        if (VEC == 8 && base + 8 <= n) {
            *reinterpret_cast<uint4 *>(dst + base) = *reinterpret_cast<const uint4 *>(src);
        } else {
#pragma unroll
            for (int e = 0; e < VEC; ++e)
                if (base + e < n) dst[base + e] = src[e];
        }
    };

    if (active && !a.first) { load(a.hist1, p1); load(a.hist2, p2); }
    for (int t = 0; t < a.B; ++t) {
        uint32_t enc = 0;
        bool nan = false;
        if (active) {
            load(a.raw + (size_t)t * n, cur);
            if (a.first && t == 0) {
#pragma unroll
                for (int e = 0; e < VEC; ++e) p1[e] = p2[e] = cur[e];
            }
#pragma unroll
            for (int e = 0; e < VEC; ++e) {
                res[e] = smooth3(cur[e], p1[e], p2[e], a.sw.w_now, a.sw.w_prev1, a.sw.w_prev2);
                float f = h2f(res[e]);
                if (base + e < n) {
                    if (f != f) nan = true; else enc = max(enc, f2ord(f));
                }
                p2[e] = p1[e];
                p1[e] = cur[e];
            }
            store(a.out + (size_t)t * n, res);
        }
        frame_max_commit(enc, nan, t, s_max, s_nan);
    }
    if (active) { store(a.hist1, p1); store(a.hist2, p2); }
    __syncthreads();
    for (int t = threadIdx.x; t < a.B; t += blockDim.x) {
        if (s_max[t]) atomicMax(&a.frame_max[t], s_max[t]);
        if (s_nan[t]) atomicOr(&a.frame_nan[t], 1u);
    }
}

// ---- packed fp16 smoothing: two pixels per instruction where the reference's rounding allows it ------
// rn16(float(x) * w) needs the fp32 product (w is not an fp16 value); the two adds are plain fp16 adds
// (an fp32 add of two fp16 values rounded to fp16 equals the correctly rounded fp16 add: 24 >= 2*11+2).
__device__ __forceinline__ __half2 mul_w(__half2 v, float w) {
    const float2 f = __half22float2(v);
    return __floats2half2_rn(__fmul_rn(f.x, w), __fmul_rn(f.y, w));
}
__device__ __forceinline__ __half2 smooth3x2(__half2 cur, __half2 p1, __half2 p2, const SmoothWeights &sw) {
    __half2 d = __hadd2(mul_w(cur, sw.w_now), mul_w(p1, sw.w_prev1));
    return __hadd2(d, mul_w(p2, sw.w_prev2));
}

// ---- depth pass: temporal smoothing + per-frame max of the SMOOTHED depth ---------------------------------
// Reads raw [B,n] once.  STORE = true (default route): also writes the smoothed depth the warp kernel consumes.
// STORE = false (option smooth_in_warp): max only; k_warp_fused<true> then recomputes the smoothing from the same
// raw rows and the smoothed depth never exists in HBM.  Writes the next batch's history (raw B-1, raw B-2) to
// hist1_out/hist2_out (the same buffers in place, or the other ping-pong set when the warp kernel still reads them).
struct DepthMaxArgs {
    __half *out;                         // [B, n] smoothed depth (STORE variant) or nullptr
    const __half *raw;                   // [B, n]
    const __half *hist1, *hist2;         // [n] raw t-1, t-2 of the previous batch
    __half *hist1_out, *hist2_out;       // [n]
    uint32_t *frame_max, *frame_nan;     // [B], pre-zeroed
    SmoothWeights sw;
    int B, first;
    size_t n;                            // H*W, multiple of 8
};

// 16-bit order-preserving key of an fp16 bit pattern (NaN -> 0x1ffff, above every number)
__device__ __forceinline__ uint32_t h16_key(uint32_t u) {
    const uint32_t k = (u & 0x8000u) ? (~u & 0xffffu) : (u | 0x8000u);
    return ((u & 0x7fffu) > 0x7c00u) ? 0x1ffffu : k;
}

template <bool STORE>
__global__ void __launch_bounds__(256) k_depth_pass(DepthMaxArgs a) {
    extern __shared__ uint32_t s_red[];          // [B] keys
    for (int i = threadIdx.x; i < a.B; i += blockDim.x) s_red[i] = 0;
    __syncthreads();
    const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = v * 8 < a.n;
    const size_t base = active ? v * 8 : 0;
    union V { uint4 u; __half2 h[4]; };
    const int lane = threadIdx.x & 31;
    const float w0 = a.sw.w_now, w1 = a.sw.w_prev1, w2 = a.sw.w_prev2;

    // Raw depth as floats in three rotating register sets X, Y, Z (each raw value is converted once and used by
    // three frames).  Frames are processed in trios so that the roles (current, t-1, t-2) rotate by renaming only.
    float X[8], Y[8], Z[8];
    auto unpack = [](const uint4 &u, float *f) {
        V t; t.u = u;
#pragma unroll
        for (int e = 0; e < 4; ++e) { const float2 g = __half22float2(t.h[e]); f[2 * e] = g.x; f[2 * e + 1] = g.y; }
    };
    auto pack = [](const float *f) -> uint4 {
        V t;
#pragma unroll
        for (int e = 0; e < 4; ++e) t.h[e] = __floats2half2_rn(f[2 * e], f[2 * e + 1]);   // exact: the values are fp16 values
        return t.u;
    };
    auto frame = [&](const float *c, const float *p1, const float *p2, int t) {
        uint32_t key = 0;
        if (active) {
            __half2 m;
            V sm;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                __half2 d = __hadd2(__floats2half2_rn(__fmul_rn(c[2 * e], w0), __fmul_rn(c[2 * e + 1], w0)),
                                    __floats2half2_rn(__fmul_rn(p1[2 * e], w1), __fmul_rn(p1[2 * e + 1], w1)));
                d = __hadd2(d, __floats2half2_rn(__fmul_rn(p2[2 * e], w2), __fmul_rn(p2[2 * e + 1], w2)));
                sm.h[e] = d;
                m = e ? __hmax2_nan(m, d) : d;
            }
            if (STORE) *reinterpret_cast<uint4 *>(a.out + (size_t)t * a.n + base) = sm.u;
            m = __hmax2_nan(m, __lowhigh2highlow(m));
            key = h16_key((uint32_t)__half_as_ushort(__low2half(m)));
        }
        key = __reduce_max_sync(0xffffffffu, key);
        if (lane == 0 && key) atomicMax(&s_red[t], key);
    };
    auto load = [&](int t) -> uint4 {
        return (active && t < a.B) ? __ldg(reinterpret_cast<const uint4 *>(a.raw + (size_t)t * a.n + base)) : make_uint4(0, 0, 0, 0);
    };
    {   // history: X = raw t-1, Y = raw t-2 (clip start: both are the first raw frame)
        const uint4 h1 = active ? __ldg(reinterpret_cast<const uint4 *>((a.first ? a.raw : a.hist1) + base)) : make_uint4(0, 0, 0, 0);
        const uint4 h2 = active ? __ldg(reinterpret_cast<const uint4 *>((a.first ? a.raw : a.hist2) + base)) : make_uint4(0, 0, 0, 0);
        unpack(h1, X);
        unpack(h2, Y);
    }
    // at the top of every trio: t-1 is in X, t-2 in Y
    for (int tb = 0; tb < a.B; tb += 6) {
        uint4 q[6];
#pragma unroll
        for (int i = 0; i < 6; ++i) q[i] = load(tb + i);       // six independent 16-byte loads in flight
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int t = tb + 3 * g;
            if (t < a.B)     { unpack(q[3 * g], Z);     frame(Z, X, Y, t); }
            if (t + 1 < a.B) { unpack(q[3 * g + 1], Y); frame(Y, Z, X, t + 1); }
            if (t + 2 < a.B) { unpack(q[3 * g + 2], X); frame(X, Y, Z, t + 2); }
        }
    }
    if (active) {
        // after B frames the newest two raw frames sit in (X,Y), (Z,X) or (Y,Z) depending on B mod 3
        const int r = a.B % 3;

        uint4 o1, o2;
        if (r == 0) { o1 = pack(X); o2 = pack(Y); } else if (r == 1) { o1 = pack(Z); o2 = pack(X); } else { o1 = pack(Y); o2 = pack(Z); }

        *reinterpret_cast<uint4 *>(a.hist1_out + base) = o1;
        *reinterpret_cast<uint4 *>(a.hist2_out + base) = o2;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < a.B; t += blockDim.x) {
        const uint32_t key = s_red[t];
        if (key == 0x1ffffu) atomicOr(&a.frame_nan[t], 1u);
        else if (key) {
            const uint32_t u = (key & 0x8000u) ? (key & 0x7fffu) : (~key & 0xffffu);
            atomicMax(&a.frame_max[t], f2ord(__half2float(__ushort_as_half((unsigned short)u))));
        }
    }
}

// ---- bicubic coefficients: ATen's cuda/UpSample.cuh arithmetic (A = -0.75) -----------------------
// `x + 1.0` is evaluated in double and narrowed, as the double literal in the ATen template forces.
// CONTRACT selects whether mul+add pairs are fused the way nvcc's default -fmad=true fuses them in
// torch's own binary (1) or kept separate (0); see DESIGN.md "bicubic parity".
template <bool CONTRACT>
struct Cubic {
    static __device__ __forceinline__ float mad(float a, float b, float c) {
        return CONTRACT ? __fmaf_rn(a, b, c) : __fadd_rn(__fmul_rn(a, b), c);
    }
    static __device__ __forceinline__ float conv1(float x) {   // ((A+2)x - (A+3)) x x + 1
        const float A = -0.75f;
        float t = mad(A + 2.f, x, -(A + 3.f));
        return mad(__fmul_rn(t, x), x, 1.f);
    }
    static __device__ __forceinline__ float conv2(float x) {   // ((A x - 5A) x + 8A) x - 4A
        const float A = -0.75f;
        float t = mad(A, x, -5.f * A);
        t = mad(t, x, 8.f * A);
        return mad(t, x, -4.f * A);
    }
    static __device__ __forceinline__ void coeffs(float t, float c[4]) {
        float x2 = (float)(1.0 - (double)t);
        c[0] = conv2((float)((double)t + 1.0));
        c[1] = conv1(t);
        c[2] = conv1(x2);
        c[3] = conv2((float)((double)x2 + 1.0));
    }
    static __device__ __forceinline__ float dot4(float v0, float v1, float v2, float v3, const float c[4]) {
        float r = __fmul_rn(v0, c[0]);
        r = mad(v1, c[1], r);
        r = mad(v2, c[2], r);
        return mad(v3, c[3], r);
    }
};

// ---- low-res DPT output in: bicubic + scaler + smoothing + max, one output pixel per thread ------
template <bool CONTRACT>
__global__ void __launch_bounds__(256) k_depth_lowres(DepthArgs a) {
    extern __shared__ uint32_t s_red[];
    uint32_t *s_max = s_red, *s_nan = s_red + a.B;
    for (int i = threadIdx.x; i < 2 * a.B; i += blockDim.x) s_red[i] = 0;
    __syncthreads();

    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const bool active = x < a.W && y < a.H;
    const size_t n = (size_t)a.H * a.W;
    const size_t pix = (size_t)y * a.W + x;

    float cx[4], cy[4];
    int ix[4], iy[4];
    {
        float rx = __fmul_rn(a.scale_x, (float)x), ry = __fmul_rn(a.scale_y, (float)y);
        int fx = (int)floorf(rx), fy = (int)floorf(ry);
        Cubic<CONTRACT>::coeffs(rx - (float)fx, cx);
        Cubic<CONTRACT>::coeffs(ry - (float)fy, cy);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            ix[k] = max(min(fx - 1 + k, a.w - 1), 0);
            iy[k] = max(min(fy - 1 + k, a.h - 1), 0) * a.w;
        }
    }
    __half p1 = __float2half(0.f), p2 = p1;
    if (active && !a.first) { p1 = a.hist1[pix]; p2 = a.hist2[pix]; }
    for (int t = 0; t < a.B; ++t) {
        uint32_t enc = 0;
        bool nan = false;
        if (active) {
            const __half *src = a.lowres + (size_t)t * a.h * a.w;
            float rows[4];
#pragma unroll
            for (int k = 0; k < 4; ++k)
                rows[k] = Cubic<CONTRACT>::dot4(h2f(__ldg(src + iy[k] + ix[0])), h2f(__ldg(src + iy[k] + ix[1])),
                                                h2f(__ldg(src + iy[k] + ix[2])), h2f(__ldg(src + iy[k] + ix[3])), cx);
            __half cur = f2h(Cubic<CONTRACT>::dot4(rows[0], rows[1], rows[2], rows[3], cy));
            if (a.scaler != 1.0f) cur = f2h(__fmul_rn(h2f(cur), a.scaler));
            if (a.first && t == 0) p1 = p2 = cur;
            __half res = smooth3(cur, p1, p2, a.sw.w_now, a.sw.w_prev1, a.sw.w_prev2);
            a.out[(size_t)t * n + pix] = res;
            float f = h2f(res);
            if (f != f) nan = true; else enc = f2ord(f);
            p2 = p1;
            p1 = cur;
        }
        frame_max_commit(enc, nan, t, s_max, s_nan);
    }
    if (active) { a.hist1[pix] = p1; a.hist2[pix] = p2; }
    __syncthreads();
    for (int t = threadIdx.x; t < a.B; t += blockDim.x) {
        if (s_max[t]) atomicMax(&a.frame_max[t], s_max[t]);
        if (s_nan[t]) atomicOr(&a.frame_nan[t], 1u);
    }
}

}  // namespace vrsbs
