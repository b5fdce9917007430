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
    pdl_launch_dependents();
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

// ---- fp32 depth (general route) ---------------------------------------------------------------------------------
// The reference's warp is dtype-agnostic, and torch >= 2.4's CUDA autocast puts upsample_bicubic2d on its fp32 list,
// so a producer running `infer_image_gpu` under autocast(fp16) hands the warp an fp32 map (measured on the B200 box).
// Smoothing then runs in fp32 (PredictAndGenerate.py:139-142: one rounding per op), the tail is bicubic on the fp32 view
// of the fp16 DPT output WITHOUT narrowing, `* scaler` in fp32.  One pixel per thread for all frames of the batch.
struct DepthArgs32 {
    const float *raw;         // [B, n] full-res raw depth (full variant)
    const __half *lowres;     // [B, h, w] DPT output (lowres variant)
    float *out;               // [B, n] smoothed
    float *hist1, *hist2;     // [n] raw depth of frames t-1, t-2
    uint32_t *frame_max, *frame_nan;
    SmoothWeights sw;
    int B, H, W, h, w, first;
    float scaler, scale_y, scale_x;
};

__device__ __forceinline__ float smooth3_f32(float cur, float p1, float p2, float w0, float w1, float w2) {
    float d = __fmul_rn(cur, w0);
    d = __fadd_rn(d, __fmul_rn(p1, w1));
    return __fadd_rn(d, __fmul_rn(p2, w2));
}

template <bool LOWRES, bool CONTRACT>
__global__ void __launch_bounds__(256) k_depth_f32(DepthArgs32 a) {
    extern __shared__ uint32_t s_red[];
    uint32_t *s_max = s_red, *s_nan = s_red + a.B;
    for (int i = threadIdx.x; i < 2 * a.B; i += blockDim.x) s_red[i] = 0;
    __syncthreads();
    const int x = blockIdx.x * 32 + (threadIdx.x & 31);
    const int y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const bool active = x < a.W && y < a.H;
    const size_t n = (size_t)a.H * a.W;
    const size_t pix = active ? (size_t)y * a.W + x : 0;
    float cx[4], cy[4];
    int ix[4], iy[4];
    if (LOWRES) {
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
    float p1 = 0.f, p2 = 0.f;
    if (active && !a.first) { p1 = a.hist1[pix]; p2 = a.hist2[pix]; }
    for (int t = 0; t < a.B; ++t) {
        uint32_t enc = 0;
        bool nan = false;
        if (active) {
            float cur;
            if (LOWRES) {
                const __half *src = a.lowres + (size_t)t * a.h * a.w;
                float rows[4];
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    rows[k] = Cubic<CONTRACT>::dot4(h2f(__ldg(src + iy[k] + ix[0])), h2f(__ldg(src + iy[k] + ix[1])),
                                                    h2f(__ldg(src + iy[k] + ix[2])), h2f(__ldg(src + iy[k] + ix[3])), cx);
                cur = Cubic<CONTRACT>::dot4(rows[0], rows[1], rows[2], rows[3], cy);
                if (a.scaler != 1.0f) cur = __fmul_rn(cur, a.scaler);
            } else {
                cur = __ldg(a.raw + (size_t)t * n + pix);
            }
            if (a.first && t == 0) p1 = p2 = cur;
            const float res = smooth3_f32(cur, p1, p2, a.sw.w_now, a.sw.w_prev1, a.sw.w_prev2);
            a.out[(size_t)t * n + pix] = res;
            if (res != res) nan = true; else enc = f2ord(res);
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

// ---- fp32 depth pass, vectorised: the fp32 counterpart of k_depth_pass<true> -------------------------------------
// Full-resolution fp32 raw depth in, smoothed fp32 depth + per-frame max out; 4 pixels (one 16-byte access) per thread and
// frame, four frames of loads in flight, history in registers.  Same arithmetic as k_depth_f32 (one fp32 rounding per op).
__global__ void __launch_bounds__(256) k_depth_pass_f32(DepthArgs32 a) {
    extern __shared__ uint32_t s_red[];
    uint32_t *s_max = s_red, *s_nan = s_red + a.B;
    pdl_launch_dependents();
    for (int i = threadIdx.x; i < 2 * a.B; i += blockDim.x) s_red[i] = 0;
    __syncthreads();
    const size_t n = (size_t)a.H * a.W;
    const size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool active = v * 4 < n;
    const size_t base = active ? v * 4 : 0;
    const float w0 = a.sw.w_now, w1 = a.sw.w_prev1, w2 = a.sw.w_prev2;
    float4 p1 = make_float4(0.f, 0.f, 0.f, 0.f), p2 = p1;
    if (active && !a.first) {
        p1 = __ldg(reinterpret_cast<const float4 *>(a.hist1 + base));
        p2 = __ldg(reinterpret_cast<const float4 *>(a.hist2 + base));
    }
    auto load = [&](int t) -> float4 {
        return (active && t < a.B) ? __ldg(reinterpret_cast<const float4 *>(a.raw + (size_t)t * n + base)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    constexpr int U = 4;
    for (int tb = 0; tb < a.B; tb += U) {
        float4 q[U];
#pragma unroll
        for (int i = 0; i < U; ++i) q[i] = load(tb + i);
#pragma unroll
        for (int i = 0; i < U; ++i) {
            const int t = tb + i;
            if (t >= a.B) break;
            uint32_t enc = 0;
            bool nan = false;
            if (active) {
                const float4 cur = q[i];
                if (a.first && t == 0) p1 = p2 = cur;
                float4 r;
                r.x = smooth3_f32(cur.x, p1.x, p2.x, w0, w1, w2);
                r.y = smooth3_f32(cur.y, p1.y, p2.y, w0, w1, w2);
                r.z = smooth3_f32(cur.z, p1.z, p2.z, w0, w1, w2);
                r.w = smooth3_f32(cur.w, p1.w, p2.w, w0, w1, w2);
                *reinterpret_cast<float4 *>(a.out + (size_t)t * n + base) = r;
                nan = (r.x != r.x) || (r.y != r.y) || (r.z != r.z) || (r.w != r.w);
                if (!nan) enc = f2ord(fmaxf(fmaxf(r.x, r.y), fmaxf(r.z, r.w)));
                p2 = p1;
                p1 = cur;
            }
            frame_max_commit(enc, nan, t, s_max, s_nan);
        }
    }
    if (active) {
        *reinterpret_cast<float4 *>(a.hist1 + base) = p1;
        *reinterpret_cast<float4 *>(a.hist2 + base) = p2;
    }
    __syncthreads();
    for (int t = threadIdx.x; t < a.B; t += blockDim.x) {
        if (s_max[t]) atomicMax(&a.frame_max[t], s_max[t]);
        if (s_nan[t]) atomicOr(&a.frame_nan[t], 1u);
    }
}

// ---- low-res DPT output in, tiled: horizontal interpolations shared by the output rows that use them --------
// Same arithmetic as k_depth_lowres (ATen's: four row interpolations, then one column interpolation, every
// mul/add in the same order), but a CTA owns a 64 x 16 output tile and first computes the row interpolation
// R[iy][x] once per (input row, output column) into shared memory: upscaling by ~2.1 means each R is used by ~8
// output rows, so the 16 gathers + 16 FMAs per output pixel shrink to ~3 + 4 LDS + 8.  One thread owns 2 x 2
// output pixels (two adjacent columns, rows ty and ty + 8) for all frames of the batch, history in registers.
constexpr int kLrTileW = 64, kLrTileH = 16;

// FULL: the whole 64 x 16 tile lies inside the frame (all but the last row of tiles at 1080p): no per-pixel bounds predicates
template <bool CONTRACT, bool INTERIOR, bool FULL = false>
__device__ __forceinline__ void depth_lowres_tile(const DepthArgs &a, int rmax, int cmax, uint8_t *lr_smem) {
    float *R = reinterpret_cast<float *>(lr_smem);                          // [rmax][64] row interpolations
    float *Tin = R + rmax * kLrTileW;                                       // [2][rmax][cmax] staged input tile (as fp32)
    uint32_t *s_key = reinterpret_cast<uint32_t *>(Tin + 2 * rmax * cmax);  // [B] 17-bit max keys (h16_key)

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tx0 = blockIdx.x * kLrTileW, x0 = tx0 + 2 * lane, ty0 = blockIdx.y * kLrTileH;
    const size_t n = (size_t)a.H * a.W;

    // input region of this tile: rows [rbase, rbase + nrows), columns [cbase, cbase + ncols)
    const int xlast = min(tx0 + kLrTileW - 1, a.W - 1), ylast = min(ty0 + kLrTileH - 1, a.H - 1);
    const int cbase = max(min((int)floorf(__fmul_rn(a.scale_x, (float)tx0)) - 1, a.w - 1), 0);
    const int cend = max(min((int)floorf(__fmul_rn(a.scale_x, (float)xlast)) + 2, a.w - 1), 0);
    const int rbase = max(min((int)floorf(__fmul_rn(a.scale_y, (float)ty0)) - 1, a.h - 1), 0);
    const int rend = max(min((int)floorf(__fmul_rn(a.scale_y, (float)ylast)) + 2, a.h - 1), 0);
    const int nrows = rend - rbase + 1, ncols = cend - cbase + 1;          // <= rmax, cmax (host)

    // columns x0, x0+1: coefficients and tap columns relative to cbase (INTERIOR: taps are ix[c][0] + k)
    float cx[2][4];
    int ix[2][INTERIOR ? 1 : 4];
    bool xin[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const int x = x0 + c;
        xin[c] = x < a.W;
        const float rx = __fmul_rn(a.scale_x, (float)min(x, a.W - 1));
        const int fx = (int)floorf(rx);
        Cubic<CONTRACT>::coeffs(rx - (float)fx, cx[c]);
#pragma unroll
        for (int k = 0; k < (INTERIOR ? 1 : 4); ++k) ix[c][k] = max(min(fx - 1 + k, a.w - 1), 0) - cbase;
    }
    // rows ty0 + warp, ty0 + warp + 8: coefficients and R offsets (INTERIOR: iy[j][0] + 64 k)
    float cy[2][4];
    int iy[2][INTERIOR ? 1 : 4];
    bool yin[2];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
        const int y = ty0 + warp + 8 * j;
        yin[j] = y < a.H;
        const float ry = __fmul_rn(a.scale_y, (float)min(y, a.H - 1));
        const int fy = (int)floorf(ry);
        Cubic<CONTRACT>::coeffs(ry - (float)fy, cy[j]);
#pragma unroll
        for (int k = 0; k < (INTERIOR ? 1 : 4); ++k) iy[j][k] = (max(min(fy - 1 + k, a.h - 1), 0) - rbase) * kLrTileW + 2 * lane;
    }
    // my (at most two) elements of the input tile
    int e_src[2], e_dst[2];
    bool e_on[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const int e = threadIdx.x + 256 * q;
        e_on[q] = e < nrows * ncols;
        const int r = e_on[q] ? e / ncols : 0, col = e_on[q] ? e - r * ncols : 0;
        e_src[q] = (rbase + r) * a.w + cbase + col;
        e_dst[q] = r * cmax + col;
    }
    const bool pair_ok = FULL || (xin[0] && xin[1]);
    auto pix = [&](int j) { return (size_t)(ty0 + warp + 8 * j) * a.W + x0; };
    auto ld2 = [&](const __half *p, size_t i) -> float2 {
        if (pair_ok) return __half22float2(*reinterpret_cast<const __half2 *>(p + i));
        return make_float2(__half2float(p[i]), 0.f);
    };
    auto st2 = [&](__half *p, size_t i, __half2 v) {
        if (pair_ok) *reinterpret_cast<__half2 *>(p + i) = v;
        else p[i] = __low2half(v);
    };
    // raw depth of t-1 / t-2 at my 2 x 2 pixels, as floats (exact fp16 values)
    float2 p1[2], p2[2];
    p1[0] = p1[1] = p2[0] = p2[1] = make_float2(0.f, 0.f);
    const bool on[2] = {FULL || (yin[0] && xin[0]), FULL || (yin[1] && xin[0])};
    if (!a.first) {
#pragma unroll
        for (int j = 0; j < 2; ++j)
            if (on[j]) { p1[j] = ld2(a.hist1, pix(j)); p2[j] = ld2(a.hist2, pix(j)); }
    }
    const float w0 = a.sw.w_now, w1 = a.sw.w_prev1, w2 = a.sw.w_prev2;
    const bool scale = a.scaler != 1.0f;
    bool init = a.first != 0;
    __half *outp[2] = {a.out + pix(0), a.out + pix(1)};

    __half nxt[2];                                                          // tile of frame t, loaded one frame ahead
#pragma unroll
    for (int q = 0; q < 2; ++q) nxt[q] = e_on[q] ? __ldg(a.lowres + e_src[q]) : __float2half(0.f);
    const __half *src = a.lowres;

    for (int t = 0; t < a.B; ++t) {
        float *Tb = Tin + (t & 1) * rmax * cmax;
#pragma unroll
        for (int q = 0; q < 2; ++q)
            if (e_on[q]) Tb[e_dst[q]] = h2f(nxt[q]);
        src += (size_t)a.h * a.w;
        if (t + 1 < a.B) {
#pragma unroll
            for (int q = 0; q < 2; ++q)
                if (e_on[q]) nxt[q] = __ldg(src + e_src[q]);
        }
        __syncthreads();
        // phase A: row interpolations of the tile's input rows at my two columns
        for (int r = warp; r < nrows; r += 8) {
            const float *row = Tb + r * cmax;
            float2 v;
            if (INTERIOR) {
                const float *q0 = row + ix[0][0], *q1 = row + ix[1][0];
                v.x = Cubic<CONTRACT>::dot4(q0[0], q0[1], q0[2], q0[3], cx[0]);
                v.y = Cubic<CONTRACT>::dot4(q1[0], q1[1], q1[2], q1[3], cx[1]);
            } else {
                v.x = Cubic<CONTRACT>::dot4(row[ix[0][0]], row[ix[0][INTERIOR ? 0 : 1]], row[ix[0][INTERIOR ? 0 : 2]], row[ix[0][INTERIOR ? 0 : 3]], cx[0]);
                v.y = Cubic<CONTRACT>::dot4(row[ix[1][0]], row[ix[1][INTERIOR ? 0 : 1]], row[ix[1][INTERIOR ? 0 : 2]], row[ix[1][INTERIOR ? 0 : 3]], cx[1]);
            }
            *reinterpret_cast<float2 *>(R + r * kLrTileW + 2 * lane) = v;
        }
        __syncthreads();
        // phase B: column interpolation, scaler, smoothing, max, store
        __half2 res[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
            float2 r0, r1, r2, r3;
            if (INTERIOR) {
                const float *q = R + iy[j][0];
                r0 = *reinterpret_cast<const float2 *>(q);                  r1 = *reinterpret_cast<const float2 *>(q + kLrTileW);
                r2 = *reinterpret_cast<const float2 *>(q + 2 * kLrTileW);   r3 = *reinterpret_cast<const float2 *>(q + 3 * kLrTileW);
            } else {
                r0 = *reinterpret_cast<const float2 *>(R + iy[j][0]);                  r1 = *reinterpret_cast<const float2 *>(R + iy[j][INTERIOR ? 0 : 1]);
                r2 = *reinterpret_cast<const float2 *>(R + iy[j][INTERIOR ? 0 : 2]);   r3 = *reinterpret_cast<const float2 *>(R + iy[j][INTERIOR ? 0 : 3]);
            }
            __half2 curh = __floats2half2_rn(Cubic<CONTRACT>::dot4(r0.x, r1.x, r2.x, r3.x, cy[j]), Cubic<CONTRACT>::dot4(r0.y, r1.y, r2.y, r3.y, cy[j]));
            float2 cur = __half22float2(curh);
            if (scale) {                                                    // `* scaler` in fp16 (PredictAndGenerate.py:55)
                curh = __floats2half2_rn(__fmul_rn(cur.x, a.scaler), __fmul_rn(cur.y, a.scaler));
                cur = __half22float2(curh);
            }
            if (init) { p1[j] = cur; p2[j] = cur; }
            __half2 d = __hadd2(__floats2half2_rn(__fmul_rn(cur.x, w0), __fmul_rn(cur.y, w0)),
                                __floats2half2_rn(__fmul_rn(p1[j].x, w1), __fmul_rn(p1[j].y, w1)));
            d = __hadd2(d, __floats2half2_rn(__fmul_rn(p2[j].x, w2), __fmul_rn(p2[j].y, w2)));
            res[j] = d;
            if (on[j]) st2(outp[j], 0, d);
            outp[j] += n;
            p2[j] = p1[j];
            p1[j] = cur;
        }
        init = false;
        // per-frame max of my pixels as an order-preserving 17-bit key (NaN on top)
        uint32_t key = 0;
        {
            if (!pair_ok) { res[0] = __low2half2(res[0]); res[1] = __low2half2(res[1]); }
            __half2 m = on[0] ? res[0] : res[1];
            if (on[0] && on[1]) m = __hmax2_nan(res[0], res[1]);
            m = __hmax2_nan(m, __lowhigh2highlow(m));
            if (on[0] || on[1]) key = h16_key((uint32_t)__half_as_ushort(__low2half(m)));
        }
        key = __reduce_max_sync(0xffffffffu, key);
        if (lane == 0 && key) atomicMax(&s_key[t], key);
    }
#pragma unroll
    for (int j = 0; j < 2; ++j)
        if (on[j]) {
            st2(a.hist1, pix(j), __floats2half2_rn(p1[j].x, p1[j].y));
            st2(a.hist2, pix(j), __floats2half2_rn(p2[j].x, p2[j].y));
        }
}

template <bool CONTRACT>
__global__ void __launch_bounds__(256, 3) k_depth_lowres_tiled(DepthArgs a, int rmax, int cmax) {
    extern __shared__ __align__(16) uint8_t lr_smem[];
    uint32_t *s_key = reinterpret_cast<uint32_t *>(reinterpret_cast<float *>(lr_smem) + rmax * kLrTileW + 2 * rmax * cmax);
    for (int i = threadIdx.x; i < a.B; i += blockDim.x) s_key[i] = 0;
    __syncthreads();
    // a tile is interior when none of its taps is clamped at the image border: the four taps are then consecutive
    const int tx0 = blockIdx.x * kLrTileW, ty0 = blockIdx.y * kLrTileH;
    const int xlast = min(tx0 + kLrTileW - 1, a.W - 1), ylast = min(ty0 + kLrTileH - 1, a.H - 1);
    const bool interior = (int)floorf(__fmul_rn(a.scale_x, (float)tx0)) >= 1 && (int)floorf(__fmul_rn(a.scale_x, (float)xlast)) + 2 <= a.w - 1 &&
                          (int)floorf(__fmul_rn(a.scale_y, (float)ty0)) >= 1 && (int)floorf(__fmul_rn(a.scale_y, (float)ylast)) + 2 <= a.h - 1;
    const bool full = tx0 + kLrTileW <= a.W && ty0 + kLrTileH <= a.H;
    if (interior && full) depth_lowres_tile<CONTRACT, true, true>(a, rmax, cmax, lr_smem);
    else if (interior) depth_lowres_tile<CONTRACT, true>(a, rmax, cmax, lr_smem);
    else depth_lowres_tile<CONTRACT, false>(a, rmax, cmax, lr_smem);
    __syncthreads();
    for (int t = threadIdx.x; t < a.B; t += blockDim.x) {
        const uint32_t key = s_key[t];
        if (key == 0x1ffffu) atomicOr(&a.frame_nan[t], 1u);
        else if (key) {
            const uint32_t u = (key & 0x8000u) ? (key & 0x7fffu) : (~key & 0xffffu);
            atomicMax(&a.frame_max[t], f2ord(__half2float(__ushort_as_half((unsigned short)u))));
        }
    }
}

}  // namespace vrsbs
