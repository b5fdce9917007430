// Stage 3b: hole blur and strip restore (PredictAndGenerate.py:191-196).
//
// The reference blurs the whole filled frame with a (2k+1)x(2k+3) fp32 gaussian (torchvision
// gaussian_blur: reflect padding, one depthwise conv, round_) and keeps the result only at hole
// pixels.  Here only hole pixels are evaluated.  The pre-blur image is reconstructed on the fly:
//     hole neighbour     -> img[y, (x - fill_off) mod W]   (what the hole fill wrote, :190)
//     painted neighbour  -> the view the warp kernel already stored in the SBS frame
// so blurred pixels can be written in place (a hole's stored value is never read).  The strip
// columns [0,strip) keep their pre-blur values until k_strip_restore runs, because holes right of
// the strip need them as neighbours.
//
// Accumulation is fp64 FMA over the fp32 weights: exact for u8 pixels (no summation-order
// dependence), then round-half-even like round_().  See DESIGN.md "blur parity".
#pragma once
#include "common.cuh"

namespace vrsbs {

struct BlurArgs {
    const uint8_t *frames;       // [B,H,W,3]
    uint8_t *sbs;                // [B,H,2W,3]
    const FrameTab *tabs;        // [B]
    const uint32_t *hole_mask;   // [B][H][Wwords]
    const float *weights;        // [ky][kx]
    int B, H, W, Wwords, kx, ky;
};

__global__ void __launch_bounds__(256) k_blur_holes(BlurArgs a) {
    extern __shared__ double s_w[];                          // [ky*kx]
    for (int i = threadIdx.x; i < a.kx * a.ky; i += blockDim.x) s_w[i] = (double)a.weights[i];
    __syncthreads();

    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long total = (long long)a.B * a.H * a.Wwords;
    const int W = a.W, H = a.H;
    for (long long g = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); g < total; g += nwarps) {
        const uint32_t m = a.hole_mask[g];
        if (m == 0) continue;
        const int wx = (int)(g % a.Wwords);
        const long long by = g / a.Wwords;
        const int y = (int)(by % H), b = (int)(by / H);
        const int x = wx * 32 + lane;
        const FrameTab *t = a.tabs + b;
        if (!((m >> lane) & 1u) || x < t->strip || x >= W) continue;
        const int fill = t->fill_off;
        double acc0 = 0.0, acc1 = 0.0, acc2 = 0.0;
        for (int i = 0; i < a.ky; ++i) {
            const int yy = reflect_idx(y + i - a.ky / 2, H);
            const size_t rb = (size_t)b * H + yy;
            const uint32_t *mrow = a.hole_mask + rb * a.Wwords;
            const uint8_t *view = a.sbs + rb * (size_t)W * 6;
            const uint8_t *src = a.frames + rb * (size_t)W * 3;
            for (int j = 0; j < a.kx; ++j) {
                const int xx = reflect_idx(x + j - a.kx / 2, W);
                const bool hole = (mrow[xx >> 5] >> (xx & 31)) & 1u;
                int xs = xx - fill;
                xs += (xs < 0) ? W : 0;
                const uint8_t *p = hole ? src + 3 * xs : view + 3 * xx;
                const double w = s_w[i * a.kx + j];
                acc0 = fma(w, (double)p[0], acc0);
                acc1 = fma(w, (double)p[1], acc1);
                acc2 = fma(w, (double)p[2], acc2);
            }
        }
        uint8_t *o = a.sbs + ((size_t)b * H + y) * (size_t)W * 6 + 3 * x;
        o[0] = (uint8_t)__double2int_rn(acc0);
        o[1] = (uint8_t)__double2int_rn(acc1);
        o[2] = (uint8_t)__double2int_rn(acc2);
    }
}

// result_img[:, 0:strip] = img[:, 0:strip]  (PredictAndGenerate.py:196); one warp per image row.
__global__ void __launch_bounds__(256) k_strip_restore(BlurArgs a) {
    const int lane = threadIdx.x & 31;
    const long long nwarps = (long long)gridDim.x * (blockDim.x >> 5);
    const long long rows = (long long)a.B * a.H;
    for (long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < rows; r += nwarps) {
        const int nbytes = a.tabs[r / a.H].strip * 3;
        const uint8_t *src = a.frames + r * (size_t)a.W * 3;
        uint8_t *dst = a.sbs + r * (size_t)a.W * 6;
        for (int c = lane; c < nbytes; c += 32) dst[c] = src[c];
    }
}

}  // namespace vrsbs
