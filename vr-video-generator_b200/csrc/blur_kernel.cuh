// Stage 3b: hole blur and strip restore (PredictAndGenerate.py:191-196).
//
// The reference blurs the whole filled frame with a (2k+1)x(2k+3) fp32 gaussian (torchvision
// gaussian_blur: reflect padding, one depthwise conv, round_) and keeps the result only at hole
// pixels.  Here only hole pixels are evaluated, tile by tile, and only tiles that contain holes are
// visited: the warp kernel appends every 16x64 tile in which it found a hole to a work list.
//
// The pre-blur image is reconstructed while a tile (+ halo) is staged in shared memory:
//     hole neighbour     -> img[y, (x - fill_off) mod W]   (what the hole fill wrote, :190)
//     painted neighbour  -> the view the warp kernel already stored in the SBS frame
// so blurred pixels can be written in place (a hole's stored value is never read by anyone).  The
// strip columns [0,strip) keep their pre-blur values until k_strip_restore runs, because holes right
// of the strip need them as neighbours.
//
// Accumulation is fp64 FMA over the fp32 weights: exact for u8 pixels (no summation-order
// dependence), then round-half-even like round_().  See DESIGN.md "blur parity".
#pragma once
#include "common.cuh"

namespace vrsbs {

constexpr int kTileH = 16, kTileW = 64;       // blur work unit; kTileW is a multiple of 32 (mask words)

struct BlurArgs {
    const uint8_t *frames;       // [B,H,W,3]
    uint8_t *sbs;                // [B,H,2W,3]
    const FrameTab *tabs;        // [B]
    const uint32_t *hole_mask;   // [B][H][Wwords]
    const float *weights;        // [ky][kx]
    const uint32_t *tile_list;   // tiles that contain at least one hole (linear tile ids)
    const uint32_t *tile_count;  // number of entries in tile_list
    int B, H, W, Wwords, kx, ky, tiles_x, tiles_y;
};

__host__ __device__ inline size_t blur_smem_bytes(int kx, int ky) {
    const size_t sw = kTileW + kx - 1, sh = kTileH + ky - 1;
    return sizeof(double) * kx * ky + sizeof(float) * 3 * sw * sh + sizeof(uint16_t) * kTileH * kTileW + 16;
}

__global__ void __launch_bounds__(256) k_blur_tiles(BlurArgs a) {
    extern __shared__ __align__(16) uint8_t blur_smem[];
    const int kx = a.kx, ky = a.ky, W = a.W, H = a.H;
    const int sw = kTileW + kx - 1, sh = kTileH + ky - 1, plane = sw * sh;
    double *s_w = reinterpret_cast<double *>(blur_smem);
    float *s_px = reinterpret_cast<float *>(s_w + kx * ky);              // [3][sh][sw]
    uint16_t *s_list = reinterpret_cast<uint16_t *>(s_px + 3 * plane);   // (ly << 8) | lx
    __shared__ int s_n;

    for (int i = threadIdx.x; i < kx * ky; i += blockDim.x) s_w[i] = (double)a.weights[i];
    const uint32_t ntiles = *a.tile_count;

    for (uint32_t ti = blockIdx.x; ti < ntiles; ti += gridDim.x) {
        const uint32_t tile = a.tile_list[ti];
        const int tx = tile % a.tiles_x, ty = (tile / a.tiles_x) % a.tiles_y, b = tile / (a.tiles_x * a.tiles_y);
        const int x0 = tx * kTileW, y0 = ty * kTileH;
        const FrameTab *t = a.tabs + b;
        const int fill = t->fill_off, strip = t->strip;
        const size_t frame_row0 = (size_t)b * H;
        __syncthreads();                       // previous tile fully consumed (also orders s_w on the first pass)
        if (threadIdx.x == 0) s_n = 0;
        // ---- stage the pre-blur tile + halo as floats, channel-planar ----
        for (int p = threadIdx.x; p < plane; p += blockDim.x) {
            const int ly = p / sw, lx = p - ly * sw;
            const int yy = reflect_idx(y0 + ly - ky / 2, H), xx = reflect_idx(x0 + lx - kx / 2, W);
            const size_t rb = frame_row0 + yy;
            const bool hole = (a.hole_mask[rb * a.Wwords + (xx >> 5)] >> (xx & 31)) & 1u;
            int xs = xx - fill;
            xs += (xs < 0) ? W : 0;
            const uint8_t *src = hole ? a.frames + (rb * W + xs) * 3 : a.sbs + (rb * 2 * W + xx) * 3;
            s_px[p] = (float)src[0];
            s_px[plane + p] = (float)src[1];
            s_px[2 * plane + p] = (float)src[2];
        }
        __syncthreads();
        // ---- list the hole pixels of this tile (right of the strip, inside the frame) ----
        for (int q = threadIdx.x; q < kTileH * (kTileW / 8); q += blockDim.x) {
            const int ly = q / (kTileW / 8), cx = (q % (kTileW / 8)) * 8;
            const int y = y0 + ly, x = x0 + cx;
            if (y >= H || x >= W) continue;
            uint32_t bits = (a.hole_mask[(frame_row0 + y) * a.Wwords + (x >> 5)] >> (x & 31)) & 0xffu;
            while (bits) {
                const int k = __ffs(bits) - 1;
                bits &= bits - 1;
                if (x + k >= strip && x + k < W) s_list[atomicAdd(&s_n, 1)] = (uint16_t)((ly << 8) | (cx + k));
            }
        }
        __syncthreads();
        const int n = s_n;
        // ---- one thread per hole pixel: ky*kx taps x 3 channels, exact fp64 accumulation ----
        for (int h = threadIdx.x; h < n; h += blockDim.x) {
            const int ly = s_list[h] >> 8, lx = s_list[h] & 0xff;
            const float *p0 = s_px + ly * sw + lx;
            double a0 = 0.0, a1 = 0.0, a2 = 0.0;
            for (int i = 0; i < ky; ++i) {
                const float *row = p0 + i * sw;
                const double *wrow = s_w + i * kx;
                for (int j = 0; j < kx; ++j) {
                    const double w = wrow[j];
                    a0 = fma(w, (double)row[j], a0);
                    a1 = fma(w, (double)row[plane + j], a1);
                    a2 = fma(w, (double)row[2 * plane + j], a2);
                }
            }
            uint8_t *o = a.sbs + ((frame_row0 + y0 + ly) * 2 * W + x0 + lx) * 3;
            o[0] = (uint8_t)__double2int_rn(a0);
            o[1] = (uint8_t)__double2int_rn(a1);
            o[2] = (uint8_t)__double2int_rn(a2);
        }
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
