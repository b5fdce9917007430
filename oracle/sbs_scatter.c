/*
 * CPU oracle #2 — the SBS warp restated in forward-scatter form, plain C + OpenMP.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Built by oracle/Makefile into
 * oracle/_build/libsbs_oracle.so and loaded with ctypes from oracle/scatter.py.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use it.
 *
 * Reference being restated (read-only /root/reference, nothing copied):
 *   PredictAndGenerate.py:139-142  temporal smoothing, fp16 with per-op rounding
 *   PredictAndGenerate.py:102      per-frame max
 *   PredictAndGenerate.py:169-183  layer loop == "every source pixel that belongs to layer k lands
 *                                  at (x + off_k) mod W; the highest k wins" (probed)
 *   PredictAndGenerate.py:184-190  holes <- roll(img, off_f), f = int(L*3/5)
 *   PredictAndGenerate.py:191-194  holes <- round(gaussian(filled image)), reflect border
 *   PredictAndGenerate.py:196-197  left strip restored from the input, [view | input] packed
 *
 * The equivalence of this scatter form with the layer loop is not assumed: tests/test_oracle_*.py
 * checks it bit-for-bit against oracle/sbs_layered.py (the literal restatement) and against the
 * reference-generated fixtures in tests/golden/.  Parity status: pinned (through those fixtures).
 *
 * Membership is evaluated by brute force over all layers (no monotonicity assumption), so this
 * file is also the judge for the CUDA kernel's binary-search + walk shortcut.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>

typedef _Float16 f16;

static inline float h2f(uint16_t bits) { f16 h; memcpy(&h, &bits, 2); return (float)h; }
static inline uint16_t f2h(float f)    { f16 h = (f16)f; uint16_t b; memcpy(&b, &h, 2); return b; }

/* PredictAndGenerate.py:139-142: d = raw*w_now; d += h1*w1; d += h0*w0, each op fp32 then ->fp16 */
void sbs_oracle_smooth_f16(const uint16_t *raw, const uint16_t *h1, const uint16_t *h0,
                           uint16_t *out, size_t n, float w_now, float w1, float w0)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        uint16_t d = f2h(h2f(raw[i]) * w_now);
        uint16_t t = f2h(h2f(h1[i]) * w1);
        d = f2h(h2f(d) + h2f(t));
        t = f2h(h2f(h0[i]) * w0);
        d = f2h(h2f(d) + h2f(t));
        out[i] = d;
    }
}

/* PredictAndGenerate.py:102: depth.max(); returns NaN if any element is NaN (torch propagates) */
float sbs_oracle_max_f16(const uint16_t *d, size_t n)
{
    float m = -INFINITY;
    int nan = 0;
#pragma omp parallel for reduction(max : m) reduction(| : nan) schedule(static)
    for (size_t i = 0; i < n; ++i) {
        float v = h2f(d[i]);
        if (v != v) nan = 1;
        else if (v > m) m = v;
    }
    return nan ? NAN : m;
}

static inline int wrap(int x, int W) { x %= W; return x < 0 ? x + W : x; }
static inline int reflect(int i, int n) { if (i < 0) i = -i; if (i >= n) i = 2 * (n - 1) - i; return i; }

/*
 * One frame.  img [H,W,3] u8, depth [H,W] fp16 bits (already smoothed), per-layer bounds lo/hi as
 * fp16 bits (already narrowed double->float->half), integer offsets off[L].
 * Outputs: sbs [H,2W,3]; optional winner [H,W] (layer index that painted the pixel, -1 = hole);
 * optional pre_blur [H,W,3] (view after hole fill, before blur and strip).
 * Returns the number of hole pixels, or -1 on bad arguments.
 */
static long warp_core(const uint8_t *img, const float *depthf, const uint16_t *depth16, int H, int W, int L,
                      const float *lo, const float *hi, const int *off,
                      int fill_layer, int strip, int kx, int ky, const float *weights,
                      uint8_t *sbs, int16_t *winner_out, uint8_t *pre_blur_out)
{
    if (H <= 0 || W <= 0 || L <= 0 || L > 32767 || fill_layer < 0 || fill_layer >= L) return -1;
    int16_t *winner = winner_out ? winner_out : malloc(sizeof(int16_t) * (size_t)H * W);
    uint8_t *pre = pre_blur_out ? pre_blur_out : malloc((size_t)H * W * 3);
    const size_t pitch = (size_t)W * 6;
    long holes = 0;

#pragma omp parallel for schedule(static) reduction(+ : holes)
    for (int y = 0; y < H; ++y) {
        int16_t *win = winner + (size_t)y * W;
        const uint8_t *src = img + (size_t)y * W * 3;
        uint8_t *p = pre + (size_t)y * W * 3;
        for (int x = 0; x < W; ++x) win[x] = -1;
        for (int x = 0; x < W; ++x) {
            /* the comparison runs in the depth dtype (:173): fp16 values and fp16-rounded bounds, or fp32 and fp32 */
            float d = depthf ? depthf[(size_t)y * W + x] : h2f(depth16[(size_t)y * W + x]);
            for (int k = 0; k < L; ++k)
                if (lo[k] <= d && d < hi[k]) {
                    int xd = wrap(x + off[k], W);
                    if (k > win[xd]) win[xd] = (int16_t)k;
                }
        }
        for (int x = 0; x < W; ++x) {
            int k = win[x];
            if (k < 0) ++holes;
            int xs = wrap(x - off[k < 0 ? fill_layer : k], W);
            p[3 * x] = src[3 * xs]; p[3 * x + 1] = src[3 * xs + 1]; p[3 * x + 2] = src[3 * xs + 2];
        }
    }

#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < H; ++y) {
        const int16_t *win = winner + (size_t)y * W;
        uint8_t *out = sbs + (size_t)y * pitch;
        const uint8_t *src = img + (size_t)y * W * 3;
        memcpy(out, pre + (size_t)y * W * 3, (size_t)W * 3);
        for (int x = 0; x < W; ++x) {
            if (win[x] >= 0) continue;
            double acc[3] = {0, 0, 0};
            for (int i = 0; i < ky; ++i) {
                const uint8_t *row = pre + (size_t)reflect(y + i - ky / 2, H) * W * 3;
                for (int j = 0; j < kx; ++j) {
                    const uint8_t *q = row + 3 * reflect(x + j - kx / 2, W);
                    double w = weights[i * kx + j];
                    acc[0] += w * q[0]; acc[1] += w * q[1]; acc[2] += w * q[2];
                }
            }
            for (int c = 0; c < 3; ++c) out[3 * x + c] = (uint8_t)nearbyint(acc[c]);
        }
        if (strip > 0) memcpy(out, src, (size_t)(strip < W ? strip : W) * 3);
        memcpy(out + (size_t)W * 3, src, (size_t)W * 3);
    }
    if (!winner_out) free(winner);
    if (!pre_blur_out) free(pre);
    return holes;
}

long sbs_oracle_warp_frame(const uint8_t *img, const uint16_t *depth, int H, int W, int L,
                           const uint16_t *lo16, const uint16_t *hi16, const int *off,
                           int fill_layer, int strip, int kx, int ky, const float *weights,
                           uint8_t *sbs, int16_t *winner_out, uint8_t *pre_blur_out)
{
    if (L <= 0 || L > 32767) return -1;
    float *lo = malloc(sizeof(float) * L), *hi = malloc(sizeof(float) * L);
    for (int k = 0; k < L; ++k) { lo[k] = h2f(lo16[k]); hi[k] = h2f(hi16[k]); }
    long r = warp_core(img, NULL, depth, H, W, L, lo, hi, off, fill_layer, strip, kx, ky, weights, sbs, winner_out, pre_blur_out);
    free(lo); free(hi);
    return r;
}

/* fp32 depth (what torch >= 2.4's CUDA autocast hands the warp: upsample_bicubic2d is on autocast's fp32 list):
 * smoothing in fp32 (:139-142 with an fp32 tensor), bounds narrowed double -> float, comparison in fp32 (:173). */
void sbs_oracle_smooth_f32(const float *raw, const float *h1, const float *h0, float *out, size_t n,
                           float w_now, float w1, float w0)
{
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; ++i) {
        float d = raw[i] * w_now;
        float t = h1[i] * w1;
        d = d + t;
        t = h0[i] * w0;
        out[i] = d + t;
    }
}

float sbs_oracle_max_f32(const float *d, size_t n)
{
    float m = -INFINITY;
    int nan = 0;
#pragma omp parallel for reduction(max : m) reduction(| : nan) schedule(static)
    for (size_t i = 0; i < n; ++i) {
        float v = d[i];
        if (v != v) nan = 1;
        else if (v > m) m = v;
    }
    return nan ? NAN : m;
}

long sbs_oracle_warp_frame_f32(const uint8_t *img, const float *depth, int H, int W, int L,
                               const float *lo, const float *hi, const int *off,
                               int fill_layer, int strip, int kx, int ky, const float *weights,
                               uint8_t *sbs, int16_t *winner_out, uint8_t *pre_blur_out)
{
    return warp_core(img, depth, NULL, H, W, L, lo, hi, off, fill_layer, strip, kx, ky, weights, sbs, winner_out, pre_blur_out);
}

int sbs_oracle_abi_version(void) { return 2; }
