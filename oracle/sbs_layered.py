"""CPU oracle #1 — the layered painter's-order SBS warp, restated in numpy.

TEST INFRASTRUCTURE ONLY.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import this file; the product path
(`vr-video-generator_b200/`) never does and fails loudly without its CUDA extension.

What it restates (reference = /root/reference, read-only, never copied):
  * temporal depth smoothing ............ PredictAndGenerate.py:131-145 (`SbsProcessor.get_depth`)
  * cutoff / step / offset tables ....... PredictAndGenerate.py:101-126 (`SbsProcessor.get_cutoff`)
  * layer loop (bin mask, roll, paint) .. PredictAndGenerate.py:169-183
  * hole fill, hole blur ................ PredictAndGenerate.py:184-194 + torchvision 0.26
                                          `transforms/v2/functional/_misc.py:86-165`
  * strip restore + SBS pack ............ PredictAndGenerate.py:196-197
  * bicubic depth tail + scaler ......... depth_anything_v2/dpt.py:196-199, PredictAndGenerate.py:27-34,55

Parity status: PINNED.  The reference ships no tests or golden vectors, so the pin
is the reference itself executed in the build container (oracle/ref_driver.py) — see
tests/golden/make_golden.py for the generator and tests/test_oracle_golden.py for the check.
Integer/byte/index stages are bit-exact against it.  The one floating-point stage whose result
depends on summation order is the hole blur: the reference accumulates 99..323 fp32 products in
whatever order the conv backend picks; this oracle accumulates the same fp32 weights in float64,
which is EXACT for u8 pixels (24-bit weight x 8-bit pixel, < 2^9 terms, narrow exponent range),
then rounds half-to-even like `round_()`.  Versus the reference's CPU conv that differs by 1 LSB
on ~1e-5 of blurred pixels (recorded per fixture).

Third-party arithmetic this depends on (not in /root/reference): torch 2.11.0 fp16 elementwise
semantics (scalar operands are applied in fp32 opmath and the result is rounded to fp16; python
float thresholds compared against an fp16 tensor are first narrowed double->float->half) and
torchvision 0.26.0 `gaussian_blur` (softmax-normalised 1-D kernels, outer product in fp32,
reflect padding, one depthwise conv, `round_`).  The reference pins torch==2.6.0+cu126
(requirements.txt:128); the versions installed in this image are the oracle's pin.
"""
import math

import numpy as np

F16 = np.float16
F32 = np.float32


# --------------------------------------------------------------------------------------
# fp16 helpers
# --------------------------------------------------------------------------------------
def py_to_f16(x):
    """python double -> float -> half, the narrowing torch applies to a python-scalar operand
    of an fp16 comparison (c10::Half has only a float constructor)."""
    return F16(F32(x))


def py_to_dtype(x, dtype):
    if dtype == np.float16:
        return py_to_f16(x)
    return np.dtype(dtype).type(x)


def _mul_scalar_f16(a16, s):
    # fp16 tensor (x) python scalar: fp32 opmath, result rounded to fp16
    return (a16.astype(F32) * F32(s)).astype(F16)


def _add_f16(a16, b16):
    return (a16.astype(F32) + b16.astype(F32)).astype(F16)


# --------------------------------------------------------------------------------------
# state + depth smoothing (PredictAndGenerate.py:63-100,131-145)
# --------------------------------------------------------------------------------------
class WarpState:
    """Per clip-range state: raw-depth history (2 frames) and the previous offset range."""

    def __init__(self, offset_fg=0.025, offset_bg=-0.01, offset_step_size=1):
        self.offset_fg = offset_fg
        self.offset_bg = offset_bg
        self.offset_step_size = offset_step_size
        self.history = []           # raw depths, oldest first
        self.last_range = None      # [bg_px, fg_px] doubles
        # weights: 0.3, 0.3*0.4 for t-1, t-2; the current frame gets 1 - their sum
        w, acc, self.taps = 0.3, 0, []
        for _ in range(2):
            acc = acc + w
            self.taps.append(w)
            w = w * 0.4
        self.w_now = 1 - acc


def smooth_depth(state, raw):
    """raw: [H,W] float16 (or float32) raw depth of the current frame. Returns the smoothed depth
    (same dtype) and pushes `raw` into the history."""
    dt = raw.dtype
    hist = state.history
    while len(hist) < 2:
        hist.append(raw.copy())
    if dt == np.float16:
        d = _mul_scalar_f16(raw, state.w_now)
        d = _add_f16(d, _mul_scalar_f16(hist[1], state.taps[0]))
        d = _add_f16(d, _mul_scalar_f16(hist[0], state.taps[1]))
    else:
        d = raw * dt.type(state.w_now)
        d = d + hist[1] * dt.type(state.taps[0])
        d = d + hist[0] * dt.type(state.taps[1])
    del hist[0]
    hist.append(raw.copy())
    return d


# --------------------------------------------------------------------------------------
# tables (PredictAndGenerate.py:101-126)
# --------------------------------------------------------------------------------------
def layer_tables(state, depth_max, height):
    """Returns (cutoffs[L+1], steps[L], offsets[L], limit, range) as python numbers and updates
    the range EMA.  All arithmetic is python double, in the reference's operation order."""
    limit = math.ceil(float(depth_max))
    fg, bg, stride = state.offset_fg, state.offset_bg, state.offset_step_size
    rng = [bg * height * limit / 14, fg * height * limit / 14]
    if state.last_range is not None:
        rng = [(state.last_range[0] + rng[0]) / 2, (state.last_range[1] + rng[1]) / 2]
    state.last_range = rng
    lo, hi = rng
    span = 0.00001 + hi - lo
    top = 0.00001 + limit

    def threshold(px):
        return (px - lo) / span * top

    marks = [threshold(px) for px in range(round(lo), 0, stride)]
    marks.append(threshold(0))
    marks += [threshold(px) for px in range(1, round(hi), stride)]
    marks.append(limit)
    marks.sort()
    marks[0] = 0
    steps = [b - a for a, b in zip(marks[:-1], marks[1:])]
    offsets = [round(m / top * span + lo) for m, _ in zip(marks, steps)]
    return marks, steps, offsets, limit, rng


def layer_bounds(marks, steps, dtype=np.float16):
    """Per-layer [lo, hi) bounds narrowed to the depth dtype (PredictAndGenerate.py:173)."""
    lo = np.array([py_to_dtype(m - 0.05 * s, dtype) for m, s in zip(marks, steps)], dtype=dtype)
    hi = np.array([py_to_dtype(m + 1.05 * s, dtype) for m, s in zip(marks, steps)], dtype=dtype)
    return lo, hi


# --------------------------------------------------------------------------------------
# gaussian weights (torchvision _misc.py:86-98) — fp32, computed with the same torch CPU ops
# --------------------------------------------------------------------------------------
def gaussian_weights(kx, ky, sigma=3.0):
    import torch

    def k1d(n):
        lim = (n - 1) / (2.0 * math.sqrt(2.0))
        x = torch.linspace(-lim, lim, steps=n, dtype=torch.float32)
        return torch.softmax(x.div(sigma).pow(2).neg(), dim=0)

    return (k1d(ky).unsqueeze(-1) * k1d(kx)).numpy().copy()      # [ky, kx]


def blur_kernel_shape(height):
    k = round(0.0036 * height)
    return k * 2 + 3, k * 2 + 1            # (kx along W, ky along H)


def _reflect(i, n):
    i = np.where(i < 0, -i, i)
    return np.where(i >= n, 2 * (n - 1) - i, i)


def blur_at(pre, ys, xs, weights):
    """Exact (float64) gaussian of `pre` [H,W,3] u8 at pixels (ys,xs); reflect borders;
    round-half-even to u8."""
    H, W, _ = pre.shape
    ky, kx = weights.shape
    acc = np.zeros((len(ys), 3), dtype=np.float64)
    w64 = weights.astype(np.float64)
    for i in range(ky):
        yy = _reflect(ys + (i - ky // 2), H)
        for j in range(kx):
            xx = _reflect(xs + (j - kx // 2), W)
            acc += w64[i, j] * pre[yy, xx, :]
    return np.rint(acc).astype(np.uint8)


# --------------------------------------------------------------------------------------
# the warp (PredictAndGenerate.py:157-198)
# --------------------------------------------------------------------------------------
def strip_width(last_offset, W):
    n = round(last_offset / 3 * 2)
    return max(0, min(W, n if n >= 0 else W + n))       # python slice 0:n on an axis of length W


def warp_frame(img, depth, marks, steps, offsets, weights=None, stages=None):
    """img [H,W,3] u8; depth [H,W] smoothed depth (fp16/fp32).  Returns the SBS frame [H,2W,3].
    `stages`, if a dict, receives the intermediate results (members, painted, pre-blur view...)."""
    H, W, _ = img.shape
    L = len(steps)
    lo, hi = layer_bounds(marks, steps, depth.dtype)
    view = np.zeros_like(img)
    painted = np.zeros((H, W), dtype=bool)
    winner = np.full((H, W), -1, dtype=np.int32)
    multiplicity = np.zeros((H, W), dtype=np.uint8)
    for k in range(L):
        member = (lo[k] <= depth) & (depth < hi[k])
        multiplicity += member
        shift = offsets[k]
        landed = np.roll(member, shift, axis=1) if shift != 0 else member
        moved = np.roll(img, shift, axis=1)
        view[landed] = moved[landed]
        winner[landed] = k
        painted |= landed
    holes = ~painted
    fill_layer = int(L * 3 / 5)
    view[holes] = np.roll(img, offsets[fill_layer], axis=1)[holes]
    pre_blur = view.copy()
    kx, ky = blur_kernel_shape(H)
    if weights is None:
        weights = gaussian_weights(kx, ky)
    ys, xs = np.nonzero(holes)
    if len(ys):
        view[ys, xs] = blur_at(pre_blur, ys, xs, weights)
    post_blur = view.copy()
    sw = strip_width(offsets[-1], W)
    view[:, :sw] = img[:, :sw]
    if stages is not None:
        stages.update(lo=lo, hi=hi, multiplicity=multiplicity, winner=winner, holes=holes,
                      pre_blur=pre_blur, post_blur=post_blur, strip=sw, fill_layer=fill_layer,
                      weights=weights)
    return np.concatenate([view, img], axis=1)


def process_frame(state, img, raw_depth, weights=None, stages=None):
    """One `left_side_sbs` call: smoothing -> tables -> warp."""
    depth = smooth_depth(state, raw_depth)
    marks, steps, offsets, limit, rng = layer_tables(state, depth.max(), depth.shape[0])
    if stages is not None:
        stages.update(depth=depth, marks=marks, steps=steps, offsets=offsets, limit=limit, range=rng)
    return warp_frame(img, depth, marks, steps, offsets, weights, stages)


# --------------------------------------------------------------------------------------
# depth tail: bicubic (A=-0.75, align_corners=True) + scaler (dpt.py:196, PredictAndGenerate.py:55)
# restating ATen's upsample_bicubic2d CUDA arithmetic (torch/include/ATen/native/cuda/UpSample.cuh)
# --------------------------------------------------------------------------------------
def _cubic_coeffs(t):
    """fp32 coefficients; the `x + 1.0` terms are evaluated in double then narrowed, as in ATen."""
    A = F32(-0.75)
    t = t.astype(F32)

    def conv1(x):
        return ((A + F32(2)) * x - (A + F32(3))) * x * x + F32(1)

    def conv2(x):
        return ((A * x - F32(5) * A) * x + F32(8) * A) * x - F32(4) * A

    x1 = t
    x2 = (1.0 - t.astype(np.float64)).astype(F32)
    c0 = conv2((x1.astype(np.float64) + 1.0).astype(F32))
    c1 = conv1(x1)
    c2 = conv1(x2)
    c3 = conv2((x2.astype(np.float64) + 1.0).astype(F32))
    return c0, c1, c2, c3


def bicubic_resize(lowres, H, W, scaler=1.0):
    """lowres [h,w] fp16 -> [H,W] fp16 (fp32 accumulate, no FMA contraction), then `* scaler`
    rounded to fp16 again."""
    h, w = lowres.shape
    src = lowres.astype(F32)
    sy = F32(h - 1) / F32(H - 1) if H > 1 else F32(0)
    sx = F32(w - 1) / F32(W - 1) if W > 1 else F32(0)
    ry = sy * np.arange(H, dtype=F32)
    rx = sx * np.arange(W, dtype=F32)
    iy = np.floor(ry).astype(np.int64)
    ix = np.floor(rx).astype(np.int64)
    cy = _cubic_coeffs(ry - iy.astype(F32))
    cx = _cubic_coeffs(rx - ix.astype(F32))
    rows = []
    for a in range(4):
        yy = np.clip(iy - 1 + a, 0, h - 1)
        taps = [src[yy][:, np.clip(ix - 1 + b, 0, w - 1)] for b in range(4)]
        r = taps[0] * cx[0][None, :]
        for b in range(1, 4):
            r = r + taps[b] * cx[b][None, :]
        rows.append(r)
    out = rows[0] * cy[0][:, None]
    for a in range(1, 4):
        out = out + rows[a] * cy[a][:, None]
    out16 = np.ascontiguousarray(out, dtype=F16)
    if scaler != 1.0:
        out16 = np.ascontiguousarray(_mul_scalar_f16(out16, scaler))
    return out16
