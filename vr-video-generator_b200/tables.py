"""Host-side mirror of the scalar logic of the reference's SbsProcessor (no tensors, no CUDA).

Used by the drop-in `SbsProcessor.get_cutoff` (which, like the reference, returns python lists)
and by the tests that pin the device-side table builder (csrc/table_kernel.cuh) — parity tier T1.
Every expression keeps the reference's operation order, because the results are compared as
exact doubles.  Reference: PredictAndGenerate.py:63-126,165,190-196.
"""
import math


def smoothing_weights(count=2, initial=0.3, ratio=0.4):
    """(weight of the current frame, [weight of t-1, weight of t-2, ...]) — PredictAndGenerate.py:70-80."""
    t, total, taps = initial, 0, []
    for _ in range(count):
        total = total + t
        taps.append(t)
        t = t * ratio
    return 1 - total, taps


def fix_offset_signs(offset_fg, offset_bg):
    """The CLI's same-sign fix-up (PredictAndGenerate.py:387-393), restated for documentation only: in the
    reference it rebinds module-level names that no worker reads (SbsProcessor takes the offsets from the untouched
    `args_god`, :92-94), so the product path never calls this."""
    if offset_bg * offset_fg > 0:
        if offset_bg >= 0:
            offset_bg = offset_bg * (-1)
        else:
            offset_fg = offset_fg * (-1)
    return offset_fg, offset_bg


def layer_tables(depth_max, height, offset_fg, offset_bg, step, last_range):
    """get_cutoff's arithmetic (PredictAndGenerate.py:101-126).

    Returns (cutoff_list, offset_range, step_list, limit_step, offset_x_list); the caller stores
    offset_range as the next `last_range`."""
    limit_step = math.ceil(depth_max)
    offset_range = [offset_bg * height * limit_step / 14, offset_fg * height * limit_step / 14]
    if last_range is not None:
        offset_range = [(last_range[0] + offset_range[0]) / 2, (last_range[1] + offset_range[1]) / 2]
    near, far = offset_range[0], offset_range[1]
    width = 0.00001 + far - near
    ceiling = 0.00001 + limit_step
    pixels = list(range(round(near), 0, step)) + [0] + list(range(1, round(far), step))
    cutoff_list = [(px - near) / width * ceiling for px in pixels]
    cutoff_list.append(limit_step)
    cutoff_list = sorted(cutoff_list)
    cutoff_list[0] = 0
    step_list = [cutoff_list[i + 1] - cutoff_list[i] for i in range(len(cutoff_list) - 1)]
    offset_x_list = [round(cutoff_list[i] / ceiling * width + near) for i in range(len(step_list))]
    return cutoff_list, offset_range, step_list, limit_step, offset_x_list


def layer_bounds(cutoff_list, step_list):
    """Python-double [lo, hi) of every layer before narrowing to the depth dtype (:173)."""
    return ([c - 0.05 * s for c, s in zip(cutoff_list, step_list)],
            [c + 1.05 * s for c, s in zip(cutoff_list, step_list)])


def fill_layer(num_layers):
    return int(num_layers * 3 / 5)                                 # PredictAndGenerate.py:190


def strip_columns(last_offset, width):
    n = round(last_offset / 3 * 2)                                 # PredictAndGenerate.py:196
    return max(0, min(width, n if n >= 0 else width + n))          # python slice 0:n


def blur_kernel_shape(height):
    """(kx, ky) handed to gaussian_blur: (2k+3, 2k+1), k = round(0.0036*H) (:165,:192)."""
    k = round(0.0036 * height)
    return k * 2 + 3, k * 2 + 1


def gaussian_weights(kx, ky, sigma=3.0, device=None):
    """fp32 [ky,kx] kernel exactly as torchvision 0.26 `_get_gaussian_kernel2d` builds it (same torch ops, same
    order): softmax(-(linspace(-lim,lim,k)/sigma)^2), outer product.  `device` = where the torch ops run: the
    reference's torchvision builds the kernel on the image's device (CUDA in production, the CPU under the golden
    fixtures' device proxy); None = CPU.  tests/test_gpu_reference_cuda.py measures whether the two differ."""
    import torch

    def one(n):
        lim = (n - 1) / (2.0 * math.sqrt(2.0))
        x = torch.linspace(-lim, lim, steps=n, dtype=torch.float32, device=device)
        return torch.softmax(x.div(sigma).pow(2).neg(), dim=0)

    return (one(ky).unsqueeze(-1) * one(kx)).contiguous().cpu().numpy()


def clip_ranges(start_frame, end_frame, video_length, num_workers):
    """main_func's shard split (PredictAndGenerate.py:274-275,303): list of (begin, end)."""
    stop = min(end_frame, video_length)
    if stop <= start_frame or num_workers < 1:
        return []
    step = math.ceil((stop - start_frame) / num_workers)
    return [(b, min(end_frame, b + step)) for b in range(start_frame, stop, step)]


# ---- host mirror of the fast-path membership tables (csrc/table_kernel.cuh: cell_entry + the LUT loop) --------
def cell_entry(lo, hi, vmin, vmax):
    """LUT value for the depth cell [vmin, vmax] given monotone fp16 bounds lo[k] <= d < hi[k]:
    e such that a value of the cell can only be in layer e-1 (iff v < hi[e-1]) or layer e (iff v >= lo[e]);
    -1 if the cell is too coarse.  Same case analysis as the device function of the same name."""
    L = len(lo)
    a = int(sum(1 for k in range(L) if hi[k] <= vmin))           # hi is monotone: first layer whose hi exceeds vmin
    if a >= L:
        return L
    if vmin >= lo[a] and (a + 2 > L - 1 or vmax < lo[a + 2]) and (a + 1 > L - 1 or vmax < hi[a + 1]):
        return a + 1
    if vmax < hi[a] and (a + 1 > L - 1 or vmax < lo[a + 1]):
        return a
    return -1


def cell_lut(lo, hi, depth_max, lut_cap=8192):
    """(shift, ncells, lut) or None.  cell = fp16 bits >> shift for non-negative values; index ncells = every
    negative value.  lo/hi: float arrays holding exact fp16 values, monotone."""
    import numpy as np
    maxbits = int(np.float16(depth_max).view(np.uint16)) if depth_max > 0 else 0
    for shift in range(9, -1, -1):
        ncells = (maxbits >> shift) + 1
        if ncells + 1 > lut_cap:
            return None
        lut = []
        for q in range(ncells + 1):
            if q < ncells:
                b0, b1 = q << shift, min(((q + 1) << shift) - 1, maxbits)
                vmin = float(np.uint16(b0).view(np.float16))
                vmax = float(np.uint16(b1).view(np.float16))
            else:
                vmin, vmax = -math.inf, 0.0
            e = cell_entry(lo, hi, vmin, vmax)
            if e < 0:
                break
            lut.append(e)
        else:
            return shift, ncells, lut
    return None


def lut_members(lo, hi, shift, ncells, lut, value):
    """Layers the device's two-compare test paints for one fp16 value (warp_fused.cuh scatter)."""
    import numpy as np
    bits = int(np.float16(value).view(np.uint16))
    e = lut[min(bits >> shift, ncells)]
    L = len(lo)
    out = []
    if e >= 1 and value < hi[e - 1]:
        out.append(e - 1)
    if e <= L - 1 and not (value < lo[e]):
        out.append(e)
    return out
