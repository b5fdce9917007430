"""Out-of-bounds writes on the caller's buffers, checked with guard bands (compute-sanitizer is closed on this GPU
pool: profiles/r02_sanitizer_closed.log).  Every device buffer handed to the C ABI sits between two 64 KiB bands of a
known pattern inside one allocation; after each route has run, the bands must be untouched and the output must still
equal the reference fixture.  Odd sizes put every buffer end at an unaligned address."""
import numpy as np
import pytest
import torch

from conftest import golden_weights, load_case
from vr_video_generator_b200 import _native, tables

pytestmark = pytest.mark.gpu
GUARD = 64 * 1024
PATTERN = 0xA5


class Guarded:
    def __init__(self, nbytes, align=16, skew=0):
        self.n = nbytes
        self.buf = torch.full((GUARD + nbytes + GUARD + 256,), PATTERN, dtype=torch.uint8, device="cuda")
        base = self.buf.data_ptr() + GUARD
        self.off = GUARD + ((-base) % align) + skew
        self.view = self.buf[self.off:self.off + nbytes]

    @property
    def ptr(self):
        return self.view.data_ptr()

    def intact(self):
        lo, hi = self.buf[:self.off], self.buf[self.off + self.n:]
        return bool((lo == PATTERN).all().item()) and bool((hi == PATTERN).all().item())


def _run(frames, raw, p, weights, opts, skew=0):
    n, H, W = raw.shape
    ctx = _native.Context(0, H, W, max(4, n), 512)
    ctx.reset(p["fg"], p["bg"], p["step"], True)
    ctx.set_blur_weights(weights)
    for k, v in opts.items():
        ctx.set_option(k, v)
    gf, gr = Guarded(frames.nbytes, skew=skew), Guarded(raw.nbytes, skew=skew)
    gd, go = Guarded(raw.nbytes, skew=skew), Guarded(n * H * 2 * W * 3, skew=skew)
    gf.view.copy_(torch.from_numpy(np.ascontiguousarray(frames)).view(-1))
    gr.view.copy_(torch.from_numpy(np.ascontiguousarray(raw)).view(torch.uint8).view(-1))
    s = torch.cuda.current_stream().cuda_stream
    ctx.process_batch(gf.ptr, gr.ptr, n, H, W, gd.ptr, go.ptr, s)
    torch.cuda.synchronize()
    ctx.close()
    ok = {name: g.intact() for name, g in (("frames", gf), ("raw", gr), ("depth_scratch", gd), ("sbs", go))}
    return go.view.cpu().numpy().reshape(n, H, 2 * W, 3), ok


ROUTES = {
    "default": {},
    "slow_membership": {"fast_tables": 0},
    "fused_barrier_kernel": {"warp_ws": 0},
    "smooth_in_warp": {"smooth_in_warp": 1},
    "row_kernel_atomic": {"fused": 0, "scatter_mode": 2},
    "row_kernel_store": {"fused": 0, "scatter_mode": 1},
    "blur_sep_per_word": {"blur_band": 0},
    "blur_sep_list_from_warp_kernel": {"blur_band": 0, "ws_no_list": 0},
    "blur_fixed": {"blur_sep": 0},
    "blur_exact_only": {"blur_screen": 0},
}


@pytest.mark.parametrize("route", sorted(ROUTES))
@pytest.mark.parametrize("name", ["small_a", "medium"])
def test_guard_bands_survive_every_route(name, route):
    meta, frames, raw, ref_left = load_case(name)
    p = meta["params"]
    out, ok = _run(frames, raw, p, golden_weights(meta), ROUTES[route])
    assert all(ok.values()), ok
    assert np.array_equal(out[:, :, :p["W"]], ref_left)
    assert np.array_equal(out[:, :, p["W"]:], frames)


@pytest.mark.parametrize("shape", [(37, 250), (40, 2561), (33, 96), (9, 2080)])
def test_guard_bands_with_odd_sizes_and_unaligned_buffers(shape):
    """Widths that are not multiples of 16 / 32 and buffers that start one byte past a 16-byte boundary take the general
    row kernel with cooperative copies; a wide odd row takes its 512- and 1024-thread instantiations."""
    H, W = shape
    n = 3
    rng = np.random.default_rng(H * 10007 + W)
    frames = rng.integers(0, 256, (n, H, W, 3), dtype=np.uint8)
    raw = (rng.random((n, H, W), dtype=np.float32) * 13.9).astype(np.float16)
    p = dict(fg=0.3, bg=-0.2, step=1)
    w = tables.gaussian_weights(*tables.blur_kernel_shape(1080))
    base, ok = _run(frames, raw, p, w, {"fused": 0, "scatter_mode": 2})
    assert all(ok.values()), ok
    for skew in (0, 2):                                   # fp16 buffers stay 2-byte aligned
        out, ok = _run(frames, raw, p, w, {}, skew=skew)
        assert all(ok.values()), (skew, ok)
        assert np.array_equal(out, base)
        assert np.array_equal(out[:, :, W:], frames)
