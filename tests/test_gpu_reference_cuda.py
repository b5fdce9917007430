"""Parity tier T5-ii: the CUDA path against the UNMODIFIED reference running on the same B200 as it ships —
torch CUDA ops, blur through `torchvision.gaussian_blur` -> cuDNN conv (PredictAndGenerate.py:157-198), with
`torch.backends.cudnn.allow_tf32` on (torch's default) and off.

The reference is loaded from oracle/_ref (staged byte for byte by oracle/stage_ref.py, shipped by gpurun; it is not in
the git history), so these tests skip where that copy is absent.

Contract checked: every byte outside the blurred hole pixels — painted pixels, strip, right half — is identical to the
reference's own CUDA output; at blurred hole pixels the integer-exact blur may differ from whatever summation order
and precision the conv backend picked, and the size of that gap is COUNTED (+-1 flips, flips > 1) and written to
gpurun_out/t5ii_reference_cuda.json; DESIGN.md quotes the table.  The bound asserted here is deliberately loose
(|d| <= 1 with TF32 off; TF32 may legitimately move more): the point of the tier is the measurement.
"""
import json
import os

import numpy as np
import pytest

from conftest import ROOT, golden_weights, load_case

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from oracle import ref_driver  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_driver.reference_available(),
                               reason="reference not staged (python oracle/stage_ref.py in the build container)")

REPORT = os.path.join(ROOT, "gpurun_out", "t5ii_reference_cuda.json")


def _record(key, value):
    os.makedirs(os.path.dirname(REPORT), exist_ok=True)
    data = {}
    if os.path.exists(REPORT):
        with open(REPORT) as f:
            data = json.load(f)
    data[key] = value
    with open(REPORT, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)


def _ours(p, frames, raw, weights=None):
    from vr_video_generator_b200 import _native, tables
    n, H, W, _ = frames.shape
    ctx = _native.Context(0, H, W, max(n, 1), 512)
    ctx.reset(p["fg"], p["bg"], p["step"], True)
    ctx.set_blur_weights(weights if weights is not None else tables.gaussian_weights(*tables.blur_kernel_shape(H)))
    f = torch.from_numpy(np.ascontiguousarray(frames)).cuda()
    r = torch.from_numpy(np.ascontiguousarray(raw)).cuda()
    out = torch.empty((n, H, 2 * W, 3), dtype=torch.uint8, device="cuda")
    dep = torch.empty((n, H, W), dtype=torch.float16, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    ctx.process_batch(f.data_ptr(), r.data_ptr(), n, H, W, dep.data_ptr(), out.data_ptr(), s)
    infos = ctx.frame_info(n, s)
    masks = ctx.hole_mask(n, H, W, s)
    torch.cuda.synchronize()
    ctx.close()
    return out.cpu().numpy(), infos, masks


def _reference_cuda(p, frames, raw, tf32):
    prev = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = tf32
    try:
        ref = ref_driver.ReferenceWarp(p["fg"], p["bg"], p["step"], device="cuda")
        outs = [ref.left_side_sbs(frames[t], torch.from_numpy(raw[t])) for t in range(len(frames))]
        torch.cuda.synchronize()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = prev
    return np.stack(outs)


def _compare(name, p, frames, raw, weights=None):
    ours, infos, masks = _ours(p, frames, raw, weights)
    n, H, W, _ = frames.shape
    row = {"frames": int(n), "H": int(H), "W": int(W)}
    for tf32 in (True, False):
        ref = _reference_cuda(p, frames, raw, tf32)
        assert ref.shape == ours.shape
        blurred = flips1 = flips_gt1 = 0
        worst = 0
        for t in range(n):
            strip = infos[t].strip
            hole = masks[t].copy()
            hole[:, :strip] = False                     # strip columns are restored from the input
            diff = ours[t].astype(np.int16) - ref[t].astype(np.int16)
            left = diff[:, :W]
            # everything that is not a blurred hole value: byte-identical to the reference's CUDA output
            assert not left[~hole].any(), f"{name}[{t}] tf32={tf32}: painted/strip pixels differ from the reference on CUDA"
            assert not diff[:, W:].any(), f"{name}[{t}]: right half differs"
            d = np.abs(left[hole])
            blurred += int(d.size)
            flips1 += int((d == 1).sum())
            flips_gt1 += int((d > 1).sum())
            worst = max(worst, int(d.max()) if d.size else 0)
        row["tf32_on" if tf32 else "tf32_off"] = {"blurred_values": blurred, "flips_pm1": flips1, "flips_gt1": flips_gt1,
                                                  "max_abs": worst, "frac": (flips1 + flips_gt1) / max(blurred, 1)}
        if not tf32:
            assert worst <= 1, f"{name}: blurred value off by {worst} against the reference's fp32 cuDNN conv"
    _record(name, row)
    print(name, json.dumps(row))
    return row


@needs_ref
@pytest.mark.parametrize("name", ["small_a", "small_b", "small_neg", "small_zero", "small_pospos", "medium"])
def test_small_cases_against_reference_on_cuda(name):
    meta, frames, raw, _ = load_case(name)
    _compare(name, meta["params"], frames, raw, golden_weights(meta))


@needs_ref
@pytest.mark.parametrize("wl", ["1080p_scene", "1080p_stress"])
def test_1080p_against_reference_on_cuda(wl):
    """The bench workloads' content (D-scene / D-stress at 1080p), 4 consecutive frames of one clip."""
    from oracle import sbs_layered as O
    from vr_video_generator_b200 import synth
    n, H, W = 4, 1080, 1920
    kind = "scene" if wl.endswith("scene") else "stress"
    frames = synth.frames_noise(n, H, W, seed=100)
    lo = synth.depth_lowres(kind, n, synth.DPT_H, synth.DPT_W, 100)
    raw = np.stack([O.bicubic_resize(lo[t], H, W, 1.0) for t in range(n)])
    p = dict(fg=0.025, bg=-0.01 if kind == "scene" else -0.015, step=1)
    row = _compare(wl, p, frames, raw)
    assert row["tf32_off"]["frac"] < 1e-3
    # the same with the gaussian kernel built on the device, as the reference's torchvision builds it there
    from vr_video_generator_b200 import tables
    w_dev = tables.gaussian_weights(*tables.blur_kernel_shape(H), 3.0, device="cuda")
    _compare(wl + "_weights_built_on_cuda", p, frames, raw, w_dev)


@needs_ref
def test_reference_gaussian_kernel_on_cuda_equals_cpu_weights():
    """ADVICE: `tables.gaussian_weights` builds the kernel with torch CPU ops, the reference's torchvision builds it on
    the image's device.  Compare the two fp32 kernels bit for bit (an impulse response through the reference's call)."""
    from torchvision.transforms.v2.functional import gaussian_blur
    from vr_video_generator_b200 import tables
    out = {}
    for H in (1080, 2160, 720):
        kx, ky = tables.blur_kernel_shape(H)
        w_cpu = tables.gaussian_weights(kx, ky, 3.0)
        from torchvision.transforms.v2.functional._misc import _get_gaussian_kernel2d
        w_gpu = _get_gaussian_kernel2d([kx, ky], [3.0, 3.0], dtype=torch.float32, device=torch.device("cuda")).cpu().numpy()
        same = bool(np.array_equal(w_cpu.view(np.uint32), w_gpu.view(np.uint32)))
        ulps = int(np.abs(w_cpu.view(np.int32).astype(np.int64) - w_gpu.view(np.int32).astype(np.int64)).max())
        out[f"{kx}x{ky}"] = {"bit_identical": same, "max_ulp": ulps}
        assert ulps <= 64, f"CUDA-built gaussian kernel differs from the CPU-built one by {ulps} ulp"
    _record("gaussian_kernel_cuda_vs_cpu", out)
    print(json.dumps(out))
    assert gaussian_blur is not None
