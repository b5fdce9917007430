"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck / synccheck / initcheck): every route of the
default and fallback kernels once, tiny frames.  modes: 0 default (k_warp_ws, LUT membership), 3 slow membership,
4 smoothing inside k_warp_fused, 5 k_warp_fused without warp specialisation, 2 general row kernel (atomicMax),
1 general row kernel (store + verify), 6 screening blur off (exact sums only), 7 k_blur_holes_fixed instead of the separable
kernels, 8 k_blur_sep (per mask word) instead of k_blur_band;
plus a 2560-pixel-wide case for the 16-warp instantiation k_warp_ws<512,8> and the host pipeline (submit / collect)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from vr_video_generator_b200 import _native, tables, synth
from conftest import load_case, golden_weights


def run(frames, raw, p, weights, mode, max_layers=512):
    n, H, W = raw.shape
    ctx = _native.Context(0, H, W, max(4, n), max_layers)
    ctx.reset(p["fg"], p["bg"], p["step"], True); ctx.set_blur_weights(weights)
    ctx.set_option("fused", 0 if mode in (1, 2) else 1)
    if mode in (1, 2): ctx.set_option("scatter_mode", mode)
    ctx.set_option("fast_tables", 0 if mode == 3 else 1)
    ctx.set_option("smooth_in_warp", 1 if mode == 4 else 0)
    ctx.set_option("warp_ws", 0 if mode == 5 else 1)
    ctx.set_option("blur_screen", 0 if mode == 6 else 1)
    ctx.set_option("blur_sep", 0 if mode == 7 else 1)
    ctx.set_option("blur_band", 0 if mode == 8 else 1)
    f = torch.from_numpy(np.ascontiguousarray(frames)).cuda(); r = torch.from_numpy(np.ascontiguousarray(raw)).cuda()
    out = torch.empty((n, H, 2 * W, 3), dtype=torch.uint8, device="cuda"); dep = torch.empty((n, H, W), dtype=torch.float16, device="cuda")
    ctx.process_batch(f.data_ptr(), r.data_ptr(), n, H, W, dep.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    res = out.cpu().numpy()
    ctx.close()
    return res


modes = [int(m) for m in os.environ.get("SAN_MODES", "0,3,4,5,2,1,6,7,8").split(",")]
for name in os.environ.get("SAN_CASES", "small_a,medium").split(","):
    meta, frames, raw, ref_left = load_case(name)
    p = meta["params"]; W = p["W"]
    for mode in modes:
        got = run(frames, raw, p, golden_weights(meta), mode)
        print(name, "mode", mode, "matches reference:", np.array_equal(got[:, :, :W], ref_left), flush=True)

if os.environ.get("SAN_WIDE", "1") == "1":
    # 16-warp instantiation: rows wider than 2048 pixels, few rows
    H, W, n = 24, 2560, 3
    frames = synth.frames_noise(n, H, W, seed=5)
    raw = (np.random.default_rng(5).random((n, H, W), dtype=np.float32) * 13.9).astype(np.float16)
    p = dict(fg=0.5, bg=-0.3, step=1)
    w = tables.gaussian_weights(*tables.blur_kernel_shape(1080))
    base = run(frames, raw, p, w, 2)
    for mode in (0, 3):
        print("wide mode", mode, "equals general row kernel:", np.array_equal(run(frames, raw, p, w, mode), base), flush=True)

if os.environ.get("SAN_HOST", "1") == "1":
    # host pipeline: submit / collect with the frames decoded in place into the right halves
    import argparse
    from vr_video_generator_b200 import SbsProcessor
    from vr_video_generator_b200.sbs import pinned_sbs_buffer
    meta, frames, raw, ref_left = load_case("small_a")
    p = meta["params"]; H, W, n = p["H"], p["W"], p["n"]
    proc = SbsProcessor(None, 0, argparse.Namespace(offset_fg=p["fg"], offset_bg=p["bg"], offset_step_size=p["step"]), device=0, max_batch=4)
    proc._context(H, W).set_blur_weights(golden_weights(meta))
    out, view, _keep = pinned_sbs_buffer(n, H, W)
    np.copyto(view, frames)
    proc.collect(proc.submit_batch(view, torch.from_numpy(np.ascontiguousarray(raw)).pin_memory(), out))
    print("host pipeline matches reference:", np.array_equal(out[:, :, :W], ref_left), flush=True)
    proc.close()
