"""Host pipeline variants on one box: frames/s of left_side_sbs_batch with pinned buffers."""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
import vr_video_generator_b200 as pkg
from vr_video_generator_b200 import synth
from oracle import sbs_layered as O
H, W, B = 1080, 1920, 64
frames = synth.frames_noise(B, H, W, 100)
lo = synth.depth_lowres("scene", B, synth.DPT_H, synth.DPT_W, 100)
raw = np.stack([O.bicubic_resize(lo[t], H, W, 1.0) for t in range(8)] * 8)
f_pin = torch.from_numpy(frames).pin_memory(); d_pin = torch.from_numpy(raw).pin_memory()
o_pin = torch.empty((B, H, 2 * W, 3), dtype=torch.uint8).pin_memory(); o_np = o_pin.numpy()
ns = argparse.Namespace(offset_fg=0.025, offset_bg=-0.01, offset_step_size=1)
for opts in ({"host_right_half": 0}, {"host_right_half": 2}, {"host_right_half": 1}, {"host_right_half": 1, "copy_threads": 12},
             {"host_right_half": 1, "host_chunk": 8}, {"host_right_half": 1, "host_chunk": 8, "copy_threads": 12},
             {"host_right_half": 1, "host_chunk": 16, "copy_threads": 12}, {"host_right_half": 1, "host_chunk": 2}):
    proc = pkg.SbsProcessor(None, 0, ns, device=0, max_batch=16)
    ctx = proc._context(H, W)
    for k, v in opts.items(): ctx.set_option(k, v)
    for _ in range(2): proc.left_side_sbs_batch(f_pin, d_pin, out=o_np)
    t0 = time.perf_counter()
    for _ in range(5): proc.left_side_sbs_batch(f_pin, d_pin, out=o_np)
    dt = (time.perf_counter() - t0) / 5
    print(opts, "fps", round(B / dt), "ms/step", round(dt * 1e3, 2))
    proc.close()
