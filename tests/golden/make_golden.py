"""Generate the golden fixtures in this directory by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py [--only NAME] [--skip-full]

What is produced (all inputs come from vr_video_generator_b200.synth, seeded, bit-reproducible):
  <case>.npz   small cases: frames, raw fp16 depth, the reference's SBS left half per frame
  <case>.json  per frame: the reference's get_cutoff lists (doubles as hex), strip, blur weights,
               sha256 of the reference SBS frame, and the list of pixels where the float64-exact
               oracle differs from the reference's fp32 conv (blurred hole pixels only, +-1)
  full_*.json  1080p / 4K cases: no arrays, only seeds + sha256 + oracle-vs-reference diff list, so
               that  oracle_output + patch(diff)  must hash to the reference's sha256
  depth_tail_*.npz  torch CPU F.interpolate(bicubic, align_corners=True) * scaler on fp16 input

The reference's right half is asserted to equal the input frame here, so only left halves are stored.
"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import sbs_layered as O            # noqa: E402
from oracle import scatter as S                # noqa: E402
from oracle.ref_driver import ReferenceWarp    # noqa: E402
from vr_video_generator_b200 import synth      # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def hexlist(xs):
    return [float(x).hex() for x in xs]


def full_depth(kind, n, H, W, seed, lowres_hw, scaler=1.0, shift=0.0):
    lo = synth.depth_lowres(kind, n, lowres_hw[0], lowres_hw[1], seed)
    raw = np.stack([O.bicubic_resize(lo[t], H, W, scaler) for t in range(n)])
    if shift:
        raw = (raw.astype(np.float32) - np.float32(shift)).astype(np.float16)
    return raw


CASES = {
    # name: (H, W, n, fg, bg, step, frames kind, depth kind, lowres hw, seed, extras)
    "small_a":     dict(H=120, W=160, n=4, fg=0.05, bg=-0.03, step=1, frames="noise", depth="stress", lo=(37, 50), seed=1),
    "small_b":     dict(H=135, W=250, n=3, fg=0.025, bg=-0.01, step=2, frames="noise", depth="scene", lo=(40, 70), seed=2),
    "small_step3": dict(H=96, W=128, n=2, fg=0.06, bg=-0.04, step=3, frames="noise", depth="stress", lo=(30, 40), seed=3),
    "small_neg":   dict(H=96, W=144, n=3, fg=0.05, bg=-0.03, step=1, frames="noise", depth="stress", lo=(30, 44), seed=4, shift=2.0),
    "small_zero":  dict(H=64, W=96, n=3, fg=0.05, bg=-0.03, step=1, frames="noise", depth="stress", lo=(20, 30), seed=5, zero=(0, 2)),
    # same-sign offsets reach the warp as typed (the CLI's fix-up at PredictAndGenerate.py:387-393 never touches args_god)
    "small_pospos": dict(H=96, W=160, n=3, fg=0.12, bg=0.04, step=1, frames="noise", depth="stress", lo=(30, 50), seed=12),
    "small_negneg": dict(H=96, W=160, n=3, fg=-0.04, bg=-0.12, step=1, frames="noise", depth="scene", lo=(30, 50), seed=13),
    # fp32 depth: what torch >= 2.4's CUDA autocast delivers (upsample_bicubic2d is on its fp32 list); values are NOT fp16-
    # representable (x 1.00037), so smoothing, the max and the bin comparisons really run in fp32
    "small_f32":   dict(H=120, W=160, n=4, fg=0.05, bg=-0.03, step=1, frames="noise", depth="stress", lo=(37, 50), seed=14, f32=True),
    "medium_f32":  dict(H=270, W=480, n=3, fg=0.025, bg=-0.01, step=1, frames="gradient", depth="scene", lo=(74, 132), seed=15, f32=True),
    "medium":      dict(H=270, W=480, n=4, fg=0.025, bg=-0.015, step=1, frames="gradient", depth="scene", lo=(74, 132), seed=6),
    "full_1080p_cfg1":  dict(H=1080, W=1920, n=3, fg=0.025, bg=-0.015, step=1, frames="noise", depth="stress", lo=(518, 924), seed=7, full=True),
    "full_1080p_step2": dict(H=1080, W=1920, n=2, fg=0.025, bg=-0.01, step=2, frames="gradient", depth="scene", lo=(518, 924), seed=8, full=True),
    "full_4k_wide":     dict(H=2160, W=3840, n=2, fg=0.05, bg=-0.03, step=1, frames="gradient", depth="scene", lo=(518, 924), seed=9, full=True),
}


def case_inputs(c):
    """Also imported by the tests: regenerates the inputs of a case from its parameters."""
    gen = synth.frames_noise if c["frames"] == "noise" else synth.frames_gradient
    frames = gen(c["n"], c["H"], c["W"], c["seed"])
    raw = full_depth(c["depth"], c["n"], c["H"], c["W"], c["seed"], c["lo"], 1.0, c.get("shift", 0.0))
    for t in c.get("zero", ()):
        raw[t] = 0
    if c.get("f32"):
        raw = raw.astype(np.float32) * np.float32(1.00037)
    return frames, raw


def run_case(name, c):
    t0 = time.time()
    frames, raw = case_inputs(c)
    H, W, n = c["H"], c["W"], c["n"]
    ref = ReferenceWarp(c["fg"], c["bg"], c["step"])
    st = O.WarpState(c["fg"], c["bg"], c["step"])
    weights = O.gaussian_weights(*O.blur_kernel_shape(H))
    meta = dict(params={k: c[k] for k in ("H", "W", "n", "fg", "bg", "step", "frames", "depth", "seed")},
                lowres=list(c["lo"]), shift=c.get("shift", 0.0), zero=list(c.get("zero", ())), f32=bool(c.get("f32", False)),
                weights=hexlist(weights.ravel()), weights_shape=list(weights.shape),
                inputs_sha=dict(frames=sha(frames), raw_depth=sha(raw)), frames=[])
    lefts = []
    for t in range(n):
        # the reference's own tables for this frame, from a twin processor fed the same state
        out = ref.left_side_sbs(frames[t], torch.from_numpy(raw[t]))
        assert out.shape == (H, 2 * W, 3) and out.dtype == np.uint8
        assert np.array_equal(out[:, W:], frames[t]), "reference right half != input frame"
        stages = {}
        mine = S.process_frame(st, frames[t], raw[t], weights=weights, stages=stages)
        diff = np.argwhere(out != mine)
        holes = stages["holes"]
        for (y, x, ch) in diff:
            assert x < W and holes[y, x] and x >= stages["strip"], f"{name}[{t}]: non-blur pixel differs at {(y, x, ch)}"
            assert abs(int(out[y, x, ch]) - int(mine[y, x, ch])) == 1
        rng = ref.proc.last_offset_range
        assert rng == stages["range"], (rng, stages["range"])
        meta["frames"].append(dict(
            sha256=sha(out), layers=len(stages["steps"]), limit=stages["limit"], strip=int(stages["strip"]),
            fill_layer=int(stages["fill_layer"]), holes=int(holes.sum()),
            blurred=int(holes[:, stages["strip"]:].sum()),
            cutoffs=hexlist(stages["marks"]), steps=hexlist(stages["steps"]), offsets=[int(o) for o in stages["offsets"]],
            range=hexlist(rng), depth_max=float(stages["depth"].max()),
            oracle_vs_reference=[[int(y), int(x), int(ch), int(out[y, x, ch]), int(mine[y, x, ch])] for (y, x, ch) in diff]))
        lefts.append(out[:, :W].copy())
        print(f"  {name}[{t}] L={len(stages['steps'])} holes={holes.mean():.4f} oracle-vs-ref diffs={len(diff)}", flush=True)
    with open(os.path.join(HERE, name + ".json"), "w") as f:
        json.dump(meta, f, indent=1)
    if not c.get("full"):
        np.savez_compressed(os.path.join(HERE, name + ".npz"), frames=frames, raw_depth=raw, ref_left=np.stack(lefts))
    print(f"{name}: done in {time.time() - t0:.1f}s", flush=True)


def depth_tail():
    """Golden for the depth tail (dpt.py:196): torch CPU bicubic, align_corners=True.

    The reference runs this op on CUDA, where ATen accumulates in fp32 and rounds once to fp16.
    torch's CPU fp16 kernel rounds its intermediate row pass to fp16 as well (measured here: 26 %
    of pixels differ by >= 1 fp16 ulp from the fp32 result), so it is NOT a model of the CUDA op.
    The fixture therefore stores the CPU result computed on the fp32 view of the same fp16 input
    (`ref32`); narrowing it to fp16 is the closest CPU-side stand-in for the CUDA kernel."""
    for name, (h, w, H, W, stride) in {"depth_tail_small": (74, 132, 270, 480, 1),
                                       "depth_tail_1080p": (518, 924, 1080, 1920, 8)}.items():
        lo = synth.depth_stress(2, h, w, seed=21)
        t = torch.from_numpy(lo).float()
        up = torch.nn.functional.interpolate(t[:, None], (H, W), mode="bicubic", align_corners=True)[:, 0].numpy()
        mine = np.stack([O.bicubic_resize(lo[i], H, W, 1.0) for i in range(2)])
        ulps = np.abs(up.astype(np.float16).view(np.int16).astype(int) - mine.view(np.int16).astype(int))
        print(f"  {name}: oracle vs fp16(torchCPU fp32): exact={np.mean(ulps == 0):.6f} max ulp={ulps.max()}")
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            params=np.array([h, w, H, W, stride, 21], dtype=np.int64), ref32=up[:, ::stride, ::stride])


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only")
    ap.add_argument("--skip-full", action="store_true")
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    if not a.only or a.only == "depth_tail":
        depth_tail()
    for name, c in CASES.items():
        if a.only and a.only != name:
            continue
        if a.skip_full and c.get("full"):
            continue
        run_case(name, c)
