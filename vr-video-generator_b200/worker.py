"""Callers of the hot path, mirrored from the reference (SURVEY.md section 8 f): the CLI parameter surface, the
per-clip-range worker loop and the sub-clip bookkeeping.  Decode, encode and the depth model stay pluggable
callables (OpenCV / ffmpeg / Depth-Anything-V2 in the reference); nothing here touches them.

  * `parse_args`      - PredictAndGenerate.py:327-365 (same flags, same defaults, unknown flags ignored); the
                        same-sign fix-up of :387-393 only rebinds module-level names there and never reaches the
                        warp, so it is NOT applied to the namespace (see `parse_args`)
  * `sbs_worker`      - PredictAndGenerate.py:200-272 (`nibba_woka`), batched: one `left_side_sbs_batch` per sub-clip
  * `subclip_*`       - naming `{last_i}_{i}.mp4` (PredictAndGenerate.py:243), numeric order (Combine_Clips.py:9-10,
                        Check_Clips.py:17-18) and the length / continuity checks of Check_Clips.py:19-37
"""
import argparse
import os

import numpy as np

from . import tables


# ---- f4: CLI parameter surface -----------------------------------------------------------------------------
def make_arg_parser():
    """The reference's parser (PredictAndGenerate.py:327-363): same names, types and defaults."""
    p = argparse.ArgumentParser()
    p.add_argument('--DebugDir', type=str, default="./Debug/")
    p.add_argument('--SubClipDir', type=str, default="./Subclip/")
    p.add_argument('--VideoDir', type=str, default="./Videos/Input/Original/Maria Nagai.mp4")
    p.add_argument('--OutputDir', type=str, default="DeleteThis.mkv")
    p.add_argument('--encoder', type=str, default='vits')
    p.add_argument('--encoder_path', type=str, default='./depth_anything_v2/checkpoints/depth_anything_v2_vits.pth')
    p.add_argument('--offset_fg', type=float, default=0.025)
    p.add_argument('--offset_bg', type=float, default=-0.01)
    p.add_argument('--offset_step_size', type=int, default=1)
    p.add_argument('--Num_Workers', type=int, default=4)
    p.add_argument('--num_gpu', type=int, default=1)
    p.add_argument('--Num_GPU_Workers', type=int, default=1)
    p.add_argument('--Max_Frame_Count', type=int, default=15)
    p.add_argument('--start_frame', type=int, default=0)
    p.add_argument('--end_frame', type=int, default=99999999999999)
    p.add_argument('--repair_mode', type=int, default=0)
    return p


def parse_args(argv=None):
    """`args_god` exactly as the reference's workers receive it: the Namespace of parse_known_args (unknown flags
    are ignored, :365), untouched.  The reference's same-sign fix-up (:387-393) rebinds the MODULE-LEVEL names
    `offset_bg` / `offset_fg` only; `main_func(args)` (:411) hands the workers the original Namespace and
    `SbsProcessor` reads `args_god.offset_fg` / `.offset_bg` (:92-94), so `--offset_fg 0.03 --offset_bg 0.01` warps
    with both offsets positive.  `tables.fix_offset_signs` restates that dead fix-up for documentation; it is not
    applied here.  Returns (namespace, discarded)."""
    args, discarded = make_arg_parser().parse_known_args(argv)
    return args, discarded


def encoder_scaler(encoder):
    """Depth scale per encoder applied by the inference worker (PredictAndGenerate.py:27-34)."""
    return {'vits': 1.618, 'vitb': 0.8, 'vitl': 0.0208}.get(encoder, 1)


# ---- f3: sub-clip bookkeeping --------------------------------------------------------------------------------
def subclip_name(last_i, i):
    return f"{last_i}_{i}.mp4"                                    # PredictAndGenerate.py:243


def subclip_sort_key(name):
    """`int(os.path.splitext(x)[0])`: python accepts '_' inside int literals, so "15_29" sorts as 1529
    (Combine_Clips.py:10, Check_Clips.py:18)."""
    return int(os.path.splitext(name)[0])


def order_subclips(names):
    """The files both tools consider (`i[-1] == '4'`) in the order they concatenate / check them."""
    return sorted([n for n in names if n[-1] == '4'], key=subclip_sort_key)


def check_subclips(entries):
    """Check_Clips.Checkin's two tests (Check_Clips.py:19-37) on (name, frame_count) pairs, without touching files.
    Returns a list of ("length", name, expected, found) and ("continuity", name, next_name, a, b) issues; like the
    reference, the last file is never examined."""
    order = order_subclips([n for n, _ in entries])
    count = dict(entries)
    issues = []
    for k in range(len(order) - 1):
        name, nxt = order[k], order[k + 1]
        first = int(name.split('_')[0])
        a = int(name.split('_')[1].split('.')[0])
        b = int(nxt.split('_')[0])
        if count[name] != a + 1 - first:
            issues.append(("length", name, a + 1 - first, count[name]))
        if a != b and a != b - 1:
            issues.append(("continuity", name, nxt, a, b))
    return issues


# ---- f1: the per-clip-range worker loop ------------------------------------------------------------------------
def sbs_worker(begin, end, read_frame, depth_for, write_subclip, args_god, video_length, height, width,
               processor=None, scaler=1.0):
    """`nibba_woka` (PredictAndGenerate.py:200-272) with the per-frame warp replaced by one batched call per
    sub-clip.

    read_frame(i)            -> BGR uint8 [H,W,3] or None (None -> black frame, :223-225)
    depth_for(rgb [n,H,W,3]) -> fp16 depth [n,H,W] (or DPT-resolution [n,h,w]; `scaler` is then applied on the device)
    write_subclip(name, sbs [n,H,2W,3] uint8) - receives exactly the frames the reference pipes into ffmpeg for
                                               `{SubClipDir}{last_i}_{i}.mp4` (:241-246)
    Sub-clip boundaries and names follow the reference loop: a flush happens when Max_Frame_Count frames have
    accumulated (frame i-1 is appended at iteration i) or at the last frame.  Returns the list of names written.
    Clip-range state (depth history, range EMA) lives in `processor` and starts fresh, like the SbsProcessor a
    worker constructs (:209)."""
    from .sbs import SbsProcessor
    stop = min(end, video_length)
    own = processor is None
    if own:
        processor = SbsProcessor(None, 0, args_god, max_batch=max(1, min(int(args_god.Max_Frame_Count) + 1, 64)))
    else:
        processor.reset_state()
    names, pending, last_i = [], [], begin
    last_img = None
    try:
        for i in range(begin, stop):
            raw = read_frame(i)
            if raw is None:
                raw = np.zeros((height, width, 3), dtype=np.uint8)
            if i != begin:
                pending.append(last_img[:, :, [2, 1, 0]])            # frame i-1, BGR -> RGB (:230-231)
            last_img = raw
            if i == stop - 1:
                pending.append(raw[:, :, [2, 1, 0]])                 # final run (:233-234)
            if len(pending) == args_god.Max_Frame_Count or i == min(end - 1, video_length - 1):
                rgb = np.ascontiguousarray(np.stack(pending)) if pending else np.zeros((0, height, width, 3), np.uint8)
                if len(rgb):
                    out = []
                    for b0 in range(0, len(rgb), processor.max_batch):      # same frames, same order, state carried over
                        chunk = rgb[b0:b0 + processor.max_batch]
                        out.append(processor.left_side_sbs_batch(chunk, depth_for(chunk), scaler=scaler))
                    sbs = np.concatenate(out) if len(out) > 1 else out[0]
                else:
                    sbs = np.zeros((0, height, 2 * width, 3), np.uint8)
                name = subclip_name(last_i, i)
                write_subclip(name, sbs)
                names.append(name)
                last_i = i + 1
                pending = []
    finally:
        if own:
            processor.close()
    return names
