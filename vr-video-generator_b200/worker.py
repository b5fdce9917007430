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


def concat_list_lines(subclip_dir, names):
    """The ffmpeg concat-demuxer list `combine_clips` writes (Combine_Clips.py:9-18): one `file '<dir>/<name>'` line per
    sub-clip, in numeric order."""
    return [f"file '{os.path.join(subclip_dir, n)}'\n" for n in order_subclips(names)]


def combine_commands(subclip_dir, original_path, output_path, just_combine=0, ffmpeg="./ffmpeg/ffmpeg",
                     list_path=os.path.join(".", "input_list.txt")):
    """The ffmpeg invocations of `combine_clips` after the list is written (Combine_Clips.py:19-42): concatenate, and unless
    `just_combine` (repair_mode 3) extract the original audio and mux it back.  Returned as argv lists; nothing is run."""
    cmds = [[ffmpeg, "-f", "concat", "-safe", "0", "-y", "-i", list_path, "-c", "copy", "temp_video.mkv"]]
    if just_combine == 1:
        return cmds
    cmds.append([ffmpeg, "-y", "-i", original_path, "-map", "0:a", "-c:a", "copy", "original_audio.mka"])
    cmds.append([ffmpeg, "-y", "-i", "temp_video.mkv", "-i", "original_audio.mka", "-map", "0:v", "-map", "1:a",
                 "-c:v", "copy", "-c:a", "copy", output_path])
    return cmds


def repair_plan(entries, repair_subclip_dir="D:/TEMP/FixxingSubclip/"):
    """What `Checkin(..., repair_mode=1)` does about each issue (Check_Clips.py:27-37), without touching files: returns
    (commands, removed) - the re-invocations of the CLI with `--repair_mode 1` for the affected frame range and the
    files it deletes.  A length issue re-runs [first, a+1) and removes the file (the removal happens in every mode,
    :32); a continuity gap re-runs [a+1, b).  `entries` = (name, frame_count) pairs."""
    order = order_subclips([n for n, _ in entries])
    count = dict(entries)
    cmds, removed = [], []
    for k in range(len(order) - 1):
        name, nxt = order[k], order[k + 1]
        first = name.split('_')[0]
        a = int(name.split('_')[1].split('.')[0])
        b = int(nxt.split('_')[0])
        if count[name] != a + 1 - int(first):
            cmds.append(["python", "PredictAndGenerate.py", "--SubClipDir", repair_subclip_dir, "--Num_Workers", "2",
                         "--start_frame", first, "--end_frame", str(a + 1), "--repair_mode", "1"])
            removed.append(name)
        if a != b and a != b - 1:
            cmds.append(["python", "PredictAndGenerate.py", "--SubClipDir", repair_subclip_dir, "--Num_Workers", "2",
                         "--start_frame", str(a + 1), "--end_frame", str(b), "--repair_mode", "1"])
    return cmds, removed


def run_plan(repair_mode):
    """What the CLI does for each `--repair_mode` (PredictAndGenerate.py:358-363,402-419): which directories are wiped,
    whether the workers run, whether the sub-clips are combined and with which `just_combine`."""
    return dict(wipe_debug=repair_mode != 1, wipe_subclips=repair_mode == 0, run_workers=repair_mode in (0, 1),
                combine=repair_mode in (0, 2, 3), just_combine=1 if repair_mode == 3 else 0)


def worker_ranges(start_frame, end_frame, video_length, num_workers, num_gpu_workers):
    """main_func's fan-out (PredictAndGenerate.py:274-275,284-286,300-306): [(begin, end, inference worker, slot)] - SBS
    worker idx is served by inference worker idx % Num_GPU_Workers, as that worker's queue number idx // Num_GPU_Workers."""
    return [(b, e, idx % num_gpu_workers, idx // num_gpu_workers)
            for idx, (b, e) in enumerate(tables.clip_ranges(start_frame, end_frame, video_length, num_workers))]


# ---- f1: the per-clip-range worker loop ------------------------------------------------------------------------
def sbs_worker(begin, end, read_frame, depth_for, write_subclip, args_god, video_length, height, width,
               processor=None, scaler=1.0, pipelined=True, ring=3, stats=None):
    """`nibba_woka` (PredictAndGenerate.py:200-272) with the per-frame warp replaced by one batched call per
    sub-clip.

    read_frame(i)            -> BGR uint8 [H,W,3] or None (None -> black frame, :223-225)
    depth_for(rgb [n,H,W,3]) -> fp16 depth [n,H,W] (or DPT-resolution [n,h,w]; `scaler` is then applied on the device);
                                numpy, CPU tensor or CUDA tensor
    write_subclip(name, sbs [n,H,2W,3] uint8) - receives exactly the frames the reference pipes into ffmpeg for
                                               `{SubClipDir}{last_i}_{i}.mp4` (:241-246); `sbs` is only valid during the call
    Sub-clip boundaries and names follow the reference loop (`shard.flush_ranges`): a flush happens when Max_Frame_Count
    frames have accumulated (frame i-1 is appended at iteration i) or at the last frame; a one-frame range writes
    nothing, like the reference (its progress print divides by zero there).  Returns the list of names written.
    Clip-range state (depth history, range EMA) lives in `processor` and starts fresh, like the SbsProcessor a
    worker constructs (:209).

    pipelined=True (default): decode, warp and encode of neighbouring sub-clips run concurrently - what north_star asks
    of the batching and what the reference's serial loop (:221-250) cannot do.  A reader thread decodes sub-clip k+2
    straight into the right halves of a page-locked SBS buffer (BGR -> RGB in the same pass, the reference's
    `raw_img[:,:,[2,1,0]]` copy), the calling thread runs the depth producer and submits sub-clip k+1 to the GPU
    (`SbsProcessor.submit_batch`: H2D, kernels, D2H of the synthesised halves only), and a writer thread hands the
    finished sub-clip k to `write_subclip`.  `ring` buffers of Max_Frame_Count + 1 frames circulate between the three.
    pipelined=False is the serial loop (read all, warp, write), kept for comparison and for pageable-only callers.
    `stats`, if a dict, receives wall-clock seconds spent reading / in the depth producer / waiting for the GPU / writing,
    `overlap` (their sum over the wall clock: > 1 means the stages ran concurrently) and the busy `intervals`."""
    from .sbs import SbsProcessor
    stop = min(end, video_length)
    own = processor is None
    if own:
        processor = SbsProcessor(None, 0, args_god, max_batch=max(1, min(int(args_god.Max_Frame_Count) + 1, 64)))
    else:
        processor.reset_state()
    try:
        if pipelined:
            return _worker_pipelined(begin, end, read_frame, depth_for, write_subclip, args_god, video_length, height, width,
                                     processor, scaler, ring, stats)
        return _worker_serial(begin, end, read_frame, depth_for, write_subclip, args_god, video_length, height, width,
                              processor, scaler)
    finally:
        if own:
            processor.close()


def _worker_serial(begin, end, read_frame, depth_for, write_subclip, args_god, video_length, height, width, processor, scaler):
    from . import shard
    names, frame = [], begin
    for (last_i, i, n) in shard.flush_ranges(begin, end, video_length, int(args_god.Max_Frame_Count)):
        pending = []
        for _ in range(n):                                        # frames reach the warp in order (:221-234)
            raw = read_frame(frame)
            if raw is None:
                raw = np.zeros((height, width, 3), dtype=np.uint8)   # failed read -> black frame (:223-225)
            pending.append(raw[:, :, [2, 1, 0]])                  # BGR -> RGB (:227-234)
            frame += 1
        rgb = np.ascontiguousarray(np.stack(pending))
        out = []
        for b0 in range(0, len(rgb), processor.max_batch):       # same frames, same order, state carried over
            chunk = rgb[b0:b0 + processor.max_batch]
            out.append(processor.left_side_sbs_batch(chunk, depth_for(chunk), scaler=scaler))
        name = subclip_name(last_i, i)
        write_subclip(name, np.concatenate(out) if len(out) > 1 else out[0])
        names.append(name)
    return names


def _worker_pipelined(begin, end, read_frame, depth_for, write_subclip, args_god, video_length, height, width, processor,
                      scaler, ring, stats):
    import queue
    import threading
    import time

    from . import shard
    from .sbs import pinned_sbs_buffer
    plan = shard.flush_ranges(begin, end, video_length, int(args_god.Max_Frame_Count))   # [(last_i, i, frames)] as the loop names them
    if not plan:
        return []
    cap = max(n for _, _, n in plan)
    ring = max(2, int(ring))
    bufs = [pinned_sbs_buffer(max(cap, 1), height, width) for _ in range(ring)]
    q_free, q_ready, q_done = queue.Queue(), queue.Queue(), queue.Queue()
    for k in range(ring):
        q_free.put(k)
    errors, t_read, t_write, busy = [], [0.0], [0.0], []
    abort = threading.Event()

    def get(q):
        while True:
            try:
                return q.get(timeout=0.2)
            except queue.Empty:
                if abort.is_set():
                    raise RuntimeError("worker pipeline aborted")

    def reader():
        try:
            frame = begin
            for (last_i, i, n) in plan:
                k = get(q_free)
                t0 = time.perf_counter()
                view = bufs[k][1]
                for j in range(n):
                    raw = read_frame(frame)                                  # frames reach the warp in order (:221-234)
                    if raw is None:
                        view[j] = 0                                           # failed read -> black frame (:223-225)
                    else:
                        np.copyto(view[j], raw[:, :, ::-1])                   # BGR -> RGB straight into the pinned right half
                    frame += 1
                t1 = time.perf_counter()
                t_read[0] += t1 - t0
                busy.append(("read", t0, t1))
                q_ready.put((k, n, subclip_name(last_i, i)))
            q_ready.put(None)
        except BaseException as e:                                            # noqa: BLE001
            errors.append(e)
            abort.set()

    def writer():
        try:
            while True:
                item = get(q_done)
                if item is None:
                    return
                k, n, name = item
                t0 = time.perf_counter()
                write_subclip(name, bufs[k][0][:n])
                t1 = time.perf_counter()
                t_write[0] += t1 - t0
                busy.append(("write", t0, t1))
                q_free.put(k)
        except BaseException as e:                                            # noqa: BLE001
            errors.append(e)
            abort.set()

    th_r, th_w = threading.Thread(target=reader, daemon=True), threading.Thread(target=writer, daemon=True)
    th_r.start(), th_w.start()
    names, prev, t_gpu, t_depth = [], None, 0.0, 0.0
    t_begin = time.perf_counter()
    try:
        while True:
            item = get(q_ready)
            if item is not None:
                k, n, name = item
                out, view, _ = bufs[k]
                ticket = None
                if n:
                    t0 = time.perf_counter()
                    depth = depth_for(view[:n])                               # the depth producer (timed separately)
                    t1 = time.perf_counter()
                    ticket = processor.submit_batch(view[:n], depth, out[:n], scaler=scaler)
                    t_depth += t1 - t0
                    busy.append(("depth", t0, t1))
                    busy.append(("submit", t1, time.perf_counter()))
            if prev is not None:                                              # sub-clip k-1 finishes while k is on the GPU
                pk, pn, pname, pticket = prev
                t0 = time.perf_counter()
                if pticket is not None:
                    processor.collect(pticket)
                t1 = time.perf_counter()
                t_gpu += t1 - t0
                busy.append(("gpu_wait", t0, t1))
                q_done.put((pk, pn, pname))
                names.append(pname)
            if item is None:
                break
            prev = (k, n, name, ticket)
        q_done.put(None)
        th_w.join()
        th_r.join()
    except BaseException:
        abort.set()
        raise
    finally:
        abort_was = abort.is_set()
        if abort_was:
            th_r.join(timeout=5), th_w.join(timeout=5)
    if errors:
        raise errors[0]
    if stats is not None:
        wall = time.perf_counter() - t_begin
        stats.update(read_s=t_read[0], write_s=t_write[0], gpu_wait_s=t_gpu, depth_s=t_depth, wall_s=wall,
                     overlap=(t_read[0] + t_write[0] + t_gpu + t_depth) / max(wall, 1e-9), intervals=list(busy))
    return names


# ---- f1: decode / encode stand-ins over OpenCV (the reference uses cv2.VideoCapture and an ffmpeg rawvideo pipe) -------
def open_video(path, begin=0):
    """`load_and_set_video` (SupportFunction.py:170-177): (cap, fps, frame count, width, height), positioned at `begin`."""
    import cv2
    cap = cv2.VideoCapture(path)
    fps = cap.get(cv2.CAP_PROP_FPS)
    n = int(cap.get(cv2.CAP_PROP_FRAME_COUNT))
    w, h = int(cap.get(cv2.CAP_PROP_FRAME_WIDTH)), int(cap.get(cv2.CAP_PROP_FRAME_HEIGHT))
    cap.set(cv2.CAP_PROP_POS_FRAMES, begin)
    return cap, fps, n, w, h


def capture_reader(cap):
    """read_frame callable over a cv2.VideoCapture positioned at the worker's `begin` (frames are read in order,
    like `cap.read()` in the loop, :222)."""
    def read(_i):
        ok, img = cap.read()
        return img if ok else None
    return read


def ffmpeg_config(width, height, fps, ffmpeg="./ffmpeg/ffmpeg"):
    """The encoder argv of `get_ffmpeg_config(..., 'cpu')` (SupportFunction.py:181-201): rawvideo rgb24 2W x H on stdin."""
    return [ffmpeg, '-y', '-f', 'rawvideo', '-vcodec', 'rawvideo', '-pix_fmt', 'rgb24', '-s', f'{2 * width}x{height}',
            '-r', str(fps), '-i', '-', '-an', '-pix_fmt', 'yuv420p', '-c:v', 'libopenh264', '-b:v', '5M', '-maxrate', '10M',
            '-bufsize', '20M']


class FfmpegPipeWriter:
    """write_subclip over the reference's wire format (PredictAndGenerate.py:241-246): one ffmpeg process per sub-clip,
    raw rgb24 frames on its stdin.  Needs an ffmpeg binary (none in this image: see Cv2SubclipWriter)."""

    def __init__(self, subclip_dir, config):
        self.dir, self.config, self.proc = subclip_dir, list(config), None

    def __call__(self, name, sbs):
        import subprocess
        if self.proc is not None:
            self.proc.wait()                                                  # :241-242
        self.proc = subprocess.Popen(self.config + [os.path.join(self.dir, name)], stdin=subprocess.PIPE)
        for frame in sbs:
            self.proc.stdin.write(memoryview(np.ascontiguousarray(frame)))
        self.proc.stdin.close()

    def close(self):
        if self.proc is not None:
            self.proc.wait()


class Cv2SubclipWriter:
    """write_subclip over cv2.VideoWriter (OpenCV's bundled libavcodec): `{SubClipDir}{name}` with 2W x H frames at the
    source fps, RGB -> BGR on the way in.  Stands in for the ffmpeg pipe where no ffmpeg binary exists."""

    def __init__(self, subclip_dir, fps, fourcc="mp4v"):
        self.dir, self.fps, self.fourcc = subclip_dir, float(fps), fourcc

    def __call__(self, name, sbs):
        import cv2
        n, h, w2, _ = sbs.shape
        wr = cv2.VideoWriter(os.path.join(self.dir, name), cv2.VideoWriter_fourcc(*self.fourcc), self.fps, (w2, h))
        if not wr.isOpened():
            raise RuntimeError(f"cannot open a {self.fourcc} writer for {name}")
        for frame in sbs:
            wr.write(np.ascontiguousarray(frame[:, :, ::-1]))
        wr.release()


def count_frames(path):
    """What Check_Clips reads per file (Check_Clips.py:21-22)."""
    import cv2
    return int(cv2.VideoCapture(path).get(cv2.CAP_PROP_FRAME_COUNT))
