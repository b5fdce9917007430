"""The asynchronous host pipeline (vrsbs_submit_host / vrsbs_collect) and the pipelined worker loop with real OpenCV
decode and encode.  Needs a B200: run with `-m gpu` under gpurun."""
import argparse
import os
import time

import numpy as np
import pytest

from conftest import golden_weights, load_case

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
from oracle import sbs_layered as O  # noqa: E402


def _args(fg=0.025, bg=-0.015, step=1, **kw):
    return argparse.Namespace(offset_fg=fg, offset_bg=bg, offset_step_size=step, **kw)


def test_submit_collect_equals_blocking_call():
    """Three batches in flight (one decoded in place into the right halves of its SBS buffer, one with DPT-resolution
    depth, one with the depth left on the device) give byte for byte what the blocking call gives; the clip state
    carries across submissions in order."""
    import vr_video_generator_b200 as pkg
    from vr_video_generator_b200 import synth
    from vr_video_generator_b200.sbs import pinned_sbs_buffer
    H, W, n = 270, 480, 7
    frames = synth.frames_noise(3 * n, H, W, seed=31)
    lo = synth.depth_scene(3 * n, 74, 132, seed=31)
    raw = np.stack([O.bicubic_resize(lo[t], H, W, 1.0) for t in range(3 * n)])
    ref = pkg.SbsProcessor(None, 0, _args(), max_batch=4)
    want = ref.left_side_sbs_batch(frames, raw)
    ref.close()
    proc = pkg.SbsProcessor(None, 0, _args(), max_batch=4)
    bufs = [pinned_sbs_buffer(n, H, W) for _ in range(3)]
    tickets = []
    # batch 0: frames decoded straight into the right halves (in place), host depth
    np.copyto(bufs[0][1], frames[:n])
    tickets.append(proc.submit_batch(bufs[0][1], raw[:n], bufs[0][0]))
    # batch 1: separate pinned frames (packed), pinned depth tensor
    f1 = torch.from_numpy(frames[n:2 * n].copy()).pin_memory()
    d1 = torch.from_numpy(raw[n:2 * n].copy()).pin_memory()
    tickets.append(proc.submit_batch(f1.numpy(), d1, bufs[1][0]))
    # batch 2: in place, depth left on the device by a producer on another stream
    np.copyto(bufs[2][1], frames[2 * n:])
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        d2 = torch.from_numpy(raw[2 * n:].copy()).cuda(non_blocking=False) * 1.0
    with torch.cuda.stream(side):
        tickets.append(proc.submit_batch(bufs[2][1], d2, bufs[2][0]))
    for k in (1, 0, 2):                                   # any collection order
        proc.collect(tickets[k])
    for k in range(3):
        got = bufs[k][0]
        assert np.array_equal(got, want[k * n:(k + 1) * n]), (k, int((got != want[k * n:(k + 1) * n]).sum()))
    with pytest.raises(Exception):
        proc.collect(tickets[0])                          # already collected
    proc.close()


def test_submit_reports_rejected_frames_per_batch():
    """A NaN depth frame (the reference raises in math.ceil) fails the batch that contains it at collect; the batches
    around it are delivered."""
    import vr_video_generator_b200 as pkg
    from vr_video_generator_b200 import _native
    from vr_video_generator_b200.sbs import pinned_sbs_buffer
    H, W, n = 96, 160, 5
    rng = np.random.default_rng(3)
    frames = rng.integers(0, 256, size=(3 * n, H, W, 3), dtype=np.uint8)
    raw = (rng.random((3 * n, H, W)) * 12).astype(np.float16)
    raw[n + 2, 5, 7] = np.float16("nan")
    proc = pkg.SbsProcessor(None, 0, _args(), max_batch=4)
    bufs = [pinned_sbs_buffer(n, H, W) for _ in range(3)]
    tickets = []
    for k in range(3):
        np.copyto(bufs[k][1], frames[k * n:(k + 1) * n])
        tickets.append(proc.submit_batch(bufs[k][1], raw[k * n:(k + 1) * n], bufs[k][0]))
    proc.collect(tickets[0])
    with pytest.raises(_native.VrsbsError) as e:
        proc.collect(tickets[1])
    assert e.value.code == -4 and "NaN" in str(e.value)
    proc.collect(tickets[2])
    assert np.array_equal(bufs[0][0][:, :, W:], frames[:n])
    # pageable buffers are refused by submit (process_host stages them)
    with pytest.raises(_native.VrsbsError):
        proc._ctx.submit_host(frames.ctypes.data, 0, 0, raw.ctypes.data, 2, H, W, 0, 0, 1.0, np.empty((2, H, 2 * W, 3), np.uint8).ctypes.data)
    proc.close()


def test_pipelined_worker_with_real_decode_and_encode(tmp_path, oracle_lib):
    """nibba_woka's loop end to end with real I/O: a lossless FFV1 clip is decoded by cv2.VideoCapture, warped, and every
    sub-clip encoded by cv2.VideoWriter (OpenCV's libavcodec; this image has no ffmpeg binary).  The pipelined loop
    (reader thread -> GPU -> writer thread) delivers the frames of the serial loop and of the per-frame oracle byte for
    byte, names the files like the reference, every file holds the frame count Check_Clips expects - and decode, GPU and
    encode intervals of different sub-clips overlap in time."""
    import cv2
    from vr_video_generator_b200 import worker
    meta, frames, raw, _ = load_case("medium")
    p = meta["params"]
    H, W = p["H"], p["W"]
    n = 26
    fr = np.concatenate([frames] * 7)[:n].copy()
    for i in range(n):
        fr[i, :8, :8] = i * 9                              # make every frame distinct
    rw = np.concatenate([raw] * 7)[:n]
    src = str(tmp_path / "clip.avi")
    wr = cv2.VideoWriter(src, cv2.VideoWriter_fourcc(*"FFV1"), 30.0, (W, H))
    assert wr.isOpened()
    for i in range(n):
        wr.write(np.ascontiguousarray(fr[i][:, :, ::-1]))  # the file holds BGR
    wr.release()
    sub = str(tmp_path / "sub") + "/"
    os.makedirs(sub)

    def depth_for(rgb):
        out = []
        for f in rgb:
            k = next(i for i in range(n) if np.array_equal(f, fr[i]))   # lossless decode: every frame is found
            out.append(rw[k])
        time.sleep(0.01 * len(rgb))                       # stands in for the depth producer's forward
        return np.stack(out)

    def run(pipelined):
        cap, fps, length, w, h = worker.open_video(src, 0)
        assert (length, w, h) == (n, W, H)
        enc = worker.Cv2SubclipWriter(sub, fps)
        kept, stats = {}, {}
        read = worker.capture_reader(cap)

        def slow_read(i):
            time.sleep(0.004)
            return read(i)

        def write(name, sbs):
            kept[name] = sbs.copy()
            enc(name, sbs)
            time.sleep(0.03)
        args = _args(p["fg"], p["bg"], p["step"], Max_Frame_Count=6)
        names = worker.sbs_worker(0, 10 ** 9, slow_read, depth_for, write, args, length, H, W, pipelined=pipelined, stats=stats)
        return names, kept, stats

    run(True)                                              # warm-up: CUDA context, page-locked ring, codec start-up are not pipeline time
    names, kept, stats = run(True)
    assert names == ["0_6.mp4", "7_12.mp4", "13_18.mp4", "19_24.mp4", "25_25.mp4"]
    # Check_Clips' own arithmetic (Check_Clips.py:23-28) on these files: every file holds what its name promises except the
    # FIRST file of a worker, which the reference's loop names "0_6" while it holds frames 0..5 (frame 6 was read ahead but
    # not yet warped, :226-231) - the reference's checker flags its own first sub-clip the same way
    assert worker.check_subclips([(nm, worker.count_frames(sub + nm)) for nm in names]) == [("length", "0_6.mp4", 7, 6)]
    assert [len(kept[nm]) for nm in names] == [6, 6, 6, 6, 2]
    names2, kept2, _ = run(False)
    assert names2 == names
    st = O.WarpState(p["fg"], p["bg"], p["step"])
    w = golden_weights(meta)
    out = np.concatenate([kept[nm] for nm in names])
    out2 = np.concatenate([kept2[nm] for nm in names])
    assert np.array_equal(out, out2)
    for i in range(n):
        assert np.array_equal(out[i], oracle_lib.process_frame(st, fr[i], rw[i], weights=w)), i
    # concurrency: the reader decodes ahead while the producer stand-in works on an earlier sub-clip, and the writer encodes a
    # sub-clip while the producer works on a later one (both follow from the pipeline's structure: the depth call of sub-clip k
    # runs on the calling thread between the reader's hand-over of k and the writer's receipt of k); summed busy time exceeds
    # the wall clock.  These are wall-clock observations on a shared box: up to three attempts, the byte-level checks above
    # hold for every one of them.
    def concurrent(st):
        iv = st["intervals"]

        def overlaps(a, b):
            return any(x[1] < y[2] and y[1] < x[2] for x in iv if x[0] == a for y in iv if y[0] == b)
        return overlaps("read", "depth") and overlaps("write", "depth") and st["overlap"] > 1.05
    ok = concurrent(stats)
    for _ in range(2):
        if ok:
            break
        names3, kept3, stats = run(True)
        assert names3 == names and all(np.array_equal(kept3[nm], kept[nm]) for nm in names)
        ok = concurrent(stats)
    assert ok, stats
    # the encoded files decode to the right size and (lossy mp4v) roughly the right content
    cap = cv2.VideoCapture(sub + names[1])
    ok, img = cap.read()
    assert ok and img.shape == (H, 2 * W, 3)
    assert np.mean(np.abs(img[:, :, ::-1].astype(int) - kept[names[1]][0].astype(int))) < 12


def test_mixed_entry_styles_share_the_clip_state(oracle_lib):
    """Device-pointer calls (caller's stream) and host calls (the library's streams) alternate on one context: the host call
    orders itself behind the device work (depth history, range EMA), and a device-pointer call is refused while a submitted
    batch has not been collected.  The frames equal the oracle's walk over the clip."""
    import argparse
    import vr_video_generator_b200 as pkg
    from vr_video_generator_b200 import _native
    from vr_video_generator_b200.sbs import pinned_sbs_buffer
    meta, frames, raw, _ = load_case("medium")
    p = meta["params"]
    H, W, n = p["H"], p["W"], p["n"]
    w = golden_weights(meta)
    proc = pkg.SbsProcessor(None, 0, argparse.Namespace(offset_fg=p["fg"], offset_bg=p["bg"], offset_step_size=p["step"]), device=0, max_batch=4)
    proc._context(H, W, False).set_blur_weights(w)
    st = O.WarpState(p["fg"], p["bg"], p["step"])
    want = [oracle_lib.process_frame(st, frames[t], raw[t], weights=w) for t in range(n)]
    side = torch.cuda.Stream()
    outs = []
    for t in range(n):
        if t % 2 == 0:                                        # device-pointer call on a side stream, not synchronised by the caller
            with torch.cuda.stream(side):
                f = torch.from_numpy(frames[t:t + 1]).cuda()
                r = torch.from_numpy(raw[t:t + 1]).cuda()
                o = proc.warp_batch_device(f, r)
            outs.append(o)
        else:                                                 # host call right behind it
            outs.append(proc.left_side_sbs_batch(frames[t:t + 1], raw[t:t + 1]))
    torch.cuda.synchronize()
    for t in range(n):
        got = outs[t].cpu().numpy()[0] if isinstance(outs[t], torch.Tensor) else outs[t][0]
        assert np.array_equal(got, want[t]), t
    # a submitted batch blocks the device-pointer entry until it is collected
    proc.reset_state()
    out, view, _keep = pinned_sbs_buffer(1, H, W)
    np.copyto(view, frames[:1])
    ticket = proc.submit_batch(view, torch.from_numpy(raw[:1]).pin_memory(), out)
    with pytest.raises(_native.VrsbsError):
        proc.warp_batch_device(torch.from_numpy(frames[:1]).cuda(), torch.from_numpy(raw[:1]).cuda())
    proc.collect(ticket)
    proc.close()


def test_many_batches_through_the_pipeline_stay_exact(oracle_lib):
    """Soak: 90 batches of varying length through submit / collect (three scratch sets in rotation, two batches in flight,
    clip state reset in between) and through the device-pointer call on one context, with the 1080p blur footprint (11 x 9:
    the band-driven blur); every batch equals the oracle's frames.  Guards the per-batch bookkeeping that must return to its
    initial state (work counters, the blur's band map)."""
    import argparse
    import vr_video_generator_b200 as pkg
    from vr_video_generator_b200.sbs import pinned_sbs_buffer
    meta, frames, raw, ref_left = load_case("medium")
    p = meta["params"]
    H, W, n = p["H"], p["W"], p["n"]
    proc = pkg.SbsProcessor(None, 0, argparse.Namespace(offset_fg=p["fg"], offset_bg=p["bg"], offset_step_size=p["step"]), device=0, max_batch=8)
    w = O.gaussian_weights(11, 9)
    proc._context(H, W, False).set_blur_weights(w)
    st = O.WarpState(p["fg"], p["bg"], p["step"])
    ref_left = np.stack([oracle_lib.process_frame(st, frames[t], raw[t], weights=w) for t in range(n)])[:, :, :W]
    ring = [pinned_sbs_buffer(n, H, W) for _ in range(2)]
    d_pin = torch.from_numpy(np.ascontiguousarray(raw)).pin_memory()
    f_dev, r_dev = torch.from_numpy(frames).cuda(), torch.from_numpy(raw).cuda()
    pending = None
    for it in range(90):
        k = 1 + it % n                                         # batch of the clip's first k frames
        if it % 5 == 4:                                        # every fifth batch through the device-pointer entry
            if pending is not None:
                t, buf, kk = pending
                proc.collect(t)
                assert np.array_equal(buf[0][:kk, :, :W], ref_left[:kk]), it
                pending = None
            proc.reset_state()
            out = proc.warp_batch_device(f_dev[:k], r_dev[:k], check=True)
            assert np.array_equal(out.cpu().numpy()[:, :, :W], ref_left[:k]), it
            continue
        buf = ring[it & 1]
        if pending is not None:                                # the clip state is per batch here: collect before the reset
            t, pbuf, kk = pending
            proc.collect(t)
            assert np.array_equal(pbuf[0][:kk, :, :W], ref_left[:kk]), it
            assert np.array_equal(pbuf[0][:kk, :, W:], frames[:kk]), it
        proc.reset_state()
        np.copyto(buf[1][:k], frames[:k])
        pending = (proc.submit_batch(buf[1][:k], d_pin[:k], buf[0][:k]), buf, k)
    if pending is not None:
        t, buf, kk = pending
        proc.collect(t)
        assert np.array_equal(buf[0][:kk, :, :W], ref_left[:kk])
    proc.close()
