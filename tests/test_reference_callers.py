"""The callers of the hot path against the REFERENCE'S OWN CODE, not a restatement: the
unmodified `nibba_woka`, `Check_Clips.Checkin` and `Combine_Clips.combine_clips` are imported (from /root/reference, or
the staged copy oracle/_ref) and run with their I/O replaced by recorders - video capture, the SbsProcessor, the ffmpeg
subprocesses - so that their control flow, file naming and command lines can be compared with `worker.py`."""
import argparse
import os
import sys
import types

import numpy as np
import pytest

from oracle import ref_driver
from vr_video_generator_b200 import worker

pytestmark = pytest.mark.skipif(not ref_driver.reference_available(), reason="reference not present / not staged")


class _Cap:
    def __init__(self, length, missing, H, W, begin):
        self.length, self.missing, self.H, self.W, self.pos = length, missing, H, W, begin

    def read(self):
        i = self.pos
        self.pos += 1
        if i in self.missing or i >= self.length:
            return False, None
        f = np.zeros((self.H, self.W, 3), np.uint8)
        f[..., 0], f[..., 1], f[..., 2] = i % 251, 7, 9              # BGR: index in B
        return True, f


class _Popen:
    log = []

    def __init__(self, argv, stdin=None):
        self.argv, self.frames = list(argv), []
        _Popen.log.append(self)
        outer = self

        class _In:
            def write(self, b):
                outer.frames.append(bytes(b))

            def close(self):
                pass
        self.stdin = _In()

    def wait(self):
        return 0


def _run_reference_loop(begin, end, length, max_count, missing, H=2, W=3):
    """The reference's nibba_woka, unmodified, with capture / processor / ffmpeg replaced by recorders."""
    PAG = ref_driver.load_reference("cpu")
    calls = []

    class _Sbs:
        def __init__(self, *a, **k):
            calls.append("init")

        def add_frame(self, img, jq, rq):
            calls.append(("add", int(img[0, 0, 2])))

        def left_side_sbs(self, img, jq, rq):
            calls.append(("sbs", int(img[0, 0, 2])))
            return np.concatenate([img, img], axis=1)
    saved = {k: getattr(PAG, k) for k in ("redirrect_stdout", "load_and_set_video", "SbsProcessor", "get_length", "print_flush",
                                          "subprocess", "random_sleep")}
    _Popen.log = []
    try:
        PAG.redirrect_stdout = lambda path: None
        PAG.print_flush = lambda *a, **k: None
        PAG.get_length = lambda path: 0.0
        PAG.random_sleep = lambda *a, **k: None                          # the error handler sleeps 9-10 s (:271)
        PAG.load_and_set_video = lambda path, b: (_Cap(length, set(missing), H, W, b), 30.0, length, W, H)
        PAG.SbsProcessor = _Sbs
        PAG.subprocess = types.SimpleNamespace(Popen=_Popen, PIPE=-1)
        args = argparse.Namespace(DebugDir="/tmp/", SubClipDir="SUB/", VideoDir="x.mp4", Max_Frame_Count=max_count)
        rc = PAG.nibba_woka(begin, end, None, None, [None, 0], args, ["ffmpeg", "-cfg"])
        assert rc == 0
    finally:
        for k, v in saved.items():
            setattr(PAG, k, v)
    out = []
    for p in _Popen.log:
        assert p.argv[:2] == ["ffmpeg", "-cfg"] and p.argv[2].startswith("SUB/")
        frames = [np.frombuffer(b, np.uint8).reshape(H, 2 * W, 3) for b in p.frames]
        out.append((p.argv[2][len("SUB/"):], frames))
    return out, calls


@pytest.mark.parametrize("pipelined", [True, False])
@pytest.mark.parametrize("begin,end,length,max_count,missing", [
    (0, 40, 100, 15, ()), (7, 23, 100, 15, (9,)), (0, 1000, 33, 15, ()), (5, 6, 100, 15, ()), (0, 31, 31, 15, ()),
    (0, 16, 100, 15, ()), (10, 41, 100, 5, (10, 40)), (3, 19, 19, 15, ()), (0, 46, 46, 15, ())])
def test_worker_loop_equals_the_references_nibba_woka(begin, end, length, max_count, missing, pipelined):
    """Same sub-clip names, same frames in the same order, same black-frame substitution as the reference's loop."""
    from test_worker_logic import _FakeProcessor
    H, W = 2, 3
    want, calls = _run_reference_loop(begin, end, length, max_count, missing, H, W)
    cap = _Cap(length, set(missing), H, W, begin)
    got = []
    args = argparse.Namespace(Max_Frame_Count=max_count)
    names = worker.sbs_worker(begin, end, lambda i: cap.read()[1], lambda rgb: np.zeros((len(rgb), H, W), np.float16),
                              lambda n, sbs: got.append((n, sbs.copy())), args, length, H, W, processor=_FakeProcessor(),
                              pipelined=pipelined)
    assert names == [n for n, _ in want]
    assert [n for n, _ in got] == names
    for (n, sbs), (_, frames) in zip(got, want):
        assert len(sbs) == len(frames)
        for k in range(len(frames)):
            assert np.array_equal(sbs[k], frames[k]), (n, k)
    # the reference warps every frame exactly once, in order, each after its add_frame
    warped = [v for c, v in (x for x in calls if x != "init") if c == "sbs"]
    assert len(warped) == min(end, length) - begin
    # ... and shard.flush_ranges names the same files with the same frame counts
    from vr_video_generator_b200 import shard
    assert [(worker.subclip_name(a, b), n) for a, b, n in shard.flush_ranges(begin, end, length, max_count)] == \
           [(n, len(f)) for n, f in want]


def _import_tool(name):
    ref_driver.load_reference("cpu")                                   # puts the reference root on sys.path
    argv = sys.argv
    sys.argv = [name + ".py"]                                          # Check_Clips parses its CLI at import time
    try:
        sys.modules.pop(name, None)
        return __import__(name)
    finally:
        sys.argv = argv


def test_check_and_repair_equal_the_references_checkin(tmp_path, monkeypatch):
    """Check_Clips.Checkin (unmodified) on a directory of real (tiny) video files with one short file and one gap: the
    issues it prints, the repair commands it runs and the files it removes are what worker.check_subclips /
    worker.repair_plan compute from (name, frame count) pairs."""
    import cv2
    CC = _import_tool("Check_Clips")
    d = str(tmp_path) + "/"
    spec = {"0_15.mp4": 16, "16_30.mp4": 12, "31_45.mp4": 15, "60_74.mp4": 15, "75_80.mp4": 6, "notes.txt": 0}
    for name, n in spec.items():
        if not name.endswith("4"):
            open(d + name, "w").close()
            continue
        wr = cv2.VideoWriter(d + name, cv2.VideoWriter_fourcc(*"mp4v"), 30.0, (32, 16))
        assert wr.isOpened()
        for i in range(n):
            wr.write(np.full((16, 32, 3), 10 * i % 255, np.uint8))
        wr.release()
    entries = [(n, worker.count_frames(d + n)) for n in spec if n.endswith("4")]
    assert dict(entries) == {n: c for n, c in spec.items() if n.endswith("4")}
    ran, removed = [], []
    monkeypatch.setattr(CC.subprocess, "run", lambda argv, **k: ran.append(list(argv)))
    monkeypatch.setattr(CC.os, "remove", lambda path: removed.append(os.path.basename(path)))
    CC.Checkin(d, repair_mode=1)
    cmds, rem = worker.repair_plan(entries)
    assert cmds == ran and rem == removed
    assert ran == [["python", "PredictAndGenerate.py", "--SubClipDir", "D:/TEMP/FixxingSubclip/", "--Num_Workers", "2",
                    "--start_frame", "16", "--end_frame", "31", "--repair_mode", "1"],
                   ["python", "PredictAndGenerate.py", "--SubClipDir", "D:/TEMP/FixxingSubclip/", "--Num_Workers", "2",
                    "--start_frame", "46", "--end_frame", "60", "--repair_mode", "1"]]
    issues = worker.check_subclips(entries)
    assert ("length", "16_30.mp4", 15, 12) in issues and ("continuity", "31_45.mp4", "60_74.mp4", 45, 60) in issues and len(issues) == 2
    # repair_mode 0 only reports (and still removes the short file, Check_Clips.py:32)
    ran.clear(), removed.clear()
    CC.Checkin(d, repair_mode=0)
    assert ran == [] and removed == ["16_30.mp4"]


@pytest.mark.parametrize("just_combine", [0, 1])
def test_concat_list_and_commands_equal_the_references_combine_clips(tmp_path, monkeypatch, just_combine):
    """Combine_Clips.combine_clips (unmodified) with ffmpeg replaced by a recorder: the concat list it writes and the
    command lines it runs are worker.concat_list_lines / worker.combine_commands."""
    CB = _import_tool("Combine_Clips")
    d = str(tmp_path / "clips") + "/"
    os.makedirs(d)
    names = ["16_30.mp4", "0_15.mp4", "100_114.mp4", "31_45.mp4", "notes.txt", "9_9.mp4"]
    for n in names:
        open(d + n, "w").close()
    ran, listing = [], []

    def fake_run(argv, **k):
        ran.append(list(argv))
        if "concat" in argv:
            listing.append(open(argv[argv.index("-i") + 1]).read())
    monkeypatch.setattr(CB.subprocess, "run", fake_run)
    monkeypatch.chdir(tmp_path)
    CB.combine_clips(d, "orig.mp4", "out.mkv", just_combine=just_combine)
    assert listing == ["".join(worker.concat_list_lines(d, names))]
    assert ran == worker.combine_commands(d, "orig.mp4", "out.mkv", just_combine)
    assert worker.order_subclips(names) == ["0_15.mp4", "9_9.mp4", "16_30.mp4", "31_45.mp4", "100_114.mp4"]   # int("0_15") = 15 < int("9_9") = 99


def test_run_plan_and_fanout_equal_the_references_main():
    """repair_mode branches (PredictAndGenerate.py:402-419) and main_func's worker fan-out (:274-306), read off the
    reference source: the wipe / run / combine decisions and the (range, inference worker, queue slot) assignment."""
    src = open(os.path.join(ref_driver.REFERENCE_ROOT, "PredictAndGenerate.py")).read()
    for frag in ("if (repair_mode != 1): #Continue mode   \n        remove_all_file (DebugDir)", "if (repair_mode == 0):\n        remove_all_file (SubClipDir)",
                 "if (repair_mode in [0, 1]):\n        main_func(args)", "if (repair_mode in [0, 2]):", "if (repair_mode in [3]):",
                 "gpu_worker_number = idx % Num_GPU_Workers", "within_gpu_worker_inference_idx = int(idx/Num_GPU_Workers)"):
        assert frag in src, frag
    assert worker.run_plan(0) == dict(wipe_debug=True, wipe_subclips=True, run_workers=True, combine=True, just_combine=0)
    assert worker.run_plan(1) == dict(wipe_debug=False, wipe_subclips=False, run_workers=True, combine=False, just_combine=0)
    assert worker.run_plan(2) == dict(wipe_debug=True, wipe_subclips=False, run_workers=False, combine=True, just_combine=0)
    assert worker.run_plan(3) == dict(wipe_debug=True, wipe_subclips=False, run_workers=False, combine=True, just_combine=1)
    # 6 SBS workers on 2 inference workers: worker idx -> (idx % 2, idx // 2), ranges as main_func splits them
    fan = worker.worker_ranges(0, 10 ** 14, 100, 6, 2)
    assert [(g, s) for _, _, g, s in fan] == [(0, 0), (1, 0), (0, 1), (1, 1), (0, 2), (1, 2)]
    assert [(b, e) for b, e, _, _ in fan] == [(0, 17), (17, 34), (34, 51), (51, 68), (68, 85), (85, 102)]
