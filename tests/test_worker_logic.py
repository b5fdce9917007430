"""Callers of the hot path (SURVEY.md section 8 f): CLI flags, sub-clip naming / ordering / checks and the worker
loop's control flow, against a restatement of the reference loop (PredictAndGenerate.py:221-250)."""
import argparse

import numpy as np
import pytest

from vr_video_generator_b200 import worker


def test_cli_defaults_and_offsets_pass_through():
    args, rest = worker.parse_args([])
    # PredictAndGenerate.py:328-357
    assert (args.DebugDir, args.SubClipDir, args.OutputDir) == ("./Debug/", "./Subclip/", "DeleteThis.mkv")
    assert (args.encoder, args.offset_fg, args.offset_bg, args.offset_step_size) == ("vits", 0.025, -0.01, 1)
    assert (args.Num_Workers, args.num_gpu, args.Num_GPU_Workers, args.Max_Frame_Count) == (4, 1, 1, 15)
    assert (args.start_frame, args.end_frame, args.repair_mode) == (0, 99999999999999, 0)
    assert rest == []
    # unknown flags are discarded, not an error (:365)
    args, rest = worker.parse_args(["--offset_fg", "0.05", "--bogus", "1"])
    assert args.offset_fg == 0.05 and rest == ["--bogus", "1"]
    # same-sign offsets reach the warp as typed: the reference's fix-up (:387-393) rebinds module-level names only,
    # main_func(args) (:411) passes the untouched Namespace and SbsProcessor reads args_god.offset_* (:92-94)
    a, _ = worker.parse_args(["--offset_fg", "0.03", "--offset_bg", "0.01"])
    assert (a.offset_fg, a.offset_bg) == (0.03, 0.01)
    a, _ = worker.parse_args(["--offset_fg", "-0.03", "--offset_bg", "-0.01"])
    assert (a.offset_fg, a.offset_bg) == (-0.03, -0.01)
    assert worker.encoder_scaler("vits") == 1.618 and worker.encoder_scaler("vitg") == 1


def test_subclip_order_and_checks():
    names = ["16_30.mp4", "0_15.mp4", "100_114.mp4", "31_45.mp4", "notes.txt", "input_list.txt"]
    assert worker.order_subclips(names) == ["0_15.mp4", "16_30.mp4", "31_45.mp4", "100_114.mp4"]
    assert worker.subclip_sort_key("15_29.mp4") == 1529
    ok = [("0_15.mp4", 16), ("16_30.mp4", 15), ("31_45.mp4", 15)]
    assert worker.check_subclips(ok) == []
    bad = [("0_15.mp4", 14), ("16_30.mp4", 15), ("40_54.mp4", 15)]
    issues = worker.check_subclips(bad)
    assert ("length", "0_15.mp4", 16, 14) in issues and ("continuity", "16_30.mp4", "40_54.mp4", 30, 40) in issues


def _reference_loop(begin, end, video_length, max_count, missing=()):
    """Control flow of nibba_woka restated with frame indices instead of images: returns [(name, [indices])]."""
    out, frame_list, last_i, last_img = [], [], begin, None
    for i in range(begin, min(end, video_length)):
        raw = -1 if i in missing else i
        if i == begin:
            last_img = raw
        else:
            frame_list.append(last_img)
            last_img = raw
        if i == min(end, video_length) - 1:
            frame_list.append(raw)
        if len(frame_list) == max_count or i == min(end - 1, video_length - 1):
            if i == begin:
                return []              # step_taken == 0: the progress print raises, the handler returns (see shard.flush_ranges)
            out.append((f"{last_i}_{i}.mp4", frame_list))
            last_i, frame_list = i + 1, []
    return out


class _FakeProcessor:
    """Stands in for SbsProcessor: marks every frame with its index so that order and grouping can be checked."""
    max_batch = 4

    def __init__(self):
        self.calls = []

    def reset_state(self):
        self.calls.append("reset")

    def left_side_sbs_batch(self, frames, depths, scaler=1.0):
        self.calls.append(len(frames))
        assert depths.shape[0] == frames.shape[0]
        return np.concatenate([frames, frames], axis=2)

    # the asynchronous pair the pipelined loop uses: the "warp" happens at collect, like a real in-flight batch
    def submit_batch(self, frames, depths, out, scaler=1.0):
        assert depths.shape[0] == frames.shape[0] and out.shape[0] == frames.shape[0]
        assert frames.base is not None and np.shares_memory(frames, out), "frames must be decoded into the SBS buffer's right half"
        self.calls.append(len(frames))
        self.pending = getattr(self, "pending", {})
        t = len(self.calls)
        self.pending[t] = (frames, out)
        return t

    def collect(self, ticket):
        frames, out = self.pending.pop(ticket)
        W = frames.shape[2]
        out[:, :, :W] = frames


@pytest.mark.parametrize("pipelined", [True, False])
@pytest.mark.parametrize("begin,end,length,max_count,missing", [
    (0, 40, 100, 15, ()), (7, 23, 100, 15, (9,)), (0, 1000, 33, 15, ()), (5, 6, 100, 15, ()), (0, 31, 31, 15, ()),
    (0, 16, 100, 15, ()), (10, 41, 100, 5, (10, 40))])
def test_worker_control_flow_matches_reference(begin, end, length, max_count, missing, pipelined):
    H, W = 2, 3
    def read(i):
        if i in missing:
            return None
        f = np.zeros((H, W, 3), np.uint8)
        f[..., 0], f[..., 1], f[..., 2] = i % 251, 7, 9             # BGR: index in B
        return f
    got = []
    args = argparse.Namespace(Max_Frame_Count=max_count)
    names = worker.sbs_worker(begin, end, read, lambda rgb: np.zeros((len(rgb), H, W), np.float16),
                              lambda n, sbs: got.append((n, sbs.copy())), args, length, H, W, processor=_FakeProcessor(),
                              pipelined=pipelined)
    want = _reference_loop(begin, end, length, max_count, missing)
    assert names == [n for n, _ in want]
    for (n, sbs), (wn, idx) in zip(got, want):
        assert n == wn and len(sbs) == len(idx)
        for k, i in enumerate(idx):
            # RGB order after the [2,1,0] swap: R=9, G=7, B=index (black frame where the read failed)
            exp = (0, 0, 0) if i < 0 else (9, 7, i % 251)
            assert tuple(sbs[k, 0, 0]) == exp and tuple(sbs[k, 0, W]) == exp


def test_bind_near_gpu_uses_sysfs_local_cpulist(tmp_path):
    """NUMA placement helper: parses the sysfs cpulist, intersects it with the CPUs this process may use, and leaves
    the affinity alone when sysfs has no entry or too few local CPUs are available."""
    import os
    from vr_video_generator_b200 import shard
    assert shard.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert shard.parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    try:
        assert shard.bind_near_gpu("0000:ff:00.0", sysfs=str(tmp_path)) is None          # no such device
        dev = tmp_path / "0000:1b:00.0"
        dev.mkdir()
        (dev / "local_cpulist").write_text("100000-100003\n")                            # CPUs we do not have
        assert shard.bind_near_gpu("0000:1B:00.0", sysfs=str(tmp_path)) is None
        assert os.sched_getaffinity(0) == before
        if len(before) >= 3:
            keep = sorted(before)[:2]
            (dev / "local_cpulist").write_text(",".join(map(str, keep)) + "\n")
            assert shard.bind_near_gpu("0000:1b:00.0", min_cpus=2, sysfs=str(tmp_path)) == set(keep)
            assert os.sched_getaffinity(0) == set(keep)
    finally:
        os.sched_setaffinity(0, before)
