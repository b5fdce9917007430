"""Host-side python logic of the product (tables.py, synth.py) — CPU only."""
import math

import numpy as np
import pytest

from conftest import F32_CASES, FULL_CASES, SMALL_CASES, load_meta, unhex
from oracle import sbs_layered as O
from vr_video_generator_b200 import synth, tables


@pytest.mark.parametrize("name", SMALL_CASES + F32_CASES + FULL_CASES)
def test_layer_tables_match_reference_lists(name):
    """T1 on the host mirror: exact doubles against the lists the reference's get_cutoff returned."""
    meta = load_meta(name)
    p = meta["params"]
    last = None
    for fm in meta["frames"]:
        cut, rng, steps, limit, offs = tables.layer_tables(fm["depth_max"], p["H"], p["fg"], p["bg"], p["step"], last)
        last = rng
        assert [float(c) for c in cut] == unhex(fm["cutoffs"])
        assert [float(s) for s in steps] == unhex(fm["steps"])
        assert offs == fm["offsets"] and limit == fm["limit"]
        assert [float(r) for r in rng] == unhex(fm["range"])
        assert tables.strip_columns(offs[-1], p["W"]) == fm["strip"]
        assert tables.fill_layer(len(steps)) == fm["fill_layer"]


def test_tables_agree_with_oracle_on_a_sweep():
    rng = np.random.default_rng(3)
    for _ in range(300):
        fg, bg = float(rng.uniform(-0.1, 0.3)), float(rng.uniform(-0.2, 0.1))
        step, H = int(rng.integers(1, 5)), int(rng.integers(16, 2200))
        st = O.WarpState(fg, bg, step)
        last = None
        for _ in range(3):
            dmax = float(np.float16(rng.uniform(-0.5, 20)))
            a = O.layer_tables(st, dmax, H)
            b = tables.layer_tables(dmax, H, fg, bg, step, last)
            last = b[1]
            assert list(a[0]) == list(b[0]) and list(a[1]) == list(b[2]) and list(a[2]) == list(b[4])
            assert a[3] == b[3] and a[4] == b[1]


def test_known_answers():
    # EMA example of SURVEY.md section 8a: previous [-10,20] + current [-16.2,27] -> [-13.1,23.5]
    cut, rng, steps, limit, offs = tables.layer_tables(14.0, 1080, 0.025, -0.015, 1, [-10.0, 20.0])
    assert limit == 14 and rng == [(-10.0 + -0.015 * 1080 * 14 / 14) / 2, (20.0 + 0.025 * 1080 * 14 / 14) / 2]
    assert offs == sorted(offs) and len(set(offs)) == len(offs)          # strictly increasing
    assert cut[0] == 0 and cut[-1] == 14 and len(cut) == len(steps) + 1
    # step 2 restarts the positive side at 1 (PredictAndGenerate.py:113)
    _, _, _, _, offs2 = tables.layer_tables(14.0, 1080, 0.025, -0.01, 2, None)
    assert offs2[:7] == [-11, -9, -7, -5, -3, -1, 0] and offs2[7:10] == [1, 3, 5]
    # max == 0: one empty layer, offset 0, no strip
    cut, rng, steps, limit, offs = tables.layer_tables(0.0, 1080, 0.025, -0.01, 1, None)
    assert (cut, steps, limit, offs) == ([0, 0], [0], 0, [0]) and tables.strip_columns(offs[-1], 1920) == 0
    # python round is half-to-even; slices with a negative end
    assert round(26.5) == 26 and round(27.5) == 28
    assert tables.strip_columns(26, 1920) == round(26 / 3 * 2) == 17
    assert tables.strip_columns(-3, 100) == 98 and tables.strip_columns(-300, 100) == 0
    assert tables.strip_columns(500, 100) == 100
    assert tables.fill_layer(43) == 25 and tables.fill_layer(1) == 0
    assert tables.blur_kernel_shape(1080) == (11, 9) and tables.blur_kernel_shape(2160) == (19, 17)
    assert tables.blur_kernel_shape(270) == (5, 3) and tables.blur_kernel_shape(100) == (3, 1)
    w_now, taps = tables.smoothing_weights()
    assert taps == [0.3, 0.3 * 0.4] and w_now == 1 - (0.3 + 0.3 * 0.4)


def test_sign_fixup_and_clip_ranges():
    assert tables.fix_offset_signs(0.025, -0.01) == (0.025, -0.01)
    assert tables.fix_offset_signs(0.025, 0.01) == (0.025, -0.01)
    assert tables.fix_offset_signs(-0.025, -0.01) == (0.025, -0.01)
    assert tables.fix_offset_signs(-0.025, 0.01) == (-0.025, 0.01)       # opposite signs: left alone
    # PredictAndGenerate.py:274-275,303
    assert tables.clip_ranges(0, 10 ** 14, 100, 4) == [(0, 25), (25, 50), (50, 75), (75, 100)]
    assert tables.clip_ranges(0, 10 ** 14, 10, 4) == [(0, 3), (3, 6), (6, 9), (9, 12)]
    assert tables.clip_ranges(5, 50, 1000, 2) == [(5, 28), (28, 50)]
    assert tables.clip_ranges(0, 10, 10, 16) == [(i, i + 1) for i in range(10)]
    assert tables.clip_ranges(10, 5, 100, 2) == []


def test_gaussian_weights_are_torchvisions():
    import torch
    from torchvision.transforms.v2.functional import gaussian_blur
    for (kx, ky) in [(11, 9), (19, 17), (5, 3), (3, 1)]:
        w = tables.gaussian_weights(kx, ky, 3.0)
        assert w.shape == (ky, kx) and w.dtype == np.float32
        # an impulse through torchvision's own blur reproduces the kernel exactly
        img = torch.zeros(1, 41, 41, dtype=torch.float32)
        img[0, 20, 20] = 1.0
        got = gaussian_blur(img, (kx, ky), sigma=3.0)[0, 20 - ky // 2:21 + ky // 2, 20 - kx // 2:21 + kx // 2].numpy()
        assert np.array_equal(got[::-1, ::-1], w)
        assert np.array_equal(w, O.gaussian_weights(kx, ky))


def test_synth_is_deterministic_and_shaped():
    a, b = synth.frames_noise(2, 8, 16, 3), synth.frames_noise(2, 8, 16, 3)
    assert np.array_equal(a, b) and a.dtype == np.uint8 and a.shape == (2, 8, 16, 3)
    d = synth.depth_stress(3, 40, 60, seed=1)
    assert d.dtype == np.float16 and d.shape == (3, 40, 60) and d.min() >= 0
    assert np.array_equal(d[0][:, 2:], d[1][:, :-2])                  # 2 px/frame drift
    s = synth.depth_scene(2)
    assert s.shape == (2, synth.DPT_H, synth.DPT_W) and float(s.max()) == pytest.approx(13.9, abs=0.01)
    assert math.ceil(float(s.max())) == 14


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the unmodified reference on the host cores; the numpy port only where no copy of the
    reference exists) runs without a GPU and prints one JSON line with the keys the driver reads; under a multi-rank
    launch only rank 0 works."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
           "--cpu-budget", "1", "--workload", "1080p_b16_cfg1"]
    env = dict(os.environ, RANK="0", WORLD_SIZE="1")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env, cwd=root)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "sbs_frames_per_sec_warp_stage" and d["unit"] == "frames/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 1
    from oracle import ref_driver
    kind = "reference" if ref_driver.reference_available() else "port"
    assert d["cpu_baseline"]["kind"] == kind and d["cpu_baseline"]["cores"] >= 1 and "frames" in d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["gpu_launches"] == 0
    if kind == "reference":                                 # a box without any copy of the reference: the port
        outp = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(env, VRSBS_NO_REFERENCE="1"), cwd=root)
        assert outp.returncode == 0, outp.stderr[-2000:]
        dp = json.loads([l for l in outp.stdout.splitlines() if l.startswith("{")][0])
        assert dp["cpu_baseline"]["kind"] == "port" and dp["value"] > 0
    # any other rank exits 0 without output
    out1 = subprocess.run(cmd, capture_output=True, text=True, timeout=120, env=dict(env, RANK="1", WORLD_SIZE="2"), cwd=root)
    assert out1.returncode == 0 and out1.stdout.strip() == ""
