"""The oracles against the reference-generated fixtures (tests/golden, made by make_golden.py
from the UNMODIFIED reference).  CPU only.  This is what pins the oracle."""
import hashlib

import numpy as np
import pytest

from conftest import F32_CASES, FULL_CASES, SMALL_CASES, golden_weights, load_case, unhex
from oracle import sbs_layered as O


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def _check_tables(stages, fm):
    # T1: exact doubles / ints, same list lengths
    assert [float(x) for x in stages["marks"]] == unhex(fm["cutoffs"])
    assert [float(x) for x in stages["steps"]] == unhex(fm["steps"])
    assert [int(x) for x in stages["offsets"]] == fm["offsets"]
    assert stages["limit"] == fm["limit"]
    assert [float(x) for x in stages["range"]] == unhex(fm["range"])
    assert int(stages["strip"]) == fm["strip"]
    assert int(stages["fill_layer"]) == fm["fill_layer"]


@pytest.mark.parametrize("name", SMALL_CASES + F32_CASES)
@pytest.mark.parametrize("impl", ["layered", "scatter"])
def test_small_cases_bit_exact(name, impl, oracle_lib):
    meta, frames, raw, ref_left = load_case(name)
    p = meta["params"]
    mod = O if impl == "layered" else oracle_lib
    st = O.WarpState(p["fg"], p["bg"], p["step"])
    w = golden_weights(meta)
    for t in range(p["n"]):
        stages = {}
        out = mod.process_frame(st, frames[t], raw[t], weights=w, stages=stages)
        fm = meta["frames"][t]
        _check_tables(stages, fm)
        assert int(stages["holes"].sum()) == fm["holes"]
        assert np.array_equal(out[:, p["W"]:], frames[t])                 # T6: right half is the input
        assert fm["oracle_vs_reference"] == []                           # small cases: no fp32-order flips
        assert np.array_equal(out[:, :p["W"]], ref_left[t]), f"{name}[{t}] {impl} differs from the reference"
        assert _sha(out) == fm["sha256"]


@pytest.mark.parametrize("name", FULL_CASES)
def test_full_size_hashes(name, oracle_lib):
    """1080p / 4K: scatter oracle + the recorded +-1 blur flips must hash to the reference's output."""
    meta, frames, raw, _ = load_case(name)
    p = meta["params"]
    st = O.WarpState(p["fg"], p["bg"], p["step"])
    w = golden_weights(meta)
    for t in range(p["n"]):
        stages = {}
        out = oracle_lib.process_frame(st, frames[t], raw[t], weights=w, stages=stages)
        fm = meta["frames"][t]
        _check_tables(stages, fm)
        assert int(stages["holes"].sum()) == fm["holes"]
        flips = fm["oracle_vs_reference"]
        assert len(flips) <= max(2, 1e-4 * 3 * fm["blurred"]), "T5: more than 1e-4 of blurred values flip"
        patched = out.copy()
        for y, x, c, ref_v, mine_v in flips:
            assert out[y, x, c] == mine_v and abs(ref_v - mine_v) == 1
            assert stages["holes"][y, x] and x >= stages["strip"]
            patched[y, x, c] = ref_v
        assert _sha(patched) == fm["sha256"], f"{name}[{t}]: oracle (+recorded blur flips) != reference"


def test_layered_equals_scatter_on_random_tables(oracle_lib):
    """The forward-scatter restatement == the literal layer loop, including wrap-around (|off| > W),
    negative depth, NaN depth and non-monotone (sign-swapped) tables."""
    rng = np.random.default_rng(0)
    for trial in range(12):
        H, W = int(rng.integers(20, 60)), int(rng.integers(33, 130))
        fg, bg = float(rng.uniform(0.01, 0.6)), -float(rng.uniform(0.01, 0.4))
        if trial % 4 == 3:
            fg, bg = bg, fg                                   # bg > 0 > fg: the CLI does not fix this up
        step = int(rng.integers(1, 4))
        img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
        depth = (rng.random((H, W)) * 14 - 1.5).astype(np.float16)
        if trial % 5 == 0:
            depth[rng.integers(0, H), rng.integers(0, W)] = np.float16("nan")
        s1, s2 = O.WarpState(fg, bg, step), O.WarpState(fg, bg, step)
        dmax = float(np.nanmax(depth))
        t1 = O.layer_tables(s1, dmax, H)
        t2 = O.layer_tables(s2, dmax, H)
        a = O.warp_frame(img, depth, *t1[:3])
        b = oracle_lib.warp_frame(img, depth, *t2[:3])
        assert np.array_equal(a, b), f"trial {trial}"


def test_bicubic_restatement_vs_torch_cpu_fp32():
    """Depth tail (dpt.py:196): the oracle's ATen-CUDA-style bicubic against torch CPU run on the fp32
    view of the same input (see make_golden.depth_tail for why not torch's CPU fp16 kernel)."""
    import os

    from conftest import GOLDEN
    from vr_video_generator_b200 import synth
    for name in ("depth_tail_small", "depth_tail_1080p"):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        h, w, H, W, stride, seed = [int(v) for v in z["params"]]
        lo = synth.depth_stress(2, h, w, seed=seed)
        for i in range(2):
            mine = O.bicubic_resize(lo[i], H, W, 1.0)[::stride, ::stride].astype(np.float32)
            ref = z["ref32"][i]
            # north_star tolerance: 1e-3 relative (fp16 has 2^-11 = 4.9e-4 relative spacing)
            assert np.all(np.abs(mine - ref) <= 1e-3 * np.abs(ref) + 1e-3)
            assert np.mean(mine.astype(np.float16) == ref.astype(np.float16)) > 0.995
