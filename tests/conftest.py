import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

SMALL_CASES = ["small_a", "small_b", "small_step3", "small_neg", "small_zero", "small_pospos", "small_negneg", "medium"]
FULL_CASES = ["full_1080p_cfg1", "full_1080p_step2", "full_4k_wide"]
F32_CASES = ["small_f32", "medium_f32"]          # fp32 depth (torch >= 2.4 CUDA autocast): run through the general row kernel


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a); run with -m gpu under gpurun")


def load_meta(name):
    with open(os.path.join(GOLDEN, name + ".json")) as f:
        return json.load(f)


def load_case(name):
    """(meta, frames, raw_depth, ref_left or None).  Full-size cases regenerate their inputs from the
    seeded generators and check them against the recorded sha256."""
    meta = load_meta(name)
    path = os.path.join(GOLDEN, name + ".npz")
    if os.path.exists(path):
        z = np.load(path)
        return meta, z["frames"], z["raw_depth"], z["ref_left"]
    frames, raw = regenerate_inputs(meta)
    return meta, frames, raw, None


def regenerate_inputs(meta):
    import hashlib

    from oracle import sbs_layered as O
    from vr_video_generator_b200 import synth

    p = meta["params"]
    gen = synth.frames_noise if p["frames"] == "noise" else synth.frames_gradient
    frames = gen(p["n"], p["H"], p["W"], p["seed"])
    lo = synth.depth_lowres(p["depth"], p["n"], meta["lowres"][0], meta["lowres"][1], p["seed"])
    raw = np.stack([O.bicubic_resize(lo[t], p["H"], p["W"], 1.0) for t in range(p["n"])])
    if meta["shift"]:
        raw = (raw.astype(np.float32) - np.float32(meta["shift"])).astype(np.float16)
    for t in meta["zero"]:
        raw[t] = 0
    if meta.get("f32"):
        raw = raw.astype(np.float32) * np.float32(1.00037)
    for key, arr in (("frames", frames), ("raw_depth", raw)):
        got = hashlib.sha256(np.ascontiguousarray(arr).tobytes()).hexdigest()
        assert got == meta["inputs_sha"][key], f"synthetic {key} not reproducible on this host"
    return frames, raw


def golden_weights(meta):
    w = np.array([float.fromhex(h) for h in meta["weights"]], dtype=np.float32)
    return w.reshape(meta["weights_shape"])


def unhex(xs):
    return [float.fromhex(h) for h in xs]


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import scatter
    scatter.build()
    return scatter
