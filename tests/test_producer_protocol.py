"""f2: `producer.inference_worker` against the reference's own `inference_worker` (PredictAndGenerate.py:23-61), both
driven by the same clients over the same queues with the same stand-in model: same results in the same order on every
client's queue, same shutdown behaviour.  The reference function runs unmodified; only its model loader, its stdout
redirection and `torch.cuda.get_device_name` are replaced (no checkpoint, no GPU here)."""
import argparse
import queue
import threading
import types

import numpy as np
import pytest
import torch

from oracle import ref_driver
from vr_video_generator_b200 import producer

pytestmark = pytest.mark.skipif(not ref_driver.reference_available(), reason="reference not present / not staged")


class _Model:
    """`infer_image_gpu(img)`: a deterministic [H,W] map of the frame, in the dtype autocast would hand over."""

    def __init__(self):
        self.calls = []

    def infer_image_gpu(self, img):
        self.calls.append(tuple(img.shape))
        x = torch.from_numpy(img[..., 0].astype(np.float32))
        return (x * 0.05 + float(img[0, 0, 1])).to(torch.float16)


def _drive(worker_fn, encoder, n_clients=2, shutdown="notify"):
    in_q = [queue.Queue() for _ in range(n_clients)]
    out_q = [queue.Queue() for _ in range(n_clients)]
    notify = queue.Queue()
    args = argparse.Namespace(encoder=encoder, encoder_path="unused.pth", DebugDir="/tmp/")
    model = _Model()
    th = threading.Thread(target=worker_fn, args=(in_q, out_q, notify, torch.device("cpu"), args, model), daemon=True)
    th.start()
    rng = np.random.default_rng(7)
    got = [[] for _ in range(n_clients)]
    # the clients interleave: 0, 1, 1, 0, 1, 0 ... and each reads its own queue, like SbsProcessor.add_frame / get_depth
    order = [0, 1, 1, 0, 1, 0, 0][: 7]
    frames = [rng.integers(0, 256, (6, 8, 3), dtype=np.uint8) for _ in order]
    for k, c in enumerate(order):
        c %= n_clients
        notify.put((c,))
        in_q[c].put((frames[k],))
    for k, c in enumerate(order):
        got[c % n_clients].append(out_q[c % n_clients].get(timeout=20))
    if shutdown == "notify":
        notify.put(None)
    else:                                  # a None job ends the loop too (:49-51)
        notify.put((0,))
        in_q[0].put(None)
    th.join(timeout=20)
    assert not th.is_alive()
    return got, model.calls


def _reference_worker():
    PAG = ref_driver.load_reference("cpu")

    def run(in_q, out_q, notify, device, args, model):
        saved = {k: getattr(PAG, k) for k in ("load_model", "redirrect_stdout", "print_flush", "torch")}
        real = saved["torch"]
        fake_cuda = types.SimpleNamespace(get_device_name=lambda d: "cpu", empty_cache=lambda: None)

        class _T:
            def __getattr__(self, name):
                return fake_cuda if name == "cuda" else getattr(real, name)
        try:
            PAG.load_model = lambda enc, path, dev: model
            PAG.redirrect_stdout = lambda path: None
            PAG.print_flush = lambda *a, **k: None
            PAG.torch = _T()
            PAG.inference_worker(in_q, out_q, notify, device, args)
        finally:
            for k, v in saved.items():
                setattr(PAG, k, v)
    return run


def _ours(in_q, out_q, notify, device, args, model):
    producer.inference_worker(in_q, out_q, notify, device, args, model=model)


@pytest.mark.parametrize("encoder", ["vits", "vitb", "vitl", "other"])
@pytest.mark.parametrize("shutdown", ["notify", "job"])
def test_inference_worker_matches_the_reference_protocol(encoder, shutdown):
    want, calls_ref = _drive(_reference_worker(), encoder, shutdown=shutdown)
    got, calls = _drive(_ours, encoder, shutdown=shutdown)
    assert calls == calls_ref and calls[0] == (1080, 1920, 3)            # warm-up on a black 1080p frame, then one call per job
    for c in range(2):
        assert len(got[c]) == len(want[c])
        for a, b in zip(got[c], want[c]):
            assert a.dtype == b.dtype and a.device == b.device and torch.equal(a, b)


def test_inference_worker_needs_a_model():
    with pytest.raises(ValueError):
        producer.inference_worker([], [], queue.Queue(), "cpu", argparse.Namespace(encoder="vits", encoder_path=""))
