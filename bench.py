#!/usr/bin/env python
"""bench.py — SBS frames/sec of the warp stage on B200 (BASELINE.json's metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A "step" is one pass of the hot path over one batch of synthetic frames + depth maps:
temporal depth smoothing + per-frame max, layer tables, layered warp, hole fill + blur, strip, SBS
pack.  `value` = frames/s with inputs resident in HBM (the device-pointer C ABI), `e2e` = the same
metric through `SbsProcessor.submit_batch` / `collect` with page-locked HOST buffers (H2D + D2H in the
timed region).  `--impl reference` times the UNMODIFIED reference (staged copy oracle/_ref behind the CPU
device proxy, one worker process per host core) on a bounded sample of the same workload.  For N > 1 launch under torchrun; every rank
owns its own clip range (independent shards, no collective in the data path).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOADS = {
    # BASELINE.json configs[1]: 1080p, 64-frame batch, code-default offsets, step 1 (the metric's config)
    "1080p_b64": dict(H=1080, W=1920, B=64, fg=0.025, bg=-0.01, step=1, depth="scene"),
    # configs[0]: the reference's own CPU-runnable case
    "1080p_b16_cfg1": dict(H=1080, W=1920, B=16, fg=0.025, bg=-0.015, step=1, depth="scene"),
    # configs[2]: 4K, wide disparity
    "4k_wide_b16": dict(H=2160, W=3840, B=16, fg=0.05, bg=-0.03, step=1, depth="scene"),
    # configs[3]: step 2 banded sweep
    "1080p_step2_b64": dict(H=1080, W=1920, B=64, fg=0.025, bg=-0.01, step=2, depth="scene"),
    # scatter / blur stress (19 % holes)
    "1080p_stress_b64": dict(H=1080, W=1920, B=64, fg=0.025, bg=-0.015, step=1, depth="stress"),
}
ranks = None
ROUTE = "default (k_depth_pass, k_build_tables, k_warp_ws, k_band_list + k_blur_band [1080p] / k_word_list + k_blur_sep [4K], k_blur_commit)"
METRIC = "sbs_frames_per_sec_warp_stage"
UNIT = "frames/s"


def algorithmic_bytes(H, W):
    """SURVEY.md section 8d, A_warp: fp16 full-res depth + RGB frame read once, SBS frame written once."""
    return H * W * 2 + H * W * 3 + H * 2 * W * 3


def measured_hbm_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons of this rank's GPU, sampled every 50 ms during the timed regions (the quantities of
    the B200_PROFILING.md nvidia-smi recipe: clocks.sm, clocks.max.sm, clocks_event_reasons.*).  Read through NVML in
    this process: one `nvidia-smi -lms` child per rank spends about a second attaching to every GPU of the box right
    when the timed region starts, and on an 8-GPU box that showed up as 8 % longer steps at N > 1.  Falls back to the
    nvidia-smi child when pynvml is missing."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        self.index, self.rows, self.proc, self.nvml, self.handle = index, [], None, None, None
        self.stop = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            visible = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if visible:
                ids = [v.strip() for v in visible.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def __enter__(self):
        if self.nvml is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return self
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _poll(self):
        n = self.nvml
        bits = [(getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), "hw_slowdown"),
                (getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40), "hw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), "sw_thermal_slowdown"),
                (getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4), "sw_power_cap")]
        while not self.stop.is_set():
            try:
                sm = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksEventReasons(self.handle))
                self.rows.append([str(sm), str(self.max_sm)] + ["Active" if mask & b else "Not Active" for b, _ in bits])
            except Exception:
                pass
            self.stop.wait(0.05)

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.nvml is not None:
            self.stop.set()
            self.thread.join(timeout=2)
        elif self.proc:
            time.sleep(0.15)
            self.proc.terminate()
            self.thread.join(timeout=2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        reasons = [n for i, n in enumerate(self.NAMES) if any(len(r) >= 6 and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": "nvml" if self.nvml is not None else "nvidia-smi"}


def make_inputs(wl, seed):
    """Seeded synthetic batch: noise frames [B,H,W,3] u8 and D-scene/D-stress low-res depth [B,518,924]."""
    from vr_video_generator_b200 import synth
    frames = synth.frames_noise(wl["B"], wl["H"], wl["W"], seed)
    lowres = synth.depth_lowres(wl["depth"], wl["B"], synth.DPT_H, synth.DPT_W, seed)
    return frames, lowres


def dist_setup():
    """(rank, world, barrier, reduce_max) via vr_video_generator_b200.shard.Ranks (NCCL under torchrun)."""
    import torch

    from vr_video_generator_b200 import shard
    rank, world, local = shard.world_from_env()
    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.set_device(local)
    r = shard.Ranks(backend="nccl" if cuda else "gloo", device=f"cuda:{local}" if cuda else "cpu")
    global ranks
    ranks = r
    return r.rank, r.world, r.barrier, r.max


# ---------------------------------------------------------------------------------------------------------
class DeviceBench:
    """Device-resident inputs of one workload + the timed step (vrsbs_process_batch, or the low-res depth route)."""

    def __init__(self, args, wl, dev, seed=100, f32=False):
        import torch

        from vr_video_generator_b200 import _native, tables
        self.torch, self.wl, self.dev = torch, wl, dev
        H, W, B = wl["H"], wl["W"], wl["B"]
        # every rank = its own clip range with its own state; the CONTENT is the same seeded clip on every rank, so that
        # the work per GPU is fixed as N grows (weak scaling) - hole counts, and with them the blur time, vary by ~10 %
        # between seeds and the job time is the max over ranks
        self.frames_h, self.lowres_h = make_inputs(wl, seed=seed)
        ctx = _native.Context(dev, H, W, B, 512)
        ctx.reset(wl["fg"], wl["bg"], wl["step"], True, _native.DEPTH_F32 if f32 else _native.DEPTH_F16)
        ctx.set_blur_weights(tables.gaussian_weights(*tables.blur_kernel_shape(H)))
        self.f32 = f32
        if args.scatter_mode:                      # 1/2: the general row kernel instead of the fused route (A/B runs)
            ctx.set_option("fused", 0)
            ctx.set_option("scatter_mode", args.scatter_mode)
        for kv in args.opt:                        # tuning experiments: --opt name=value
            k, v = kv.split("=")
            ctx.set_option(k, int(v))
        self.ctx = ctx
        stream = self.stream = torch.cuda.current_stream().cuda_stream
        # device-resident inputs: frames + RAW full-res fp16 depth (what SbsProcessor.get_depth receives),
        # produced once, untimed, by the library's own depth tail from the low-res synthetic DPT maps
        self.frames_d = torch.from_numpy(self.frames_h).cuda()
        self.lowres_d = torch.from_numpy(self.lowres_h).cuda()
        self.raw_d = torch.empty((B, H, W), dtype=torch.float16, device="cuda")
        prep = _native.Context(dev, H, W, 1, 64)
        for t in range(B):                                            # B=1 with a fresh clip each: raw depth, not smoothed
            prep.reset(wl["fg"], wl["bg"], wl["step"], False)
            prep.depth_from_lowres(self.lowres_d[t].data_ptr(), 1, self.lowres_h.shape[1], self.lowres_h.shape[2], 1.0, H, W,
                                   self.raw_d[t].data_ptr(), stream)
            # first frame of a clip: smoothed = 0.58d + 0.3d + 0.12d != d in fp16; good enough as "raw" input
        torch.cuda.synchronize()
        prep.close()
        self.raw_h = self.raw_d.cpu().numpy()
        if f32:                                    # the same depth values as an fp32 clip (what torch >= 2.4 autocast hands over)
            self.raw_d = self.raw_d.float()
        self.scratch_d = torch.empty_like(self.raw_d)
        self.sbs_d = torch.empty((B, H, 2 * W, 3), dtype=torch.uint8, device="cuda")
        self.lowres = args.depth_input == "lowres"

    def step(self):
        wl, ctx, s = self.wl, self.ctx, self.stream
        H, W, B = wl["H"], wl["W"], wl["B"]
        if self.lowres:
            # the depth tail on the device: DPT-resolution map in, bicubic + smoothing + max fused (A_fused)
            lh, lw = self.lowres_h.shape[1], self.lowres_h.shape[2]
            ctx.depth_from_lowres(self.lowres_d.data_ptr(), B, lh, lw, 1.0, H, W, self.scratch_d.data_ptr(), s)
            ctx.build_tables(B, H, W, s)
            ctx.warp_batch(self.frames_d.data_ptr(), self.scratch_d.data_ptr(), B, H, W, self.sbs_d.data_ptr(), s)
        else:
            ctx.process_batch(self.frames_d.data_ptr(), self.raw_d.data_ptr(), B, H, W, self.scratch_d.data_ptr(),
                              self.sbs_d.data_ptr(), s)

    def timed(self, steps, warmup, barrier=lambda: None, on_start=lambda: None):
        """(ms for `steps` steps on this rank, host issue ms per step, stage times, launches, frame infos)."""
        torch, ctx = self.torch, self.ctx
        for _ in range(warmup):
            self.step()
        torch.cuda.synchronize()
        infos = ctx.frame_info(self.wl["B"], self.stream)
        ctx.stage_times()
        ctx.set_option("stage_timing", 1)
        launches0 = ctx.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        on_start()
        e0.record()
        t_host = time.perf_counter()
        for _ in range(steps):
            self.step()
        e1.record()
        host_issue_ms = (time.perf_counter() - t_host) * 1e3 / steps      # CPU time to issue one step (launch bound if ~ ms_per_step)
        torch.cuda.synchronize()
        barrier()
        ms = e0.elapsed_time(e1)
        stage = ctx.stage_times()
        ctx.set_option("stage_timing", 0)
        # a second, untimed-by-events pass WITHOUT the per-kernel event pairs: the value the line reports
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        torch.cuda.synchronize()
        e2.record()
        for _ in range(steps):
            self.step()
        e3.record()
        torch.cuda.synchronize()
        barrier()
        ms_clean = e2.elapsed_time(e3)
        return ms, ms_clean, host_issue_ms, stage, (ctx.launch_count() - launches0) // 2, infos

    def workload_text(self, name, infos):
        wl = self.wl
        return (f"{name}: {wl['W']}x{wl['H']}, {wl['B']}-frame batch per GPU, fg={wl['fg']} bg={wl['bg']} step={wl['step']}, "
                f"D-{wl['depth']} depth (limit_step {infos[0].limit_step}, {infos[0].layers} layers, "
                f"{100.0 * infos[0].holes / (wl['H'] * wl['W']):.2f}% holes)")

    def close(self):
        self.ctx.close()


def stage_roofline(name, wl, ms_per_step, stage, steps, peak, peak_src, kernel):
    """roofline of the dominant kernel (algorithmic bytes per launch / mean launch duration) AND of the whole warp stage
    (`stage_frac` = frames/s x A_warp / peak: what north_star's >= 60 % target is about)."""
    H, W, B = wl["H"], wl["W"], wl["B"]
    warp_ms, warp_n = stage["warp"]
    abytes = algorithmic_bytes(H, W) * B
    achieved = abytes / (warp_ms / max(warp_n, 1) * 1e-3) / 1e9 if warp_n else None
    traffic = stage_traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic, stage_traffic = tj.get(name), tj.get(name + "_stage")
    stage_gbs = abytes / (ms_per_step * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": (achieved / peak) if achieved else None, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": abytes, "launch_ms": warp_ms / max(warp_n, 1),
            "stage_achieved": stage_gbs, "stage_frac": stage_gbs / peak, "stage_frac_vs_nominal_8TBs": stage_gbs / 8000.0,
            "stage_traffic": stage_traffic,
            "stage_note": "stage_* = all kernels of one step (depth pass, tables, warp, blur, commit): A_warp bytes x frames / step time"}


def pcie_probe(world, barrier, reduce_max, mb=256, reps=4, mix=None):
    """Aggregate host<->device copy bandwidth of the job: every rank copies `mb` MiB of page-locked memory to its GPU and
    back, `reps` times, one direction at a time and both at once (two streams); GB/s summed over the ranks, from the
    slowest rank's time.  Gives the DMA ceiling the end-to-end figure can be held against (at N > 1 the ranks share the
    host's memory system and PCIe roots)."""
    import torch
    n = mb << 20
    h_in, h_out = torch.empty(n, dtype=torch.uint8).pin_memory(), torch.empty(n, dtype=torch.uint8).pin_memory()
    d_a, d_b = torch.empty(n, dtype=torch.uint8, device="cuda"), torch.empty(n, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize()
        dt = reduce_max(time.perf_counter() - t0)
        barrier()
        return world * reps * n / dt / 1e9

    def h2d():
        with torch.cuda.stream(s1):
            d_a.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_b, non_blocking=True)

    def both():
        h2d()
        d2h()
    out = {"h2d_gbs": timed(h2d), "d2h_gbs": timed(d2h), "both_each_way_gbs": timed(both), "mib": mb, "reps": reps}
    if mix:
        # the pipeline's own byte mix: per frame `mix[0]` bytes in and `mix[1]` bytes out, both directions at once, plain
        # contiguous copies of page-locked memory and nothing else - what the box's host <-> device path can carry for
        # this workload when every rank asks at the same time (frames/s summed over the ranks)
        k = max(1, n // max(mix))
        a_in, a_out = int(mix[0] * k), int(mix[1] * k)

        def mixed():
            with torch.cuda.stream(s1):
                d_a[:a_in].copy_(h_in[:a_in], non_blocking=True)
            with torch.cuda.stream(s2):
                h_out[:a_out].copy_(d_b[:a_out], non_blocking=True)
        mixed()
        torch.cuda.synchronize()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            mixed()
        torch.cuda.synchronize()
        dt = reduce_max(time.perf_counter() - t0)
        barrier()
        out["mix_frames_per_sec"] = world * reps * k / dt
        out["mix_h2d_gbs"], out["mix_d2h_gbs"] = world * reps * a_in / dt / 1e9, world * reps * a_out / dt / 1e9
    del h_in, h_out, d_a, d_b
    return out


def reference_cuda_fps(wl, frames, raw, n=6):
    """The UNMODIFIED reference on this B200 as it ships (torch CUDA ops, cuDNN blur, pageable H2D, blocking D2H per
    frame; PredictAndGenerate.py:157-198), fed through a plain queue: a second, more telling baseline."""
    import torch

    from oracle import ref_driver
    if not ref_driver.reference_available():
        return None
    n = min(n, len(frames) - 1)
    ref = ref_driver.ReferenceWarp(wl["fg"], wl["bg"], wl["step"], device="cuda")
    ref.left_side_sbs(frames[0], torch.from_numpy(raw[0]))             # warm-up: cuDNN plan, allocator
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for t in range(1, 1 + n):
        ref.left_side_sbs(frames[t], torch.from_numpy(raw[t]))
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    torch.cuda.empty_cache()
    return {"value": n / dt, "unit": UNIT, "frames": n, "ms_per_frame": 1e3 * dt / n,
            "what": "unmodified reference SbsProcessor.left_side_sbs on cuda:0 (host numpy frame + CPU fp16 depth in, host SBS frame out), "
                    f"cudnn.allow_tf32={torch.backends.cudnn.allow_tf32}, torch {torch.__version__}"}


def producer_timing(frames, encoders=("vits", "vitb", "vitl"), n=4):
    """Depth producer, timed SEPARATELY (north_star; never part of the warp figure): the reference's Depth-Anything-V2
    module with random-initialised weights (no checkpoints offline) under fp16 autocast.  `reference_per_frame` = what
    inference_worker does per request (PredictAndGenerate.py:54-55): `infer_image_gpu(img) * scaler` (cv2 preprocessing
    on the host, batch-1 forward, bicubic tail to frame size) plus the `.to('cpu')`.  `batched` = producer.DepthProducer:
    one forward for n frames, DPT-resolution output left on the device (the bicubic tail and the scaler then run inside
    the warp's depth pass)."""
    import torch

    from oracle import ref_driver
    from vr_video_generator_b200.producer import DepthProducer
    from vr_video_generator_b200.worker import encoder_scaler
    if not ref_driver.reference_available():
        return None
    out = {}
    dev = torch.device("cuda", torch.cuda.current_device())
    for enc in encoders:
        model = ref_driver.depth_model(enc, dev)
        scaler = encoder_scaler(enc)
        with torch.no_grad():
            model.infer_image_gpu(frames[0])                                   # warm-up outside autocast, like :37
            with torch.autocast(device_type="cuda", dtype=torch.float16):
                (model.infer_image_gpu(frames[0]) * scaler).to("cpu")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for t in range(n):
                with torch.autocast(device_type="cuda", dtype=torch.float16):
                    d = (model.infer_image_gpu(frames[t % len(frames)]) * scaler).to(torch.device("cpu"))
            per_frame = (time.perf_counter() - t0) / n
            x, _ = model.image2tensor(frames[0], 518)
            with torch.autocast(device_type="cuda", dtype=torch.float16):
                model.forward(x)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(n):
                    model.forward(x)
                e1.record()
                torch.cuda.synchronize()
            fwd1 = e0.elapsed_time(e1) / n
        prod = DepthProducer(model, enc, max_forward_batch=n)
        prod(frames[:n])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        lo = prod(frames[:n])
        torch.cuda.synchronize()
        batched = (time.perf_counter() - t0) / n
        xb = torch.cat([x] * n)
        with torch.no_grad(), torch.autocast(device_type="cuda", dtype=torch.float16):
            model.forward(xb)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            model.forward(xb)
            e1.record()
            torch.cuda.synchronize()
        out[enc] = {"reference_per_frame_ms": 1e3 * per_frame, "reference_per_frame_fps": 1.0 / per_frame,
                    "forward_batch1_ms": fwd1, f"forward_batch{n}_ms_per_frame": e0.elapsed_time(e1) / n,
                    "batched_producer_ms_per_frame": 1e3 * batched, "batched_producer_fps": 1.0 / batched,
                    "out_dtype": str(d.dtype), "lowres_shape": list(lo.shape[1:]), "params_M": sum(p.numel() for p in model.parameters()) / 1e6}
        del model, prod, lo, xb
        torch.cuda.empty_cache()
    out["note"] = ("random-init weights (timing only; random-init depth is degenerate for the warp, SURVEY 8d); 1080p frames; "
                   "host preprocessing (cv2 INTER_AREA + float64 normalise, dpt.py:204-228) is inside reference_per_frame and "
                   "batched_producer, not inside forward_*")
    return out


def run_ours(args, wl, name):
    import torch

    import vr_video_generator_b200 as pkg
    from vr_video_generator_b200 import tables

    rank, world, barrier, reduce_max = dist_setup()
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback in the product path)"
    dev = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(dev)
    H, W, B = wl["H"], wl["W"], wl["B"]
    binding = None
    if args.bind:
        from vr_video_generator_b200 import shard
        pr = torch.cuda.get_device_properties(dev)
        bdf = "%04x:%02x:%02x.0" % (getattr(pr, "pci_domain_id", 0), getattr(pr, "pci_bus_id", 0), getattr(pr, "pci_device_id", 0))
        cpus = shard.bind_near_gpu(bdf)
        binding = f"{bdf}: {len(cpus)} local cpus" if cpus else f"{bdf}: unchanged ({len(os.sched_getaffinity(0))} cpus)"

    db = DeviceBench(args, wl, dev)
    frames_h, lowres_h, raw_h, lowres_d, sbs_d = db.frames_h, db.lowres_h, db.raw_h, db.lowres_d, db.sbs_d
    clocks = ClockSampler(dev)               # NVML attach happens here, outside the timed region

    def start_clocks():
        clocks.__enter__()                   # sampled over the device-timed region AND the e2e region
        time.sleep(0.25)
    ms_rank_staged, ms_rank, host_issue_ms, stage, launches, infos = db.timed(args.steps, args.warmup, barrier, start_clocks)
    ms_total = reduce_max(ms_rank)
    per_rank = ranks.gather({"rank": rank, "ms_per_step": ms_rank / args.steps, "host_issue_ms_per_step": host_issue_ms})
    value = world * B * args.steps / (ms_total * 1e-3)
    peak, peak_src = measured_hbm_peak()
    kernel = "k_warp_rows" if args.scatter_mode else ("k_warp_ws" if W % 32 == 0 else "k_warp_fused")
    roof = stage_roofline(name, wl, ms_total / args.steps, stage, args.steps, peak, peak_src, kernel)

    # end to end: the public batch API with pinned host buffers (H2D + D2H inside the timed region)
    ns = argparse.Namespace(offset_fg=wl["fg"], offset_bg=wl["bg"], offset_step_size=wl["step"])
    proc = pkg.SbsProcessor(None, 0, ns, device=dev, max_batch=16)
    if args.host_chunk:
        proc._context(H, W).set_option("host_chunk", args.host_chunk)
    # the host copy threads of all ranks share the box's cores
    proc._context(H, W).set_option("copy_threads", max(2, min(8, (os.cpu_count() or 8) // world)))
    for kv in args.host_opt:
        k, v = kv.split("=")
        proc._context(H, W).set_option(k, int(v))
    # Two page-locked SBS buffers; the frames live in their right halves, where a decoder would put them (the reference's
    # loop copies every decoded frame once for the BGR -> RGB swap: that copy lands here).  A step = submit batch k, then
    # collect batch k-1: H2D of the frames (pitched) + depth, kernels, D2H of the synthesised halves.
    from vr_video_generator_b200.sbs import pinned_sbs_buffer
    ring = [pinned_sbs_buffer(B, H, W) for _ in range(2)]
    for out_k, view_k, _t in ring:
        np.copyto(view_k, frames_h)
    d_pin = torch.from_numpy(raw_h).pin_memory()
    l_pin = torch.from_numpy(lowres_h).pin_memory()

    def run_async(depth, steps):
        """frames/s of `steps` batches through submit_batch / collect, two batches in flight"""
        proc.reset_state()
        prev = None
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for k in range(steps):
            out_k, view_k, _t = ring[k & 1]
            t = proc.submit_batch(view_k, depth, out_k)
            if prev is not None:
                proc.collect(prev)
            prev = t
        proc.collect(prev)
        dt = reduce_max(time.perf_counter() - t0)
        barrier()
        return world * B * steps / dt

    e2e_steps = max(2, min(args.steps, args.e2e_steps))
    run_async(d_pin, 2)                                       # warm-up: slot allocation, first-touch of the pinned pages
    e2e_value = run_async(d_pin, e2e_steps)
    clocks.__exit__()
    e2e = {"value": e2e_value, "unit": UNIT,
           "h2d_bytes_per_step": int(frames_h.nbytes + raw_h.nbytes), "d2h_bytes_per_step": int(ring[0][0].nbytes // 2),
           "host_to_host_bytes_per_step": 0, "steps": e2e_steps,
           "api": "SbsProcessor.submit_batch / collect (vrsbs_submit_host), page-locked buffers, frames decoded in place into the "
                  "right halves of the SBS buffer, full-resolution fp16 depth on the host, two batches in flight"}
    per_frame_in, per_frame_out = (frames_h.nbytes + raw_h.nbytes) / B, ring[0][0].nbytes / 2 / B
    pcie = pcie_probe(world, barrier, reduce_max, mix=(per_frame_in, per_frame_out))
    e2e["pcie"] = dict(pcie, dma_ceiling_fps=pcie["mix_frames_per_sec"],
                       note="plain page-locked copies, aggregate over the ranks (slowest rank's time): each direction alone, both at once "
                            "with equal bytes, and `mix` = this workload's bytes per frame in and out at once; dma_ceiling_fps = the mix "
                            "figure: what the box's host <-> device path carries for this byte mix with no kernels and no pipeline")
    e2e["frac_of_dma_ceiling"] = e2e_value / e2e["pcie"]["dma_ceiling_fps"]
    # the blocking call of round 1 (separate frame buffer, right halves copied host to host by the library), for comparison
    f_pin = torch.from_numpy(frames_h).pin_memory()
    o_np = ring[0][0]
    proc.reset_state()
    proc.left_side_sbs_batch(f_pin, d_pin, out=o_np)
    barrier()
    t0 = time.perf_counter()
    for _ in range(2):
        proc.left_side_sbs_batch(f_pin, d_pin, out=o_np)
    e2e["blocking_call_value"] = world * B * 2 / reduce_max(time.perf_counter() - t0)
    barrier()
    # the e2e output of the whole batch (every frame) against the device-resident output of the same clip, both from a
    # clean clip state, untimed
    db.ctx.reset(wl["fg"], wl["bg"], wl["step"], True)
    db.step()
    proc.reset_state()
    proc.collect(proc.submit_batch(ring[1][1], d_pin, ring[1][0]))
    torch.cuda.synchronize()
    dev_out = sbs_d.cpu().numpy()
    same = bool(np.array_equal(ring[1][0], dev_out))
    same_frames = int(sum(np.array_equal(ring[1][0][t], dev_out[t]) for t in range(B)))
    del dev_out
    # the same call with ordinary (pageable) numpy arrays, as nibba_woka's FrameList holds them: one step, reported aside
    if args.pageable:
        o_pg = np.empty_like(o_np)
        proc.left_side_sbs_batch(frames_h, raw_h, out=o_pg)
        t0 = time.perf_counter()
        proc.left_side_sbs_batch(frames_h, raw_h, out=o_pg)
        e2e["pageable_value"] = world * B / reduce_max(time.perf_counter() - t0)

    # the same pipeline when the depth producer hands over the DPT-resolution map (bicubic on the device): 4x less depth H2D
    run_async(l_pin, 1)
    e2e["lowres_depth_value"] = run_async(l_pin, max(2, e2e_steps // 2))
    # ... and when the producer runs in this process and leaves that map on the device: frames are the only H2D traffic
    run_async(lowres_d, 1)
    e2e["device_depth_value"] = run_async(lowres_d, max(2, e2e_steps // 2))

    # the reference's own per-frame call (left_side_sbs with the depth arriving on a queue), as nibba_woka makes it
    import queue
    q = queue.Queue()
    proc.reset_state()
    nper = min(B, 16)
    for t in range(nper + 2):
        q.put(torch.from_numpy(raw_h[t % B]))
    proc.left_side_sbs(frames_h[0], None, q)
    proc.left_side_sbs(frames_h[1], None, q)
    t0 = time.perf_counter()
    for t in range(nper):
        proc.left_side_sbs(frames_h[t], None, q)
    e2e["per_frame_call_fps"] = nper / (time.perf_counter() - t0)

    # BASELINE.json configs[4]: a synthetic N-frame video sharded by clip range over the ranks (main_func's split,
    # PredictAndGenerate.py:274-275), every rank streaming its range through the host API from a cycled pinned pool
    video = None
    if args.video_frames:
        ranges = tables.clip_ranges(0, args.video_frames, args.video_frames, world)
        begin, end = ranges[rank] if rank < len(ranges) else (0, 0)
        proc.reset_state()
        barrier()
        t0 = time.perf_counter()
        prev, k = None, 0
        for b0 in range(begin, end, B):
            n = min(B, end - b0)
            out_k, view_k, _t = ring[k & 1]
            t = proc.submit_batch(view_k[:n], d_pin[:n], out_k[:n])
            if prev is not None:
                proc.collect(prev)
            prev, k = t, k + 1
        if prev is not None:
            proc.collect(prev)
        dt = reduce_max(time.perf_counter() - t0)
        barrier()
        video = {"frames": args.video_frames, "seconds": dt, "frames_per_sec": args.video_frames / dt,
                 "ranges": [list(r) for r in ranges], "batch": B,
                 "what": "BASELINE.json configs[4] (bounded): synthetic 1080p video sharded by clip range over the ranks "
                         "(main_func's split), submit_batch / collect with two page-locked SBS buffers cycled from one batch"}
    proc.close()
    del f_pin, d_pin, o_np, l_pin, ring

    # BASELINE.json's metric names 1080p AND 4K: configs[2] (4K, wide disparity) device-resident on the same box
    extra = {}
    if args.also_4k and name != "4k_wide_b16" and not args.scatter_mode:
        wl4 = WORKLOADS["4k_wide_b16"]
        frames_h = lowres_h = raw_h = None
        db.close()
        del db, sbs_d, lowres_d
        torch.cuda.empty_cache()
        db4 = DeviceBench(args, wl4, dev)
        steps4 = max(3, min(args.steps, 10))
        _ms_s, ms4, _hi, stage4, _l4, infos4 = db4.timed(steps4, 3, barrier)
        ms4 = reduce_max(ms4)
        roof4 = stage_roofline("4k_wide_b16", wl4, ms4 / steps4, stage4, steps4, peak, peak_src, "k_warp_ws")
        extra["4k_wide_b16"] = {"workload": db4.workload_text("4k_wide_b16", infos4), "value": world * wl4["B"] * steps4 / (ms4 * 1e-3),
                                "unit": UNIT, "steps": steps4, "ms_per_step": ms4 / steps4,
                                "stage_ms_per_step": {k: v[0] / steps4 for k, v in stage4.items()},
                                "roofline": {k: roof4[k] for k in ("achieved", "frac", "stage_achieved", "stage_frac", "launch_ms", "traffic", "stage_traffic")}}
        db4.close()
        del db4
        torch.cuda.empty_cache()
    else:
        db.close()
    # the same headline workload with fp32 depth (the dtype the reference's producer hands over under torch >= 2.4: fp32
    # smoothing, fp32 comparison) - reported aside, the metric's config is fp16
    if args.also_f32 and not args.scatter_mode and args.depth_input == "full":
        torch.cuda.empty_cache()
        db32 = DeviceBench(args, wl, dev, f32=True)
        steps32 = max(3, min(args.steps, 10))
        _ms_s, ms32, _hi, stage32, _l32, _infos32 = db32.timed(steps32, 3, barrier)
        ms32 = reduce_max(ms32)
        extra[name + "_f32_depth"] = {"value": world * B * steps32 / (ms32 * 1e-3), "unit": UNIT, "steps": steps32, "ms_per_step": ms32 / steps32,
                                     "stage_ms_per_step": {k: v[0] / steps32 for k, v in stage32.items()},
                                     "route": "k_depth_pass_f32, k_build_tables (fp32 bounds + fp32 cell LUT), k_warp_ws<256,4,F32>, k_blur_sep, k_blur_commit"}
        db32.close()
        del db32
        torch.cuda.empty_cache()

    cpu = ref_cuda = producer = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        # the baseline legs sample the first frames of the headline workload (same seed, same generators)
        from oracle import sbs_layered as O
        nb = min(B, 16)
        if frames_h is None:
            frames_h, lo = make_inputs(dict(wl, B=nb), seed=100)
            raw_h = np.stack([O.bicubic_resize(lo[t], H, W, 1.0) for t in range(nb)])
        ref_cuda = reference_cuda_fps(wl, frames_h, raw_h)
        if not args.no_producer:
            producer = producer_timing(frames_h)
        cpu = cpu_baseline(wl, frames_h, raw_h, budget_s=args.cpu_budget)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8 pixels / fp16 depth compares / integer-exact blur / f64 layer tables", "data": "synthetic",
            "config": {"workload": DeviceBench.workload_text(argparse.Namespace(wl=wl), name, infos),
                       "timed_region": "depth smoothing+max pass, device tables, warp+fill+pack, hole blur, commit+strip; inputs/outputs in HBM",
                       "l2": f"inputs per step {int((B * H * W * 5) / 2**20)} MiB + outputs "
                             f"{int(B * H * W * 6 / 2**20)} MiB per GPU > 126 MB L2 (no flush needed)",
                       "sharding": "independent clip range per GPU (same seeded content on every rank: fixed work per GPU), no collective", "depth_input": args.depth_input, "cpu_binding": binding,
                       "route": f"general row kernel (scatter_mode={args.scatter_mode})" if args.scatter_mode else ROUTE},
            "clocks": clocks.summary(), "e2e": e2e, "gpu_launches": int(launches), "per_rank": per_rank,
            "roofline": roof,
            "stage_ms_per_step": {k: v[0] / args.steps for k, v in stage.items()},
            "ms_per_step_with_stage_events": ms_rank_staged / args.steps,
            "e2e_equals_device_output": same, "e2e_frames_equal": f"{same_frames}/{B}",
        }
        if extra:
            line["configs"] = extra
        if video:
            line["video"] = video
        if cpu:
            line["cpu_baseline"] = cpu
        if ref_cuda:
            line["reference_cuda"] = ref_cuda
        if producer:
            line["producer"] = producer
        print(json.dumps(line))
    ranks.close()


# ---------------------------------------------------------------------------------------------------------
# CPU baselines.  "reference" = the UNMODIFIED reference class (oracle/_ref staged copy, or /root/reference in the
# build container) behind the CPU device proxy of oracle/ref_driver.py, parallelised the way the reference itself
# parallelises: Num_Workers processes, each owning a contiguous clip range with its own state (main_func,
# PredictAndGenerate.py:274-275,300-306).  "port" = oracle/sbs_layered.py, used only when no reference copy exists.
def _ref_cpu_worker(idx, path, lo, hi, wl, threads, barrier, out_q):
    try:
        import numpy as np
        import torch
        torch.set_num_threads(threads)
        from oracle.ref_driver import ReferenceWarp
        z = np.load(path, mmap_mode="r")
        frames, raw = z["frames"], z["raw"]
        warm = ReferenceWarp(wl["fg"], wl["bg"], wl["step"])
        warm.left_side_sbs(np.ascontiguousarray(frames[lo]), torch.from_numpy(np.ascontiguousarray(raw[lo])))   # untimed: imports, allocator
        ref = ReferenceWarp(wl["fg"], wl["bg"], wl["step"])          # fresh clip-range state, like a worker's SbsProcessor (:209)
        barrier.wait()
        t0 = time.perf_counter()
        for t in range(lo, hi):
            ref.left_side_sbs(np.ascontiguousarray(frames[t]), torch.from_numpy(np.ascontiguousarray(raw[t])))
        out_q.put((idx, hi - lo, time.perf_counter() - t0, None))
    except Exception as e:                                           # noqa: BLE001
        try:
            barrier.abort()
        except Exception:
            pass
        out_q.put((idx, 0, 0.0, repr(e)))


def reference_cpu_fps(wl, frames, raw, budget_s=20.0, workers=None, threads=None):
    """frames/s of the unmodified reference warp on this box's host cores, on a bounded sample of the workload.
    The worker processes start, import the reference and warm up OUTSIDE the timed region."""
    import multiprocessing as mp
    import tempfile

    from oracle import ref_driver
    if not ref_driver.reference_available():
        return None
    cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    threads = threads or 1            # measured: one torch thread per worker and a worker per core beats 2x4 / 4x2 splits
    workers = workers or max(1, cores // threads)
    # ~0.4-1.5 s per 1080p frame per worker (4x that at 4K): size every worker's range to the budget
    per_frame = 1.2 * (wl["H"] * wl["W"]) / (1080 * 1920) * (4.0 if wl["W"] > 2048 else 1.0)
    per_worker = int(max(1, min(8, budget_s / per_frame)))
    n = min(len(frames), workers * per_worker)
    workers = min(workers, n)
    bounds = [n * i // workers for i in range(workers + 1)]
    tmp = tempfile.NamedTemporaryFile(suffix=".npz", dir="/dev/shm" if os.path.isdir("/dev/shm") else None, delete=False)
    tmp.close()
    try:
        np.savez(tmp.name, frames=frames[:n], raw=raw[:n])
        ctx = mp.get_context("spawn")
        barrier, q = ctx.Barrier(workers + 1), ctx.Queue()
        procs = [ctx.Process(target=_ref_cpu_worker, args=(i, tmp.name, bounds[i], bounds[i + 1], wl, threads, barrier, q))
                 for i in range(workers)]
        for p_ in procs:
            p_.start()
        try:
            barrier.wait(timeout=600)
        except Exception:
            errs = []
            while not q.empty():
                errs.append(q.get()[3])
            for p_ in procs:
                p_.join(timeout=5)
            raise RuntimeError(f"reference CPU workers failed to start: {errs}")
        t0 = time.perf_counter()
        res = [q.get(timeout=3600) for _ in procs]
        dt = time.perf_counter() - t0
        for p_ in procs:
            p_.join()
    finally:
        os.unlink(tmp.name)
    bad = [r[3] for r in res if r[3]]
    if bad:
        raise RuntimeError(f"reference CPU worker error: {bad[0]}")
    import torch
    return {"value": n / dt, "unit": UNIT, "cores": workers * threads, "kind": "reference",
            "sample": f"{n} frames of the same workload ({wl['W']}x{wl['H']}) through the UNMODIFIED reference "
                      f"SbsProcessor.left_side_sbs on the CPU (oracle/ref_driver.py device proxy; torch {torch.__version__}), "
                      f"{workers} worker processes x {threads} torch threads on {cores} available host cores, one clip range "
                      f"per worker like main_func's Num_Workers; start-up and one warm-up frame per worker untimed; "
                      f"slowest worker {max(r[2] for r in res):.2f} s for {max(r[1] for r in res)} frames",
            "workers": workers, "threads_per_worker": threads, "frames": n, "seconds": dt}


def _cpu_worker(job):
    from oracle import sbs_layered as O
    img, depth, marks, steps, offsets, weights = job
    return O.warp_frame(img, depth, marks, steps, offsets, weights)[0, 0, 0]


def cpu_port_fps(wl, frames, raw, budget_s=20.0, procs=None):
    """Fallback when no copy of the reference exists: oracle/sbs_layered.py (numpy layer loop, a port).  Smoothing +
    tables run serially (they carry state), the per-frame warp — >99 % of the time — runs one frame per process; the
    pool is started before the timed region."""
    import multiprocessing as mp

    from oracle import sbs_layered as O
    cores = os.cpu_count() or 1
    procs = procs or cores
    st = O.WarpState(wl["fg"], wl["bg"], wl["step"])
    weights = O.gaussian_weights(*O.blur_kernel_shape(wl["H"]))
    d = O.smooth_depth(st, raw[0])
    tabs = O.layer_tables(st, d.max(), d.shape[0])
    t0 = time.perf_counter()
    O.warp_frame(frames[0], d, tabs[0], tabs[1], tabs[2], weights)
    one = time.perf_counter() - t0
    n = int(max(procs, min(len(frames) - 1, procs * max(1, int(budget_s / max(one, 1e-3))))))
    n = min(n, len(frames) - 1)
    with mp.get_context("fork").Pool(min(procs, n)) as pool:
        pool.map(abs, range(procs))                                  # workers up before the clock starts
        jobs = []
        t0 = time.perf_counter()
        for t in range(1, 1 + n):
            d = O.smooth_depth(st, raw[t])
            tabs = O.layer_tables(st, d.max(), d.shape[0])
            jobs.append((frames[t], d, tabs[0], tabs[1], tabs[2], weights))
        pool.map(_cpu_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    return {"value": n / dt, "unit": UNIT, "cores": min(procs, n), "kind": "port",
            "sample": f"{n} frames of the same workload ({wl['W']}x{wl['H']}), oracle/sbs_layered.py numpy layer loop, "
                      f"one frame per process on {min(procs, n)} of {cores} host cores; single-frame latency {one:.2f} s"}


def cpu_baseline(wl, frames, raw, budget_s):
    r = reference_cpu_fps(wl, frames, raw, budget_s=budget_s)
    return r if r is not None else cpu_port_fps(wl, frames, raw, budget_s=budget_s)


def run_reference(args, wl, name):
    """--impl reference: the reference's own CPU implementation of the path on this box's host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    B = min(wl["B"], max(8, cores * 4))
    sub = dict(wl, B=B)
    frames, lowres = make_inputs(sub, seed=100)
    from oracle import sbs_layered as O
    raw = np.stack([O.bicubic_resize(lowres[t], wl["H"], wl["W"], 1.0) for t in range(B)])
    vals = []
    for i in range(args.warmup + args.steps):
        if i < args.warmup and i > 0:
            continue                                                 # one warm-up pass is enough for a CPU arm
        r = cpu_baseline(sub, frames, raw, budget_s=args.cpu_budget / max(1, args.steps))
        if i >= args.warmup:
            vals.append(r)
    v = float(np.mean([r["value"] for r in vals]))
    n_frames = sum(int(r.get("frames", r["sample"].split()[0])) for r in vals)
    cpu = dict(vals[-1], value=v)
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * n_frames / max(v, 1e-9) / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8 pixels / fp16 depth", "data": "synthetic",
        "config": {"workload": f"{name}: {wl['W']}x{wl['H']}, fg={wl['fg']} bg={wl['bg']} step={wl['step']}, D-{wl['depth']} depth; "
                               f"each step = a bounded sample of the batch on the host CPU"},
        "cpu_baseline": cpu, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="1080p_b64", choices=sorted(WORKLOADS))
    ap.add_argument("--scatter-mode", type=int, default=0)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--depth-input", default="full", choices=["full", "lowres"],
                    help="full: raw full-resolution fp16 depth (the metric's config); lowres: DPT-resolution map, bicubic on the device")
    ap.add_argument("--host-chunk", type=int, default=0)
    ap.add_argument("--bind", action="store_true",
                    help="bind the process to the CPUs sysfs lists as local to its GPU (shard.bind_near_gpu); off by default: "
                         "not yet measured on a multi-socket 8-GPU box")
    ap.add_argument("--video-frames", type=int, default=3600,
                    help="stream an N-frame synthetic video sharded by clip range over the ranks (BASELINE.json configs[4]; default "
                         "3600 = 2 min of 1080p30 so that the default run stays short, 18000 = the full 10 minutes, 0 = skip)")
    ap.add_argument("--no-4k", dest="also_4k", action="store_false", help="skip the 4k_wide_b16 sub-record (configs[2])")
    ap.add_argument("--no-f32", dest="also_f32", action="store_false", help="skip the fp32-depth sub-record")
    ap.add_argument("--no-producer", action="store_true", help="skip the depth producer timing")
    ap.add_argument("--host-opt", action="append", default=[], help="library option name=value for the host-API context")
    ap.add_argument("--opt", action="append", default=[], help="library option name=value (vrsbs_set_option)")
    ap.add_argument("--pageable", action="store_true", help="also time the host API with pageable numpy buffers")
    ap.add_argument("--cpu-budget", type=float, default=20.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, wl, args.workload)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_ours(args, wl, args.workload)


if __name__ == "__main__":
    main()
