"""Drop-in `SbsProcessor` backed by the sm_100a CUDA library (no CPU fallback).

Mirrors the reference class (PredictAndGenerate.py:63-198): same constructor, `add_frame`,
`get_depth`, `get_cutoff`, `left_side_sbs`, same parameters read from `args_god`
(`offset_fg`, `offset_bg`, `offset_step_size`), same per-instance clip-range state (two raw
depth frames of history + the previous offset range).  Batched / device-resident entry points are
additions (`left_side_sbs_batch`, `warp_batch_device`).

PyTorch is used for device memory and streams only.
"""
import numpy as np
import torch

from . import _native, tables

DEFAULT_MAX_LAYERS = 512


class SbsProcessor:
    def __init__(self, gpu_notify_queue, gpu_notify_worker_idx, args_god, debug_config=[None],
                 device=None, max_batch=64, max_layers=DEFAULT_MAX_LAYERS, depth_dtype=None):
        self.debug_filePrefix = debug_config[0]
        self.gpu_notify_queue = gpu_notify_queue
        self.gpu_notify_worker_idx = gpu_notify_worker_idx
        self.args_god = args_god
        self.offset_step_size = args_god.offset_step_size
        self.offset_bg = args_god.offset_bg
        self.offset_fg = args_god.offset_fg
        self.sigmaboi = 3
        self.depth_dampening_count = 2
        self.depth_dampening_ratio = 0.4
        self.depth_dampening_initial_value = 0.3
        self.depth_dampening_original_ratio, _ = tables.smoothing_weights(
            self.depth_dampening_count, self.depth_dampening_initial_value, self.depth_dampening_ratio)

        if not torch.cuda.is_available():
            raise RuntimeError("vr-video-generator_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.max_batch, self.max_layers = max_batch, max_layers
        self._ctx = None
        self._shape = None          # (H, W) the context was created for
        self._blur = True
        self._ctx_f32 = False
        # Element type of the full-resolution depth of this clip range.  The reference is dtype-agnostic (smoothing, max and
        # bin comparison run in the tensor's dtype): fp16 is what autocast produced under the reference's pinned torch,
        # torch >= 2.4 on CUDA hands over fp32 (upsample_bicubic2d is on autocast's fp32 list).  None = taken from the
        # first full-resolution depth seen (fp16 when only DPT-resolution maps are given).
        self._f32 = None if depth_dtype is None else _is_f32(depth_dtype)
        self._inflight = {}         # ticket -> arrays of a submitted batch (kept alive until collected)
        self._depth_pool = []       # recycled page-locked staging buffers for host depth given to submit_batch

    # ------------------------------------------------------------------------------------------
    def _context(self, H, W, f32=None):
        """The native context for (H, W) and the clip's depth dtype; a new size or dtype starts a new context (and, like a
        new reference SbsProcessor, a fresh clip state)."""
        if f32 is not None and self._f32 is None:
            self._f32 = bool(f32)
        elif f32 is not None and bool(f32) != self._f32:
            self._f32 = bool(f32)
            self.close()
        if self._ctx is None or self._shape != (H, W):
            if self._ctx is not None:
                self._ctx.close()
            self._ctx = _native.Context(self.device.index, H, W, self.max_batch, self.max_layers)
            self._ctx_f32 = bool(self._f32)
            self._ctx.reset(self.offset_fg, self.offset_bg, self.offset_step_size, self._blur,
                            _native.DEPTH_F32 if self._ctx_f32 else _native.DEPTH_F16)
            kx, ky = tables.blur_kernel_shape(H)
            self._ctx.set_blur_weights(tables.gaussian_weights(kx, ky, float(self.sigmaboi)))
            self._shape = (H, W)
        elif bool(self._f32) != self._ctx_f32:              # dtype became known after the context was made
            self._ctx_f32 = bool(self._f32)
            self._ctx.reset(self.offset_fg, self.offset_bg, self.offset_step_size, self._blur,
                            _native.DEPTH_F32 if self._ctx_f32 else _native.DEPTH_F16)
        return self._ctx

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    def close(self):
        if self._ctx is not None:
            self._ctx.close()
            self._ctx = None

    def reset_state(self):
        """Forget depth history and range EMA (what a new reference SbsProcessor starts with)."""
        if self._ctx is not None:
            self._ctx.reset(self.offset_fg, self.offset_bg, self.offset_step_size, self._blur,
                            _native.DEPTH_F32 if self._ctx_f32 else _native.DEPTH_F16)

    @property
    def last_offset_range(self):
        return None if self._ctx is None else self._ctx.get_range_state()

    @last_offset_range.setter
    def last_offset_range(self, value):
        if self._ctx is None:
            raise RuntimeError("no frame size known yet; process a frame or call _context(H, W) first")
        self._ctx.set_range_state(value)

    # ---- the reference's entry points -------------------------------------------------------
    def add_frame(self, raw_img, job_queue, result_queue):
        """Ask the depth producer for this frame (PredictAndGenerate.py:127-129)."""
        self.gpu_notify_queue.put((self.gpu_notify_worker_idx,))
        job_queue.put((raw_img,))

    def get_depth(self, raw_img, job_queue, result_queue):
        """Pop the producer's raw depth and return the temporally smoothed depth as a CUDA fp16
        tensor [H,W] (PredictAndGenerate.py:131-145).  Advances the depth history."""
        raw = result_queue.get()
        return self._smooth(self._raw_to_device(raw))[0]

    def get_cutoff(self, depth):
        """Python lists exactly like the reference's get_cutoff (PredictAndGenerate.py:101-126);
        updates the shared range state."""
        ctx = self._context(depth.shape[0], depth.shape[1])
        out = tables.layer_tables(float(depth.max()), depth.shape[0], self.offset_fg, self.offset_bg,
                                  self.offset_step_size, ctx.get_range_state())
        ctx.set_range_state(out[1])
        return out

    def left_side_sbs(self, raw_img, job_queue, result_queue):
        """numpy [H,W,3] uint8 RGB -> numpy [H,2W,3] uint8 SBS frame (PredictAndGenerate.py:157-198).
        The depth popped from `result_queue` is a CPU tensor in the reference (inference_worker does `.to('cpu')`,
        :55-56): that case goes through the host pipeline (pinned staging, only the synthesised half crosses PCIe
        on the way back).  A CUDA depth tensor is used where it is."""
        H, W, _ = raw_img.shape
        raw = result_queue.get()
        if isinstance(raw, tuple):
            # (DPT-resolution map, scaler) from producer.inference_worker(lowres=True): bicubic + scaler on the device
            lo, sc = raw
            return self.left_side_sbs_batch(np.ascontiguousarray(raw_img)[None], lo if lo.dim() == 3 else lo[None], scaler=sc)[0]
        if not (isinstance(raw, torch.Tensor) and raw.is_cuda):
            d = _as_numpy(raw)
            _check_depth_dtype(d.dtype)
            ctx = self._context(H, W, d.dtype == np.float32)
            if d.shape != (H, W):
                raise ValueError(f"depth {d.shape} does not match the frame {(H, W)}")
            out = np.empty((H, 2 * W, 3), dtype=np.uint8)
            f = np.ascontiguousarray(raw_img)
            ctx.process_host(f.ctypes.data, d.ctypes.data, 1, H, W, 0, 0, 1.0, out.ctypes.data)
            return out
        with torch.cuda.device(self.device):
            img = torch.from_numpy(np.ascontiguousarray(raw_img)).to(self.device, non_blocking=True)
            depth = self._smooth(self._raw_to_device(raw))
            ctx = self._ctx
            st = self._stream()
            ctx.build_tables(1, H, W, st)
            sbs = torch.empty((1, H, 2 * W, 3), dtype=torch.uint8, device=self.device)
            ctx.warp_batch(img.data_ptr(), depth.data_ptr(), 1, H, W, sbs.data_ptr(), st)
            out = sbs[0].cpu().numpy()           # blocking D2H, like the reference
            ctx.frame_info(1, st)                # raises if the device rejected the frame
        return out

    # ---- batched additions ---------------------------------------------------------------------
    def left_side_sbs_batch(self, frames, depths, scaler=1.0, out=None):
        """Host-buffer batch: frames [B,H,W,3] uint8 and depths [B,H,W] fp16 (raw, full-res) or
        [B,h,w] fp16 (DPT low-res; bicubic + `scaler` applied on the device).  numpy arrays or CPU
        tensors, pinned or pageable; the depths may also be a CUDA tensor (the producer's output left on
        the device).  Returns numpy [B,H,2W,3].  Pipelined (pinned double buffering)."""
        f = _as_numpy(frames)
        B, H, W, _ = f.shape
        ctx = self._context(H, W, _full_res_f32(depths, H, W))
        dshape, dptr, _keep = self._depth_arg(ctx, depths, B)
        lowres = tuple(dshape[1:]) != (H, W)
        if out is None:
            out = np.empty((B, H, 2 * W, 3), dtype=np.uint8)
        ctx.process_host(f.ctypes.data, dptr, B, H, W, dshape[1] if lowres else 0,
                         dshape[2] if lowres else 0, float(scaler), out.ctypes.data)
        return out

    def submit_batch(self, frames, depths, out, scaler=1.0):
        """Asynchronous `left_side_sbs_batch` for PAGE-LOCKED buffers (see `pinned_sbs_buffer`): enqueues the copies and
        kernels of this batch and returns a ticket; `collect(ticket)` blocks until `out` holds the SBS frames.  Batches
        run in submission order and share the clip-range state, so the caller can decode the next sub-clip and encode the
        previous one while this one is on the GPU (vrsbs_submit_host).  `frames` may be `out[:, :, W:, :]` - the caller
        decoded straight into the right halves of the SBS buffer: nothing is copied on the host then, and only the
        synthesised left halves cross PCIe on the way back.  The arrays must stay alive and untouched until collected."""
        B, H, W, _ = frames.shape
        if out.shape != (B, H, 2 * W, 3) or out.dtype != np.uint8 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous uint8 [B,H,2W,3] array")
        if frames.dtype != np.uint8 or frames.strides[2:] != (3, 1):
            raise ValueError("frames must be uint8 with packed pixels and rows")
        ctx = self._context(H, W, _full_res_f32(depths, H, W))
        dshape, dptr, keep = self._depth_arg(ctx, depths, B)
        staged = None
        if not (isinstance(depths, torch.Tensor) and (depths.is_cuda or depths.is_pinned())):
            # pageable host depth (what a result queue delivers): one copy into a recycled page-locked buffer
            staged = self._pinned_depth(keep.nbytes)
            view = staged.numpy()[:keep.nbytes].view(keep.dtype).reshape(keep.shape)
            np.copyto(view, keep)
            dptr, keep = view.ctypes.data, (staged, view)
        lowres = tuple(dshape[1:]) != (H, W)
        fptr = frames.ctypes.data
        in_place = fptr == out.ctypes.data + 3 * W and frames.strides[:2] == (out.strides[0], out.strides[1])
        t = ctx.submit_host(fptr, frames.strides[1], frames.strides[0], dptr, B, H, W, dshape[1] if lowres else 0,
                            dshape[2] if lowres else 0, float(scaler), out.ctypes.data,
                            _native.HOST_RIGHT_IN_PLACE if in_place else 0)
        self._inflight[t] = (frames, keep, out, staged)
        return t

    def collect(self, ticket):
        """Block until the batch behind `ticket` is complete; raises if the device rejected one of its frames."""
        try:
            self._ctx.collect(ticket)
        finally:
            entry = self._inflight.pop(ticket, None)
            if entry is not None and entry[3] is not None:
                self._depth_pool.append(entry[3])

    def _pinned_depth(self, nbytes):
        for i, t in enumerate(self._depth_pool):
            if t.numel() >= nbytes:
                return self._depth_pool.pop(i)
        return torch.empty((nbytes,), dtype=torch.uint8).pin_memory()

    def _depth_arg(self, ctx, depths, B):
        """(shape, pointer, object to keep alive) of a depth batch given as numpy / CPU tensor / CUDA tensor."""
        if isinstance(depths, torch.Tensor) and depths.is_cuda:
            # the depth producer's output is still on the device (same process): used where it is, no H2D of depth; the
            # library's kernels are ordered behind the producer's stream by an event (the host does not wait)
            _check_depth_dtype(depths.dtype)
            _check_cuda(depths, depths.dtype)
            ctx.host_depends_on(torch.cuda.current_stream(depths.device).cuda_stream)
            dshape, dptr, keep = tuple(depths.shape), depths.data_ptr(), depths
        else:
            d = _as_numpy(depths)
            _check_depth_dtype(d.dtype)
            dshape, dptr, keep = d.shape, d.ctypes.data, d
        if len(dshape) != 3 or dshape[0] != B:
            raise ValueError(f"depth {dshape} does not match {B} frames")
        return dshape, dptr, keep

    def warp_batch_device(self, frames, raw_depth, out=None, depth_scratch=None, check=False):
        """Device-resident batch on the current stream, asynchronous: frames [B,H,W,3] uint8 CUDA,
        raw_depth [B,H,W] fp16 / fp32 CUDA (raw, full-res) -> sbs [B,H,2W,3] uint8 CUDA.  A frame the device rejects (NaN
        depth - the reference raises in `math.ceil` - or more layers than `max_layers`) leaves garbage in its SBS frame:
        `check=True` waits for the stream and raises like the host calls do; otherwise call `frame_status(B)` once the
        stream has been synchronised.  After such an error the clip state is undefined (`reset_state`)."""
        B, H, W, _ = frames.shape
        _check_depth_dtype(raw_depth.dtype)
        ctx = self._context(H, W, raw_depth.dtype == torch.float32)
        if out is None:
            out = torch.empty((B, H, 2 * W, 3), dtype=torch.uint8, device=self.device)
        if depth_scratch is None:
            depth_scratch = torch.empty((B, H, W), dtype=raw_depth.dtype, device=self.device)
        _check_cuda(frames, torch.uint8), _check_cuda(raw_depth, raw_depth.dtype)
        ctx.process_batch(frames.data_ptr(), raw_depth.data_ptr(), B, H, W, depth_scratch.data_ptr(),
                          out.data_ptr(), self._stream())
        if check:
            ctx.frame_info(B, self._stream())      # synchronises the stream; raises VrsbsError(E_FRAME) for a rejected frame
        return out

    def frame_status(self, B):
        """Per-frame records of the last device-resident batch (layers, limit_step, strip, holes, status); waits for the
        current stream and raises if the device rejected a frame."""
        return self._ctx.frame_info(B, self._stream())

    # ---- helpers ---------------------------------------------------------------------------------
    def _raw_to_device(self, raw):
        if isinstance(raw, np.ndarray):
            raw = torch.from_numpy(raw)
        _check_depth_dtype(raw.dtype)
        return raw.to(self.device, non_blocking=True).contiguous()

    def _smooth(self, raw_dev):
        """raw_dev [H,W] or [B,H,W] fp16 CUDA -> smoothed [B,H,W]; leaves the frame maxima on the device."""
        if raw_dev.dim() == 2:
            raw_dev = raw_dev[None]
        B, H, W = raw_dev.shape
        ctx = self._context(H, W, raw_dev.dtype == torch.float32)
        out = torch.empty_like(raw_dev)
        ctx.depth_from_full(raw_dev.data_ptr(), B, H, W, out.data_ptr(), self._stream())
        return out


def pinned_sbs_buffer(n, H, W):
    """Page-locked SBS buffer for `submit_batch`: (out [n,H,2W,3] uint8 numpy, frames view out[:, :, W:, :], owner).
    Decode into the frames view (e.g. `np.copyto(frames[i], bgr[:, :, ::-1])`), submit, and after `collect` hand `out`
    to the encoder - the right halves were never copied.  Keep `owner` (the torch tensor) alive as long as the arrays."""
    t = torch.empty((n, H, 2 * W, 3), dtype=torch.uint8)
    if torch.cuda.is_available():             # (CPU-only hosts run the worker's control-flow tests with a fake processor)
        t = t.pin_memory()
    out = t.numpy()
    return out, out[:, :, W:, :], t


def _as_numpy(x):
    if isinstance(x, torch.Tensor):
        x = x.numpy()
    return np.ascontiguousarray(x)


def _is_f32(dt):
    return str(dt).replace("torch.", "") in ("float32", "<class 'numpy.float32'>") or dt in (np.float32, torch.float32)


def _check_depth_dtype(dt):
    if not (dt in (np.float16, np.float32, torch.float16, torch.float32)):
        raise TypeError(f"depth must be float16 or float32 (what the producer's autocast hands over), got {dt}")


def _full_res_f32(depths, H, W):
    """True / False when `depths` is a full-resolution map (its dtype is the clip's depth dtype), None for a
    DPT-resolution map (always fp16; the clip's dtype then decides what the device tail produces)."""
    shape = tuple(depths.shape)
    if len(shape) == 3 and shape[1:] == (H, W):
        return _is_f32(depths.dtype)
    if True:
        if depths.dtype not in (np.float16, torch.float16):
            raise TypeError(f"a DPT-resolution depth map must be float16 (the model's autocast output), got {depths.dtype}")
    return None


def _check_cuda(t, dtype):
    if not (t.is_cuda and t.is_contiguous() and t.dtype == dtype):
        raise TypeError(f"expected a contiguous CUDA {dtype} tensor")
