"""Depth producer hand-off: the other side of the queue, batched and kept on the device.

The reference's `inference_worker` (PredictAndGenerate.py:23-61) runs the Depth-Anything-V2 model on ONE frame per
call, up-samples its output to frame size (`dpt.py:196`), multiplies by the per-encoder scaler (:27-34), copies the
result to the CPU (:55) and the SBS worker copies it back to the GPU (:133).  Here the model is still the caller's
(`model.image2tensor` and `model.forward` are the reference's own methods, `dpt.py:204-228` and `:180-188`; the ViT is
outside this package and timed separately), but

  * the frames of a sub-clip go through ONE forward,
  * the DPT-resolution output stays on the device, and
  * the bicubic up-sampling and the scaler run inside the warp's depth pass (`vrsbs_depth_from_lowres`), in the
    reference's order: bicubic in fp16, then `* scaler` in fp16.

`DepthProducer` is a callable with the signature `worker.sbs_worker` expects for `depth_for`, so
`sbs_worker(..., depth_for=producer, scaler=producer.scaler)` is the reference's two processes in one.
A batched forward is not bit-identical to B forwards of one frame in general (the GEMM libraries pick other tiles);
that is model-side numerics, outside the parity scope of the warp (same depth in -> same frame out).
"""
import numpy as np
import torch

from .worker import encoder_scaler


class DepthProducer:
    def __init__(self, model, encoder="vits", input_size=518, max_forward_batch=16):
        """model: an object with the reference's `image2tensor(raw_image, input_size) -> (tensor [1,3,h,w], (H, W))`
        and `forward(x [B,3,h,w]) -> [B,h,w]` (ReLU'd DPT output); `encoder` picks the scaler like :27-34."""
        self.model = model
        self.scaler = encoder_scaler(encoder)
        self.input_size = int(input_size)
        self.max_forward_batch = int(max_forward_batch)

    def __call__(self, frames):
        """frames [n,H,W,3] uint8, exactly what the SBS worker puts on the job queue (:128) -> DPT-resolution depth
        [n,h,w] fp16 on the device (not up-sampled, not scaled)."""
        frames = np.asarray(frames)
        if frames.ndim != 4 or frames.shape[-1] != 3 or frames.dtype != np.uint8:
            raise TypeError("frames must be uint8 [n,H,W,3]")
        outs = []
        with torch.no_grad(), torch.autocast(device_type="cuda", dtype=torch.float16):
            for b0 in range(0, len(frames), self.max_forward_batch):
                x = torch.cat([self.model.image2tensor(f, self.input_size)[0] for f in frames[b0:b0 + self.max_forward_batch]])
                d = self.model.forward(x)
                if d.dim() != 3 or d.shape[0] != x.shape[0]:
                    raise ValueError(f"model.forward returned {tuple(d.shape)} for {tuple(x.shape)}")
                if d.dtype != torch.float16:
                    raise TypeError(f"depth must be float16 (the dtype autocast produces at :54-55), got {d.dtype}")
                outs.append(d)
        return (torch.cat(outs) if len(outs) > 1 else outs[0]).contiguous()


def reference_depth(model, frame, scaler, input_size=518):
    """What the reference's inference worker puts on the result queue for one frame (:54-55), for comparisons:
    `model.infer_image_gpu(img) * scaler` under fp16 autocast, restated with the two calls used above."""
    with torch.no_grad(), torch.autocast(device_type="cuda", dtype=torch.float16):
        x, (h, w) = model.image2tensor(frame, input_size)
        d = model.forward(x)
        d = torch.nn.functional.interpolate(d[:, None], (h, w), mode="bicubic", align_corners=True)[0][0]
        return d * scaler


# ---- f2: the reference's producer process, protocol for protocol ------------------------------------------------------
def inference_worker(in_queue_list, out_queue_list, notify_queue_list, DEVICE, args_god, model=None, load_model=None,
                     on_device=False, lowres=False, input_size=518):
    """Mirror of `inference_worker` (PredictAndGenerate.py:23-61): one depth model serving several SBS workers.

    Protocol, unchanged: an SBS worker announces a job with `notify_queue.put((idx,))` and puts `(img,)` on its own job
    queue (`SbsProcessor.add_frame`, :127-129); this loop pops the notification, pops that worker's job, runs the model
    under `no_grad` + fp16 autocast, multiplies by the per-encoder scaler (:27-34) and puts the result on that worker's
    result queue; `None` on the notify queue or on a job queue ends the loop (:46-51, the shutdown `main_func` sends at
    :316-319).  A two-deep ring per client keeps the tensor a client may still be reading alive (:40,53-56).  The warm-up
    forward on a black 1080p frame runs outside autocast like :37.

    `model` (or `load_model(encoder, encoder_path, DEVICE)`, the reference's `SupportFunction.load_model`) supplies
    `infer_image_gpu(img)`; nothing of the ViT lives in this package.

    on_device=False is the reference's hand-off: the full-resolution depth goes `.to('cpu')` and through the queue (the
    consumer's `left_side_sbs` then takes the host pipeline).  on_device=True is for SBS workers that are THREADS of this
    process (queue.Queue): the tensor stays on the GPU - what the author tried and gave up on across processes (:54) - and
    is complete before it is put (an event is waited for on the producer's side, never on the consumer's).  With
    lowres=True the DPT-resolution map (`image2tensor` + `forward`, dpt.py:180-228) is handed over instead and the bicubic
    tail and the scaler run inside the warp's depth pass (`left_side_sbs_batch(..., scaler=producer scaler)`); the queue
    item is then `(tensor, scaler)`.  Returns the number of frames served."""
    scaler = encoder_scaler(args_god.encoder)
    dev = DEVICE if isinstance(DEVICE, torch.device) else torch.device(DEVICE)
    if model is None:
        if load_model is None:
            raise ValueError("inference_worker needs a model or a load_model callable")
        model = load_model(args_god.encoder, args_god.encoder_path, dev)
    warm = model.infer_image_gpu(np.zeros((1080, 1920, 3), dtype=np.uint8))
    ring = [[warm.detach().clone(), warm.detach().clone()] for _ in out_queue_list]
    del warm
    served = 0
    while True:
        queue_idx = notify_queue_list.get()
        if queue_idx is None:
            break
        task = in_queue_list[queue_idx[0]].get()
        if task is None:
            break
        img = task[0]
        del ring[queue_idx[0]][0]
        with torch.no_grad(), torch.autocast(device_type=dev.type, dtype=torch.float16):
            if lowres:
                x, _hw = model.image2tensor(img, input_size)
                d = model.forward(x)[0].contiguous()
            else:
                d = model.infer_image_gpu(img) * scaler
            if not on_device:
                d = d.to(torch.device("cpu"))
            elif d.is_cuda:
                ev = torch.cuda.Event()
                ev.record(torch.cuda.current_stream(d.device))
                ev.synchronize()
        ring[queue_idx[0]].append(d)
        out_queue_list[queue_idx[0]].put((d, scaler) if lowres else d)
        served += 1
    return served
