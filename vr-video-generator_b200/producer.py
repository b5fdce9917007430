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
