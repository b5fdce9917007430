"""vr-video-generator_b200 — B200-native (sm_100a) SBS stereo warp: the hot path of
Gia-Huynh/VR-Video-Generator's `SbsProcessor.left_side_sbs`, behind the reference's own entry
points.  See DESIGN.md / INTEGRATION.md.  Import name: `vr_video_generator_b200`."""
__version__ = "0.1.0"

from . import tables, synth  # noqa: F401  (pure python, importable without a GPU)


def __getattr__(name):
    # the CUDA-backed pieces are imported lazily so that `tables`/`synth` work on a CPU-only box,
    # while any use of the warp itself fails loudly when the extension or the GPU is missing
    if name in ("SbsProcessor",):
        from .sbs import SbsProcessor
        return SbsProcessor
    if name in ("_native",):
        import importlib
        return importlib.import_module(__name__ + "._native")
    raise AttributeError(name)
