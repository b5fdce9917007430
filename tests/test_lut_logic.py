"""The fast-path membership tables (cell LUT + two thresholds per entry) classify EVERY fp16 depth value
exactly like the reference's bin test `lo_k <= d < hi_k` (PredictAndGenerate.py:173).  Host mirror of
csrc/table_kernel.cuh (cell_entry, LUT loop) and of the two-compare test in csrc/warp_fused.cuh, checked
exhaustively over all fp16 values up to the frame maximum and all negative values."""
import numpy as np
import pytest

from oracle import sbs_layered as O
from vr_video_generator_b200 import tables


def _bounds(fg, bg, step, H, dmax, last=None):
    cut, rng, steps, limit, offs = tables.layer_tables(dmax, H, fg, bg, step, last)
    lo, hi = O.layer_bounds(cut, steps, np.float16)
    return lo.astype(np.float64), hi.astype(np.float64), rng


def _all_values(dmax):
    pos = np.arange(0, int(np.float16(dmax).view(np.uint16)) + 1, dtype=np.uint16).view(np.float16)
    neg = np.arange(0x8000, 0xFC01, dtype=np.uint16).view(np.float16)          # -0.0 .. -inf
    return np.concatenate([pos, neg]).astype(np.float64)


@pytest.mark.parametrize("fg,bg,step,H,dmax", [
    (0.025, -0.01, 1, 1080, 13.9), (0.025, -0.015, 1, 1080, 16.4), (0.025, -0.01, 2, 1080, 13.9),
    (0.05, -0.03, 1, 2160, 16.9), (0.025, -0.015, 3, 270, 6.5), (0.025, -0.01, 1, 1080, 0.06)])
def test_lut_equals_bin_test(fg, bg, step, H, dmax):
    last = None
    for frame_max in (dmax, dmax * 0.8):                     # second pass exercises the EMA'd range
        lo, hi, last = _bounds(fg, bg, step, H, frame_max, last)
        assert np.all(np.diff(lo) >= 0) and np.all(np.diff(hi) >= 0)
        got = tables.cell_lut(lo, hi, frame_max)
        assert got is not None, "real tables must validate"
        shift, ncells, lut = got
        vals = _all_values(frame_max)
        member = (lo[None, :] <= vals[:, None]) & (vals[:, None] < hi[None, :])
        assert member.sum(axis=1).max() <= 2
        for i in range(0, len(vals), 7):                     # every 7th value + the bin edges below
            want = list(np.nonzero(member[i])[0])
            assert tables.lut_members(lo, hi, shift, ncells, lut, vals[i]) == want, (vals[i], want)
        edges = np.unique(np.concatenate([lo, hi, np.nextafter(lo.astype(np.float16), np.float16(-np.inf)).astype(np.float64),
                                          np.nextafter(hi.astype(np.float16), np.float16(-np.inf)).astype(np.float64)]))
        for v in edges[(edges <= frame_max)]:
            want = list(np.nonzero((lo <= v) & (v < hi))[0])
            assert tables.lut_members(lo, hi, shift, ncells, lut, v) == want, (v, want)


def test_coarse_cells_are_rejected():
    """A LUT that is too coarse for the bins must fail validation instead of misclassifying."""
    lo = np.array([0.0, 0.9, 1.9, 2.9]); hi = np.array([1.1, 2.1, 3.1, 4.1])
    # a cell spanning three bins has no valid entry
    assert tables.cell_entry(lo, hi, 0.5, 3.0) == -1
    assert tables.cell_entry(lo, hi, 1.2, 1.8) in (1, 2)
    # non-overlapping gap: value below every lo -> virtual entry 0
    assert tables.cell_entry(lo, hi, -np.inf, -0.0) == 0
