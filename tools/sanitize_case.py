"""Small end-to-end cases for compute-sanitizer (memcheck / racecheck): every route once, tiny frames."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from vr_video_generator_b200 import _native, tables
from conftest import load_case, golden_weights

for name in ("small_a", "medium"):
    meta, frames, raw, ref_left = load_case(name)
    p = meta["params"]; H, W, n = p["H"], p["W"], p["n"]
    for mode in (0, 3, 4, 2):
        ctx = _native.Context(0, H, W, 4, 512)
        ctx.reset(p["fg"], p["bg"], p["step"], True); ctx.set_blur_weights(golden_weights(meta))
        ctx.set_option("fused", 0 if mode == 2 else 1); ctx.set_option("fast_tables", 0 if mode == 3 else 1)
        ctx.set_option("smooth_in_warp", 1 if mode == 4 else 0)
        f = torch.from_numpy(np.ascontiguousarray(frames)).cuda(); r = torch.from_numpy(np.ascontiguousarray(raw)).cuda()
        out = torch.empty((n, H, 2 * W, 3), dtype=torch.uint8, device="cuda"); dep = torch.empty((n, H, W), dtype=torch.float16, device="cuda")
        ctx.process_batch(f.data_ptr(), r.data_ptr(), n, H, W, dep.data_ptr(), out.data_ptr(), torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        ok = np.array_equal(out.cpu().numpy()[:, :, :W], ref_left)
        print(name, "mode", mode, "matches reference:", ok)
        ctx.close()
