"""Stage-by-stage diff of the fused route against the C oracle on small cases (GPU box only)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from oracle import sbs_layered as O, scatter as S
from vr_video_generator_b200 import _native
from conftest import load_case, golden_weights


def run(name, mode, blur):
    meta, frames, raw, _ = load_case(name)
    p = meta["params"]; w = golden_weights(meta)
    H, W, n = p["H"], p["W"], p["n"]
    ctx = _native.Context(0, H, W, 4, 512)
    ctx.reset(p["fg"], p["bg"], p["step"], blur); ctx.set_blur_weights(w)
    ctx.set_option("fused", 0 if mode in (1, 2) else 1); ctx.set_option("fast_tables", 0 if mode in (3, 5) else 1); ctx.set_option("smooth_in_warp", 1 if mode in (4, 5) else 0)
    if mode in (1, 2): ctx.set_option("scatter_mode", mode)
    f = torch.from_numpy(np.ascontiguousarray(frames)).cuda(); r = torch.from_numpy(np.ascontiguousarray(raw)).cuda()
    out = torch.zeros((n, H, 2 * W, 3), dtype=torch.uint8, device="cuda"); dep = torch.empty((n, H, W), dtype=torch.float16, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    ctx.process_batch(f.data_ptr(), r.data_ptr(), n, H, W, dep.data_ptr(), out.data_ptr(), s)
    torch.cuda.synchronize()
    infos = ctx.frame_info(n, s); masks = ctx.hole_mask(n, H, W, s)
    sbs = out.cpu().numpy()
    st = O.WarpState(p["fg"], p["bg"], p["step"])
    for t in range(n):
        stg = {}
        want = S.process_frame(st, frames[t], raw[t], weights=w, stages=stg)
        ref = want if blur else np.concatenate([stg["pre_blur"], frames[t]], axis=1)
        mm = masks[t] != stg["holes"]
        im = (sbs[t] != ref).any(axis=2)
        print(f"{name} mode {mode} blur {blur} t {t}: L {infos[t].layers}/{len(stg['steps'])} holes {infos[t].holes}/{int(stg['holes'].sum())} "
              f"mask-mism {int(mm.sum())} left-mism {int(im[:, :W].sum())} right-mism {int(im[:, W:].sum())} "
              f"mism-at-holes {int((im[:, :W] & stg['holes']).sum())} mism-at-painted {int((im[:, :W] & ~stg['holes']).sum())}")
        if im[:, :W].any():
            ys, xs = np.nonzero(im[:, :W]); print("   first mism (y,x):", list(zip(ys[:8], xs[:8])), "got", sbs[t][ys[0], xs[0]], "want", ref[ys[0], xs[0]])
    ctx.close()


if __name__ == "__main__":
    for name in ("small_a", "small_neg", "small_zero", "medium"):
        for mode in (0, 3):
            for blur in (False, True):
                try:
                    run(name, mode, blur)
                except Exception as e:
                    print(name, mode, blur, "EXC", e)
