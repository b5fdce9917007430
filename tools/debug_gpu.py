import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from oracle import sbs_layered as O, scatter as S
from vr_video_generator_b200 import _native, synth, tables
from conftest import load_case, golden_weights

def lowres_dbg():
    h, w, H, W = 74, 132, 270, 480
    lo = synth.depth_stress(2, h, w, seed=21)
    want = np.stack([O.bicubic_resize(lo[i], H, W, 1.0) for i in range(2)])
    lo_t = torch.from_numpy(np.ascontiguousarray(lo)).cuda()
    s = torch.cuda.current_stream().cuda_stream
    for contract in (1, 0, 1):
        for B in (1, 2):
            out = torch.zeros((2, H, W), dtype=torch.float16, device="cuda")
            ctx = _native.Context(0, H, W, 8, 64)
            ctx.reset(0.025, -0.01, 1, False)
            ctx.set_option("bicubic_contract", contract)
            ctx.depth_from_lowres(lo_t.data_ptr(), B, h, w, 1.0, H, W, out.data_ptr(), s)
            torch.cuda.synchronize()
            got = out.cpu().numpy()
            st = O.WarpState(); sm = np.stack([O.smooth_depth(st, want[i]) for i in range(B)])
            print("contract", contract, "B", B, "exact", [float(np.mean(got[i] == sm[i])) for i in range(B)],
                  "got[0,0,:4]", got[0, 0, :4], "want", sm[0, 0, :4], "got[0,1,:2]", got[0, 1, :2], "want", sm[0, 1, :2])
            ctx.close()

def full_dbg(name, mode):
    meta, frames, raw, _ = load_case(name)
    p = meta["params"]; w = golden_weights(meta)
    H, W, n = p["H"], p["W"], p["n"]
    ctx = _native.Context(0, H, W, 4, 512)
    ctx.reset(p["fg"], p["bg"], p["step"], True); ctx.set_blur_weights(w); ctx.set_option("scatter_mode", mode)
    f = torch.from_numpy(np.ascontiguousarray(frames)).cuda(); r = torch.from_numpy(np.ascontiguousarray(raw)).cuda()
    out = torch.empty((n, H, 2 * W, 3), dtype=torch.uint8, device="cuda"); dep = torch.empty((n, H, W), dtype=torch.float16, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    ctx.process_batch(f.data_ptr(), r.data_ptr(), n, H, W, dep.data_ptr(), out.data_ptr(), s)
    infos = ctx.frame_info(n, s); masks = ctx.hole_mask(n, H, W, s)
    torch.cuda.synchronize()
    sbs = out.cpu().numpy(); depg = dep.cpu().numpy()
    st = O.WarpState(p["fg"], p["bg"], p["step"])
    for t in range(n):
        stg = {}
        want = S.process_frame(st, frames[t], raw[t], weights=w, stages=stg)
        dm = (depg[t].view(np.uint16) != stg["depth"].view(np.uint16))
        mm = masks[t] != stg["holes"]
        im = (sbs[t] != want).any(axis=2)
        rows = np.nonzero(mm.any(axis=1))[0]
        print(name, "mode", mode, "t", t, "L", infos[t].layers, len(stg["steps"]), "holes dev", infos[t].holes, int(masks[t].sum()), "oracle", int(stg["holes"].sum()),
              "depth mism", int(dm.sum()), "mask mism", int(mm.sum()), "img mism px", int(im.sum()), "right-half mism", int(im[:, W:].sum()),
              "rows with mask mism", len(rows), rows[:12], "extra holes", int((masks[t] & ~stg["holes"]).sum()), "missing holes", int((~masks[t] & stg["holes"]).sum()))
        if len(rows):
            y = rows[0]; xs = np.nonzero(mm[y])[0]
            print("   first bad row", y, "cols", xs[:20], "n", len(xs), " row mod grid?", y % 148)
    ctx.close()

if __name__ == "__main__":
    lowres_dbg()
    for mode in (1, 2):
        full_dbg("full_1080p_step2", mode)
    full_dbg("full_1080p_cfg1", 2)
