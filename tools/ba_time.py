#!/usr/bin/env python
"""Device time of the whole S4 solve (one cooperative launch of k_ba_solve), CUDA events on the launch stream."""
import importlib, json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
pkg = importlib.import_module("low-cost-mocap_b200")
synth = pkg.synth

def case(C, F, seed=9, reps=5, **kw):
    obs_obj, poses, K, pts = synth.make_tracks(C, F, seed=seed, missing_frac=0.1)
    start = synth.perturb_poses(poses, seed=seed + 1)
    obs = np.array([[[-1 if v is None else v for v in cam] for cam in fr] for fr in obs_obj], dtype=np.float64)
    mask = np.array([[cam[0] is not None for cam in fr] for fr in obs_obj], dtype=np.uint8)
    ctx = pkg.MocapContext(C)
    ctx.set_cameras([K] * C, start)
    d_obs, d_mask = torch.from_numpy(obs).cuda(), torch.from_numpy(mask).cuda()
    R0 = torch.from_numpy(np.stack([p["R"] for p in start])).cuda().contiguous()
    t0 = torch.from_numpy(np.stack([np.asarray(p["t"]).reshape(3) for p in start])).cuda().contiguous()
    ms = []
    rep = None
    for _ in range(reps + 1):
        R, t = R0.clone(), t0.clone()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        rep = ctx.bundle_adjust_dev(d_obs, d_mask, R, t, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    r = ctx.decode_ba_report(rep)
    return {"cameras": C, "points": F, "options": kw, "ms": float(np.median(ms[1:])), "ms_all": ms[1:], **r}

if __name__ == "__main__":
    out = [case(8, 16000), case(8, 32000), case(16, 6400), case(4, 4000), case(8, 16000, prefit=False, reps=2),
           case(8, 16000, prefit_max_iter=1, max_nfev=1)]
    print(json.dumps(out, indent=1))
