#!/usr/bin/env python
"""Which capacity flags does the config-3/4 workload raise, per pool seed and per limit set?"""
import importlib.util, importlib, os, sys, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py")); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
pkg = importlib.import_module("low-cost-mocap_b200")
b.N_CAM, b.N_MARKERS, b.BATCH, b.POOL, b.MAX_ROOTS = 8, 16, 4000, 4000, 64
dev = torch.device("cuda", 0)
out = {}
for seed in (0, 6, 3):
    frames, truth, poses, K = b.render_pool_on_device(torch, dev, 4000, seed=seed)
    for lim in ({"max_roots": 64}, {"max_roots": 64, "max_cands": 16}, {"max_roots": 64, "max_cands": 16, "max_groups": 65536}, {"max_roots": 128, "max_cands": 16, "max_groups": 65536}):
        ctx = pkg.MocapContext(8, 640, 480, **lim)
        ctx.set_cameras([K] * 8, poses)
        o = ctx.pipeline(frames)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(5): ctx.pipeline(frames, out=o)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / 5 * 1e3
        fl = o["flags"].cpu().numpy()
        hist = {int(v): int((fl == v).sum()) for v in np.unique(fl)}
        out[f"seed{seed} {lim}"] = {"flags_hist": hist, "ms": round(ms, 3), "max_n": int(o["n"].max().item())}
        print(seed, lim, hist, round(ms, 3), int(o["n"].max().item()), flush=True)
        del ctx
    del frames
