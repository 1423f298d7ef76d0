#!/usr/bin/env python
"""Chunked matcher (MOCAP_MATCH_CHUNK = candidate groups per work item, 0 = one warp per frame-set): time per batch of the
three-kernel pipeline and bit-equality of every output with the unchunked matcher, on the config-3 shape (8 cameras x 16
markers, two marker layouts) and on config 2 (4 x 4)."""
import importlib.util, importlib, os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py")); b = importlib.util.module_from_spec(spec); spec.loader.exec_module(b)
pkg = importlib.import_module("low-cost-mocap_b200")
dev = torch.device("cuda", 0)
res = {}


def run(tag, ncam, nmark, nsets, limits, variants, seed):
    b.N_CAM, b.N_MARKERS = ncam, nmark
    frames, truth, poses, K = b.render_pool_on_device(torch, dev, nsets, seed=seed)
    base = None
    for mode, chunk in variants:
        os.environ["MOCAP_MATCH_CHUNK"] = str(chunk)
        os.environ["MOCAP_PIPELINE"] = mode
        ctx = pkg.MocapContext(ncam, 640, 480, **limits)
        ctx.set_cameras([K] * ncam, poses)
        o = ctx.pipeline(frames, want_tracks=True)
        for _ in range(5): ctx.pipeline(frames, out=o)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); e0.record()
        for _ in range(20): ctx.pipeline(frames, out=o)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        # the matcher alone, on the blob lists of the same batch
        d = ctx.detect(frames)
        m = ctx.match_triangulate(d["xy"], d["n"])
        torch.cuda.synchronize(); e0.record()
        for _ in range(10): ctx.match_triangulate(d["xy"], d["n"])
        e1.record(); torch.cuda.synchronize()
        ms_match = e0.elapsed_time(e1) / 10
        cur = {k: o[k].clone() for k in ("n", "flags", "obj", "err", "track_xy")}
        if base is None:
            base, same = cur, None
        else:
            n = base["n"]; R = cur["obj"].shape[1]
            live = torch.arange(R, device=dev)[None, :] < n[:, None]
            same = bool(torch.equal(cur["n"], n) and torch.equal(cur["flags"], base["flags"])
                        and torch.equal(cur["obj"][live], base["obj"][live]) and torch.equal(cur["err"][live], base["err"][live])
                        and torch.equal(cur["track_xy"][live], base["track_xy"][live]))
        res[f"{tag} {mode} chunk={chunk}"] = {"ms": round(ms, 4), "matcher_ms": round(ms_match, 4), "equal_to_first": same, "flagged": int((cur["flags"] != 0).sum())}
        print(tag, mode, chunk, round(ms, 4), round(ms_match, 4), same, flush=True)
        del ctx
    del frames


lim8 = {"max_roots": 64, "max_groups": 65536}
run("c8m16 seed0", 8, 16, 4000, lim8, [("split", 0), ("split", 256), ("split", 512), ("split", 1024), ("split", 2048)], 0)
run("c8m16 seed6", 8, 16, 4000, lim8, [("split", 0), ("split", 256), ("split", 512), ("split", 1024), ("split", 2048)], 6)
run("c4m4 seed0", 4, 4, 8192, {"max_roots": 16}, [("fused", 0), ("fused", 1024), ("split", 0), ("split", 1024)], 0)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "matcher_probe.json"), "w"), indent=1)
