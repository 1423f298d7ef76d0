#!/usr/bin/env python
"""Secondary (not HBM-bound) figures of SURVEY.md §8(d): per-stage throughput on the shapes of
BASELINE configs 3 and 5.  Prints one JSON object; run on the GPU box:

    python tests/stage_bench.py > gpurun_out/stage_bench.json
"""
import importlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))   # tests/ -> repo root
sys.path.insert(0, ROOT)
import torch  # noqa: E402

pkg = importlib.import_module("low-cost-mocap_b200")
synth = pkg.synth
from oracle.ref_port import RefPort  # noqa: E402  (CPU baseline leg only)


def timed(fn, reps=5):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def config3_pipeline():
    C, M, P, B = 8, 16, 64, 4000
    frames, truth, poses, K = synth.make_frame_pool(C, M, P, seed=1)
    ctx = pkg.MocapContext(C, 640, 480, max_roots=64)
    ctx.set_cameras([K] * C, poses)
    pool = torch.from_numpy(frames).cuda()
    batch = pool[torch.arange(B, device="cuda") % P].contiguous()     # 9.8 GB
    out = ctx.alloc_tracks(B)
    ms = timed(lambda: ctx.pipeline(batch, out=out))
    n = out["n"].cpu().numpy()
    flags = out["flags"].cpu().numpy()
    # CPU reference on a few frame-sets
    port = RefPort([K] * C)
    t0 = time.perf_counter()
    for b in range(3):
        pts = [port.find_dot(np.repeat(frames[b, c][:, :, None], 3, axis=2)) for c in range(C)]
        port.match_and_triangulate(pts, poses)
    cpu_ms = (time.perf_counter() - t0) / 3 * 1e3
    return {"workload": "8 cameras, 16 markers, 4000 frame-sets of 640x480 (9.8 GB resident)", "ms_per_batch": ms,
            "frame_sets_per_s": B / ms * 1e3, "hbm_gbs": B * C * 307200 / ms / 1e6, "points_per_frame_set": float(n.mean()),
            "overflow_flags": int((flags != 0).sum()), "cpu_reference_ms_per_frame_set_1core": cpu_ms,
            "pipeline": os.environ.get("MOCAP_PIPELINE", "auto (single-pass kernel first, three-kernel pipeline after a heavy batch)")}


def colour_pipeline():
    """The layout _find_dot actually receives (helpers.py:143-145): 3 interleaved channels per pixel.  Config-2
    shape, every channel carrying the synthetic grey frame, compared with the 1-channel run of the same frames."""
    C, M, P, B = 4, 4, 64, 2000
    frames, truth, poses, K = synth.make_frame_pool(C, M, P, seed=1)
    ctx = pkg.MocapContext(C, 640, 480, max_roots=16)
    ctx.set_cameras([K] * C, poses)
    pool = torch.from_numpy(frames).cuda()
    grey = pool[torch.arange(B, device="cuda") % P].contiguous()
    colour = grey[..., None].expand(-1, -1, -1, -1, 3).contiguous()          # 7.4 GB
    out1, out3 = ctx.alloc_tracks(B), ctx.alloc_tracks(B)
    ctx.pipeline(grey, out=out1)
    ms = timed(lambda: ctx.pipeline(colour, out=out3))
    valid = torch.arange(out1["obj"].shape[1], device="cuda")[None, :] < out1["n"][:, None]
    same = bool(torch.equal(out1["n"], out3["n"]) and torch.equal(out1["obj"][valid], out3["obj"][valid]))
    return {"workload": "4 cameras, 4 markers, 2000 frame-sets of 640x480x3 interleaved (7.4 GB resident), " + ("three-kernel path" if os.environ.get("MOCAP_PIPELINE") == "split" else "single-pass kernel"),
            "ms_per_batch": ms, "frame_sets_per_s": B / ms * 1e3, "hbm_gbs": B * C * 921600 / ms / 1e6,
            "equals_one_channel_run": same}


def dlt_rate():
    C, F = 8, 1_000_000
    obs_obj, poses, K, pts = synth.make_tracks(C, 2000, seed=3)
    obs = np.array([[[-1 if v is None else v for v in cam] for cam in fr] for fr in obs_obj], dtype=np.float64)
    mask = np.array([[cam[0] is not None for cam in fr] for fr in obs_obj], dtype=np.uint8)
    ctx = pkg.MocapContext(C)
    ctx.set_cameras([K] * C, poses)
    reps = F // len(obs)
    d_obs = torch.from_numpy(np.tile(obs, (reps, 1, 1))).cuda()
    d_mask = torch.from_numpy(np.tile(mask, (reps, 1))).cuda()
    X = torch.empty((F, 3), dtype=torch.float64, device="cuda")
    err = torch.empty((F,), dtype=torch.float64, device="cuda")
    valid = torch.empty((F,), dtype=torch.uint8, device="cuda")
    import ctypes as Ct
    p = lambda t: Ct.c_void_p(t.data_ptr())
    ctx.use_current_stream()
    fn = lambda: ctx._check(ctx.lib.mocap_triangulate_dev(ctx.h, p(d_obs), p(d_mask), F, p(X), p(err), p(valid)))
    ms = timed(fn)
    return {"workload": "1e6 points x 8 views, DLT + reprojection error", "ms": ms, "dlt_solves_per_s": F / ms * 1e3}


def ba_case(C, F, label, cpu_eval=True):
    obs_obj, poses, K, pts = synth.make_tracks(C, F, seed=9, missing_frac=0.1)
    start = synth.perturb_poses(poses, seed=10)
    obs = np.array([[[-1 if v is None else v for v in cam] for cam in fr] for fr in obs_obj], dtype=np.float64)
    mask = np.array([[cam[0] is not None for cam in fr] for fr in obs_obj], dtype=np.uint8)
    ctx = pkg.MocapContext(C)
    ctx.set_cameras([K] * C, start)
    ctx.bundle_adjust(obs, mask, start)                         # warm-up (allocations)
    t0 = time.perf_counter()
    out, rep = ctx.bundle_adjust(obs, mask, start)
    wall = time.perf_counter() - t0
    res = {"workload": label, "points": F, "cameras": C, "wall_s": wall, **rep}
    # residual evaluations per second: one launch evaluates 1 + 6(C-1) residual vectors
    r0 = time.perf_counter()
    for _ in range(20):
        ctx.ba_residuals(obs, mask, start)
    res["single_residual_vector_ms_incl_copies"] = (time.perf_counter() - r0) / 20 * 1e3
    if cpu_eval:
        port = RefPort([K] * C)
        sub = obs_obj[: min(F, 200)]
        t0 = time.perf_counter()
        port.ba_residuals(port.poses_to_params(start), sub)
        per_point = (time.perf_counter() - t0) / len(sub)
        n_params = 1 + 7 * (C - 1)
        res["cpu_reference_s_per_residual_vector_1core"] = per_point * F
        res["cpu_reference_s_per_jacobian_1core"] = per_point * F * (n_params + 1)
    return res


def preprocess_rate():
    C, B = 4, 2000
    K = np.array([[320.0, 0, 160], [0, 320, 160], [0, 0, 1]])
    dist = [-1.26372388e-01, 2.62661497e-01, 1.21306197e-03, 2.24507008e-04, -2.48534118e-01]
    ctx = pkg.MocapContext(C, 320, 320)
    ctx.set_preprocess(320, 240, [0] * C, [K] * C, [dist] * C)
    raw = torch.randint(0, 256, (B, C, 240, 320, 3), dtype=torch.uint8, device="cuda")
    ms = timed(lambda: ctx.preprocess(raw))
    port = RefPort([K] * C)
    frame = raw[0, 0].cpu().numpy()
    t0 = time.perf_counter()
    for _ in range(20):
        port.preprocess(frame, 0, dist, 0)
    cpu_ms = (time.perf_counter() - t0) / 20 * 1e3
    n_img = B * C
    # raw frames -> tracks in one call (preprocess + S1-S3); noise frames are thresholded to nothing, so
    # use dark frames with a few bright spots per camera
    poses = [{"R": np.eye(3), "t": np.array([-0.4 * c, 0.0, 0.0])} for c in range(C)]
    ctx.set_cameras([K] * C, poses)
    dark = torch.randint(0, 25, (B, C, 240, 320, 3), dtype=torch.uint8, device="cuda")
    yy, xx = torch.meshgrid(torch.arange(240, device="cuda"), torch.arange(320, device="cuda"), indexing="ij")
    gen = np.random.default_rng(3)
    for m in range(4):
        X = np.array([gen.uniform(-0.3, 0.9), gen.uniform(-0.25, 0.25), gen.uniform(2.0, 3.0)])
        for c in range(C):
            pc = X + poses[c]["t"]
            u, v = 320 * pc[0] / pc[2] + 160, 320 * pc[1] / pc[2] + 160 - 40
            spot = (255 * torch.exp(-((yy - v) ** 2 + (xx - u) ** 2) / (2 * 2.0 ** 2))).to(torch.uint8)
            dark[:, c] = torch.maximum(dark[:, c], spot[None, :, :, None])
    raw_ms = timed(lambda: ctx.pipeline_raw(dark))
    tracks = ctx.pipeline_raw(dark)
    return {"workload": "8000 raw 320x240x3 frames -> 320x320x3 (undistort + Gaussian 9x9 + 5x5 filter), one kernel",
            "ms": ms, "frames_per_s": n_img / ms * 1e3, "algorithmic_gbs": n_img * (240 * 320 * 3 + 320 * 320 * 3) / ms / 1e6,
            "cpu_reference_ms_per_frame_1core": cpu_ms,
            "raw_to_tracks_ms": raw_ms, "raw_to_tracks_frame_sets_per_s": B / raw_ms * 1e3,
            "raw_to_tracks_points_per_frame_set": float(tracks["n"].float().mean().item())}


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "config3":      # for ncu: just the 8-camera pipeline
        print(json.dumps(config3_pipeline()))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "colour":
        print(json.dumps(colour_pipeline()))
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "preprocess":
        print(json.dumps(preprocess_rate()))
        sys.exit(0)
    out = {"preprocess": preprocess_rate(), "colour_pipeline": colour_pipeline(),"config3_pipeline": config3_pipeline(), "dlt": dlt_rate(),
           "ba_config3_batch": ba_case(8, 16000, "config 3 per-batch BA: 8 cameras, 1000 frames x 16 markers = 16000 tracked points"),
           "ba_config5": ba_case(16, 6400, "config 5 cold start: 16 cameras, 64 markers x 100 frames = 6400 tracked points")}
    print(json.dumps(out, indent=1))
