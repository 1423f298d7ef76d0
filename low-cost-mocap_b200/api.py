"""Host-side mirror of the reference's hot-path call surface on top of libmocap_b200.so.

Two layers:

* :class:`MocapContext` -- the batched API the benchmark and the multi-GPU driver use:
  device (torch) tensors in, device tensors out, no synchronisation.
* module-level functions with the reference's own names, arguments and return values
  (computer_code/api/helpers.py), bound to a :class:`MocapSession` that plays the part
  of the reference's ``Cameras`` singleton for this path.  ``install_into(helpers)``
  monkey-patches a loaded reference ``helpers`` module so that ``index.py``, the UI and
  the drone loop run unchanged on the CUDA path.

torch is used for device memory and streams only.  There is no CPU fallback: every
function raises :class:`MocapError` when the library or the GPU is missing.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib
from ._lib import MocapError, Config, BAOptions, BAReport, check

THRESHOLD = 51   # cv.threshold(grey, 255*0.2, 255, THRESH_BINARY) on uint8 == pix > 51 (helpers.py:146)

# MOCAP_F_* bits of include/mocap_b200.h
F_SEGMENTS, F_BLOBS, F_ROOTS, F_CANDS, F_GROUPS, F_HOLES = 1, 2, 4, 8, 16, 32
_FLAG_NAMES = {F_SEGMENTS: "max_segments", F_BLOBS: "max_blobs", F_ROOTS: "max_roots", F_CANDS: "max_cands",
               F_GROUPS: "max_groups"}
# the reference is unbounded; the drop-in mirrors run with the compile-time maxima and raise on overflow
MIRROR_LIMITS = dict(max_blobs=64, max_segments=4096, max_roots=128, max_cands=16, max_groups=1 << 16)


def raise_on_overflow(flags, what):
    """The reference keeps every contour / root / candidate group; a result truncated at a configured
    capacity would differ from it silently, so the mirrors turn any MOCAP_F_* bit into an error."""
    flags = int(flags)
    if flags & F_HOLES:
        raise MocapError(-1, f"{what}: a blob with a hole is too large for the RETR_TREE slow path (wider or taller than 62 pixels, "
                             f"or more than 64 holes in the image); cv.findContours would give each hole a contour of its own "
                             f"and fill the outer one (MOCAP_F_HOLES)")
    if flags:
        names = [n for b, n in _FLAG_NAMES.items() if flags & b]
        raise MocapError(-1, f"{what}: capacity overflow ({', '.join(names)}); results would be truncated "
                             f"where the reference is unbounded")


def _torch():
    import torch
    return torch


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _np_ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else C.c_void_p(0)


def _report_dict(rep):
    d = {f: getattr(rep, f) for f, _ in BAReport._fields_}
    d["phase_ms"] = [float(v) for v in rep.phase_ms]
    return d


class MocapContext:
    """One libmocap_b200 context (one CUDA device, one stream, one camera rig)."""

    def __init__(self, n_cam, width=640, height=480, device=0, max_blobs=None, max_segments=None,
                 max_roots=None, max_cands=None, max_groups=None):
        self.lib = _lib.load()
        cfg = Config()
        self.lib.mocap_default_config(C.byref(cfg), n_cam, width, height)
        cfg.device = device
        for k, v in dict(max_blobs=max_blobs, max_segments=max_segments, max_roots=max_roots,
                         max_cands=max_cands, max_groups=max_groups).items():
            if v is not None:
                setattr(cfg, k, v)
        self.cfg = cfg
        h = C.c_void_p()
        check(self.lib.mocap_create(C.byref(h), C.byref(cfg)))
        self.h = h
        self.n_cam, self.width, self.height = n_cam, width, height
        self.device = device
        self._tdev = None
        self._pp_in = None

    # -- lifetime -------------------------------------------------------------------------
    def close(self):
        if getattr(self, "h", None):
            self.lib.mocap_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, st):
        check(st, self.h)

    @property
    def torch_device(self):
        if self._tdev is None:
            self._tdev = _torch().device("cuda", self.device)
        return self._tdev

    def use_current_stream(self):
        """Enqueue on torch's current stream of this device."""
        s = _torch().cuda.current_stream(self.torch_device)
        self._check(self.lib.mocap_set_stream(self.h, C.c_void_p(s.cuda_stream)))

    # -- session state --------------------------------------------------------------------
    def set_cameras(self, intrinsics, poses):
        """intrinsics: C 3x3 matrices; poses: list of {"R": 3x3, "t": 3} as the reference passes them."""
        n = self.n_cam
        K = np.ascontiguousarray(np.stack([np.asarray(k, dtype=np.float64).reshape(3, 3) for k in intrinsics]))
        R = np.ascontiguousarray(np.stack([np.asarray(p["R"], dtype=np.float64).reshape(3, 3) for p in poses]))
        t = np.ascontiguousarray(np.stack([np.asarray(p["t"], dtype=np.float64).reshape(3) for p in poses]))
        if K.shape[0] != n or R.shape[0] != n:
            raise ValueError(f"context was created for {n} cameras")
        self._check(self.lib.mocap_set_cameras(self.h, _np_ptr(K), _np_ptr(R), _np_ptr(t)))

    def set_world_transform(self, M):
        if M is None:
            self._check(self.lib.mocap_set_world_transform(self.h, C.c_void_p(0)))
        else:
            M = np.ascontiguousarray(np.asarray(M, dtype=np.float64).reshape(4, 4))
            self._check(self.lib.mocap_set_world_transform(self.h, _np_ptr(M)))

    # -- capture-side preprocessing (helpers.py:70-82) -------------------------------------
    def set_preprocess(self, in_width, in_height, rotations, intrinsics, distortions):
        rot = np.ascontiguousarray(np.asarray(rotations, dtype=np.int32).reshape(self.n_cam))
        K = np.ascontiguousarray(np.stack([np.asarray(k, dtype=np.float64).reshape(3, 3) for k in intrinsics]))
        D = np.ascontiguousarray(np.stack([np.asarray(d, dtype=np.float64).reshape(5) for d in distortions]))
        self._check(self.lib.mocap_set_preprocess(self.h, int(in_width), int(in_height), _np_ptr(rot), _np_ptr(K), _np_ptr(D)))
        self._pp_in = (int(in_height), int(in_width))

    def preprocess(self, raw):
        """raw uint8 cuda tensor [..., in_h, in_w, 3] (whole frame-sets) -> uint8 [N, S, S, 3]."""
        torch = _torch()
        if self._pp_in is None:
            raise MocapError(-5, "mocap_set_preprocess has not been called (MocapContext.set_preprocess)")
        h, w = self._pp_in
        n = raw.numel() // (h * w * 3)
        out = torch.empty((n, self.height, self.width, 3), dtype=torch.uint8, device=raw.device)
        self.use_current_stream()
        self._check(self.lib.mocap_preprocess_dev(self.h, _ptr(raw.contiguous()), n, _ptr(out)))
        return out

    def pipeline_raw(self, raw, threshold=THRESHOLD, want_frames=False):
        """Raw camera frames uint8 cuda [B, C, in_h, in_w, 3] -> tracks (and the processed frames)."""
        torch = _torch()
        if self._pp_in is None:
            raise MocapError(-5, "mocap_set_preprocess has not been called (MocapContext.set_preprocess)")
        h, w = self._pp_in
        B = raw.numel() // (self.n_cam * h * w * 3)
        out = self.alloc_tracks(B, raw.device)
        frames = torch.empty((B, self.n_cam, self.height, self.width, 3), dtype=torch.uint8, device=raw.device) if want_frames else None
        self.use_current_stream()
        self._check(self.lib.mocap_pipeline_raw_dev(self.h, _ptr(raw.contiguous()), B, int(threshold), _ptr(frames),
                                                    _ptr(out["obj"]), _ptr(out["err"]), _ptr(out["n"]), _ptr(out["flags"])))
        if want_frames:
            out["frames"] = frames
        return out

    def undistort_map(self, cam):
        m1 = np.empty((self.height, self.width, 2), dtype=np.int16)
        m2 = np.empty((self.height, self.width), dtype=np.uint16)
        self._check(self.lib.mocap_get_undistort_map(self.h, int(cam), _np_ptr(m1), _np_ptr(m2)))
        return m1, m2

    # -- batched device API ---------------------------------------------------------------
    def detect(self, frames, threshold=THRESHOLD, want_moments=False):
        """frames: uint8 cuda tensor [..., H, W] or [..., H, W, 3].  Returns dict of cuda tensors:
        xy int32 [N, max_blobs, 2], n int32 [N], flags int32 [N], (mom int64 [N, max_blobs, 4])."""
        torch = _torch()
        ch = 3 if (frames.dim() >= 3 and frames.shape[-1] == 3 and frames.shape[-2] == self.width) else 1
        per = self.width * self.height * ch
        if not frames.is_contiguous() or frames.dtype != torch.uint8 or frames.numel() % per:
            raise ValueError("frames must be a contiguous uint8 tensor of whole images")
        n = frames.numel() // per
        dev = frames.device
        xy = torch.empty((n, self.cfg.max_blobs, 2), dtype=torch.int32, device=dev)
        cnt = torch.empty((n,), dtype=torch.int32, device=dev)
        flags = torch.empty((n,), dtype=torch.int32, device=dev)
        mom = torch.empty((n, self.cfg.max_blobs, 4), dtype=torch.int64, device=dev) if want_moments else None
        self.use_current_stream()
        self._check(self.lib.mocap_detect_dev(self.h, _ptr(frames), n, ch, int(threshold), _ptr(xy), _ptr(cnt), _ptr(mom), _ptr(flags)))
        out = {"xy": xy, "n": cnt, "flags": flags}
        if want_moments:
            out["mom"] = mom
        return out

    def match_triangulate(self, xy, n, want_chosen=False):
        """xy int32 [B*C, max_blobs, 2], n int32 [B*C] (output of detect).  Returns dict: obj f64
        [B, max_roots, 3], err f64 [B, max_roots], n int32 [B], flags int32 [B], (chosen int32 [B, max_roots, C])."""
        torch = _torch()
        B = n.numel() // self.n_cam
        dev = xy.device
        R = self.cfg.max_roots
        obj = torch.empty((B, R, 3), dtype=torch.float64, device=dev)
        err = torch.empty((B, R), dtype=torch.float64, device=dev)
        cnt = torch.empty((B,), dtype=torch.int32, device=dev)
        flags = torch.empty((B,), dtype=torch.int32, device=dev)
        chosen = torch.empty((B, R, self.n_cam), dtype=torch.int32, device=dev) if want_chosen else None
        self.use_current_stream()
        self._check(self.lib.mocap_match_triangulate_dev(self.h, _ptr(xy), _ptr(n), B, _ptr(obj), _ptr(err), _ptr(cnt), _ptr(flags), _ptr(chosen)))
        out = {"obj": obj, "err": err, "n": cnt, "flags": flags}
        if want_chosen:
            out["chosen"] = chosen
        return out

    def tracks_to_observations(self, xy, n_obj, chosen):
        """Matcher output -> explicit correspondences for S4 (BASELINE config 3: bundle adjustment on the
        tracks of a batch).  xy int32 [B*C, max_blobs, 2] (detect), n_obj int32 [B], chosen int32
        [B, max_roots, C] (match_triangulate(want_chosen=True)).  Returns host arrays obs f64 [P, C, 2],
        mask uint8 [P, C] with one row per triangulated point, in frame order."""
        torch = _torch()
        B, R, Cn = chosen.shape
        MB = xy.shape[1]
        keep = torch.arange(R, device=chosen.device)[None, :] < n_obj[:, None].to(torch.int64)      # [B, R]
        ch = chosen[keep].to(torch.int64)                                                          # [P, C]
        set_idx = torch.arange(B, device=chosen.device)[:, None].expand(B, R)[keep]                # [P]
        img = set_idx[:, None] * Cn + torch.arange(Cn, device=chosen.device)[None, :]              # [P, C]
        mask = ch >= 0
        pts = xy.view(-1, MB, 2)[img, ch.clamp(min=0)]                                             # [P, C, 2]
        obs = torch.where(mask[:, :, None], pts.to(torch.float64), torch.zeros((), dtype=torch.float64, device=pts.device))
        return obs.cpu().numpy(), mask.to(torch.uint8).cpu().numpy()

    def alloc_tracks(self, n_sets, device=None):
        torch = _torch()
        dev = device or self.torch_device
        R = self.cfg.max_roots
        return {"obj": torch.empty((n_sets, R, 3), dtype=torch.float64, device=dev),
                "err": torch.empty((n_sets, R), dtype=torch.float64, device=dev),
                "n": torch.empty((n_sets,), dtype=torch.int32, device=dev),
                "flags": torch.empty((n_sets,), dtype=torch.int32, device=dev)}

    def pipeline(self, frames, threshold=THRESHOLD, out=None, want_tracks=False):
        """S1+S2+S3 on a device tensor of frame-sets [B, C, H, W] (or [B, C, H, W, 3]).  want_tracks (or an ``out``
        that holds "track_xy"): also the winners' pixels per camera, int32 [B, max_roots, C, 2], (-1, -1) = no view."""
        torch = _torch()
        ch = 3 if (frames.shape[-1] == 3 and frames.shape[-2] == self.width) else 1
        B = frames.numel() // (self.n_cam * self.width * self.height * ch)
        if out is None:
            out = self.alloc_tracks(B, frames.device)
        if want_tracks and "track_xy" not in out:
            out["track_xy"] = torch.empty((B, self.cfg.max_roots, self.n_cam, 2), dtype=torch.int32, device=frames.device)
        self.use_current_stream()
        self._check(self.lib.mocap_pipeline_tracks_dev(self.h, _ptr(frames), B, ch, int(threshold), _ptr(out["obj"]), _ptr(out["err"]),
                                                       _ptr(out["n"]), _ptr(out["flags"]), _ptr(out.get("track_xy"))))
        return out

    def tracks_to_observations_dev(self, tracks, max_err=0.0, capacity=None, out=None):
        """Matcher output of a batch (``pipeline(..., want_tracks=True)``) -> the explicit correspondences S4 consumes,
        on the device, no synchronisation: dict obs f64 [capacity, C, 2], mask uint8 [capacity, C], n int32 [1]."""
        torch = _torch()
        B = tracks["n"].numel()
        dev = tracks["n"].device
        cap = int(capacity or B * self.cfg.max_roots)
        if out is None:
            out = {"obs": torch.empty((cap, self.n_cam, 2), dtype=torch.float64, device=dev),
                   "mask": torch.empty((cap, self.n_cam), dtype=torch.uint8, device=dev),
                   "n": torch.zeros((1,), dtype=torch.int32, device=dev)}
        self.use_current_stream()
        self._check(self.lib.mocap_tracks_to_observations_dev(self.h, _ptr(tracks["track_xy"]), _ptr(tracks["n"]), _ptr(tracks["err"]), B,
                                                              float(max_err), _ptr(out["obs"]), _ptr(out["mask"]), _ptr(out["n"]), cap))
        return out

    def set_ba_grid(self, n_ctas=0):
        """CTAs of the device-resident bundle adjustment of this context (0: one per SM).  Independent solves finish sooner
        side by side: K contexts on K streams with SMs // K CTAs each (mocap_set_ba_grid, include/mocap_b200.h)."""
        self._check(self.lib.mocap_set_ba_grid(self.h, int(n_ctas)))

    def bundle_adjust_dev(self, obs, mask, R, t, n_points=None, report=None, ftol=1e-2, max_nfev=0, prefit=True, jacobian=1,
                          prefit_max_iter=50):
        """S4 wholly on the device (mocap_bundle_adjust_dev): obs f64 [P, C, 2], mask uint8 [P, C], R f64 [C, 3, 3] and
        t f64 [C, 3] cuda tensors (R, t updated in place), n_points an int32 cuda tensor [1] or None.  One cooperative
        launch, no synchronisation.  Returns the report as a uint8 cuda tensor (decode with ``decode_ba_report``)."""
        torch = _torch()
        opt = BAOptions()
        self.lib.mocap_ba_default_options(C.byref(opt))
        opt.ftol, opt.max_nfev, opt.prefit, opt.jacobian, opt.prefit_max_iter = ftol, max_nfev, 1 if prefit else 0, jacobian, prefit_max_iter
        if report is None:
            report = torch.zeros((C.sizeof(BAReport),), dtype=torch.uint8, device=obs.device)
        self.use_current_stream()
        self._check(self.lib.mocap_bundle_adjust_dev(self.h, _ptr(obs), _ptr(mask), obs.shape[0], _ptr(n_points), _ptr(R), _ptr(t),
                                                     C.byref(opt), _ptr(report)))
        return report

    @staticmethod
    def decode_ba_report(report):
        rep = BAReport.from_buffer_copy(report.cpu().numpy().tobytes())
        return _report_dict(rep)

    def pipeline_host(self, frames, threshold=THRESHOLD, out=None):
        """Same through HOST memory: frames is a (preferably pinned) uint8 cpu tensor / ndarray;
        returns cpu tensors.  Blocks until the results are in host memory."""
        torch = _torch()
        if isinstance(frames, np.ndarray):
            frames = torch.from_numpy(frames)
        ch = 3 if (frames.shape[-1] == 3 and frames.shape[-2] == self.width) else 1
        B = frames.numel() // (self.n_cam * self.width * self.height * ch)
        R = self.cfg.max_roots
        if out is None:
            out = {"obj": torch.empty((B, R, 3), dtype=torch.float64).pin_memory(),
                   "err": torch.empty((B, R), dtype=torch.float64).pin_memory(),
                   "n": torch.empty((B,), dtype=torch.int32).pin_memory(),
                   "flags": torch.empty((B,), dtype=torch.int32).pin_memory()}
        self.use_current_stream()
        self._check(self.lib.mocap_pipeline_host(self.h, _ptr(frames), B, ch, int(threshold),
                                                 _ptr(out["obj"]), _ptr(out["err"]), _ptr(out["n"]), _ptr(out["flags"])))
        return out

    def locate_objects(self, obj, err, n, max_objects=8):
        """obj f64 [B, max_roots, 3], err f64 [B, max_roots], n int32 [B] (matcher output, cuda tensors) ->
        dict: objects f64 [B, max_objects, 5] = x y z heading error, drone_index int32 [B, max_objects], n int32 [B]."""
        torch = _torch()
        B = n.numel()
        dev = obj.device
        out = torch.empty((B, max_objects, 5), dtype=torch.float64, device=dev)
        di = torch.empty((B, max_objects), dtype=torch.int32, device=dev)
        cnt = torch.empty((B,), dtype=torch.int32, device=dev)
        self.use_current_stream()
        self._check(self.lib.mocap_locate_objects_dev(self.h, _ptr(obj), _ptr(err), _ptr(n), B, max_objects, _ptr(out), _ptr(di), _ptr(cnt)))
        return {"objects": out, "drone_index": di, "n": cnt}

    # -- explicit correspondences (host arrays) -------------------------------------------
    def triangulate(self, obs, mask, want_err=True):
        """obs float64 [F, C, 2], mask uint8 [F, C] -> (X [F,3], err [F] or None, valid [F])."""
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        F = obs.shape[0]
        X = np.empty((F, 3))
        err = np.empty((F,)) if want_err else None
        valid = np.empty((F,), dtype=np.uint8)
        self._check(self.lib.mocap_triangulate_host(self.h, _np_ptr(obs), _np_ptr(mask), F, _np_ptr(X), _np_ptr(err), _np_ptr(valid)))
        return X, err, valid

    def reprojection_errors(self, obs, mask, X):
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        X = np.ascontiguousarray(X, dtype=np.float64)
        F = obs.shape[0]
        err = np.empty((F,))
        valid = np.empty((F,), dtype=np.uint8)
        self._check(self.lib.mocap_reprojection_errors_host(self.h, _np_ptr(obs), _np_ptr(mask), _np_ptr(X), F, _np_ptr(err), _np_ptr(valid)))
        return err, valid

    def calibrate_init(self, obs, mask, F_given=None):
        """Chain of relative poses from 2D tracks (index.py:229-270).  Returns (poses, F_used [C-1,3,3], votes [C-1,4])."""
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        Cn = self.n_cam
        R = np.empty((Cn, 3, 3)); t = np.empty((Cn, 3))
        Fu = np.empty((Cn - 1, 3, 3)); votes = np.empty((Cn - 1, 4), dtype=np.int32)
        Fg = None if F_given is None else np.ascontiguousarray(np.asarray(F_given, dtype=np.float64).reshape(Cn - 1, 3, 3))
        self._check(self.lib.mocap_calibrate_init_host(self.h, _np_ptr(obs), _np_ptr(mask), obs.shape[0], _np_ptr(Fg),
                                                       _np_ptr(R), _np_ptr(t), _np_ptr(Fu), _np_ptr(votes)))
        return [{"R": R[i].copy(), "t": t[i].copy()} for i in range(Cn)], Fu, votes

    def ba_residuals(self, obs, mask, poses):
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        R = np.ascontiguousarray(np.stack([np.asarray(p["R"], dtype=np.float64).reshape(3, 3) for p in poses]))
        t = np.ascontiguousarray(np.stack([np.asarray(p["t"], dtype=np.float64).reshape(3) for p in poses]))
        F = obs.shape[0]
        r = np.empty((F,), dtype=np.float32)
        valid = np.empty((F,), dtype=np.uint8)
        nv = C.c_int(0)
        self._check(self.lib.mocap_ba_residuals_host(self.h, _np_ptr(obs), _np_ptr(mask), F, _np_ptr(R), _np_ptr(t), _np_ptr(r), _np_ptr(valid), C.byref(nv)))
        return r[valid.astype(bool)]

    def bundle_adjust(self, obs, mask, poses, ftol=1e-2, max_nfev=0, engine=0, prefit=True, jacobian=1):
        obs = np.ascontiguousarray(obs, dtype=np.float64)
        mask = np.ascontiguousarray(mask, dtype=np.uint8)
        R = np.ascontiguousarray(np.stack([np.asarray(p["R"], dtype=np.float64).reshape(3, 3) for p in poses]))
        t = np.ascontiguousarray(np.stack([np.asarray(p["t"], dtype=np.float64).reshape(3) for p in poses]))
        opt = BAOptions()
        self.lib.mocap_ba_default_options(C.byref(opt))
        opt.ftol = ftol
        opt.max_nfev = max_nfev
        opt.engine, opt.prefit, opt.jacobian = engine, 1 if prefit else 0, jacobian
        rep = BAReport()
        self._check(self.lib.mocap_bundle_adjust_host(self.h, _np_ptr(obs), _np_ptr(mask), obs.shape[0], _np_ptr(R), _np_ptr(t), C.byref(opt), C.byref(rep)))
        out = [{"R": R[i].copy(), "t": t[i].copy()} for i in range(R.shape[0])]
        return out, _report_dict(rep)

    # -- accounting -----------------------------------------------------------------------
    def launch_count(self):
        return int(self.lib.mocap_launch_count(self.h))

    def enable_kernel_timing(self, on=True):
        self._check(self.lib.mocap_enable_kernel_timing(self.h, 1 if on else 0))

    def detect_kernel_ms(self, reset=True):
        ms, n = C.c_double(0), C.c_int(0)
        self._check(self.lib.mocap_detect_kernel_ms(self.h, 1 if reset else 0, C.byref(ms), C.byref(n)))
        return ms.value, n.value


# =================================================================================================
# Reference-signature mirror
# =================================================================================================
def _split_observations(image_points):
    """(F, C, 2) list/object array with None for missing views -> float64 obs + uint8 mask."""
    arr = np.array(image_points, dtype=object)
    if arr.ndim != 3:
        raise ValueError("image_points must be (F, C, 2)")
    F, Cn, _ = arr.shape
    mask = np.empty((F, Cn), dtype=np.uint8)
    obs = np.zeros((F, Cn, 2), dtype=np.float64)
    for f in range(F):
        for c in range(Cn):
            a, b = arr[f, c]
            if a is None and b is None:
                mask[f, c] = 0
            else:
                mask[f, c] = 1
                obs[f, c, 0], obs[f, c, 1] = a, b
    return obs, mask


class MocapSession:
    """What the reference keeps in its ``Cameras`` singleton for this path: the intrinsics
    (helpers.py:19-22) -- plus lazily created CUDA contexts per camera count.  Thread safe
    (one lock per session; the reference's singleton is unsynchronised, Singleton.py:3)."""

    _default = None

    def __init__(self, intrinsics, width=640, height=480, device=0):
        self.intrinsics = [np.asarray(k, dtype=np.float64).reshape(3, 3) for k in intrinsics]
        self.width, self.height, self.device = width, height, device
        self._ctxs = {}
        self._lock = threading.RLock()

    @classmethod
    def default(cls):
        if cls._default is None:
            raise MocapError(-5, "no MocapSession installed: call MocapSession.install(intrinsics) or install_into(helpers)")
        return cls._default

    @classmethod
    def install(cls, intrinsics, **kw):
        cls._default = cls(intrinsics, **kw)
        return cls._default

    def ctx(self, n_cam, **kw):
        with self._lock:
            key = (n_cam, tuple(sorted(kw.items())))
            if key not in self._ctxs:
                self._ctxs[key] = MocapContext(n_cam, self.width, self.height, self.device, **kw)
            return self._ctxs[key]

    def ctx_with_poses(self, camera_poses):
        n = len(camera_poses)
        if n > len(self.intrinsics):
            raise ValueError("more camera poses than intrinsics")
        c = self.ctx(n, **MIRROR_LIMITS)
        c.set_cameras(self.intrinsics[:n], camera_poses)
        return c


def find_dot(img, session=None):
    """Mirror of ``Cameras._find_dot(self, img)`` (helpers.py:143-163): returns
    ``(img, image_points)`` with image_points a list of ``[x, y]`` or ``[[None, None]]``.
    The contour / text overlay the reference draws into ``img`` (helpers.py:148,156-157) is
    cosmetic and is not reproduced; a 1-px centre dot is drawn instead."""
    torch = _torch()
    s = session or MocapSession.default()
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w = img.shape[:2]
    with s._lock:
        ctx = s._ctxs.get(("detect", w, h))
        if ctx is None:
            ctx = MocapContext(1, w, h, s.device, max_blobs=MIRROR_LIMITS["max_blobs"],
                               max_segments=MIRROR_LIMITS["max_segments"])
            s._ctxs[("detect", w, h)] = ctx
        d = ctx.detect(torch.from_numpy(img).to(ctx.torch_device))
        raise_on_overflow(d["flags"][0].item(), "find_dot")
        n = int(d["n"][0].item())
        pts = d["xy"][0, :n].cpu().numpy().tolist()
    for x, y in pts:
        if 0 <= y < h and 0 <= x < w:
            img[y, x] = (100, 255, 100) if img.ndim == 3 else 255
    return img, (pts if pts else [[None, None]])


def triangulate_points(image_points, camera_poses, session=None):
    """Mirror of helpers.py:330-336: ndarray (F, 3); rows of ``[None]*3`` (object dtype) where a
    point has fewer than two views."""
    s = session or MocapSession.default()
    if len(image_points) == 0:
        return np.array([])
    obs, mask = _split_observations(image_points)
    with s._lock:
        X, _, valid = s.ctx_with_poses(camera_poses).triangulate(obs, mask, want_err=False)
    if valid.all():
        return X
    out = np.empty((len(X), 3), dtype=object)
    for f in range(len(X)):
        out[f] = list(X[f]) if valid[f] else [None, None, None]
    return out


def triangulate_point(image_points, camera_poses, session=None):
    """Mirror of helpers.py:293-327 (one point)."""
    r = triangulate_points([image_points], camera_poses, session)[0]
    return list(r) if r[0] is None else np.asarray(r, dtype=np.float64)


def calculate_reprojection_errors(image_points, object_points, camera_poses, session=None):
    """Mirror of helpers.py:203-211: float64 vector; points with <= 1 view are skipped."""
    s = session or MocapSession.default()
    if len(image_points) == 0:
        return np.array([])
    obs, mask = _split_observations(image_points)
    X = np.array([[np.nan] * 3 if p[0] is None else [float(v) for v in p] for p in object_points], dtype=np.float64)
    with s._lock:
        err, valid = s.ctx_with_poses(camera_poses).reprojection_errors(obs, mask, X)
    return err[valid.astype(bool)]


def calculate_reprojection_error(image_points, object_point, camera_poses, session=None):
    """Mirror of helpers.py:214-241 (one point): float or None."""
    e = calculate_reprojection_errors([image_points], [object_point], camera_poses, session)
    return float(e[0]) if len(e) else None


def find_point_correspondance_and_object_points(image_points, camera_poses, frames, session=None):
    """Mirror of helpers.py:339-421: returns ``(errors (K,), object_points (K,3), frames)``.
    Like the reference it removes the ``[None, None]`` sentinels from ``image_points`` in
    place (helpers.py:342-346); the epipolar lines the reference draws into ``frames``
    (helpers.py:365) are cosmetic and are not drawn."""
    torch = _torch()
    s = session or MocapSession.default()
    for pts in image_points:
        try:
            pts.remove([None, None])
        except ValueError:
            pass
    n_cam = len(camera_poses)
    with s._lock:
        ctx = s.ctx_with_poses(camera_poses)
        MB = ctx.cfg.max_blobs
        xy = np.zeros((n_cam, MB, 2), dtype=np.int32)
        n = np.zeros((n_cam,), dtype=np.int32)
        for c in range(n_cam):
            k = len(image_points[c])
            if k > MB:
                raise MocapError(-1, f"camera {c} has {k} points; context keeps {MB}")
            n[c] = k
            if k:
                xy[c, :k] = np.asarray(image_points[c], dtype=np.int32)
        d = ctx.match_triangulate(torch.from_numpy(xy).to(ctx.torch_device), torch.from_numpy(n).to(ctx.torch_device))
        raise_on_overflow(d["flags"][0].item(), "find_point_correspondance_and_object_points")
        k = int(d["n"][0].item())
        errors = d["err"][0, :k].cpu().numpy()
        object_points = d["obj"][0, :k].cpu().numpy()
    return errors, object_points, frames


def locate_objects(object_points, errors, session=None):
    """Mirror of helpers.py:424-480: list of {"pos": ndarray(3), "heading", "error", "droneIndex"}."""
    torch = _torch()
    s = session or MocapSession.default()
    pts = np.asarray(object_points, dtype=np.float64).reshape(-1, 3)
    errs = np.asarray(errors, dtype=np.float64).reshape(-1)
    K = pts.shape[0]
    if K == 0:
        return []
    with s._lock:
        ctx = s.ctx(len(s.intrinsics), **MIRROR_LIMITS)
        R = ctx.cfg.max_roots
        if K > R:
            raise MocapError(-1, f"{K} points; context keeps {R}")
        obj = np.zeros((1, R, 3)); err = np.zeros((1, R))
        obj[0, :K] = pts; err[0, :K] = errs
        d = ctx.locate_objects(torch.from_numpy(obj).to(ctx.torch_device), torch.from_numpy(err).to(ctx.torch_device),
                               torch.tensor([K], dtype=torch.int32, device=ctx.torch_device), max_objects=K)      # only i is screened: up to one object per point (helpers.py:433-436)
        k = int(d["n"][0].item())
        rec = d["objects"][0, :k].cpu().numpy()
        di = d["drone_index"][0, :k].cpu().numpy()
    return [{"pos": rec[i, :3].copy(), "heading": float(rec[i, 3]), "error": float(rec[i, 4]), "droneIndex": int(di[i])}
            for i in range(k)]


def bundle_adjustment(image_points, camera_poses, socketio, session=None):
    """Mirror of helpers.py:244-290: returns the list of ``{"R": ndarray 3x3, "t": ndarray (3,)}``.
    ``socketio.emit("camera-pose", ...)`` fires once with the result (the reference emits on
    every residual evaluation, helpers.py:274; the UI only renders the latest, App.tsx:254-263)."""
    s = session or MocapSession.default()
    obs, mask = _split_observations(image_points)
    with s._lock:
        out, _ = s.ctx_with_poses(camera_poses).bundle_adjust(obs, mask, camera_poses)
    if socketio is not None:
        socketio.emit("camera-pose", {"camera_poses": [{"R": p["R"].tolist(), "t": p["t"].tolist()} for p in out]})
    return out


def calculate_camera_poses(image_points, socketio=None, session=None):
    """The computation of the reference's ``calculate-camera-pose`` handler (index.py:229-277): cold-start
    chain of relative poses, then bundle adjustment.  ``image_points`` is the (F, C, 2) list the UI sends
    (``data["cameraPoints"]``) with ``None`` for missing views.  Returns the list of {"R", "t"}."""
    s = session or MocapSession.default()
    obs, mask = _split_observations(image_points)
    n_cam = obs.shape[1]
    with s._lock:
        ctx = s.ctx(n_cam)
        ident = [{"R": np.eye(3), "t": np.zeros(3)} for _ in range(n_cam)]
        ctx.set_cameras(s.intrinsics[:n_cam], ident)
        start, _, _ = ctx.calibrate_init(obs, mask)
        ctx.set_cameras(s.intrinsics[:n_cam], start)
        out, _ = ctx.bundle_adjust(obs, mask, start)
    if socketio is not None:
        socketio.emit("camera-pose", {"camera_poses": [{"R": p["R"].tolist(), "t": p["t"].tolist()} for p in out]})
    return out


PATCHED_NAMES = ("triangulate_point", "triangulate_points", "calculate_reprojection_error",
                 "calculate_reprojection_errors", "find_point_correspondance_and_object_points",
                 "bundle_adjustment", "locate_objects")


def install_into(helpers_module, *also, session=None):
    """Point a loaded reference ``helpers`` module at the CUDA path.

    ``also``: modules that imported the hot-path names BY VALUE -- the reference's ``index.py`` does
    (``from helpers import ... bundle_adjustment, triangulate_points, calculate_reprojection_errors``,
    index.py:1), so ``calculate_camera_pose`` (index.py:254,272,274,275) would keep the CPU functions.
    Every name of PATCHED_NAMES such a module holds is re-bound as well:
    ``install_into(helpers, sys.modules[__name__])`` from inside index.py."""
    cams = helpers_module.Cameras.instance()
    s = session or MocapSession.install([np.asarray(p["intrinsic_matrix"], dtype=np.float64) for p in cams.camera_params])
    # helpers.Cameras is a Singleton WRAPPER object (Singleton.py:17-37); _camera_read looks _find_dot up on the
    # decorated class of the instance, so that is where the replacement goes
    type(cams)._find_dot = lambda self, img: find_dot(img, s)
    repl = {
        "triangulate_point": lambda ip, cp: triangulate_point(ip, cp, s),
        "triangulate_points": lambda ip, cp: triangulate_points(ip, cp, s),
        "calculate_reprojection_error": lambda ip, op, cp: calculate_reprojection_error(ip, op, cp, s),
        "calculate_reprojection_errors": lambda ip, op, cp: calculate_reprojection_errors(ip, op, cp, s),
        "find_point_correspondance_and_object_points":
            lambda ip, cp, fr: find_point_correspondance_and_object_points(ip, cp, fr, s),
        "bundle_adjustment": lambda ip, cp, sio: bundle_adjustment(ip, cp, sio, s),
        "locate_objects": lambda op, er: locate_objects(op, er, s),
    }
    for name, fn in repl.items():
        fn.__name__ = name
        fn.__mocap_b200__ = True
        setattr(helpers_module, name, fn)
        for mod in also:
            if hasattr(mod, name):
                setattr(mod, name, fn)
    return s
