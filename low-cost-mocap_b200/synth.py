"""Seeded synthetic N-camera marker streams.

The reference ships no recorded frames, so every parity check and
every benchmark line runs on streams produced here: a ring (C < 8) or a 200° arc
(C >= 8) of pin-hole cameras of radius 3 looking at a 1 m cube in which M
markers random-walk; each marker is rendered as a solid Gaussian spot (peak 255,
sigma 1.2..2.5 px) on a 640x480 uint8 image over sub-threshold uniform noise.
Poses are expressed relative to camera 0, which the reference pins at (I, 0)
(helpers.py:250-253, index.py:235-238).
"""
from __future__ import annotations

import numpy as np

WIDTH = 640
HEIGHT = 480
K_DEFAULT = np.array([[600.0, 0.0, 320.0], [0.0, 600.0, 240.0], [0.0, 0.0, 1.0]])
THRESHOLD = 51  # cv.threshold(grey, 255*0.2, ...) on 8-bit data == (pix > 51), helpers.py:146


def _rot_y(a: float) -> np.ndarray:
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0.0, s], [0.0, 1.0, 0.0], [-s, 0.0, c]])


def make_rig(num_cameras: int, radius: float = 3.0):
    """Return (poses, K) with poses = [{"R": 3x3, "t": (3,)}] relative to camera 0."""
    C = num_cameras
    if C < 8:
        angles = [2.0 * np.pi * i / C for i in range(C)]
    else:  # 200 degree arc: avoids exactly opposed cameras (degenerate epipoles)
        span = np.deg2rad(200.0)
        angles = [span * i / (C - 1) for i in range(C)]
    world = []
    for i, a in enumerate(angles):
        centre = np.array([radius * np.sin(a), 0.3 * (i % 2), -radius * np.cos(a)])
        R = _rot_y(a)
        world.append((R, -R @ centre))
    R0, t0 = world[0]
    poses = []
    for R, t in world:
        Rr = R @ R0.T
        poses.append({"R": Rr, "t": t - Rr @ t0})
    poses[0] = {"R": np.eye(3), "t": np.zeros(3)}
    return poses, K_DEFAULT.copy()


def project(points_cam0: np.ndarray, pose, K: np.ndarray) -> np.ndarray:
    """Sub-pixel pin-hole projection of (M,3) camera-0-frame points -> (M,2)."""
    pc = points_cam0 @ np.asarray(pose["R"]).T + np.asarray(pose["t"]).reshape(1, 3)
    uv = pc @ K.T
    return uv[:, :2] / uv[:, 2:3]


class MarkerStream:
    """Random-walking markers + their rendered camera images.

    Markers live in the camera-0 frame: cube [-0.5,0.5]^3 centred 3 units in
    front of camera 0.  ``next_frame_set`` returns the true 3D points, the
    sub-pixel projections and (optionally) the C rendered uint8 images.
    """

    def __init__(self, num_cameras: int, num_markers: int, seed: int = 0,
                 noise_max: int = 40, min_sep_px: float = 12.0, step_sigma: float = 0.005):
        self.C, self.M = num_cameras, num_markers
        self.rng = np.random.default_rng(seed)
        self.poses, self.K = make_rig(num_cameras)
        self.noise_max = noise_max
        self.min_sep = min_sep_px
        self.step_sigma = step_sigma
        self.centre = np.array([0.0, 0.0, 3.0])
        self.sigmas = self.rng.uniform(1.2, 2.5, size=(num_cameras, num_markers))
        self.pos = None

    def _separated(self, pos) -> bool:
        for pose in self.poses:
            uv = project(pos + self.centre, pose, self.K)
            if self.M > 1:
                d = uv[:, None, :] - uv[None, :, :]
                dist = np.sqrt((d ** 2).sum(-1)) + np.eye(self.M) * 1e9
                if dist.min() < self.min_sep:
                    return False
            if (uv[:, 0] < 16).any() or (uv[:, 0] > WIDTH - 16).any() \
                    or (uv[:, 1] < 16).any() or (uv[:, 1] > HEIGHT - 16).any():
                return False
        return True

    def _advance(self):
        for _ in range(10000):
            if self.pos is None:
                cand = self.rng.uniform(-0.5, 0.5, size=(self.M, 3))
            else:
                cand = self.pos + self.rng.normal(0.0, self.step_sigma, size=(self.M, 3))
                cand = np.where(cand > 0.5, 1.0 - cand, cand)    # reflect at the cube
                cand = np.where(cand < -0.5, -1.0 - cand, cand)
            if self._separated(cand):
                self.pos = cand
                return
        raise RuntimeError("could not place separated markers")

    def render(self, uv: np.ndarray, cam: int) -> np.ndarray:
        if self.noise_max > 0:
            img = self.rng.integers(0, self.noise_max + 1, size=(HEIGHT, WIDTH), dtype=np.uint8)
        else:
            img = np.zeros((HEIGHT, WIDTH), dtype=np.uint8)
        r = 9
        for m in range(self.M):
            u, v = uv[m]
            s = self.sigmas[cam, m]
            x0, y0 = int(np.floor(u)) - r, int(np.floor(v)) - r
            xs = np.arange(max(x0, 0), min(x0 + 2 * r + 2, WIDTH))
            ys = np.arange(max(y0, 0), min(y0 + 2 * r + 2, HEIGHT))
            if len(xs) == 0 or len(ys) == 0:
                continue
            d2 = (xs[None, :] - u) ** 2 + (ys[:, None] - v) ** 2
            spot = np.floor(255.0 * np.exp(-d2 / (2.0 * s * s))).astype(np.uint8)
            patch = img[ys[0]:ys[-1] + 1, xs[0]:xs[-1] + 1]
            np.maximum(patch, spot, out=patch)
        return img

    def next_frame_set(self, render: bool = True):
        self._advance()
        pts = self.pos + self.centre
        uvs = [project(pts, pose, self.K) for pose in self.poses]
        imgs = None
        if render:
            imgs = np.stack([self.render(uvs[c], c) for c in range(self.C)])
        return pts.copy(), np.stack(uvs), imgs


def clutter(shape, max_value: int = 40, salt: int = 0) -> np.ndarray:
    """Deterministic sub-threshold clutter in [0, max_value] from an integer hash of
    the pixel index (no RNG, so committed golden frames can be stored clean and
    compressible and the clutter re-applied bit-identically anywhere)."""
    n = int(np.prod(shape))
    i = np.arange(n, dtype=np.uint64) + np.uint64(salt) * np.uint64(0x9E3779B97F4A7C15 & 0xFFFFFFFF)
    h = (i * np.uint64(2654435761)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(15)
    h = (h * np.uint64(2246822519)) & np.uint64(0xFFFFFFFF)
    h ^= h >> np.uint64(13)
    return (h % np.uint64(max_value + 1)).astype(np.uint8).reshape(shape)


def add_clutter(frames: np.ndarray, max_value: int = 40, salt: int = 0) -> np.ndarray:
    return np.maximum(frames, clutter(frames.shape, max_value, salt))


def make_frame_pool(num_cameras: int, num_markers: int, num_frame_sets: int, seed: int = 0,
                    noise_max: int = 40):
    """(frames uint8 [B,C,H,W], truth f64 [B,M,3], poses, K)."""
    st = MarkerStream(num_cameras, num_markers, seed=seed, noise_max=noise_max)
    frames = np.empty((num_frame_sets, num_cameras, HEIGHT, WIDTH), dtype=np.uint8)
    truth = np.empty((num_frame_sets, num_markers, 3))
    for b in range(num_frame_sets):
        pts, _, imgs = st.next_frame_set(render=True)
        frames[b] = imgs
        truth[b] = pts
    return frames, truth, st.poses, st.K


def make_tracks(num_cameras: int, num_points: int, seed: int = 0, missing_frac: float = 0.1,
                round_to_int: bool = True):
    """Known-correspondence 2D tracks for bundle adjustment (BASELINE config 5).

    Returns (image_points object array [F,C,2] with None for missing views,
    true poses, K, true points [F,3]).  Each 3D point is an independent uniform
    sample of the cube (cold-start calibration waves one marker around).
    """
    rng = np.random.default_rng(seed)
    poses, K = make_rig(num_cameras)
    pts = rng.uniform(-0.5, 0.5, size=(num_points, 3)) + np.array([0.0, 0.0, 3.0])
    obs = np.empty((num_points, num_cameras, 2), dtype=object)
    for c, pose in enumerate(poses):
        uv = project(pts, pose, K)
        for f in range(num_points):
            if round_to_int:
                obs[f, c, 0], obs[f, c, 1] = int(uv[f, 0]), int(uv[f, 1])
            else:
                obs[f, c, 0], obs[f, c, 1] = float(uv[f, 0]), float(uv[f, 1])
    drop = rng.uniform(size=(num_points, num_cameras)) < missing_frac
    for f in range(num_points):
        if (~drop[f]).sum() < 2:
            drop[f, :] = False
        for c in range(num_cameras):
            if drop[f, c]:
                obs[f, c, 0] = None
                obs[f, c, 1] = None
    return obs, poses, K, pts


def perturb_poses(poses, seed: int = 1, rot_sigma: float = 0.03, t_sigma: float = 0.05):
    """Camera 0 stays (I,0); others get a rotvec/translation perturbation."""
    from scipy.spatial.transform import Rotation
    rng = np.random.default_rng(seed)
    out = [{"R": np.eye(3), "t": np.zeros(3)}]
    for p in poses[1:]:
        dR = Rotation.from_rotvec(rng.normal(0.0, rot_sigma, size=3)).as_matrix()
        out.append({"R": dR @ np.asarray(p["R"]), "t": np.asarray(p["t"]).reshape(3) + rng.normal(0.0, t_sigma, size=3)})
    return out


def make_drone_points(num_drones: int, num_clutter: int, seed: int = 0, jitter: float = 0.004):
    """3D point sets for locate_objects (helpers.py:424-480): per drone two markers 0.15 apart and a
    third 0.095 from both, random pose in a 2 m box, Gaussian jitter, plus clutter points; shuffled.
    Returns (points [K,3], errors [K])."""
    rng = np.random.default_rng(seed)
    pts = []
    half = 0.075
    h = np.sqrt(0.095 ** 2 - half ** 2)
    for _ in range(num_drones):
        centre = rng.uniform(-1.0, 1.0, size=3)
        yaw = rng.uniform(-np.pi, np.pi)
        side = rng.choice([-1.0, 1.0])
        c, s_ = np.cos(yaw), np.sin(yaw)
        local = np.array([[half, 0, 0], [-half, 0, 0], [0, side * h, 0]])
        Rz = np.array([[c, -s_, 0], [s_, c, 0], [0, 0, 1]])
        pts.extend(list(local @ Rz.T + centre + rng.normal(0, jitter, size=(3, 3))))
    pts.extend(list(rng.uniform(-1.0, 1.0, size=(num_clutter, 3))))
    pts = np.array(pts, dtype=np.float64).reshape(-1, 3)
    order = rng.permutation(len(pts))
    return pts[order], rng.uniform(0.05, 2.0, size=len(pts))
