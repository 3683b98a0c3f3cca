"""Synthetic KITTI-shaped scans: a spinning multi-beam lidar over a ground plane and
axis-aligned boxes, range noise, optional voxel-grid downsample.  Pure numpy (input generation, not part of the hot path).
"""
import numpy as np


def _ray_box(o, d, lo, hi):
    """Slab test for rays o + t d (o (3,), d (R,3)) against the box [lo, hi]; returns t (inf when missed)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        t0 = (lo - o) * inv
        t1 = (hi - o) * inv
    tmin = np.minimum(t0, t1).max(axis=1)
    tmax = np.maximum(t0, t1).min(axis=1)
    hit = (tmax >= np.maximum(tmin, 0.0)) & (tmin > 0.0)
    return np.where(hit, tmin, np.inf)


def _make_world(rng, n_boxes=40):
    boxes = []
    while len(boxes) < n_boxes:
        c = rng.uniform(-80, 80, size=2)
        if np.hypot(*c) < 12.0:
            continue
        wl = rng.uniform(2, 20, size=2)
        h = rng.uniform(2, 12)
        boxes.append((np.array([c[0] - wl[0] / 2, c[1] - wl[1] / 2, -1.73]), np.array([c[0] + wl[0] / 2, c[1] + wl[1] / 2, -1.73 + h])))
    return boxes


def scan(world, beams, az_steps, sensor_pose, noise_rng, max_range=120.0, noise=0.02):
    """One revolution seen from `sensor_pose` (4x4, sensor -> world); returns points in the SENSOR frame, float32."""
    el = np.radians(np.linspace(-24.8, 2.0, beams))
    az = np.linspace(-np.pi, np.pi, az_steps, endpoint=False)
    ce, se = np.cos(el)[:, None], np.sin(el)[:, None]
    d_s = np.stack([ce * np.cos(az)[None, :], ce * np.sin(az)[None, :], np.broadcast_to(se, (beams, az_steps))], axis=-1).reshape(-1, 3)
    R, o = sensor_pose[:3, :3], sensor_pose[:3, 3]
    d = d_s @ R.T
    t = np.full(len(d), np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        tg = (-1.73 - o[2]) / d[:, 2]  # ground plane z = -1.73
    t = np.where((tg > 0) & np.isfinite(tg), tg, t)
    for lo, hi in world:
        t = np.minimum(t, _ray_box(o, d, lo, hi))
    ok = np.isfinite(t) & (t < max_range) & (t > 1.0)
    r = t[ok] + noise_rng.normal(0.0, noise, size=int(ok.sum()))
    return (d_s[ok] * r[:, None]).astype(np.float32)


def voxel_downsample(pts, leaf):
    """Exact centroid voxel grid (pcl::VoxelGrid-like), output ordered by voxel index."""
    if leaf <= 0:
        return pts
    ijk = np.floor(pts / np.float32(leaf)).astype(np.int64)
    ijk -= ijk.min(axis=0)
    dims = ijk.max(axis=0) + 1
    idx = ijk[:, 0] + dims[0] * (ijk[:, 1] + dims[1] * ijk[:, 2])
    order = np.argsort(idx, kind="stable")
    idx_s = idx[order]
    starts = np.flatnonzero(np.r_[True, idx_s[1:] != idx_s[:-1]])
    sums = np.add.reduceat(pts[order].astype(np.float64), starts, axis=0)
    counts = np.diff(np.r_[starts, len(idx_s)])[:, None]
    return (sums / counts).astype(np.float32)


def kitti_like_pair(beams=64, az_steps=2083, seed=42, pose=(0.8, 0.05, 0.7), downsample=0.25, max_points=None):
    """(target, source, T_gt): target scanned at the origin, source from a sensor moved by pose=(tx, ty, yaw_deg).
    T_gt maps source-frame points into the target frame (what registration should recover)."""
    rng = np.random.default_rng(seed)
    world = _make_world(rng)
    T = np.eye(4)
    yaw = np.radians(pose[2])
    T[:3, :3] = [[np.cos(yaw), -np.sin(yaw), 0], [np.sin(yaw), np.cos(yaw), 0], [0, 0, 1]]
    T[:3, 3] = [pose[0], pose[1], 0.0]
    tgt = scan(world, beams, az_steps, np.eye(4), np.random.default_rng(seed))
    src = scan(world, beams, az_steps, T, np.random.default_rng(seed + 1))
    if max_points:
        tgt, src = tgt[:max_points], src[:max_points]
    return voxel_downsample(tgt, downsample), voxel_downsample(src, downsample), T
