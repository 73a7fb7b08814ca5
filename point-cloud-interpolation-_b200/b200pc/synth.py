"""Deterministic synthetic LiDAR sweeps shaped like a Velodyne HDL-64 (KITTI) frame.

SURVEY.md section 8(d) fixes the recipe so that every test, golden fixture and bench line is
regenerable from a seed: 64 beams with elevation uniform in [-24.8, +2.0] degrees, 2083
azimuth steps, sensor 1.73 m above a ground plane, a piece-wise constant wall profile
(40 angular segments, 4..60 m), a few boxes, an 80 m range cap, N(0, 0.02 m) range noise.
Frame 2 of a pair is the same scene re-scanned after ego motion (1.2 m forward, 1 degree yaw).
Real input shapes being mimicked: PointINet20230424/data/interpolation_data.py:66-77 (random
sub-sample to `npoints`) in the reference.

numpy only; no torch import here so the oracle side can use it too.
"""
import numpy as np

N_BEAMS = 64
N_AZIMUTH = 2083
SENSOR_HEIGHT = 1.73
MAX_RANGE = 80.0


def _scene(rng):
    walls = rng.uniform(4.0, 60.0, size=40).astype(np.float64)
    n_box = 6
    boxes = np.stack([rng.uniform(-30, 30, n_box), rng.uniform(-30, 30, n_box),          # centre x,y
                      rng.uniform(1.0, 3.0, n_box), rng.uniform(1.0, 3.0, n_box)], 1)    # half sizes
    return walls, boxes


def _scan(walls, boxes, pose, rng):
    """one sweep from pose = (tx, ty, yaw) -> [n_returns, 3] float32 in the sensor frame."""
    tx, ty, yaw = pose
    elev = np.deg2rad(np.linspace(-24.8, 2.0, N_BEAMS))
    azim = np.linspace(-np.pi, np.pi, N_AZIMUTH, endpoint=False)
    el, az = np.meshgrid(elev, azim, indexing="ij")
    el = el.ravel(); az = az.ravel()
    cos_el = np.cos(el); sin_el = np.sin(el)
    # ground-plane hit (only for rays pointing down)
    with np.errstate(divide="ignore", invalid="ignore"):
        r_ground = np.where(sin_el < -1e-6, SENSOR_HEIGHT / -sin_el, np.inf)
    # wall profile is attached to the world: look it up with the world azimuth
    world_az = (az + yaw + np.pi) % (2 * np.pi)
    seg = np.minimum((world_az / (2 * np.pi) * walls.size).astype(np.int64), walls.size - 1)
    # walls are "cylindrical" around the world origin; correct the horizontal range for the offset
    r_wall = np.maximum(walls[seg] - (tx * np.cos(az + yaw) + ty * np.sin(az + yaw)), 1.0) / np.maximum(cos_el, 1e-3)
    rng_m = np.minimum(np.minimum(r_ground, r_wall), MAX_RANGE)
    # boxes: horizontal slab test along the ray in the world frame
    dxw = np.cos(az + yaw) * cos_el; dyw = np.sin(az + yaw) * cos_el
    for cx, cy, hx, hy in boxes:
        with np.errstate(divide="ignore", invalid="ignore"):
            t1 = (cx - hx - tx) / dxw; t2 = (cx + hx - tx) / dxw
            t3 = (cy - hy - ty) / dyw; t4 = (cy + hy - ty) / dyw
        tn = np.maximum(np.minimum(t1, t2), np.minimum(t3, t4))
        tf = np.minimum(np.maximum(t1, t2), np.maximum(t3, t4))
        hit = (tf >= tn) & (tn > 0.5) & np.isfinite(tn)
        z_at = SENSOR_HEIGHT + tn * sin_el
        hit &= (z_at > 0.0) & (z_at < 2.0)
        rng_m = np.where(hit & (tn < rng_m), tn, rng_m)
    keep = rng_m < MAX_RANGE - 1e-3
    r = rng_m[keep] + rng.normal(0.0, 0.02, size=int(keep.sum()))
    x = r * cos_el[keep] * np.cos(az[keep])
    y = r * cos_el[keep] * np.sin(az[keep])
    z = r * sin_el[keep]
    return np.stack([x, y, z], 1).astype(np.float32)


def frame_pair(i, npoints):
    """pair i -> (frame1, frame2) each [npoints, 3] float32.
    scene/noise seed 1000+i, sub-sample seed 2000+i (without replacement)."""
    rng = np.random.default_rng(1000 + i)
    walls, boxes = _scene(rng)
    f1 = _scan(walls, boxes, (0.0, 0.0, 0.0), rng)
    f2 = _scan(walls, boxes, (1.2, 0.0, np.deg2rad(1.0)), rng)
    sub = np.random.default_rng(2000 + i)
    out = []
    for f in (f1, f2):
        if f.shape[0] >= npoints:
            sel = sub.choice(f.shape[0], npoints, replace=False)
        else:  # pad by re-drawing, like the reference's loader does for short frames
            sel = np.concatenate([np.arange(f.shape[0]), sub.choice(f.shape[0], npoints - f.shape[0], replace=True)])
        out.append(np.ascontiguousarray(f[sel]))
    return out[0], out[1]


def batch_pairs(first, count, npoints):
    """`count` pairs starting at index `first` -> ([count,npoints,3], [count,npoints,3])."""
    a, b = zip(*(frame_pair(first + j, npoints) for j in range(count)))
    return np.stack(a), np.stack(b)


def grid_snapped(seed, B, N, extent=64.0, step=2.0 ** -6, span=None):
    """tie stress set: coordinates are multiples of `step` so that all fp32 arithmetic on them
    is exact and exactly equal distances are common.  `span` (in steps) narrows the cloud to
    force many duplicates / ties."""
    rng = np.random.default_rng(seed)
    hi = int(extent / step) if span is None else int(span)
    pts = rng.integers(-hi, hi + 1, size=(B, N, 3)).astype(np.float32) * np.float32(step)
    return pts
