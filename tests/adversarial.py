"""Adversarial inputs shared by the CPU and GPU tests (test infrastructure)."""
import numpy as np


def _sq_unfused(d):
    xx = d[:, 0] * d[:, 0]; yy = d[:, 1] * d[:, 1]; zz = d[:, 2] * d[:, 2]
    return (xx + yy) + zz


def _sq_fused(d):
    """fma(dz,dz,fma(dy,dy,dx*dx)) emulated in float64 (exact products; a 24+48-bit sum of this size rounds once)"""
    t = (d[:, 0].astype(np.float64) ** 2).astype(np.float32)
    t = (d[:, 1].astype(np.float64) ** 2 + t.astype(np.float64)).astype(np.float32)
    return (d[:, 2].astype(np.float64) ** 2 + t.astype(np.float64)).astype(np.float32)


def fps_rot90_cloud(seed, n_pairs, scale=40.0):
    """A cloud whose first FPS round a fused multiply-add gets wrong.  Point 0 is the origin (the start) and
    every other point p = (x, y, z) comes with its quarter turn p' = (-y, x, z).  From the origin,
    (x*x + y*y) + z*z and (y*y + x*x) + z*z are the same float (addition commutes), so the reference's
    first arg-max picks the LOWER index of the farthest pair; fma(z,z,fma(y,y,x*x)) and fma(z,z,fma(x,x,y*y))
    round differently for ~12 % of the pairs.  The farthest pair is chosen to be one where the fused value of the
    HIGHER index is the larger one (pairs beyond it are dropped), so a contracted kernel picks the wrong point in
    round 1 and every later pick diverges.  -> [1 + 2*n_pairs, 3] float32"""
    rng = np.random.default_rng(seed)
    m = n_pairs + max(64, n_pairs // 4)
    p = (rng.random((m, 3), dtype=np.float32) * 2 - 1) * np.float32(scale)
    q = np.stack([-p[:, 1], p[:, 0], p[:, 2]], axis=1)
    both = np.stack([p, q], axis=1)                         # [m, 2, 3]
    swap = rng.random(m) < 0.5                              # which of the two comes first
    both[swap] = both[swap][:, ::-1]
    lo, hi = _sq_fused(both[:, 0]), _sq_fused(both[:, 1])
    u = _sq_unfused(both[:, 0])
    assert np.array_equal(u, _sq_unfused(both[:, 1]))
    trap = np.flatnonzero(hi > lo)                          # a fused kernel prefers the higher index here
    star = trap[np.argmax(u[trap])]
    keep = np.flatnonzero((u < u[star]) | (np.arange(m) == star))
    rest = keep[keep != star]
    assert rest.size >= n_pairs - 1, "not enough pairs left"
    chosen = np.sort(np.concatenate([rest[:n_pairs - 1], [star]]))
    pts = both[chosen].reshape(-1, 3)
    return np.concatenate([np.zeros((1, 3), np.float32), pts], 0).astype(np.float32)


def first_round_pick(cloud, fused):
    """index FPS picks after the start at point 0; fused=True emulates fma(dz,dz,fma(dy,dy,dx*dx))
    (float64 holds the exact products and a 24+48-bit sum of this size rounds once)."""
    d = cloud.astype(np.float32) - cloud[0]
    return int(np.argmax(_sq_fused(d) if fused else _sq_unfused(d)))
