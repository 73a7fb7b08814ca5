"""PolyPCI's polynomial trajectory fit on the device.

Reference: PolyPCI/Models/Models_V1.py:116-124 (`fitting_and_predict`) and its call site :191-219.  For every batch
item and coordinate the reference copies the stacked frames [F,N] to the host, calls `np.polyfit(T, frames, degree)`
(an independent least-squares fit per point), evaluates the polynomial at t and copies the result back.  Least squares
is linear in the data, so all N fits of a batch item share ONE weight vector

    value[n] = sum_f w[f] * frames[f, n],      w = [t^d, ..., t, 1] . polyfit(T, I_F, d)

`poly_weights` obtains w from the very same numpy call the reference makes (on the identity instead of the data), in
float64; `fit_and_predict` applies it with one kernel (`b200pc_poly_predict`: float64 accumulation, one rounding to
fp32) -- the point data never leaves the device.
"""
import numpy as np
import torch

from . import ops


def poly_weights(T, t, degree):
    """T: the F time stamps of the stacked frames; t: the query time; -> float64 [F]."""
    T = np.asarray(T, dtype=np.float64).reshape(-1)
    coef = np.polyfit(T, np.eye(T.shape[0]), int(degree))                 # [degree+1, F], highest power first
    powers = np.ones(int(degree) + 1, dtype=np.float64)
    for i in range(int(degree) - 1, -1, -1):                              # [t^d, ..., t, 1] by repeated multiplication,
        powers[i] = powers[i + 1] * float(t)                              # like PolynomialFeatures + np.flip
    return powers @ coef


def fit_and_predict(frames, T_list, t, degree):
    """frames: list of F CUDA tensors [B,3,N] in the reference's stacking order (key, forward 0, backward 0, ...);
    T_list[b]: the F time stamps of batch item b; t[b]: its query time.  -> [B,3,N] fp32 on the same device
    (what PolyPCI.forward returns, Models_V1.py:187-219)."""
    if not frames:
        raise ValueError("fit_and_predict: no frames")
    dev = frames[0].device
    if dev.type != "cuda":
        raise RuntimeError("b200pc: frames are on %s; this library has no CPU path (CUDA tensors only)" % dev)
    B = frames[0].shape[0]
    per_batch = int(frames[0][0].numel())
    fr = []
    for f in frames:
        if f.shape != frames[0].shape or f.device != dev:
            raise ValueError("fit_and_predict: all frames must share shape and device")
        fr.append(f.float().contiguous())
    F = len(fr)
    t_host = [float(x) for x in (t.detach().reshape(-1).tolist() if isinstance(t, torch.Tensor) else np.asarray(t).reshape(-1))]
    if len(t_host) != B or len(T_list) != B:
        raise ValueError("fit_and_predict: need one time stamp list and one query time per batch item")
    w = np.stack([poly_weights(T_list[b], t_host[b], degree) for b in range(B)])            # [B,F] float64
    if w.shape[1] != F:
        raise ValueError("fit_and_predict: %d time stamps for %d frames" % (w.shape[1], F))
    wd = torch.from_numpy(np.ascontiguousarray(w)).to(dev)
    return ops.poly_predict(fr, wd)
