"""TEST INFRASTRUCTURE ONLY -- restatement of PolyPCI's host-side polynomial fit (PolyPCI/Models/Models_V1.py:116-124,
call site :191-219), step by step as the reference does it: np.polyfit through the F time stamps for each point,
PolynomialFeatures(degree) of t, np.flip, np.matmul, then float32.  Pinned by tests/golden (the real function's source
is executed there by tests/golden/make_golden.py)."""
import numpy as np


def _poly_features(t, degree):
    try:
        from sklearn.preprocessing import PolynomialFeatures           # what the reference imports (Models_V1.py:7)
        return PolynomialFeatures(degree=degree).fit_transform(np.array(t, dtype=np.float64).reshape(-1, 1))
    except ImportError:                                                  # same values: [1, t, t^2, ...]
        t = float(np.asarray(t).reshape(-1)[0])
        cols = [1.0]
        for _ in range(degree):
            cols.append(cols[-1] * t)
        return np.array([cols], dtype=np.float64)


def fitting_and_predict(x, y, t, degree):
    """Models_V1.py:116-124.  x: time stamps [F]; y: frames [F,N] (float32, as np.array(tensor.cpu())); -> [1,N] float64."""
    coefficients = np.polyfit(x, y, degree)
    X_poly = np.flip(_poly_features(t, degree), axis=1)
    return np.matmul(X_poly, coefficients)


def forward_tail(frames, T_list, t, degree):
    """Models_V1.py:187-219.  frames: list of F arrays [B,3,N] float32 -> [B,3,N] float32."""
    B = frames[0].shape[0]
    out = []
    for i in range(B):
        T = np.array(T_list[i]).reshape(-1)
        rows = []
        for c in range(3):
            ys = np.stack([f[i, c, :] for f in frames], axis=0)          # xs_all[i] / ys_all[i] / zs_all[i]: [F,N]
            rows.append(fitting_and_predict(T, ys, t[i], degree).astype(np.float32))
        out.append(np.concatenate(rows, axis=0))
    return np.stack(out, axis=0)
