"""ORACLE — TEST INFRASTRUCTURE ONLY.  **PARITY UNPINNED** (see oracle/tf_semantics.py header).

Independent direct-loop numpy (float64) restatement of the two TF ops whose semantics are easy to get wrong:
``tf.nn.conv2d(padding='SAME')`` (gym/envs/mujoco/arm_shaping.py:27) and ``tf.nn.conv2d_transpose``
(arm_shaping.py:71), written from TensorFlow's published definitions:

  conv2d:            out[n,p,q,o] = sum_{r,s,c} x[n, p*st + r - lo_h, q*st + s - lo_w, c] * w[r,s,c,o]
  conv2d_transpose:  the gradient of conv2d w.r.t. its input, i.e. a *scatter* of every input pixel through
                     the filter:  y[n, p*st + r - lo_h, q*st + s - lo_w, o] += x[n,p,q,i] * w[r,s,o,i]
  with SAME padding  out = ceil(in/st), total = max((out-1)*st + k - in, 0), lo = total // 2.

It shares no code with tf_semantics.py (which goes through torch's conv kernels), so agreement between the
two pins the padding / crop / layout conventions.  Pure-Python loops: use only on tiny shapes.
"""
import numpy as np


def same_geometry(size, k, st):
    out = -(-size // st)
    total = max((out - 1) * st + k - size, 0)
    lo = total // 2
    return out, lo


def conv2d_same(x, w, b, st):
    N, H, W, C = x.shape
    kh, kw, C2, O = w.shape
    assert C == C2
    P, lo_h = same_geometry(H, kh, st)
    Q, lo_w = same_geometry(W, kw, st)
    y = np.zeros((N, P, Q, O), dtype=np.float64)
    for p in range(P):
        for q in range(Q):
            for r in range(kh):
                h = p * st + r - lo_h
                if h < 0 or h >= H:
                    continue
                for s in range(kw):
                    ww = q * st + s - lo_w
                    if ww < 0 or ww >= W:
                        continue
                    y[:, p, q, :] += x[:, h, ww, :].astype(np.float64) @ w[r, s].astype(np.float64)
    if b is not None:
        y += b
    return y


def conv2d_transpose_same(x, w, b, out_hw, st):
    N, P, Q, Ci = x.shape
    kh, kw, O, Ci2 = w.shape
    assert Ci == Ci2
    H, W = out_hw
    P2, lo_h = same_geometry(H, kh, st)
    Q2, lo_w = same_geometry(W, kw, st)
    assert (P2, Q2) == (P, Q)
    y = np.zeros((N, H, W, O), dtype=np.float64)
    for p in range(P):
        for q in range(Q):
            for r in range(kh):
                h = p * st + r - lo_h
                if h < 0 or h >= H:
                    continue
                for s in range(kw):
                    ww = q * st + s - lo_w
                    if ww < 0 or ww >= W:
                        continue
                    y[:, h, ww, :] += x[:, p, q, :].astype(np.float64) @ w[r, s].astype(np.float64).T
    if b is not None:
        y += b
    return y


def conv2d_wgrad(big, small, kh, kw, st):
    """dw[r,s,a,b] = sum_{n,p,q} big[n, p*st+r-lo_h, q*st+s-lo_w, a] * small[n,p,q,b]  (filter gradient of
    conv2d_same with x=big, dy=small; also of conv2d_transpose_same with dy=big, x=small)."""
    N, H, W, A = big.shape
    _, P, Q, B = small.shape
    P2, lo_h = same_geometry(H, kh, st)
    Q2, lo_w = same_geometry(W, kw, st)
    assert (P2, Q2) == (P, Q)
    dw = np.zeros((kh, kw, A, B), dtype=np.float64)
    for p in range(P):
        for q in range(Q):
            for r in range(kh):
                h = p * st + r - lo_h
                if h < 0 or h >= H:
                    continue
                for s in range(kw):
                    ww = q * st + s - lo_w
                    if ww < 0 or ww >= W:
                        continue
                    dw[r, s] += big[:, h, ww, :].astype(np.float64).T @ small[:, p, q, :].astype(np.float64)
    return dw
