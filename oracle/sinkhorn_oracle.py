"""CPU oracle for EMLight's Sinkhorn earth-mover loss (test infrastructure).

Restates in numpy the tensorized debiased Sinkhorn divergence of the vendored GeomLoss 0.2.3
(``RegressionNetwork/geomloss``) as the reference calls it from train.py:61,90-92 with
``SamplesLoss("sinkhorn", p=2, blur=.025, batchsize=B)`` on x,y of shape (B,N,1):

* cost               geomloss/utils.py:85-99 + samples_loss.py:82  C^{ab}_ij = (0.1 (a_i-b_j)^2 + M_ij) / 2,
                     M_ij = |anchor_i - anchor_j| over the fp32 Fibonacci lattice (utils.py:64-81).
                     (the reference expands (a-b)^2 as a^2 - 2ab + b^2; ``expanded=True`` reproduces that)
* weights            samples_loss.py:62-70  uniform 1/N ; log_weights sinkhorn_divergence.py:47-50
* schedule           sinkhorn_divergence.py:9-36  diameter = |max - min| over ALL x and y values of the batch,
                     eps_s = [d^2] + exp(arange(2 ln d, 2 ln blur, 2 ln scaling)) + [blur^2]
* loop               sinkhorn_divergence.py:72-109 (init at eps_s[0]; symmetric averaging; final extrapolation)
* softmin            samples_loss.py:75-77   -eps * logsumexp_j(h_j - C_ij / eps)
* cost               sinkhorn_divergence.py:65-69  sum_i a_i (b_x - a_x)_i + sum_j b_j (a_y - b_y)_j

Gradient w.r.t. x (only C_xx's first argument and C_xy carry gradient; second arguments are detached at
samples_loss.py:85-86 and utils.py:88; the last softmins use detached potentials, sinkhorn_divergence.py:102-107):
    dL_b/dx_i = (1/N) [ sum_j P^{xy}_ij 0.1 (x_i - y_j) - sum_j P^{xx}_ij 0.1 (x_i - x_j) ],
    P^{ab}_i. = softmax_j(w_j - C^{ab}_ij / eps).

The gmloss variant (RegressionNetwork/gmloss/utils.py:63-108) only changes the anchors:
(r_k cos(theta_k), r_k sin(theta_k), z_k) with per-anchor radius ``geometry`` -- see ``geometric_anchor_distances``.
"""
import numpy as np

from .render_oracle import sphere_points


def anchor_distances(n, dtype=np.float32):
    """M (n,n): chord lengths between fp32 anchors (geomloss/utils.py:66-77)."""
    a = sphere_points(n).astype(np.float32)
    d = a[:, None, :] - a[None, :, :]
    return np.sqrt((d * d).sum(-1)).astype(dtype)


def geometric_anchor_distances(geometry, dtype=np.float32):
    """gmloss/utils.py:63-93: anchors scaled in the xy-plane by a per-anchor radius."""
    n = len(geometry)
    golden_angle = np.pi * (3 - np.sqrt(5))
    theta = golden_angle * np.arange(n)
    z = np.linspace(1 - 1.0 / n, 1.0 / n - 1, n)
    a = np.stack((np.asarray(geometry, np.float64) * np.cos(theta), np.asarray(geometry, np.float64) * np.sin(theta), z), 1)
    a = a.astype(np.float32)
    d = a[:, None, :] - a[None, :, :]
    return np.sqrt((d * d).sum(-1)).astype(dtype)


def epsilon_schedule(diameter, blur=0.025, scaling=0.5, p=2):
    return [diameter ** p] + [float(np.exp(e)) for e in
                              np.arange(p * np.log(diameter), p * np.log(blur), p * np.log(scaling))] + [blur ** p]


def _lse(a, axis):
    m = a.max(axis=axis, keepdims=True)
    return (m + np.log(np.exp(a - m).sum(axis=axis, keepdims=True))).squeeze(axis)


def sinkhorn_loss(x, y, M=None, blur=0.025, scaling=0.5, dtype=np.float64, expanded=False, diameter=None,
                  return_grad=True):
    """x,y: (B,N) or (B,N,1).  Returns (loss (B,), grad_x (B,N))."""
    x = np.asarray(x, dtype).reshape(np.shape(x)[0], -1)
    y = np.asarray(y, dtype).reshape(np.shape(y)[0], -1)
    B, N = x.shape
    if M is None:
        M = anchor_distances(N)
    M = np.asarray(M, dtype)
    if M.ndim == 2:
        M = M[None]

    def cost(a, b):
        if expanded:
            sq = (a * a)[:, :, None] - 2 * a[:, :, None] * b[:, None, :] + (b * b)[:, None, :]
        else:
            sq = (a[:, :, None] - b[:, None, :]) ** 2
        return (sq * dtype(0.1) + M) / dtype(2)

    C_xx, C_yy, C_xy, C_yx = cost(x, x), cost(y, y), cost(x, y), cost(y, x)
    if diameter is None:
        lo = min(x.min(), y.min()); hi = max(x.max(), y.max())
        diameter = float(abs(dtype(hi) - dtype(lo)))       # D == 1: the norm of a 1-vector
    eps_s = epsilon_schedule(diameter, blur, scaling)
    lw = dtype(np.log(1.0 / N))

    def softmin(eps, C, h):                                  # h: (B,N) over columns j
        return -eps * _lse(h[:, None, :] - C / eps, 2)

    eps = dtype(eps_s[0])
    lwv = np.full((B, N), lw, dtype)
    a_x = softmin(eps, C_xx, lwv); b_y = softmin(eps, C_yy, lwv)
    a_y = softmin(eps, C_yx, lwv); b_x = softmin(eps, C_xy, lwv)
    for e in eps_s:
        eps = dtype(e)
        at_x = softmin(eps, C_xx, lw + a_x / eps); bt_y = softmin(eps, C_yy, lw + b_y / eps)
        at_y = softmin(eps, C_yx, lw + b_x / eps); bt_x = softmin(eps, C_xy, lw + a_y / eps)
        a_x, b_y = dtype(.5) * (a_x + at_x), dtype(.5) * (b_y + bt_y)
        a_y, b_x = dtype(.5) * (a_y + at_y), dtype(.5) * (b_x + bt_x)
    w_x, w_y, w_bx, w_by = lw + a_x / eps, lw + b_y / eps, lw + b_x / eps, lw + a_y / eps
    f_ax = softmin(eps, C_xx, w_x); f_by = softmin(eps, C_yy, w_y)
    f_ay = softmin(eps, C_yx, w_bx); f_bx = softmin(eps, C_xy, w_by)
    loss = ((f_bx - f_ax) + (f_ay - f_by)).sum(1) / N
    if not return_grad:
        return loss, None

    def softmax_rows(C, w):
        t = w[:, None, :] - C / eps
        t = t - t.max(2, keepdims=True)
        p = np.exp(t)
        return p / p.sum(2, keepdims=True)

    P_xy = softmax_rows(C_xy, w_by); P_xx = softmax_rows(C_xx, w_x)
    g = (P_xy * (x[:, :, None] - y[:, None, :])).sum(2) - (P_xx * (x[:, :, None] - x[:, None, :])).sum(2)
    return loss, dtype(0.1) * g / N
