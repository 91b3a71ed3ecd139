"""ORACLE (test infrastructure, NOT product code) -- numpy restatement of the PoseNDF hot path.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this module.  The product path (posendf_b200/) never does; it fails loudly without the CUDA library.

What is restated, and from where (all paths relative to /root/reference):
  normalise_columns   model/posendf.py:71                (F.normalize(pose, dim=1) on (B,21,4))
  PARENTS             model/network/net_utils.py:44-50   (table at :46)
  bone MLP            model/network/net_modules.py:75-111
  structure encoder   model/network/net_modules.py:140-170 (child input = cat(quat_i, feat_parent), :167)
  DFNet               model/network/net_modules.py:9-72
  forward             model/posendf.py:62-76,100-101
  gradient            model/posendf.py:18-27 (autograd there; closed form here, SURVEY Appx A)
  projection step     experiments/sample_poses.py:70-74
  prior term          experiments/motion_denoise.py:81-83 (+ pytorch3d 0.7.2 axis_angle_to_quaternion,
                      source NOT in /root/reference: restated from its published formula -- that one
                      boundary is "parity unpinned")
  train-mode losses   model/posendf.py:78-99

Pinning: tests/golden/*.npz were produced by tests/golden/make_golden.py, which imports the REAL
reference modules from /root/reference (fp32 and fp64) in the build container; tests/test_oracle.py
checks this restatement against those vectors, so the oracle is pinned for everything except the
pytorch3d axis-angle boundary named above.

dtype follows the inputs: float32 in -> float32 arithmetic (numpy/BLAS), float64 in -> float64.
"""
from __future__ import annotations

import numpy as np

PARENTS = (-1, -1, -1, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19)
NJ = 21
EPS_NORMALIZE = 1e-12          # torch.nn.functional.normalize default eps
LRELU_SLOPE = 0.01             # nn.LeakyReLU() default
SOFTPLUS_THRESHOLD = 20.0      # nn.Softplus default threshold


# ----------------------------------------------------------------------------- activations
def act(x, kind, beta):
    if kind == "relu":
        return np.maximum(x, 0)
    if kind == "lrelu":
        return np.where(x > 0, x, x * x.dtype.type(LRELU_SLOPE))
    if kind == "softplus":
        b = x.dtype.type(beta)
        bx = x * b
        with np.errstate(over="ignore"):
            soft = np.log1p(np.exp(bx)) / b
        return np.where(bx > SOFTPLUS_THRESHOLD, x, soft)
    raise ValueError(kind)


def dact(x, kind, beta):
    """derivative w.r.t. the pre-activation, torch conventions at the kinks."""
    one = x.dtype.type(1)
    if kind == "relu":
        return np.where(x > 0, one, x.dtype.type(0))
    if kind == "lrelu":
        return np.where(x > 0, one, x.dtype.type(LRELU_SLOPE))
    if kind == "softplus":
        b = x.dtype.type(beta)
        bx = x * b
        with np.errstate(over="ignore"):
            sig = one / (one + np.exp(-bx))
        return np.where(bx > SOFTPLUS_THRESHOLD, one, sig)
    raise ValueError(kind)


def out_act_kind(df_act):
    """relu / lrelu configs end in nn.ReLU, softplus ends in Softplus(beta) (net_modules.py:30-41)."""
    return "softplus" if df_act == "softplus" else "relu"


def default_cfg(**kw):
    cfg = dict(use_enc=True, enc_act="lrelu", enc_beta=100.0, df_act="lrelu", df_beta=100.0)
    cfg.update(kw)
    return cfg


def _p(params, name, dtype):
    return np.asarray(params[name]).astype(dtype, copy=False)


def num_df_layers(params):
    n = 0
    while f"dfnet.lin{n}.weight" in params:
        n += 1
    return n


# ----------------------------------------------------------------------------- forward pieces
def normalise_columns(x):
    """x (B,21,4): L2-normalise each of the 4 quaternion COMPONENTS across the 21 joints."""
    n = np.sqrt(np.sum(x * x, axis=1, keepdims=True))
    n = np.maximum(n, x.dtype.type(EPS_NORMALIZE))
    return x / n, n


def encoder_forward(params, q, cfg):
    """q (B,21,4) -> features (B,126), cache of pre-activations for the backward."""
    dt = q.dtype
    feats = [None] * NJ
    cache = []
    for i in range(NJ):
        par = PARENTS[i]
        u = q[:, i, :] if par < 0 else np.concatenate([q[:, i, :], feats[par]], axis=1)
        w1 = _p(params, f"enc.net.{i}.net.0.weight", dt); b1 = _p(params, f"enc.net.{i}.net.0.bias", dt)
        w2 = _p(params, f"enc.net.{i}.net.2.weight", dt); b2 = _p(params, f"enc.net.{i}.net.2.bias", dt)
        pre1 = u @ w1.T + b1
        h = act(pre1, cfg["enc_act"], cfg["enc_beta"])
        pre2 = h @ w2.T + b2
        feats[i] = act(pre2, cfg["enc_act"], cfg["enc_beta"])
        cache.append((pre1, pre2))
    return np.concatenate(feats, axis=1), cache


def dfnet_forward(params, z, cfg):
    dt = z.dtype
    L = num_df_layers(params)
    pres = []
    for l in range(L):
        w = _p(params, f"dfnet.lin{l}.weight", dt); b = _p(params, f"dfnet.lin{l}.bias", dt)
        pre = z @ w.T + b
        pres.append(pre)
        if l < L - 1:
            z = act(pre, cfg["df_act"], cfg["df_beta"])
    d = act(pres[-1], out_act_kind(cfg["df_act"]), cfg["df_beta"])
    return d, pres


def forward(params, pose, cfg, normalise=True):
    """PoseNDF.forward(train=False): pose (...) -> d (B,1).  normalise=False is the manifold branch of
    the train path (model/posendf.py:80-83)."""
    x = np.asarray(pose).reshape(-1, NJ, 4)
    q = normalise_columns(x)[0] if normalise else x
    z = encoder_forward(params, q, cfg)[0] if cfg["use_enc"] else q.reshape(len(q), -1)
    return dfnet_forward(params, z, cfg)[0]


def _dact_flip(pre, kind, beta, units):
    """dact with the derivative branch of the listed units inverted (piecewise-linear activations only): what an fp32
    evaluation computes when rounding puts a pre-activation that is ~0 on the other side of its kink."""
    d = dact(pre, kind, beta)
    if units:
        assert kind in ("relu", "lrelu")
        lo = pre.dtype.type(0.0 if kind == "relu" else LRELU_SLOPE)
        for u in units:
            d[:, u] = np.where(d[:, u] == 1, lo, pre.dtype.type(1))
    return d


def forward_grad(params, pose, cfg, g_up=None, normalise=True, flip=None):
    """d (B,1) and  g_up[b] * dd_b/dpose_b  (B,21,4), closed form (SURVEY Appx A).
    flip (tests only): {("df", l): [units], ("enc", i, 0|1): [units]} -- evaluate the gradient with those units on the other
    branch of their kink (see _dact_flip)."""
    flip = flip or {}
    x = np.asarray(pose).reshape(-1, NJ, 4)
    dt = x.dtype
    B = len(x)
    if normalise:
        q, n = normalise_columns(x)
    else:
        q, n = x, None
    if cfg["use_enc"]:
        z0, ecache = encoder_forward(params, q, cfg)
    else:
        z0, ecache = q.reshape(B, -1), None
    d, pres = dfnet_forward(params, z0, cfg)
    L = len(pres)

    g = np.ones((B, 1), dtype=dt) if g_up is None else np.asarray(g_up, dtype=dt).reshape(B, 1)
    g = g * dact(pres[-1], out_act_kind(cfg["df_act"]), cfg["df_beta"])
    g = g @ _p(params, f"dfnet.lin{L-1}.weight", dt)
    for l in range(L - 2, -1, -1):
        g = (g * _dact_flip(pres[l], cfg["df_act"], cfg["df_beta"], flip.get(("df", l)))) @ _p(params, f"dfnet.lin{l}.weight", dt)

    if cfg["use_enc"]:
        fbar = [g[:, 6 * i:6 * i + 6].copy() for i in range(NJ)]
        qbar = np.zeros_like(q)
        for i in range(NJ - 1, -1, -1):
            pre1, pre2 = ecache[i]
            w1 = _p(params, f"enc.net.{i}.net.0.weight", dt); w2 = _p(params, f"enc.net.{i}.net.2.weight", dt)
            t = (fbar[i] * _dact_flip(pre2, cfg["enc_act"], cfg["enc_beta"], flip.get(("enc", i, 1)))) @ w2
            ubar = (t * _dact_flip(pre1, cfg["enc_act"], cfg["enc_beta"], flip.get(("enc", i, 0)))) @ w1
            qbar[:, i, :] += ubar[:, :4]
            if PARENTS[i] >= 0:
                fbar[PARENTS[i]] += ubar[:, 4:10]
    else:
        qbar = g.reshape(B, NJ, 4)

    if normalise:
        # Jacobian of x/max(|x|,eps) per component column.  Where the clamp is active the map is
        # linear (x/eps) -- unreachable for real poses, kept for completeness.
        dot = np.sum(q * qbar, axis=1, keepdims=True)
        clamped = (np.sqrt(np.sum(x * x, axis=1, keepdims=True)) < EPS_NORMALIZE)
        xbar = np.where(clamped, qbar / n, (qbar - q * dot) / n)
    else:
        xbar = qbar
    return d, xbar


def kink_units(params, pose, cfg, eps=1e-5, normalise=True):
    """Piecewise-linear units of ONE pose whose pre-activation sits within eps of its kink, relative to the magnitude of the
    terms it sums (sum |w_k z_k| + |b|, the scale an fp32 rounding error lives on): [(key, unit, margin)] sorted by margin,
    key as in forward_grad's `flip`.  Evaluate in float64."""
    x = np.asarray(pose, dtype=np.float64).reshape(1, NJ, 4)
    q = normalise_columns(x)[0] if normalise else x
    out = []

    def scan(key, pre, zin, w, b, kind):
        if kind == "softplus":
            return
        denom = np.abs(zin) @ np.abs(w).T + np.abs(b)
        m = (np.abs(pre) / np.maximum(denom, 1e-300))[0]
        out.extend((key, int(u), float(m[u])) for u in np.nonzero(m < eps)[0])

    if cfg["use_enc"]:
        feats = [None] * NJ
        for i in range(NJ):
            par = PARENTS[i]
            u = q[:, i, :] if par < 0 else np.concatenate([q[:, i, :], feats[par]], axis=1)
            w1 = _p(params, f"enc.net.{i}.net.0.weight", np.float64); b1 = _p(params, f"enc.net.{i}.net.0.bias", np.float64)
            w2 = _p(params, f"enc.net.{i}.net.2.weight", np.float64); b2 = _p(params, f"enc.net.{i}.net.2.bias", np.float64)
            pre1 = u @ w1.T + b1
            scan(("enc", i, 0), pre1, u, w1, b1, cfg["enc_act"])
            h = act(pre1, cfg["enc_act"], cfg["enc_beta"])
            pre2 = h @ w2.T + b2
            scan(("enc", i, 1), pre2, h, w2, b2, cfg["enc_act"])
            feats[i] = act(pre2, cfg["enc_act"], cfg["enc_beta"])
        z = np.concatenate(feats, axis=1)
    else:
        z = q.reshape(1, -1)
    L = num_df_layers(params)
    for l in range(L - 1):
        w = _p(params, f"dfnet.lin{l}.weight", np.float64); b = _p(params, f"dfnet.lin{l}.bias", np.float64)
        pre = z @ w.T + b
        scan(("df", l), pre, z, w, b, cfg["df_act"])
        z = act(pre, cfg["df_act"], cfg["df_beta"])
    return sorted(out, key=lambda t: t[2])


def project(params, pose, cfg, steps=10, renorm=False, return_traj=False):
    """experiments/sample_poses.py:70-74: x <- x - d * dd/dx, `steps` times; no renormalisation in the
    reference (renorm=True is the north-star option: per-quaternion renormalise after each step)."""
    x = np.array(pose, copy=True).reshape(-1, NJ, 4)
    traj = []
    d = None
    for _ in range(steps):
        d, g = forward_grad(params, x, cfg)
        x = x - d.reshape(-1, 1, 1) * g
        if renorm:
            x = x / np.sqrt(np.sum(x * x, axis=2, keepdims=True))
        if return_traj:
            traj.append(x.copy())
    return (x, d, traj) if return_traj else (x, d)


# ----------------------------------------------------------------------------- motion-denoise prior term
def axis_angle_to_quaternion(aa):
    """pytorch3d 0.7.2 transforms.axis_angle_to_quaternion (not vendored in /root/reference; formula as
    published): angle=|aa|; k = sin(angle/2)/angle, small-angle (|angle|<1e-6) k = 1/2 - angle^2/48;
    quaternion = (cos(angle/2), k*aa), real part first.  PARITY UNPINNED (no copy of pytorch3d here)."""
    aa = np.asarray(aa)
    dt = aa.dtype
    ang = np.sqrt(np.sum(aa * aa, axis=-1, keepdims=True))
    half = ang * dt.type(0.5)
    small = np.abs(ang) < 1e-6
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(small, dt.type(0.5) - ang * ang / dt.type(48), np.sin(half) / np.where(small, dt.type(1), ang))
    return np.concatenate([np.cos(half), aa * k], axis=-1)


def axis_angle_to_quaternion_vjp(aa, qbar):
    """VJP of the map above; at angle->0 uses the analytic limit (the reference's autograd gives NaN
    there) -- documented deviation."""
    aa = np.asarray(aa); dt = aa.dtype
    ang2 = np.sum(aa * aa, axis=-1, keepdims=True)
    ang = np.sqrt(ang2)
    half = ang * dt.type(0.5)
    small = ang < 1e-6
    sa = np.where(small, dt.type(1), ang)
    s, c = np.sin(half), np.cos(half)
    k = np.where(small, dt.type(0.5) - ang2 / dt.type(48), s / sa)
    # dk/dang / ang  (so that dk/daa = (dk/dang/ang) * aa)
    dk_over = np.where(small, dt.type(-1.0 / 24.0), (dt.type(0.5) * c * sa - s) / (sa * sa * sa))
    qw, qv = qbar[..., :1], qbar[..., 1:]
    dot = np.sum(qv * aa, axis=-1, keepdims=True)
    # d cos(ang/2)/daa = -sin(ang/2)/(2 ang) * aa = -k/2 * aa
    return qw * (-dt.type(0.5) * k) * aa + k * qv + dk_over * dot * aa


def prior_loss_grad(params, aa, cfg, weight=1.0):
    """loss = weight * mean(d(aa->quat))  and dloss/daa  (motion_denoise.py:81-83 with the upstream
    scalar folded in by the caller)."""
    aa = np.asarray(aa).reshape(-1, NJ, 3)
    B = len(aa)
    quat = axis_angle_to_quaternion(aa)
    g_up = np.full((B, 1), weight / B, dtype=aa.dtype)
    d, qbar = forward_grad(params, quat, cfg, g_up=g_up)
    return aa.dtype.type(weight) * d.mean(), axis_angle_to_quaternion_vjp(aa, qbar), d


# ----------------------------------------------------------------------------- train-mode losses
def train_losses(params, pose, dist_gt, man_poses, cfg, loss_type="l1", eikonal=1.0):
    """model/posendf.py:78-99 (values only; parameter gradients are checked through torch autograd in
    oracle/posendf_torch.py)."""
    d, g = forward_grad(params, pose, cfg)
    dist_gt = np.asarray(dist_gt).reshape(-1).astype(d.dtype)
    diff = d[:, 0] - dist_gt
    loss = np.abs(diff).mean() if loss_type == "l1" else (diff * diff).mean()
    out = {"dist": loss}
    if eikonal > 0.0:
        d_man = forward(params, man_poses, cfg, normalise=False)
        out["man_loss"] = np.abs(d_man).mean()
        gn = np.sqrt(np.sum(g * g, axis=-1))
        out["eikonal"] = ((gn - 1) ** 2).mean()
    return loss, out


def denoise_prior(params, aa, cfg, iterations=10, steps_per_iter=50, lr=0.02, betas=(0.9, 0.999), eps=1e-8):
    """experiments/motion_denoise.py:70,74-83,97-99 restricted to the prior term: for every sequence s
    (aa: (S,T,21,3)) Adam(lr) on loss_s = 1e7/(1+it) * mean_t(dist)^2, torch.optim.Adam's update order."""
    aa = np.array(aa, copy=True)
    S, T = aa.shape[:2]
    x = aa.reshape(S, T, 21, 3)
    m = np.zeros_like(x); v = np.zeros_like(x)
    hist = []
    t = 0
    d = None
    for it in range(iterations):
        w = 1e7 / (1.0 + it)
        for _ in range(steps_per_iter):
            t += 1
            quat = axis_angle_to_quaternion(x.reshape(S * T, 21, 3))
            d, qbar = forward_grad(params, quat, cfg)
            graw = axis_angle_to_quaternion_vjp(x.reshape(S * T, 21, 3), qbar).reshape(S, T, 21, 3)
            c = d.reshape(S, T).mean(axis=1)
            hist.append(w * c * c)
            g = (w * 2.0 * c / T).reshape(S, 1, 1, 1) * graw
            m = betas[0] * m + (1 - betas[0]) * g
            v = betas[1] * v + (1 - betas[1]) * g * g
            step = lr / (1 - betas[0] ** t)
            denom = np.sqrt(v) / np.sqrt(1 - betas[1] ** t) + eps
            x = x - step * (m / denom)
    return x, d.reshape(S, T), np.array(hist)


# ----------------------------------------------------------------------------- distance-label rerank
JOINT_RANK = (7, 7, 7, 6, 6, 6, 5, 5, 5, 4, 4, 4, 4, 4, 3, 3, 3, 2, 2, 1, 1)     # data/dist_utils.py:15,37


def knn_rerank(query, database, cand_idx, metric="geo", weighted=False, k=5):
    """data/dist_utils.py:19-30 (euc) / :41-50 (geo) + torch.topk(k, largest=False): for every query pose the k
    nearest of its candidate poses.  query (Q,21,4), database (N,21,4), cand_idx (Q,K) -> (values (Q,k) ascending,
    positions inside the candidate list (Q,k))."""
    q = np.asarray(query)
    cand = np.asarray(database)[np.asarray(cand_idx)]                      # (Q,K,21,4)
    w = np.asarray(JOINT_RANK, dtype=q.dtype)
    w = w / np.maximum(np.sqrt(np.sum(w * w)), 1e-12)                      # F.normalize(joint_rank, dim=0)
    if metric == "geo":
        per_joint = 1 - np.abs(np.sum(cand * q[:, None], axis=3))          # (Q,K,21)
    elif metric == "euc":
        diff = q[:, None] - cand
        per_joint = np.sqrt(np.sum(diff * diff, axis=3))
    else:
        raise ValueError(metric)
    dis = np.sum(w * per_joint, axis=2) if weighted else np.mean(per_joint, axis=2)
    order = np.argsort(dis, axis=1, kind="stable")[:, :k]
    return np.take_along_axis(dis, order, axis=1), order


def knn_exact(query, database, metric="geo", weighted=False, k=5, chunk=64):
    """The reference's dist_calc (data/dist_utils.py:19-50) with the WHOLE database as the candidate list of every
    query: exact k nearest database poses (values (Q,k) ascending, database row indices (Q,k); ties by lower index).
    What data/prepare_traindata.py:138-170 approximates through its faiss candidate stage."""
    q = np.asarray(query)
    n = len(database)
    vals, idxs = [], []
    for c0 in range(0, len(q), chunk):
        qq = q[c0:c0 + chunk]
        cand = np.broadcast_to(np.arange(n), (len(qq), n))
        v, p = knn_rerank(qq, database, cand, metric, weighted, k)
        vals.append(v)
        idxs.append(p)
    return np.concatenate(vals), np.concatenate(idxs)


def quaternion_to_axis_angle(quat):
    """pytorch3d 0.7.2 transforms.quaternion_to_axis_angle (not vendored in /root/reference; formula as published):
    norms = |q[1:]|, half = atan2(norms, q[0]), angle = 2 half, aa = q[1:] / (sin(half)/angle), small-angle
    (|angle| < 1e-6) denominator 1/2 - angle^2/48.  PARITY UNPINNED (no copy of pytorch3d here)."""
    q = np.asarray(quat)
    dt = q.dtype
    nrm = np.sqrt(np.sum(q[..., 1:] * q[..., 1:], axis=-1, keepdims=True))
    half = np.arctan2(nrm, q[..., :1])
    ang = dt.type(2) * half
    small = np.abs(ang) < 1e-6
    with np.errstate(invalid="ignore", divide="ignore"):
        k = np.where(small, dt.type(0.5) - ang * ang / dt.type(48), np.sin(half) / np.where(small, dt.type(1), ang))
    return q[..., 1:] / k
