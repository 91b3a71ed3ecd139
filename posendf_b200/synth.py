"""Deterministic synthetic weights / poses for tests and benchmarks.

Nothing here computes the distance field; it only manufactures inputs.  The generator is a
counter-based splitmix64 hash so that the exact same bits come out on every box, numpy / torch
version and process (the GPU box has no access to the reference or to its RNG state).

Shapes follow the reference's state_dict (SURVEY Appx B):
  enc.net.{i}.net.0.weight (10, 4|10)   enc.net.{i}.net.0.bias (10)     /root/reference/model/network/net_modules.py:75-111
  enc.net.{i}.net.2.weight (6, 10)      enc.net.{i}.net.2.bias (6)
  dfnet.lin{l}.weight (out, in)         dfnet.lin{l}.bias (out)         /root/reference/model/network/net_modules.py:9-41
"""
from __future__ import annotations

import numpy as np

# /root/reference/model/network/net_utils.py:46 -- SMPL table with the root removed, indices NOT shifted.
PARENTS = (-1, -1, -1, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19)
NUM_JOINTS = 21
QUAT = 4
FEAT = 6
HID = QUAT + FEAT  # 10
AMASS_DIMS = (256, 512, 1024, 512, 256, 64)

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = x
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def uniform01(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """n doubles in [0,1), a pure function of (seed, stream, index)."""
    with np.errstate(over="ignore"):
        base = _splitmix64(np.array([seed * 0x1000003 + stream * 0x10001 + 0x5851F42D], dtype=np.uint64))[0]
        idx = np.arange(n, dtype=np.uint64)
        bits = _splitmix64((idx * np.uint64(0xD1342543DE82EF95) + base) & _M64)
    return (bits >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def normal(seed: int, n: int, stream: int = 0) -> np.ndarray:
    """Box-Muller on the hash stream (float64)."""
    m = (n + 1) // 2
    u1 = uniform01(seed, m, stream * 2 + 101)
    u2 = uniform01(seed, m, stream * 2 + 102)
    r = np.sqrt(-2.0 * np.log(1.0 - u1))
    out = np.concatenate([r * np.cos(2 * np.pi * u2), r * np.sin(2 * np.pi * u2)])
    return out[:n]


def param_shapes(in_dim: int = 126, dims=AMASS_DIMS, use_enc: bool = True):
    """Ordered (name, shape) list == the reference state_dict order (enc first, then dfnet)."""
    out = []
    if use_enc:
        for i in range(NUM_JOINTS):
            fin = QUAT if PARENTS[i] < 0 else HID
            out.append((f"enc.net.{i}.net.0.weight", (HID, fin)))
            out.append((f"enc.net.{i}.net.0.bias", (HID,)))
            out.append((f"enc.net.{i}.net.2.weight", (FEAT, HID)))
            out.append((f"enc.net.{i}.net.2.bias", (FEAT,)))
    d = [in_dim] + list(dims) + [1]
    for l in range(len(d) - 1):
        out.append((f"dfnet.lin{l}.weight", (d[l + 1], d[l])))
        out.append((f"dfnet.lin{l}.bias", (d[l + 1],)))
    return out


def make_params(seed: int, in_dim: int = 126, dims=AMASS_DIMS, use_enc: bool = True,
                sensitised: bool = True, dtype=np.float32) -> dict:
    """nn.Linear-style init U(+-1/sqrt(fan_in)) from the hash stream.

    sensitised=True applies SURVEY 8(d)'s recipe (all weights x1.6, last bias 0.5) so that d>0 and
    actually varies with the pose; with the plain default init the last pre-activation is almost
    constant and often negative, which makes d == 0 everywhere."""
    params = {}
    for k, (name, shape) in enumerate(param_shapes(in_dim, dims, use_enc)):
        fan_in = shape[1] if len(shape) == 2 else None
        if fan_in is None:
            # bias: bound from the matching weight's fan_in (previous entry)
            fan_in = prev_fan_in
        prev_fan_in = fan_in
        bound = 1.0 / np.sqrt(fan_in)
        n = int(np.prod(shape))
        v = (uniform01(seed, n, stream=k) * 2.0 - 1.0) * bound
        if sensitised and name.endswith("weight"):
            v = v * 1.6
        params[name] = v.reshape(shape).astype(dtype)
    if sensitised:
        last = f"dfnet.lin{len(dims)}.bias"
        params[last] = np.full(params[last].shape, 0.5, dtype=dtype)
    return params


def make_poses(seed: int, batch: int, kind: str = "randn", sigma: float = 0.1, dtype=np.float32) -> np.ndarray:
    """(batch,21,4) poses.

    randn  : normal then per-quaternion normalise (experiments/sample_poses.py:96-97 does this with rand)
    rand   : uniform[0,1) then per-quaternion normalise (exactly sample_poses.py:96-97)
    noisy  : per-quaternion-normalised base + sigma * uniform noise, renormalised (data/create_data.py:51,89-90)
    raw    : plain normal, NOT normalised (exercises the column normalisation with arbitrary scale)
    """
    n = batch * NUM_JOINTS * QUAT
    if kind == "rand":
        x = uniform01(seed, n, stream=7).reshape(batch, NUM_JOINTS, QUAT)
    else:
        x = normal(seed, n, stream=7).reshape(batch, NUM_JOINTS, QUAT)
    if kind == "raw":
        return x.astype(dtype)
    x = x / np.linalg.norm(x, axis=2, keepdims=True)
    if kind == "noisy":
        x = x + sigma * uniform01(seed, n, stream=8).reshape(x.shape)
        x = x / np.linalg.norm(x, axis=2, keepdims=True)
    return x.astype(dtype)


def make_axis_angle(seed: int, batch: int, std: float = 0.3, dtype=np.float32) -> np.ndarray:
    """(batch,21,3) axis-angle poses ~ N(0, std^2)  (SURVEY 8(d) config C4)."""
    return (normal(seed, batch * NUM_JOINTS * 3, stream=9) * std).reshape(batch, NUM_JOINTS, 3).astype(dtype)


def flat_param_count(in_dim: int = 126, dims=AMASS_DIMS, use_enc: bool = True) -> int:
    return sum(int(np.prod(s)) for _, s in param_shapes(in_dim, dims, use_enc))


def flatten_params(params: dict, in_dim: int = 126, dims=AMASS_DIMS, use_enc: bool = True) -> np.ndarray:
    """Canonical flat fp32 vector handed to pndf_set_weights (state_dict order, row-major)."""
    return np.concatenate([np.asarray(params[n], dtype=np.float32).reshape(-1)
                           for n, _ in param_shapes(in_dim, dims, use_enc)])
