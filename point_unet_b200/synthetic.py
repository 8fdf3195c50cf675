"""Synthetic Pancreas- / BraTS-shaped point clouds (there is no network for real data).

Shapes and construction mirror how the reference turns a volume into a cloud:

* Pancreas (``utils/dataPreparePancreas.py:132-169``, ``runPancreas.py:96-114``): all foreground
  voxels first, then background voxels sampled without replacement up to ``n_points``, NOT
  shuffled; ``xyz = uint16 voxel.astype(f32) / shape.astype(f32)``; one intensity channel.
* BraTS (``utils/dataPrepareBraTS.py:75-116``, ``runBraTS.py:100-119``): brain voxels only; all
  tumour voxels plus randomly drawn non-tumour voxels up to ``n_points``, then shuffled;
  ``xyz = (voxel_f64 / shape).astype(f32)``; four modality channels; labels {0,1,2,3}.

Everything here is host-side numpy with explicit seeds so the oracle and the CUDA path see the
same bytes.
"""
from __future__ import annotations

import numpy as np

PANCREAS_SHAPE = (512, 512, 240)
BRATS_SHAPE = (240, 240, 155)


def _ellipsoid_voxels(center, semi_axes, shape):
    """Integer voxel coordinates [M,3] inside an axis-aligned ellipsoid, x-major order."""
    c = np.asarray(center, dtype=np.float64)
    r = np.asarray(semi_axes, dtype=np.float64)
    lo = np.maximum(np.floor(c - r).astype(np.int64), 0)
    hi = np.minimum(np.ceil(c + r).astype(np.int64) + 1, np.asarray(shape))
    gx, gy, gz = np.meshgrid(np.arange(lo[0], hi[0]), np.arange(lo[1], hi[1]), np.arange(lo[2], hi[2]),
                             indexing="ij")
    vox = np.stack([gx.ravel(), gy.ravel(), gz.ravel()], axis=1)
    inside = (((vox - c) / r) ** 2).sum(axis=1) <= 1.0
    return vox[inside]


def _sample_excluding(rng, shape, n, exclude_keys):
    """n distinct voxels drawn uniformly from the volume, none in ``exclude_keys`` (linearised)."""
    total = int(shape[0]) * int(shape[1]) * int(shape[2])
    got = np.empty(0, dtype=np.int64)
    while got.size < n:
        cand = rng.integers(0, total, size=int((n - got.size) * 1.2) + 64, dtype=np.int64)
        cand = cand[~np.isin(cand, exclude_keys)]
        got = np.unique(np.concatenate([got, cand]))  # unique => without replacement
    got = rng.permutation(got)[:n]
    x = got // (shape[1] * shape[2])
    y = (got // shape[2]) % shape[1]
    z = got % shape[2]
    return np.stack([x, y, z], axis=1)


def pancreas_cloud(n_points: int = 180000, seed: int = 0, shape=PANCREAS_SHAPE, max_foreground: int = 60000):
    """One Pancreas-shaped cloud.

    Returns dict(xyz f32 [N,3], features f32 [N,1], labels int32 [N], xyz_origin uint16 [N,3]).
    Foreground-first, unshuffled (``dataPreparePancreas.py:154-159``).
    """
    rng = np.random.default_rng(seed)
    shape_a = np.asarray(shape, dtype=np.int64)
    # scale the organ with the requested cloud size so small test clouds keep the fg/bg mix
    scale = min(1.0, (n_points / 180000.0) ** (1.0 / 3.0))
    semi = np.array([40.0, 25.0, 20.0]) * scale
    center = shape_a * np.array([0.45, 0.55, 0.5]) + rng.uniform(-10, 10, size=3)
    fg = _ellipsoid_voxels(center, semi, shape)
    cap = min(max_foreground, n_points // 3)
    if fg.shape[0] > cap:
        fg = fg[np.sort(rng.choice(fg.shape[0], size=cap, replace=False))]
    fg_keys = (fg[:, 0] * shape[1] + fg[:, 1]) * shape[2] + fg[:, 2]
    bg = _sample_excluding(rng, shape, n_points - fg.shape[0], fg_keys)
    vox = np.concatenate([fg, bg], axis=0).astype(np.uint16)
    xyz = vox.astype(np.float32) / shape_a.astype(np.float32)  # fp32 division, dataPreparePancreas.py:163
    labels = np.zeros(n_points, dtype=np.int32)
    labels[: fg.shape[0]] = 1
    feats = rng.standard_normal((n_points, 1)).astype(np.float32)
    return dict(xyz=xyz, features=feats, labels=labels, xyz_origin=vox)


def brats_cloud(n_points: int = 180000, seed: int = 0, shape=BRATS_SHAPE):
    """One BraTS-shaped cloud: all tumour voxels + random brain voxels, shuffled (``runBraTS.py:107-114``)."""
    rng = np.random.default_rng(seed)
    shape_a = np.asarray(shape, dtype=np.int64)
    scale = min(1.0, (n_points / 180000.0) ** (1.0 / 3.0))
    brain_c = shape_a * 0.5
    brain_r = shape_a * np.array([0.36, 0.42, 0.40])
    tum_c = brain_c + rng.uniform(-0.15, 0.15, size=3) * shape_a
    tum_r = np.array([26.0, 22.0, 18.0]) * scale
    tum = _ellipsoid_voxels(tum_c, tum_r, shape)
    inside_brain = (((tum - brain_c) / brain_r) ** 2).sum(axis=1) <= 1.0
    tum = tum[inside_brain]
    if tum.shape[0] > n_points // 2:
        tum = tum[np.sort(rng.choice(tum.shape[0], size=n_points // 2, replace=False))]
    # nested labels: 3 (core) inside 1 inside 2 (edema), by normalised radius
    rad = np.sqrt((((tum - tum_c) / tum_r) ** 2).sum(axis=1))
    tum_lab = np.where(rad < 0.45, 3, np.where(rad < 0.75, 1, 2)).astype(np.int32)
    tum_keys = (tum[:, 0] * shape[1] + tum[:, 1]) * shape[2] + tum[:, 2]
    # non-tumour brain voxels, uniform inside the brain ellipsoid
    need = n_points - tum.shape[0]
    got = np.empty((0, 3), dtype=np.int64)
    while got.shape[0] < need:
        cand = _sample_excluding(rng, shape, int(need * 2.2) + 64, tum_keys)
        ok = (((cand - brain_c) / brain_r) ** 2).sum(axis=1) <= 1.0
        got = np.unique(np.concatenate([got, cand[ok]]), axis=0)
    got = rng.permutation(got)[:need]
    vox = np.concatenate([tum, got], axis=0)
    labels = np.concatenate([tum_lab, np.zeros(need, dtype=np.int32)])
    perm = rng.permutation(n_points)  # DP.shuffle_idx, runBraTS.py:114
    vox, labels = vox[perm], labels[perm]
    xyz = (vox.astype(np.float64) / shape_a.astype(np.float64)).astype(np.float32)  # dataPrepareBraTS.py:85-89
    feats = rng.standard_normal((n_points, 4)).astype(np.float32)
    return dict(xyz=xyz, features=feats, labels=labels, xyz_origin=vox.astype(np.int32))


def uniform_cloud(n_points: int, seed: int | None = None):
    """``rng.random((N,3), float32)`` -- mirrors nearest_neighbors/test.py:8; tie-free in practice."""
    rng = np.random.default_rng(n_points if seed is None else seed)
    return rng.random((n_points, 3), dtype=np.float32)


def jittered_lattice_cloud(n_points: int, seed: int = 0, shape=PANCREAS_SHAPE):
    """Pancreas-shaped density, voxel + U(-0.45,0.45) jitter: same distribution, tie-free by construction."""
    c = pancreas_cloud(n_points, seed, shape)
    rng = np.random.default_rng(seed + 7919)
    vox = c["xyz_origin"].astype(np.float64) + rng.uniform(-0.45, 0.45, size=(n_points, 3))
    return (vox / np.asarray(shape, dtype=np.float64)).astype(np.float32)


def batch(fn, batch_size: int, n_points: int, seed0: int = 0):
    """Stack ``batch_size`` clouds made by ``fn`` (seeds seed0, seed0+1, ...) into [B,N,*] arrays."""
    cs = [fn(n_points, seed0 + i) for i in range(batch_size)]
    return {k: np.stack([c[k] for c in cs], axis=0) for k in cs[0]}
