"""Synthetic correspondences shared by the tests, the golden-vector generator and bench.py.

Image size 1024 x 1024, f = 1000, principal point (512, 512); inliers follow a random model with Gaussian noise
(sigma = 0.5 px), outliers are uniform in both images; rows are shuffled.  Everything is numpy / FP64, seeded.
"""
import numpy as np

SIZE = 1024.0
FOCAL = 1000.0
K_MAT = np.array([[FOCAL, 0, SIZE / 2], [0, FOCAL, SIZE / 2], [0, 0, 1.0]])


def _rotation(rng, max_deg=20.0):
    axis = rng.normal(size=3)
    axis /= np.linalg.norm(axis)
    ang = np.deg2rad(rng.uniform(-max_deg, max_deg))
    Kx = np.array([[0, -axis[2], axis[1]], [axis[2], 0, -axis[0]], [-axis[1], axis[0], 0]])
    return np.eye(3) + np.sin(ang) * Kx + (1 - np.cos(ang)) * Kx @ Kx


def random_homography(rng):
    R = _rotation(rng)
    t = rng.normal(size=3) * 0.2
    nrm = np.array([0, 0, 1.0]) + rng.normal(size=3) * 0.1
    nrm /= np.linalg.norm(nrm)
    d = rng.uniform(4, 8)
    Hn = R + np.outer(t, nrm) / d
    Hm = K_MAT @ Hn @ np.linalg.inv(K_MAT)
    return Hm / Hm[2, 2]


def homography_pair(n, outlier_ratio, seed, sigma=0.5):
    rng = np.random.default_rng(seed)
    Hm = random_homography(rng)
    x1 = rng.uniform(0, SIZE, size=(n, 2))
    p = np.c_[x1, np.ones(n)] @ Hm.T
    x2 = p[:, :2] / p[:, 2:] + rng.normal(0, sigma, size=(n, 2))
    n_out = int(round(n * outlier_ratio))
    out = rng.permutation(n)[:n_out]
    x1[out] = rng.uniform(0, SIZE, size=(n_out, 2))
    x2[out] = rng.uniform(0, SIZE, size=(n_out, 2))
    perm = rng.permutation(n)
    pts = np.ascontiguousarray(np.c_[x1, x2][perm])
    inl = np.ones(n, bool)
    inl[out] = False
    return pts, Hm, inl[perm]


def relative_pose(rng):
    R = _rotation(rng)
    t = rng.normal(size=3)
    t /= np.linalg.norm(t)
    return R, t


def _skew(t):
    return np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])


def epipolar_pair(n, outlier_ratio, seed, sigma=0.5):
    """Returns pixel correspondences, F (x2^T F x1 = 0), E, K, inlier mask."""
    rng = np.random.default_rng(seed)
    R, t = relative_pose(rng)
    # 3-D points in front of both cameras
    x1 = rng.uniform(0, SIZE, size=(n, 2))
    depth = rng.uniform(4, 8, size=n)
    rays = np.c_[x1, np.ones(n)] @ np.linalg.inv(K_MAT).T
    X = rays * depth[:, None]
    X2 = X @ R.T + t
    p2 = X2 @ K_MAT.T
    x2 = p2[:, :2] / p2[:, 2:] + rng.normal(0, sigma, size=(n, 2))
    n_out = int(round(n * outlier_ratio))
    out = rng.permutation(n)[:n_out]
    x1[out] = rng.uniform(0, SIZE, size=(n_out, 2))
    x2[out] = rng.uniform(0, SIZE, size=(n_out, 2))
    perm = rng.permutation(n)
    pts = np.ascontiguousarray(np.c_[x1, x2][perm])
    Em = _skew(t) @ R
    Ki = np.linalg.inv(K_MAT)
    Fm = Ki.T @ Em @ Ki
    inl = np.ones(n, bool)
    inl[out] = False
    return pts, Fm / Fm[2, 2], Em, K_MAT.copy(), inl[perm]


def epipolar_pair_planar(n, outlier_ratio, plane_fraction, seed, sigma=0.5):
    """As epipolar_pair, but `plane_fraction` of the inliers lie on ONE scene plane (a dominant plane: the case
    DEGENSAC, fundamental_estimator.h:341-572, exists for).  Returns pixel correspondences, F, inlier mask, on-plane mask."""
    rng = np.random.default_rng(seed)
    R, t = relative_pose(rng)
    x1 = rng.uniform(0, SIZE, size=(n, 2))
    rays = np.c_[x1, np.ones(n)] @ np.linalg.inv(K_MAT).T
    depth = rng.uniform(4, 8, size=n)
    on_plane = rng.uniform(size=n) < plane_fraction
    nrm = np.array([0.15, -0.1, 1.0])
    nrm /= np.linalg.norm(nrm)
    depth[on_plane] = 6.0 / (rays[on_plane] @ nrm)                 # n . (depth * ray) = 6
    X = rays * depth[:, None]
    X2 = X @ R.T + t
    p2 = X2 @ K_MAT.T
    x2 = p2[:, :2] / p2[:, 2:] + rng.normal(0, sigma, size=(n, 2))
    n_out = int(round(n * outlier_ratio))
    out = rng.permutation(n)[:n_out]
    x1[out] = rng.uniform(0, SIZE, size=(n_out, 2))
    x2[out] = rng.uniform(0, SIZE, size=(n_out, 2))
    perm = rng.permutation(n)
    pts = np.ascontiguousarray(np.c_[x1, x2][perm])
    Ki = np.linalg.inv(K_MAT)
    Fm = Ki.T @ (_skew(t) @ R) @ Ki
    inl = np.ones(n, bool)
    inl[out] = False
    on_plane[out] = False
    return pts, Fm / Fm[2, 2], inl[perm], on_plane[perm]


def pnp_scene(n, outlier_ratio, seed, sigma=0.5 / FOCAL):
    """Rows u v X Y Z with (u, v) K-normalised; returns points, P = [R|t] (3x4), inlier mask."""
    rng = np.random.default_rng(seed)
    R, t = relative_pose(rng)
    t = t * 0.5 + np.array([0, 0, 6.0])
    X = rng.uniform(-2, 2, size=(n, 3))
    pc = X @ R.T + t
    uv = pc[:, :2] / pc[:, 2:] + rng.normal(0, sigma, size=(n, 2))
    n_out = int(round(n * outlier_ratio))
    out = rng.permutation(n)[:n_out]
    uv[out] = rng.uniform(-0.5, 0.5, size=(n_out, 2))
    perm = rng.permutation(n)
    pts = np.ascontiguousarray(np.c_[uv, X][perm])
    inl = np.ones(n, bool)
    inl[out] = False
    return pts, np.c_[R, t], inl[perm]


def perturbed_models(model, count, rng, rel=2e-3):
    """`count` models around `model` (entry-wise relative perturbation) plus pure-noise ones: a mix of good and bad
    hypotheses like a RANSAC batch."""
    model = np.asarray(model, dtype=np.float64)
    out = np.empty((count,) + model.shape)
    scale = np.abs(model) + np.abs(model).mean() * 1e-3
    for i in range(count):
        if i % 4 == 3:
            out[i] = model + rng.normal(size=model.shape) * scale * 0.3
        else:
            out[i] = model + rng.normal(size=model.shape) * scale * rel * rng.uniform(0, 1)
    out[0] = model
    return out
