"""Generates tests/golden/pose6dscene.npz from the reference's own absolute-pose example data
(build/data/pose6dscene/: 96 2D-3D matches, intrinsics, ground-truth pose; examples/example_absolute_pose.ipynb), run in
the build container where /root/reference and cv2 exist:

    python oracle/gen_pnp_fixture.py

Contents: `pts` [384, 5] = the 96 matches + 3 x as many uniform outliers (the notebook's recipe, seeded), image points
K-normalised; `inlier` mask; `K`; `gt_pose`; `threshold` (2 px normalised as the notebook does); and golden vectors for the
non-minimal solver, independent of this repository's code: `epnp_lm_all` / `epnp_lm_21` = OpenCV's EPnP
(cv2.solvePnP(SOLVEPNP_EPNP), the reference's initialiser, solver_pnp_bundle_adjustment.h:150-158) on all 96 inliers / the
first 21, refined to the reprojection-error minimum by scipy.optimize.least_squares."""
import os

import cv2
import numpy as np
from scipy.optimize import least_squares
from scipy.spatial.transform import Rotation

REF = os.environ.get("GCR_REFERENCE_DIR", "/root/reference")
D = os.path.join(REF, "build", "data", "pose6dscene")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pose6dscene.npz")


def refine(pts, R0, t0):
    def res(x):
        R = Rotation.from_rotvec(x[:3]).as_matrix() @ R0
        Z = pts[:, 2:5] @ R.T + (t0 + x[3:])
        return (Z[:, :2] / Z[:, 2:] - pts[:, :2]).ravel()
    sol = least_squares(res, np.zeros(6), xtol=1e-15, ftol=1e-15, gtol=1e-15)
    return np.c_[Rotation.from_rotvec(sol.x[:3]).as_matrix() @ R0, t0 + sol.x[3:]]


def epnp_lm(pts):
    ok, rvec, tvec = cv2.solvePnP(pts[:, 2:5].copy(), pts[:, :2].copy(), np.eye(3), None, flags=cv2.SOLVEPNP_EPNP)
    assert ok
    R0, _ = cv2.Rodrigues(rvec)
    return refine(pts, R0, tvec.ravel())


def main():
    corr = np.loadtxt(os.path.join(D, "pose6dscene_points.txt"), skiprows=1)   # first line: the row count
    K = np.loadtxt(os.path.join(D, "pose6dscene.K"))
    gt = np.loadtxt(os.path.join(D, "pose6dscene_gt.txt"))
    rng = np.random.default_rng(6)
    n_out = 3 * len(corr)
    mins, maxs = corr.min(0), corr.max(0)
    mins[:2] = 0
    outliers = rng.uniform(mins, maxs, size=(n_out, 5))
    pts = np.concatenate([corr, outliers])
    inlier = np.r_[np.ones(len(corr), bool), np.zeros(n_out, bool)]
    perm = rng.permutation(len(pts))
    pts, inlier = pts[perm], inlier[perm]
    pts[:, 0] = (pts[:, 0] - K[0, 2]) / K[0, 0]
    pts[:, 1] = (pts[:, 1] - K[1, 2]) / K[1, 1]
    threshold = 2.0 / (K[0, 0] + K[1, 1]) / 2.0
    idx = np.nonzero(inlier)[0]
    np.savez_compressed(OUT, pts=pts, inlier=inlier, K=K, gt_pose=gt, threshold=threshold,
                        epnp_lm_all=epnp_lm(pts[idx]), epnp_lm_21=epnp_lm(pts[idx[:21]]))
    print("wrote", OUT, pts.shape, int(inlier.sum()))


if __name__ == "__main__":
    main()
