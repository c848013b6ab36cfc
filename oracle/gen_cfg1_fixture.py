"""Generates tests/golden/cfg1_matches.npz: the SIFT tentative correspondences of the reference's bundled pair
tests/test_data/004520.png <-> 004523.png (BASELINE.json configs[0]), extracted exactly as the reference's own tests do
(tests/test_homography.py:52-69: cv2.SIFT_create(8000), BF kNN, SNN ratio 0.8, sorted by ratio, float32 N x 4), plus the
result of the reference's GCRANSAC<>::run on them (oracle/_ref, verbatim driver + restated estimators):
findHomography settings threshold 3.0, conf 0.99, max_iters 10000, min_iters 50, spatial_coherence_weight 0.975,
neighborhood_size 4, lo_number 50, use_sprt False, use_space_partitioning False.

Run here (needs /root/reference and cv2):  python oracle/gen_cfg1_fixture.py
"""
import os
import sys

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
import oracle as O  # noqa: E402

REF = os.environ.get("GCR_REFERENCE_DIR", "/root/reference")


def main():
    img1 = cv2.imread(os.path.join(REF, "tests/test_data/004520.png"))
    img2 = cv2.imread(os.path.join(REF, "tests/test_data/004523.png"))
    det = cv2.SIFT_create(8000)
    k1, d1 = det.detectAndCompute(cv2.cvtColor(img1, cv2.COLOR_BGR2GRAY), None)
    k2, d2 = det.detectAndCompute(cv2.cvtColor(img2, cv2.COLOR_BGR2GRAY), None)
    matches = cv2.BFMatcher().knnMatch(d1, d2, k=2)
    good = sorted(((a.distance / b.distance, a) for a, b in matches if a.distance < 0.8 * b.distance), key=lambda x: x[0])
    pts = np.float32([list(k1[a.queryIdx].pt) + list(k2[a.trainIdx].pt) for _, a in good])
    h1, w1 = img1.shape[:2]
    h2, w2 = img2.shape[:2]
    ref = O.Ref()
    cells = 4
    eps = np.finfo(np.float64).eps
    cs = np.array([(w1 + eps) / cells, (h1 + eps) / cells, (w2 + eps) / cells, (h2 + eps) / cells])  # gcransac_python.cpp:4290-4297
    runs = {}
    for seed in (1, 2, 3):
        r = ref.run(O.H, pts.astype(np.float64), cs, cells, 3.0, conf=0.99, spatial_weight=0.975, max_iters=10000,
                    min_iters=50, lo_number=50, seed=seed)
        runs["model_%d" % seed] = r["model"]
        runs["mask_%d" % seed] = r["mask"]
        runs["stats_%d" % seed] = np.array(list(r["stats"].values()), dtype=np.uint64)
        print("seed", seed, "inliers", r["inliers"], r["stats"])
        # BASELINE configs[0] passes sampler=1: PROSAC over the SNN-ratio order (py/gcransac_python.cpp:4320-4323)
        r = ref.run(O.H, pts.astype(np.float64), cs, cells, 3.0, conf=0.99, spatial_weight=0.975, max_iters=10000,
                    min_iters=50, lo_number=50, seed=seed, sampler=1)
        runs["prosac_model_%d" % seed] = r["model"]
        runs["prosac_mask_%d" % seed] = r["mask"]
        runs["prosac_stats_%d" % seed] = np.array(list(r["stats"].values()), dtype=np.uint64)
        print("seed", seed, "PROSAC inliers", r["inliers"], r["stats"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "cfg1_matches.npz"), pts=pts, size=np.array([h1, w1, h2, w2]),
                        cell_sizes=cs, **runs)


if __name__ == "__main__":
    main()
