"""Generates tests/golden/fountain_matches.npz: the SIFT tentative correspondences of the reference's fountain pair
(build/data/fountain/fountain1.jpg <-> fountain2.jpg, intrinsics fountain{1,2}.K), extracted exactly as the reference's
tests/test_essential_matrix.py:62-83 does (cv2.SIFT_create(8000), BF kNN, SNN ratio 0.8, sorted by ratio, float32 N x 4).

Run here (needs /root/reference and cv2):  python oracle/gen_fountain_fixture.py
"""
import os

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = os.environ.get("GCR_REFERENCE_DIR", "/root/reference")


def main():
    d = os.path.join(REF, "build/data/fountain")
    img1 = cv2.cvtColor(cv2.imread(os.path.join(d, "fountain1.jpg")), cv2.COLOR_BGR2RGB)
    img2 = cv2.cvtColor(cv2.imread(os.path.join(d, "fountain2.jpg")), cv2.COLOR_BGR2RGB)
    K1, K2 = np.loadtxt(os.path.join(d, "fountain1.K")), np.loadtxt(os.path.join(d, "fountain2.K"))
    det = cv2.SIFT_create(8000)
    k1, d1 = det.detectAndCompute(img1, None)
    k2, d2 = det.detectAndCompute(img2, None)
    matches = cv2.BFMatcher().knnMatch(d1, d2, k=2)
    good = [(a.distance / b.distance, a) for a, b in matches if a.distance < 0.8 * b.distance]
    order = np.argsort([g[0] for g in good])
    pts = np.float32([list(k1[good[i][1].queryIdx].pt) + list(k2[good[i][1].trainIdx].pt) for i in order])
    print("correspondences", pts.shape, "image sizes", img1.shape, img2.shape)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "fountain_matches.npz"), pts=pts, K1=K1, K2=K2,
                        size=np.array([img1.shape[0], img1.shape[1], img2.shape[0], img2.shape[1]]))


if __name__ == "__main__":
    main()
