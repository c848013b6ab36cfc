"""Pins the oracle solvers that are Eigen-dependent in the reference against independent numpy/scipy statements of the same
problems (tests/test_reference_solver_sources.py holds them against the reference's own source).  CPU only."""
import numpy as np
import oracle as O
from pygcransac import synth

from np_five_point import five_point


def _norm(E):
    E = E / np.linalg.norm(E)
    k = np.argmax(np.abs(E))
    return E * np.sign(E.flat[k])


def test_e5_nister_constraints_and_ground_truth(oracle):
    pts, Fm, Em, K, inl = synth.epipolar_pair(400, 0.0, seed=3, sigma=0.0)
    npts = oracle.normalize_correspondences(pts, K, K)
    # utils.h:65-110: K^-1 applied to both images
    want = np.c_[(np.c_[pts[:, :2], np.ones(400)] @ np.linalg.inv(K).T)[:, :2], (np.c_[pts[:, 2:], np.ones(400)] @ np.linalg.inv(K).T)[:, :2]]
    assert np.allclose(npts, want, rtol=1e-13, atol=1e-15)
    rng = np.random.default_rng(0)
    En = _norm(Em)
    good = total = recovered = 0
    for t in range(150):
        s = rng.choice(400, 5, replace=False).astype(np.uint32)
        Es = oracle.solve_e5(npts, s)
        assert len(Es) <= 10
        x1 = np.c_[npts[s, :2], np.ones(5)]
        x2 = np.c_[npts[s, 2:], np.ones(5)]
        ok_any = False
        for E in Es:
            total += 1
            # any combination of the null vectors meets the epipolar constraints exactly
            assert np.abs(np.einsum("ij,jk,ik->i", x2, E, x1)).max() < 1e-9
            assert abs(np.linalg.norm(E) - 1.0) < 1.0   # e is normalised by sqrt(x^2 + y^2 + z^2 + 1), not Frobenius
            En_ = E / np.linalg.norm(E)
            tr = np.abs(2 * En_ @ En_.T @ En_ - np.trace(En_ @ En_.T) * En_).max()
            good += tr < 1e-6
            ok_any |= np.abs(_norm(E) - En).max() < 1e-6
        recovered += ok_any
    # The FullPivLU kernel basis of the reference (:324-330) is not orthonormal: for some samples the degree-10
    # polynomial has clustered roots and the solutions lose accuracy -- inherited behaviour, counted here.
    assert good > 0.85 * total and recovered > 0.75 * 150, (good, total, recovered)


def test_e5_matches_numpy_on_well_conditioned_samples(oracle):
    pts, Fm, Em, K, inl = synth.epipolar_pair(400, 0.0, seed=4, sigma=0.3)
    npts = oracle.normalize_correspondences(pts, K, K)
    rng = np.random.default_rng(1)
    compared = 0
    for t in range(60):
        s = rng.choice(400, 5, replace=False).astype(np.uint32)
        mine = [_norm(E) for E in oracle.solve_e5(npts, s)]
        ref, _ = five_point(npts[s])
        ref = [_norm(E) for E in ref]
        if len(mine) != len(ref):
            continue  # root-count disagreements are confined to near-multiple roots
        errs = [min(np.abs(m - r).max() for r in ref) for m in mine]
        if errs and max(errs) < 1e-7:
            compared += 1
    assert compared >= 40, compared


def test_e5_orientation_filter_is_a_subset(oracle):
    pts, Fm, Em, K, inl = synth.epipolar_pair(300, 0.3, seed=5)
    npts = oracle.normalize_correspondences(pts, K, K)
    s_all = oracle.sample_uniform(3, 0, 200, 300, 5)
    kept = total = 0
    for s in s_all:
        a, b = oracle.solve_e5(npts, s), oracle.estimate_e5(npts, s)
        total += len(a)
        kept += len(b)
        for E in b:
            assert any(np.array_equal(E, A) for A in a)
    assert 0 < kept < total


def test_f7_against_numpy_null_space(oracle):
    pts, Fm, Em, K, inl = synth.epipolar_pair(300, 0.0, seed=6, sigma=0.0)
    rng = np.random.default_rng(2)
    Fn = Fm / np.linalg.norm(Fm)
    hits = 0
    for t in range(50):
        s = rng.choice(300, 7, replace=False).astype(np.uint32)
        Fs = oracle.solve_f7(pts, s)
        assert 1 <= len(Fs) <= 3
        x1 = np.c_[pts[s, :2], np.ones(7)]
        x2 = np.c_[pts[s, 2:], np.ones(7)]
        best = 1.0
        for F in Fs:
            assert np.abs(np.einsum("ij,jk,ik->i", x2, F, x1)).max() < 1e-6 * np.abs(F).max() * 1e6
            assert abs(np.linalg.det(F / np.linalg.norm(F))) < 1e-10
            Fq = F / np.linalg.norm(F)
            best = min(best, np.abs(Fq - Fn).max(), np.abs(Fq + Fn).max())
        hits += best < 1e-7
    assert hits >= 48


def test_p3p_lambda_twist(oracle):
    pts, P, inl = synth.pnp_scene(300, 0.0, seed=3, sigma=0.0)
    rng = np.random.default_rng(0)
    hit = total = 0
    for t in range(200):
        s = rng.choice(300, 3, replace=False).astype(np.uint32)
        Ps = oracle.solve_p3p(pts, s)
        assert 1 <= len(Ps) <= 4
        best = 1.0
        for Q in Ps:
            total += 1
            R = Q[:, :3]
            assert np.abs(R @ R.T - np.eye(3)).max() < 1e-8 and np.linalg.det(R) > 0.99
            pc = pts[s, 2:] @ R.T + Q[:, 3]
            assert np.abs(pc[:, :2] / pc[:, 2:] - pts[s, :2]).max() < 1e-8      # the three sample points reproject exactly
            assert (pc[:, 2] > 0).all()                                         # lambda_i > 0: in front of the camera
            best = min(best, np.abs(Q - P).max())
        hit += best < 1e-7
    assert hit == 200 and total > 300


def test_dense_kernels_against_numpy(oracle):
    """oracle/linalg.h: non-symmetric eigenvalues (Hessenberg + Francis QR), eigenvectors by inverse iteration, and the
    Eigen-style column-pivoted Householder Q -- the stand-ins for EigenSolver / ColPivHouseholderQR in the Stewenius
    solver -- against numpy.linalg.eig and LAPACK dgeqp3 (scipy)."""
    from scipy.linalg import qr
    rng = np.random.default_rng(0)
    for t in range(100):
        A = rng.standard_normal((10, 10))
        ok, w = oracle.eigenvalues(A)
        assert ok
        assert np.allclose(np.sort_complex(w), np.sort_complex(np.linalg.eigvals(A)), atol=1e-9)
        for lam in w[w.imag == 0].real:
            v = oracle.eigenvector(A, lam)
            assert abs(np.linalg.norm(v) - 1) < 1e-12 and np.linalg.norm(A @ v - lam * v) < 1e-9 * np.linalg.norm(A)
    for t in range(50):
        B = rng.standard_normal((30, 9))
        A = B.T @ B
        Q = oracle.colpiv_q(A)
        Qs = qr(A, pivoting=True)[0]
        assert np.allclose(Q.T @ Q, np.eye(9), atol=1e-12)
        assert np.allclose(Q[:, 5:] @ Q[:, 5:].T, Qs[:, 5:] @ Qs[:, 5:].T, atol=1e-9)   # the same 4-D complement


def test_refit_e_stewenius_against_numpy(oracle):
    """EssentialMatrixEstimator::estimateModelNonminimal (Stewenius action matrix) against tests/np_five_point.stewenius:
    same model count, every model within 1e-6 of a numpy one; recovers the ground truth from noisy inliers."""
    from np_five_point import stewenius
    pts, Fm, Em, K, inl = synth.epipolar_pair(2048, 0.5, seed=3)
    npts = oracle.normalize_correspondences(pts, K, K)
    rng = np.random.default_rng(1)
    inl_idx = np.flatnonzero(inl)
    En = Em / np.linalg.norm(Em)
    En = -En if En[2, 2] < 0 else En
    recovered = 0
    for t in range(60):
        n = 35 if t % 2 == 0 else int(rng.integers(6, 200))
        pool = inl_idx if t % 4 else np.arange(len(pts))
        s = rng.choice(pool, size=n, replace=False).astype(np.uint32)
        w = None if t % 3 else rng.uniform(0.1, 1, len(pts))
        got = oracle.refit_e(npts, s, w)
        want = stewenius(npts[s], None if w is None else w[s])      # the solver reads weights[sample[i]]
        assert len(got) == len(want) and 1 <= len(got) <= 10
        for E in got:
            assert abs(np.linalg.norm(E) - 1) < 1e-12 and E[2, 2] >= 0
            assert min(np.abs(E - q).max() for q in want) < 1e-6
        if t % 4 and n >= 35:
            recovered += min(np.abs(E - En).max() for E in got) < 2e-2
    assert recovered >= 20
    assert len(oracle.refit_e(npts, inl_idx[:4].astype(np.uint32))) == 0       # below nonMinimalSampleSize()
    # n == 5 takes the LU null space (stewenius.h:175-180): same models as the minimal solver up to normalisation
    s5 = inl_idx[:5].astype(np.uint32)
    a = oracle.refit_e(npts, s5)
    b = oracle.solve_e5(npts, s5)
    assert len(a) == len(b)
    for E in a:
        assert min(np.abs(_norm(E) - _norm(q)).max() for q in b) < 1e-6


def _skew(v):
    return np.array([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def test_plane_parallax_solver_against_numpy(oracle):
    """solver_fundamental_matrix_plane_and_parallax.h:109-160 stated with numpy: e = (Hx1 x x1') x (Hx2 x x2'), F = [e]x H.
    With the true plane homography and two off-plane points the result is the true F (up to scale)."""
    pts, Fm, inl, on_plane = synth.epipolar_pair_planar(400, 0.0, 0.5, seed=21, sigma=0.0)
    plane_idx = np.flatnonzero(on_plane)[:40].astype(np.uint32)
    Hm = oracle.refit(O.H, pts, plane_idx)                      # exact plane homography (noise-free points)
    off = np.flatnonzero(~on_plane)
    for a, b in ((off[0], off[1]), (off[5], off[9]), (off[2], off[30])):
        got = oracle.solve_f_plane_parallax(pts, [a, b], Hm)
        l = [np.cross(Hm @ np.r_[pts[i, :2], 1], np.r_[pts[i, 2:], 1]) for i in (a, b)]
        want = _skew(np.cross(l[0], l[1])) @ Hm
        assert np.abs(got - want).max() <= 1e-12 * np.abs(want).max()
        g, w = got / np.linalg.norm(got), Fm / np.linalg.norm(Fm)
        assert min(np.abs(g - w).max(), np.abs(g + w).max()) < 1e-6
    # two points ON the plane: H x = x', both lines vanish, no epipole -> no model


def test_degensac_sample_test_against_numpy(oracle):
    """fundamental_estimator.h:341-500 stated with numpy (SVD epipole, H = A - e (M^-1 b)^T): a 7-point sample with
    >= 5 points on one plane is H-degenerate and the implied homography maps the plane's points; a general sample is not."""
    pts, Fm, inl, on_plane = synth.epipolar_pair_planar(600, 0.0, 0.5, seed=22, sigma=0.05)
    P, Q = np.flatnonzero(on_plane), np.flatnonzero(~on_plane)
    rng = np.random.default_rng(3)
    hits = 0
    for trial in range(40):
        k = [7, 6, 5, 2, 0][trial % 5]                          # points of the sample on the plane
        sample = np.r_[rng.choice(P, k, replace=False), rng.choice(Q, 7 - k, replace=False)].astype(np.uint32)
        rng.shuffle(sample)
        models = oracle.solve_f7(pts, sample)
        if len(models) == 0:
            continue
        F = models[trial % len(models)]
        got = oracle.degensac_sample_homography(F, pts, sample)
        # numpy statement
        U = np.linalg.svd(F)[0]
        e = U[:, 2] / U[2, 2]
        A = _skew(e) @ F
        want = None
        for t in ((0, 1, 2), (3, 4, 5), (0, 1, 6), (3, 4, 6), (2, 5, 6)):
            idx = sample[list(t)]
            x1 = np.c_[pts[idx, :2], np.ones(3)]
            x2 = np.c_[pts[idx, 2:], np.ones(3)]
            c = [np.cross(x2[q], e) for q in range(3)]
            b = np.array([np.cross(x2[q], A @ x1[q]) @ c[q] / (c[q] @ c[q]) for q in range(3)])
            Hm = A - np.outer(e, np.linalg.inv(x1) @ b)
            others = [i for i in sample if i not in idx]
            proj = np.c_[pts[others, :2], np.ones(len(others))] @ Hm.T
            r2 = ((proj[:, :2] / proj[:, 2:] - pts[others, 2:]) ** 2).sum(1)
            if 3 + (r2 < 4.0).sum() >= 5:
                want = Hm
                break
        assert (got is None) == (want is None), (trial, k)
        if got is not None:
            hits += 1
            assert np.abs(got - want).max() <= 1e-6 * np.abs(want).max()
    assert hits >= 8


def _sampson_residuals(F, p):
    x1 = np.c_[p[:, :2], np.ones(len(p))]
    x2 = np.c_[p[:, 2:], np.ones(len(p))]
    Fx, Ftx = x1 @ F.T, x2 @ F
    return (x2 * Fx).sum(1) / np.sqrt(Fx[:, 0] ** 2 + Fx[:, 1] ** 2 + Ftx[:, 0] ** 2 + Ftx[:, 1] ** 2)


def test_f_bundle_adjustment_against_scipy(oracle):
    """relative_pose/bundle.cpp:253-339 (lm_F_impl) + jacobian_impl.h:472-606 restated; pinned against an independent
    scipy least-squares minimisation of the same Sampson cost over rank-2 matrices U diag(1, s, 0) V^T."""
    from scipy.optimize import least_squares
    from scipy.spatial.transform import Rotation
    pts, Fm, Em, Kc, _ = synth.epipolar_pair(400, 0.0, seed=31, sigma=0.8)
    idx = np.arange(150, dtype=np.uint32)
    npts, T1, T2 = oracle.normalize_points(pts, idx)
    F0 = np.linalg.inv(T2).T @ oracle.refit(O.F, pts, idx) @ np.linalg.inv(T1)               # 8-point result, Hartley frame
    F1, iters = oracle.refine_f_lm(npts, F0)
    r0, r1 = _sampson_residuals(F0, npts), _sampson_residuals(F1, npts)
    assert 0 < iters <= 25 and (r1 ** 2).sum() < (r0 ** 2).sum()
    assert abs(np.linalg.det(F1 / np.linalg.norm(F1))) < 1e-14                               # rank 2 by construction
    U, s, Vt = np.linalg.svd(F1)

    def unpack(x):
        Ux = Rotation.from_rotvec(x[:3]).as_matrix() @ U
        Vx = Rotation.from_rotvec(x[3:6]).as_matrix() @ Vt.T
        return np.outer(Ux[:, 0], Vx[:, 0]) + (s[1] / s[0] + x[6]) * np.outer(Ux[:, 1], Vx[:, 1])

    sol = least_squares(lambda x: _sampson_residuals(unpack(x), npts), np.zeros(7), xtol=1e-15, ftol=1e-15, gtol=1e-15)
    # the LM result is (numerically) a stationary point: scipy cannot lower the cost from it by more than 1e-9 relative
    assert (sol.fun ** 2).sum() >= (r1 ** 2).sum() * (1 - 1e-6)
    G, W = unpack(sol.x), F1 / s[0]
    assert np.abs(G / np.linalg.norm(G) - W / np.linalg.norm(W)).max() < 1e-4
    # weights: zero-weight points do not influence the fit
    w = np.ones(len(idx))
    w[100:] = 0.0
    Fa = oracle.refit_f_ba(pts, idx, weights=w)
    pts2 = pts.copy()
    pts2[100:150, 2:] += 50.0
    Fb = oracle.refit_f_ba(pts2, idx, weights=w)
    # (the Hartley normalisation sees the moved points, so the frames differ: compare the fits on the weighted points)
    ra, rb = _sampson_residuals(Fa, pts[:100]), _sampson_residuals(Fb, pts[:100])
    assert abs((ra ** 2).sum() - (rb ** 2).sum()) < 0.05 * (ra ** 2).sum()


def test_e_bundle_adjustment_against_scipy(oracle):
    """relative_pose/bundle.cpp:176-250 (lm_5dof_impl) + jacobian_impl.h:307-460 + essential.cpp:44-117 restated
    (oracle/gcr_oracle_eba.cpp); pinned against an independent scipy least-squares minimisation of the same Sampson cost over
    essential matrices [t]_x R (rotation vector + translation), started from the oracle's result."""
    from scipy.optimize import least_squares
    from scipy.spatial.transform import Rotation
    px, Fm, Em, Kc, _ = synth.epipolar_pair(500, 0.0, seed=33, sigma=0.8)
    pts = oracle.normalize_correspondences(px, Kc, Kc)
    idx = np.arange(200, dtype=np.uint32)
    init = oracle.refit_e(pts, idx)
    assert len(init) >= 1
    costs0 = [(_sampson_residuals(E, pts[:200]) ** 2).sum() for E in init]
    E0 = init[int(np.argmin(costs0))]
    refined = oracle.refine_e_lm(pts, idx, E0)
    assert 1 <= len(refined) <= 4
    costs = [(_sampson_residuals(E, pts[:200]) ** 2).sum() for E in refined]
    E1 = refined[int(np.argmin(costs))]
    assert min(costs) <= min(costs0) * (1 + 1e-12)
    # an essential matrix by construction: two equal singular values, one zero
    sv = np.linalg.svd(E1, compute_uv=False)
    assert abs(sv[0] - sv[1]) < 1e-9 * sv[0] and sv[2] < 1e-12 * sv[0]
    # scipy from the oracle's optimum: E = [t]_x R with R = exp(w) R1 and t on the sphere
    U, s, Vt = np.linalg.svd(E1)
    W = np.array([[0, -1, 0], [1, 0, 0], [0, 0, 1.0]])
    best = None
    for R1 in (U @ W @ Vt, U @ W.T @ Vt):
        R1 = R1 * np.sign(np.linalg.det(R1))
        for t1 in (U[:, 2], -U[:, 2]):
            tx = np.array([[0, -t1[2], t1[1]], [t1[2], 0, -t1[0]], [-t1[1], t1[0], 0]])
            Ec = tx @ R1
            err = min(np.abs(Ec / np.linalg.norm(Ec) - E1 / np.linalg.norm(E1)).max(),
                      np.abs(Ec / np.linalg.norm(Ec) + E1 / np.linalg.norm(E1)).max())
            if best is None or err < best[0]:
                best = (err, R1, t1)
    assert best[0] < 1e-9
    _, R1, t1 = best

    def unpack(x):
        R = Rotation.from_rotvec(x[:3]).as_matrix() @ R1
        t = t1 + x[3:6]
        return np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]]) @ R

    sol = least_squares(lambda x: _sampson_residuals(unpack(x), pts[:200]), np.zeros(6), xtol=1e-15, ftol=1e-15, gtol=1e-15)
    assert (sol.fun ** 2).sum() >= min(costs) * (1 - 1e-6)                  # the LM result is a stationary point
    G = unpack(sol.x)
    assert min(np.abs(G / np.linalg.norm(G) - E1 / np.linalg.norm(E1)).max(),
               np.abs(G / np.linalg.norm(G) + E1 / np.linalg.norm(E1)).max()) < 1e-4
    # ... close to the ground truth, and the whole post-run polish returns it among its models
    Et = Em / np.linalg.norm(Em)
    assert min(np.abs(E1 / np.linalg.norm(E1) - Et).max(), np.abs(E1 / np.linalg.norm(E1) + Et).max()) < 5e-3
    w = np.linspace(1.0, 0.2, len(idx))
    models = oracle.polish_e_ba(pts, idx, weights=w)
    assert 1 <= len(models) <= 40
    assert min(min(np.abs(M / np.linalg.norm(M) - Et).max(), np.abs(M / np.linalg.norm(M) + Et).max()) for M in models) < 5e-3
    assert len(oracle.polish_e_ba(pts, idx[:5])) == 0                       # sampleSize() of the BA solver is 6


def test_pnp_nonminimal_against_opencv_epnp_golden(oracle):
    """The stated non-minimal absolute-pose solver (oracle/gcr_oracle_pba.cpp: normalised DLT + the LM of refine_pnp) against
    golden poses made WITHOUT this repository's code (oracle/gen_pnp_fixture.py: OpenCV's EPnP -- the reference's
    initialiser -- refined by scipy least-squares) on the reference's own example scene.  The LM stops on the reference's
    tolerances (gradient / step 1e-8), hence 1e-4 rather than 1e-9."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "pose6dscene.npz"))
    pts, idx = g["pts"], np.nonzero(g["inlier"])[0].astype(np.uint32)

    def cost(P, ii):
        Z = pts[ii, 2:5] @ P[:, :3].T + P[:, 3]
        return ((Z[:, :2] / Z[:, 2:] - pts[ii, :2]) ** 2).sum()

    for ii, key in ((idx, "epnp_lm_all"), (idx[:21], "epnp_lm_21")):
        P, G = oracle.refit_pnp(pts, ii), g[key]
        assert np.abs(P[:, :3] - G[:, :3]).max() < 1e-4 and np.abs(P[:, 3] - G[:, 3]).max() < 1e-4 * np.abs(G[:, 3]).max()
        assert abs(np.linalg.det(P[:, :3]) - 1) < 1e-12 and np.abs(P[:, :3] @ P[:, :3].T - np.eye(3)).max() < 1e-12
        assert cost(P, ii) < 1e-10                                         # (the scene is noise-free: both sit in the minimum)
        P2, iters = oracle.refine_pnp_lm(pts, ii, G)                       # the golden pose is a fixed point of the LM
        assert iters == 0 and np.abs(P2 - G).max() < 1e-12
    assert np.abs(oracle.refit_pnp(pts, idx)[:, :3] - g["gt_pose"][:, :3]).max() < 1e-4
    assert oracle.refit_pnp(pts, idx[:5]) is None                          # fewer than six points: no model (stated)
    # noisy synthetic scene: DLT + LM and EPnP-like start + LM meet in the same minimum; weights (by position) act
    spts, P0, inl = synth.pnp_scene(400, 0.0, seed=12)
    si = np.arange(120, dtype=np.uint32)
    Pa = oracle.refit_pnp(spts, si)
    Pb, _ = oracle.refine_pnp_lm(spts, si, P0)
    assert np.abs(Pa - Pb).max() < 1e-7 and np.abs(Pa - P0).max() < 5e-3
    w = np.ones(len(si))
    w[60:] = 0.0
    moved = spts.copy()
    moved[60:120, :2] += 0.3
    assert np.abs(oracle.refit_pnp(moved, si, weights=w) - oracle.refit_pnp(spts, si[:60])).max() < 1e-6
