"""K6 (weighted least-squares refits) and K7 (grid neighbourhood) parity through the C ABI against the oracle.

Tolerances (BASELINE.json north_star): final models <= 1e-5 relative, compared up to scale; grid cell keys, cell
lists, filled-cell count and neighbor_number bit-exact."""
import numpy as np
import oracle as O
import pytest
from pygcransac import _cabi, synth

pytestmark = pytest.mark.gpu
RTOL_MODEL = 1e-5


def _up_to_scale(A, B):
    A = A / np.linalg.norm(A)
    B = B / np.linalg.norm(B)
    if (A * B).sum() < 0:
        B = -B
    return np.abs(A - B).max() / np.abs(A).max()


def _index_sets(rng, inl_idx, sizes):
    return [rng.choice(inl_idx, size=s, replace=False).astype(np.uint32) for s in sizes]


def test_refit_h_matches_oracle(gcr_ctx, oracle):
    pts, Hm, inl = synth.homography_pair(3000, 0.3, seed=51)
    gcr_ctx.points_upload(pts)
    rng = np.random.default_rng(1)
    inl_idx = np.nonzero(inl)[0]
    # LO-sized sets (7 m = 28), the minimal n == 4 branch, a large "all inliers" set, and one with outliers mixed in
    sets = _index_sets(rng, inl_idx, [28] * 20 + [4] * 6 + [5, 9, 200, len(inl_idx)])
    sets.append(rng.choice(3000, size=40, replace=False).astype(np.uint32))
    models, ok = gcr_ctx.refit(_cabi.H, sets)
    worst = 0.0
    for s, mdl, k in zip(sets, models, ok):
        want = oracle.refit(O.H, pts, s)
        assert bool(k) == (want is not None)
        if want is not None:
            worst = max(worst, _up_to_scale(want, mdl))
    assert worst < RTOL_MODEL, worst


def test_refit_h_weighted_both_index_modes(gcr_ctx, oracle):
    pts, Hm, inl = synth.homography_pair(1200, 0.2, seed=52)
    gcr_ctx.points_upload(pts)
    rng = np.random.default_rng(2)
    idx = np.sort(rng.choice(np.nonzero(inl)[0], size=300, replace=False)).astype(np.uint32)
    w_point = rng.uniform(0.05, 1.0, size=1200)
    # reference behaviour: the solver reads weights[i] by POSITION in the set
    got, ok = gcr_ctx.refit(_cabi.H, [idx], weights=w_point, weight_mode=_cabi.WEIGHT_POSITIONAL)
    want = oracle.refit(O.H, pts, idx, weights=w_point)
    assert ok[0] and _up_to_scale(want, got[0]) < RTOL_MODEL
    # by-point indexing == positional indexing of the gathered weights
    got2, ok2 = gcr_ctx.refit(_cabi.H, [idx], weights=w_point, weight_mode=_cabi.WEIGHT_BY_POINT)
    want2 = oracle.refit(O.H, pts, idx, weights=np.ascontiguousarray(w_point[idx]))
    assert ok2[0] and _up_to_scale(want2, got2[0]) < RTOL_MODEL
    assert _up_to_scale(want, want2) > 1e-9  # the two modes really differ


def test_refit_f8_matches_oracle(gcr_ctx, oracle):
    pts, Fm, Em, Kc, inl = synth.epipolar_pair(2500, 0.4, seed=53)
    gcr_ctx.points_upload(pts)
    rng = np.random.default_rng(3)
    inl_idx = np.nonzero(inl)[0]
    sets = _index_sets(rng, inl_idx, [49] * 20 + [8, 8, 9, 15, 300, len(inl_idx)])
    models, ok = gcr_ctx.refit(_cabi.F, sets)
    worst = 0.0
    for s, mdl, k in zip(sets, models, ok):
        want = oracle.refit(O.F, pts, s)
        assert bool(k) == (want is not None)
        worst = max(worst, _up_to_scale(want, mdl))
    assert worst < RTOL_MODEL, worst
    w = rng.uniform(0.05, 1.0, size=2500)
    got, ok = gcr_ctx.refit(_cabi.F, [sets[-1]], weights=w, weight_mode=_cabi.WEIGHT_POSITIONAL)
    assert ok[0] and _up_to_scale(oracle.refit(O.F, pts, sets[-1], weights=w), got[0]) < RTOL_MODEL


def test_refit_f_bundle_adjustment_matches_oracle(gcr_ctx, oracle):
    """K6 + LM bundle adjustment (gcr_refit_f_ba: the reference's default non-minimal F fit, 8-point + lm_F_impl) against
    the oracle's restatement: Hartley frame as the estimator calls it, pixel frame as the post-run polish does, weights."""
    pts, Fm, Em, Kc, inl = synth.epipolar_pair(2500, 0.4, seed=54)
    gcr_ctx.points_upload(pts)
    rng = np.random.default_rng(5)
    inl_idx = np.nonzero(inl)[0]
    sets = _index_sets(rng, inl_idx, [49] * 20 + [8, 9, 15, 300, len(inl_idx)])
    sets.append(rng.choice(2500, size=49, replace=False).astype(np.uint32))                  # outliers mixed in
    sets.append(inl_idx[:7].astype(np.uint32))                                               # n < 8: no model
    models, ok = gcr_ctx.refit(_cabi.F, sets, ba_frame=_cabi.BA_HARTLEY)
    plain, _ = gcr_ctx.refit(_cabi.F, sets)
    worst, moved = 0.0, 0.0
    for s, mdl, lin, k in zip(sets, models, plain, ok):
        want = oracle.refit_f_ba(pts, s)
        assert bool(k) == (want is not None)
        if want is not None:
            worst = max(worst, _up_to_scale(want, mdl))
            moved = max(moved, _up_to_scale(lin, mdl))
            assert abs(np.linalg.det(mdl / np.linalg.norm(mdl))) < 1e-12                      # rank 2 after the refinement
    assert worst < RTOL_MODEL, worst
    assert moved > 1e-6                                                                       # the refinement does something
    w = rng.uniform(0.05, 1.0, size=2500)
    got, ok = gcr_ctx.refit(_cabi.F, [sets[24]], weights=w, weight_mode=_cabi.WEIGHT_POSITIONAL, ba_frame=_cabi.BA_HARTLEY)
    assert ok[0] and _up_to_scale(oracle.refit_f_ba(pts, sets[24], weights=w), got[0]) < RTOL_MODEL
    # the post-run polish (py/gcransac_python.cpp:1832-1858): pixel frame, weights 1 - r^2 / T by inlier position
    got, ok = gcr_ctx.refit(_cabi.F, [sets[24]], weights=w, weight_mode=_cabi.WEIGHT_POSITIONAL, ba_frame=_cabi.BA_PIXEL)
    assert ok[0] and _up_to_scale(oracle.refit_f_ba(pts, sets[24], weights=w, polish=True), got[0]) < RTOL_MODEL


def _e_set_distance(got, want):
    """every model of `got` has a partner in `want` (normalised, E33 >= 0 on both sides): worst partner distance"""
    return max(min(np.abs(E - q).max() for q in want) for E in got) if len(got) else 0.0


def test_refit_e_stewenius_matches_oracle(gcr_ctx, oracle):
    """K6 for E (pivoted-QR null space of A^T A + the 5-point kernel's Sturm back end) against the oracle's restatement of
    the reference's Stewenius action-matrix solver: same model set to 1e-5 (SURVEY a13's acceptance), LO-sized sets,
    small sets, all inliers, weights read by POINT (stewenius.h:143) and the n == 5 LU branch."""
    pts, Fm, Em, Kc, inl = synth.epipolar_pair(2048, 0.5, seed=57)
    npts = oracle.normalize_correspondences(pts, Kc, Kc)
    gcr_ctx.points_upload(npts)
    rng = np.random.default_rng(4)
    inl_idx = np.nonzero(inl)[0]
    sets = _index_sets(rng, inl_idx, [35] * 40 + [5, 5, 6, 7, 12, 100, len(inl_idx)])
    sets += [rng.choice(2048, size=35, replace=False).astype(np.uint32) for _ in range(6)]       # outliers mixed in
    models, ok = gcr_ctx.refit(_cabi.E, sets)
    assert models.shape == (len(sets), 10, 3, 3) and ok.shape == (len(sets), 10)
    worst, exact_sets, best_worst = 0.0, 0, 0.0
    thr = 0.75 / 1000.0
    for s, mdl, k in zip(sets, models, ok):
        want = oracle.refit_e(npts, s)
        got = mdl[k.astype(bool)]
        assert np.all(np.diff(k.astype(int)) <= 0)                    # valid slots first
        for E in got:
            assert abs(np.linalg.norm(E) - 1) < 1e-12 and E[2, 2] >= 0
        if len(got) == len(want):
            exact_sets += 1
            worst = max(worst, _e_set_distance(got, want))
        # what the LO step uses -- the best-scoring model of the set -- always agrees
        if len(want):
            sw = [oracle.score(O.E, O.MSAC, npts, E, thr)[0] for E in want]
            sg = [oracle.score(O.E, O.MSAC, npts, E, thr)[0] for E in got]
            if max(sw) > 10:
                bw, bg = want[int(np.argmax(sw))], got[int(np.argmax(sg))]
                best_worst = max(best_worst, np.abs(bw - bg).max())
    assert exact_sets >= len(sets) - 2, exact_sets      # a near-double root may split differently in the two root finders
    assert worst < RTOL_MODEL, worst
    assert best_worst < RTOL_MODEL, best_worst
    # weights by point
    w = rng.uniform(0.05, 1.0, size=2048)
    got, k = gcr_ctx.refit(_cabi.E, [sets[0], sets[-7]], weights=w, weight_mode=_cabi.WEIGHT_BY_POINT)
    for i, s in enumerate((sets[0], sets[-7])):
        want = oracle.refit_e(npts, s, w)
        g = got[i][k[i].astype(bool)]
        assert len(g) == len(want) and _e_set_distance(g, want) < RTOL_MODEL
    # too small
    got, k = gcr_ctx.refit(_cabi.E, [inl_idx[:4].astype(np.uint32)])
    assert not k.any()


def test_refit_too_small_sets_and_errors(gcr_ctx):
    pts, Hm, inl = synth.homography_pair(100, 0.0, seed=54)
    gcr_ctx.points_upload(pts)
    models, ok = gcr_ctx.refit(_cabi.H, [np.arange(3, dtype=np.uint32), np.arange(10, dtype=np.uint32)])
    assert list(ok) == [0, 1] and not models[0].any()
    models, ok = gcr_ctx.refit(_cabi.F, [np.arange(7, dtype=np.uint32)])
    assert list(ok) == [0]
    with pytest.raises(_cabi.GcrError):
        gcr_ctx.refit(_cabi.PNP, [np.arange(10, dtype=np.uint32)])


@pytest.mark.parametrize("n,cells", [(762, 4), (4096, 8), (5000, 20), (100000, 3), (1, 4), (257, 64)])
def test_grid_bit_exact(gcr_ctx, oracle, n, cells):
    pts = synth.homography_pair(n, 0.5, seed=60 + cells)[0]
    gcr_ctx.points_upload(pts)
    cs = np.array([synth.SIZE / cells] * 4) + (1e-9 if cells == 4 else 0.0)
    want = oracle.grid_build(pts, cs, cells)
    got = gcr_ctx.grid_build(cs, cells)
    assert np.array_equal(got["keys"], want["keys"])
    assert np.array_equal(got["sorted_idx"], want["sorted_idx"])
    assert got["n_cells"] == want["n_cells"] and got["neighbor_number"] == want["neighbor_number"]
    assert np.array_equal(got["cell_offsets"], want["cell_offsets"][: want["n_cells"] + 1])


def test_grid_golden_fixture(gcr_ctx):
    """The committed fixture was generated from the reference's own GridNeighborhoodGraph (oracle/gen_golden.py)."""
    import os
    from conftest import GOLDEN
    g = np.load(os.path.join(GOLDEN, "grid.npz"))
    gcr_ctx.points_upload(g["pts"])
    for Cn in (3, 4, 8):
        got = gcr_ctx.grid_build(g["cell_sizes_%d" % Cn], Cn)
        assert np.array_equal(got["keys"], g["keys_%d" % Cn])
        assert np.array_equal(got["sorted_idx"], g["sorted_idx_%d" % Cn])
        assert [got["n_cells"], got["neighbor_number"]] == list(g["meta_%d" % Cn])


def test_grid_second_pair_and_errors(gcr_ctx, oracle):
    pairs = np.stack([synth.homography_pair(900, 0.3, seed=s)[0] for s in (71, 72)])
    gcr_ctx.points_upload(pairs, n_pairs=2)
    cs = np.array([synth.SIZE / 5.0] * 4)
    got = gcr_ctx.grid_build(cs, 5, pair=1)
    want = oracle.grid_build(pairs[1], cs, 5)
    assert np.array_equal(got["sorted_idx"], want["sorted_idx"]) and got["neighbor_number"] == want["neighbor_number"]
    with pytest.raises(_cabi.GcrError):
        gcr_ctx.grid_build(np.array([0.0, 1.0, 1.0, 1.0]), 4)


def test_radius_graph_bit_exact(gcr_ctx, ref):
    """neighborhood=1: the exact radius neighbourhood (FlannNeighborhoodGraph's contract without FLANN's approximation)
    against the graph class the verbatim reference driver runs on, and against numpy."""
    for n, radius, seed in ((700, 20.0, 61), (3000, 35.0, 62), (257, 500.0, 63)):
        pts, _, _ = synth.homography_pair(n, 0.4, seed=seed)
        pts[5] = pts[6]                                              # a duplicate point is a neighbour at distance 0
        gcr_ctx.points_upload(pts)
        off, nbr = gcr_ctx.radius_graph(radius)
        want_off, want_nbr = ref.radius_graph(pts, radius)
        assert np.array_equal(off, want_off) and np.array_equal(nbr, want_nbr)
        f = pts.astype(np.float32)
        d = f[:, None, :] - f[None, :, :]
        d2 = ((d[..., 0] ** 2 + d[..., 1] ** 2) + d[..., 2] ** 2) + d[..., 3] ** 2
        A = (d2 < np.float32(radius) ** 2) & ~np.eye(n, dtype=bool)
        assert np.array_equal(np.nonzero(A)[1].astype(np.uint32), nbr) and len(nbr) > 0
    with pytest.raises(_cabi.GcrError):
        gcr_ctx.radius_graph(0.0)


def _grid_edges_numpy(key):
    """labeling()'s pairs: for i ascending, the points j > i of i's cell, ascending (GCRANSAC.h:963-1008 with the grid);
    `key` = the reference's cell key per point (points outside the image can share a key across cells: kept as it is)."""
    edges = []
    members = {}
    for i in range(len(key)):
        members.setdefault(int(key[i]), []).append(int(i))
    for i in range(len(key)):
        for j in members[int(key[i])]:
            if j > i:
                edges.append((i, j))
    return np.array(edges, dtype=np.uint32).reshape(-1, 2)


@pytest.mark.parametrize("n", [700, 20000])
def test_grid_edge_list_matches_reference_order(gcr_ctx, oracle, n):
    """gcr_grid_edges: every unordered pair of one cell once, (i, j > i), i ascending then j ascending -- the pairs and the
    order the reference's labeling() produces with its N x N used_edges matrix."""
    pts = synth.homography_pair(n, 0.3, seed=8)[0]
    gcr_ctx.points_upload(pts)
    cs = np.array([synth.SIZE / 8] * 4)
    edges, off = gcr_ctx.grid_edges(cs, 8)
    want = _grid_edges_numpy(oracle.grid_build(pts, cs, 8)["keys"])
    assert np.array_equal(edges, want)
    assert off[-1] == len(want) and np.array_equal(np.diff(off.astype(np.int64)), np.bincount(want[:, 0], minlength=n))


def test_csr_edge_list_and_labeling_terms(gcr_ctx, oracle):
    """gcr_csr_edges over the radius graph, and gcr_labeling_terms: unary / pairwise capacities bit-identical to the host
    statement of GCRANSAC.h:933-1003 on the oracle's FP64 residuals."""
    pts, Hm, _ = synth.homography_pair(1500, 0.3, seed=9)
    gcr_ctx.points_upload(pts)
    offsets, nbrs = gcr_ctx.radius_graph(40.0)
    edges, eoff = gcr_ctx.csr_edges(offsets, nbrs)
    want = np.array([(i, j) for i in range(len(pts)) for j in nbrs[offsets[i]:offsets[i + 1]] if j > i], dtype=np.uint32).reshape(-1, 2)
    assert len(want) > 1000 and np.array_equal(edges, want) and eoff[-1] == len(want)
    thr, lam = 2.0, 0.975
    unary, pairwise, dist = gcr_ctx.labeling_terms(_cabi.H, Hm, thr, lam, edges)
    r2 = oracle.residuals(O.H, pts, Hm)
    T = thr * thr * 9 / 4
    d = np.clip(r2 / T, 0.0, 1.0)
    tmp = 1 - d
    src = np.where(r2 <= T, (1.0 - lam) * tmp, 0.0)
    snk = np.where(r2 <= T, 0.0, (1.0 - lam) * (1 - tmp))
    assert np.array_equal(unary[:, 0], src) and np.array_equal(unary[:, 1], snk)
    assert np.array_equal(pairwise, 0.5 * (d[want[:, 0]] + d[want[:, 1]]) * lam) and np.array_equal(dist, d)


def test_refit_e_bundle_adjustment_matches_oracle(gcr_ctx, oracle):
    """gcr_refit_e_ba (post-run polish of findEssentialMatrix_: Stewenius -> motion_from_essential + cheirality -> 5-dof LM)
    against the oracle restatement: the same models in the same slots, <= 1e-6 up to scale (tolerance 1e-5), with and
    without weights; fewer than 6 points: nothing."""
    px, Fm, Em, Kc, inl = synth.epipolar_pair(900, 0.3, seed=44)
    pts = oracle.normalize_correspondences(px, Kc, Kc)
    gcr_ctx.points_upload(pts)
    rng = np.random.default_rng(3)
    idx = np.nonzero(inl)[0].astype(np.uint32)
    for weights in (None, rng.uniform(0.1, 1.0, size=len(idx))):
        got, ok = gcr_ctx.refit_e_ba(idx, weights)
        want = oracle.polish_e_ba(pts, idx, weights)
        assert ok.sum() == len(want) and len(want) >= 1
        for M, Wm in zip(got[ok.astype(bool)], want):
            assert _up_to_scale(Wm, M) < 1e-6, _up_to_scale(Wm, M)
        Et = Em / np.linalg.norm(Em)
        assert min(_up_to_scale(Et, M) for M in got[ok.astype(bool)]) < 5e-3
    got, ok = gcr_ctx.refit_e_ba(idx[:5])
    assert not ok.any()


def test_refit_pnp_matches_oracle(gcr_ctx, oracle):
    """gcr_refit with GCR_PNP (normalised DLT + the LM of refine_pnp, one CTA per index set) against the oracle: synthetic
    scenes with noise, several sets per launch, weights in both index modes, the reference's example scene; < 6 points: none."""
    import os
    from conftest import GOLDEN
    pts, P0, inl = synth.pnp_scene(1200, 0.3, seed=17)
    gcr_ctx.points_upload(pts)
    rng = np.random.default_rng(2)
    inl_idx = np.nonzero(inl)[0]
    sets = _index_sets(rng, inl_idx, [21, 21, 60, 400, 6, 5])
    got, ok = gcr_ctx.refit(_cabi.PNP, sets)
    assert list(ok) == [1, 1, 1, 1, 1, 0]
    for s, M in zip(sets[:5], got):
        want = oracle.refit_pnp(pts, s)
        assert np.abs(M - want).max() < 1e-7, np.abs(M - want).max()
    assert np.abs(got[3] - P0).max() < 5e-3
    w = rng.uniform(0.2, 1.0, size=len(pts))
    s = sets[3]
    got_p, _ = gcr_ctx.refit(_cabi.PNP, [s], weights=w[: len(s)], weight_mode=_cabi.WEIGHT_POSITIONAL)
    got_b, _ = gcr_ctx.refit(_cabi.PNP, [s], weights=w, weight_mode=_cabi.WEIGHT_BY_POINT)
    assert np.abs(got_p[0] - oracle.refit_pnp(pts, s, weights=w[: len(s)])).max() < 1e-7
    assert np.abs(got_b[0] - oracle.refit_pnp(pts, s, weights=w[s])).max() < 1e-7
    g = np.load(os.path.join(GOLDEN, "pose6dscene.npz"))
    gcr_ctx.points_upload(g["pts"])
    idx = np.nonzero(g["inlier"])[0].astype(np.uint32)
    got, ok = gcr_ctx.refit(_cabi.PNP, [idx, idx[:21]])
    assert ok.all()
    for M, key in zip(got, ("epnp_lm_all", "epnp_lm_21")):
        G = g[key]
        assert np.abs(M[:, :3] - G[:, :3]).max() < 1e-4 and np.abs(M[:, 3] - G[:, 3]).max() < 1e-4 * np.abs(G[:, 3]).max()


def test_radius_graph_five_columns(gcr_ctx):
    """find6DPose_'s neighbourhood: the radius graph over the whole 5-column row (u v X Y Z)."""
    pts, _, _ = synth.pnp_scene(900, 0.3, seed=5)
    gcr_ctx.points_upload(pts)
    off, nbr = gcr_ctx.radius_graph(0.8)
    f = pts.astype(np.float32)
    d = f[:, None, :] - f[None, :, :]
    d2 = (((d[..., 0] ** 2 + d[..., 1] ** 2) + d[..., 2] ** 2) + d[..., 3] ** 2) + d[..., 4] ** 2
    A = (d2 < np.float32(0.8) ** 2) & ~np.eye(len(pts), dtype=bool)
    assert np.array_equal(np.nonzero(A)[1].astype(np.uint32), nbr) and len(nbr) > 1000
    assert np.array_equal(off, np.concatenate([[0], np.cumsum(A.sum(1))]).astype(np.uint32))
