"""Degenerate and hostile inputs through the three entry points: minimal sizes, identical / collinear points, NaN / inf
coordinates, extreme magnitudes, pure noise, float32 input, ragged and large N.  Nothing may raise, hang or abort (the
reference's max-flow asserts on the NaN capacities such inputs produce, energy.h:228); the result is a (model | None,
bool mask of length n) pair as the reference's bindings return it (py/bindings.cpp:973-975)."""
import numpy as np
import pytest
from pygcransac import synth

pytestmark = [pytest.mark.gpu, pytest.mark.timeout(120)]


@pytest.fixture(scope="module")
def pyg():
    import pygcransac
    try:
        pygcransac._load_driver()
    except Exception as e:
        pytest.skip("host driver not available: %s" % e)
    return pygcransac


def _cases(m, planar):
    rng = np.random.default_rng(m)
    good = (synth.homography_pair if planar else synth.epipolar_pair)
    yield "n = m", rng.uniform(0, 1000, (m, 4)), None
    yield "n = m + 1", rng.uniform(0, 1000, (m + 1, 4)), None
    yield "identical points", np.tile(np.array([[10.0, 20.0, 30.0, 40.0]]), (100, 1)), 0
    a = np.arange(200.0)
    yield "collinear", np.c_[a, a, a + 1, a + 1], None
    p = rng.uniform(0, 1000, (300, 4))
    p[::7] = np.nan
    yield "NaN rows", p, None
    p = rng.uniform(0, 1000, (300, 4))
    p[3] = np.inf
    yield "an inf row", p, None
    yield "huge coordinates", rng.uniform(0, 1e12, (300, 4)), None
    yield "tiny coordinates", rng.uniform(0, 1e-12, (300, 4)), None
    yield "pure noise", rng.uniform(0, 1000, (2000, 4)), None
    yield "float32", good(500, 0.3, 3)[0].astype(np.float32), 300
    yield "ragged n = 257", good(257, 0.3, 5)[0], 150
    yield "n = 20000", good(20000, 0.5, 6)[0], 9000


@pytest.mark.parametrize("problem", ["H", "F", "E"])
def test_degenerate_inputs_return(pyg, problem):
    K = synth.K_MAT
    fn = {"H": lambda p: pyg.findHomography(p, 1000, 1000, 1000, 1000, [], threshold=3.0, max_iters=2000),
          "F": lambda p: pyg.findFundamentalMatrix(p, 1000, 1000, 1000, 1000, [], threshold=0.75, max_iters=2000),
          "E": lambda p: pyg.findEssentialMatrix(p, K, K, 1000, 1000, 1000, 1000, [], threshold=0.75, max_iters=2000)}[problem]
    for name, pts, floor in _cases({"H": 4, "F": 7, "E": 5}[problem], problem == "H"):
        model, mask = fn(pts)
        assert mask.dtype == bool and mask.shape == (len(pts),), name
        assert model is None or (model.shape == (3, 3) and np.isfinite(model).all()), name
        assert (model is None) == (mask.sum() == 0), name
        if floor is not None:
            assert mask.sum() >= floor, (name, int(mask.sum()))
