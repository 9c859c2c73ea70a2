"""The oracle's C restatement (oracle/c) against the NumPy oracle: bit-exact."""
import numpy as np
import pytest

import parity_utils  # noqa: F401
from oracle import ckernels

pytestmark = pytest.mark.skipif(not ckernels.available(), reason="oracle/c not built (run __graft_entry__.build())")


def test_scan_hist_and_cooc():
    rng = np.random.default_rng(0)
    n, dom = 50001, [2, 7, 33, 80]
    cols = [rng.integers(-1, d, size=n).astype(np.int32) for d in dom]
    h = ckernels.scan_hist(cols, dom)
    off = 0
    for c, d in zip(cols, dom):
        assert np.array_equal(h[off:off + d + 1], np.bincount(c.astype(np.int64) + 1, minlength=d + 1))
        off += d + 1
    px, py = [0, 1, 3], [2, 3, 0]
    out, toff = ckernels.cooc(cols, dom, px, py)
    for q, (x, y) in enumerate(zip(px, py)):
        want = np.bincount((cols[x].astype(np.int64) + 1) * (dom[y] + 1) + cols[y] + 1,
                           minlength=(dom[x] + 1) * (dom[y] + 1))
        assert np.array_equal(out[toff[q]:toff[q + 1]], want)
    assert ckernels.num_threads() >= 1


@pytest.mark.parametrize("n_classes", [1, 2, 5])
def test_forest_margins(n_classes):
    from oracle.forest import forest_margins
    from tools.randforest import random_forest
    rng = np.random.default_rng(n_classes)
    f = random_forest(9, n_classes, 30, [[-0.5, 0.5, 1.5, 2.5]] * 9, rng, leaf_scale=0.3)
    X = rng.integers(-2, 4, size=(2000, 9)).astype(np.float64)
    X[rng.random(X.shape) < 0.1] = np.nan
    assert np.array_equal(ckernels.forest_margins(f, X), forest_margins(f, X))
