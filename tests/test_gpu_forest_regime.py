"""The rank-coded forest kernel in the regime the benchmark runs it in: every CTA owns several
tiles, so the chunk stream index runs across tiles, the mbarrier parities flip many times and the
refill of a buffer for the NEXT tile's first chunk is issued while slow warps are still inside the
current tile.  Margins must be bit-identical to the oracle's C forest (tree-order float64 sums) and
the filled labels identical, for every shared-memory layout of the kernel, with one / an odd / an
even number of forest chunks and one / several class sequences (model.py:1107-1135 of the
reference is the loop this replaces)."""
import numpy as np
import pytest

import parity_utils  # noqa: F401  (sys.path)

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")


@pytest.fixture(scope="module")
def ctx():
    from repair._native import Context
    c = Context(0)
    yield c
    c.close()


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def build_case(rng, doms, n_classes, n_iter, n_rows):
    from repair.forest import DeviceModel, encoder_width
    from tools.randforest import random_forest
    k = len(doms)
    cols = []
    for d in doms:
        c = rng.integers(0, d, size=n_rows).astype(np.int32)
        c[rng.random(n_rows) < 0.06] = -1
        cols.append(c)
    tile_np = np.stack(cols, axis=1).astype(np.int32)
    names = ["a%d" % i for i in range(k)]
    encoders = []
    for i in range(1, k):
        cats = list(rng.permutation(doms[i]))[: doms[i] - (i % 3 == 0)]       # sometimes an unseen value
        if i % 2:
            cats.insert(1, -1)                                                 # NULL seen in training
        encoders.append({"attr": names[i], "type": "sum" if doms[i] < 12 else "ordinal",
                         "categories": [int(c) for c in cats]})
    n_feat = sum(encoder_width(e) for e in encoders)
    thr = []
    for e in encoders:
        kk = len(e["categories"])
        thr += [[-0.5, 0.5]] * (kk - 1) if e["type"] == "sum" else [[-1.0] + [j + 0.5 for j in range(1, kk)]]
    forest = random_forest(n_feat, n_classes, n_iter, thr, rng, leaf_scale=0.1)
    spec = {"forest": forest, "encoders": encoders, "class_codes": list(range(max(n_classes, 2))), "integral": False}
    tile_cols = {nm: i for i, nm in enumerate(names)}
    dict_sizes = dict(zip(names, doms))
    dm = DeviceModel(spec, tile_cols, dict_sizes, {}, torch.device("cuda", 0))
    return tile_np, names, encoders, forest, dm, tile_cols, dict_sizes, n_feat


# (classes, boosting rounds, extra attributes): chunk counts 1 / even / odd / one per sequence, and a
# feature count past what the 512-cell wide tile holds (so that the 256-cell kernels run)
CASES = [(2, 9, []), (2, 300, []), (3, 40, []), (5, 70, []), (7, 150, []), (4, 35, [11, 10, 9, 8, 7])]


@pytest.mark.parametrize("n_classes,n_iter,extra", CASES)
def test_ranked_forest_many_tiles_per_cta(ctx, n_classes, n_iter, extra):
    from oracle import ckernels
    from oracle.forest import forest_predict
    from repair.forest import encode_matrix
    assert ckernels.available(), "oracle/c is not built (run __graft_entry__.build())"
    sm = torch.cuda.get_device_properties(0).multi_processor_count
    n_rows = sm * 512 * 3 + 4096 + 17
    rng = np.random.default_rng(1000 * n_classes + n_iter)
    doms = [7, 4, 30, 3, 9, 64, 2, 12] + extra
    tile_np, names, encoders, forest, dm, tile_cols, dict_sizes, n_feat = build_case(rng, doms, n_classes, n_iter,
                                                                                   n_rows)
    assert dm.ranked is not None
    n_chunks = dm.ranked.n_chunks
    cells = np.sort(rng.choice(n_rows, size=sm * 512 * 3 + 17, replace=False)).astype(np.int32)
    assert len(cells) >= sm * 512 * 3 + 17                       # >= 3 tiles per CTA even at 512 cells / tile
    X = encode_matrix(encoders, {nm: tile_np[cells, tile_cols[nm]] for nm in names[1:]}, {}, dict_sizes)
    want_m = ckernels.forest_margins(forest, X)
    # oracle label rule (oracle/forest.py forest_predict): binary margin > 0, multiclass argmax, ties -> lowest
    lab = (want_m[:, 0] > 0).astype(np.int32) if want_m.shape[1] == 1 else np.argmax(want_m, axis=1).astype(np.int32)
    assert np.array_equal(forest_predict(forest, X[:2048]).astype(np.int32), lab[:2048])
    variants = ["auto", "bytes", "wide8", "wide16"]
    assert (dm.ranked.n_slots > 72) == bool(extra)               # past the 512-cell tile: the 256-cell kernels
    k = len(doms)
    for variant in variants:
        dm.ranked.layout = {"auto": 0, "bytes": 1, "wide8": 2, "wide16": 3}[variant]
        for with_margin in (True, False):
            tile = dev(tile_np)
            margins = torch.empty((len(cells), dm.n_seq), dtype=torch.float64, device="cuda") if with_margin else None
            dm.predict(ctx, tile, k, None, 0, dev(cells), len(cells), 0, margins)
            got_lab = tile.cpu().numpy()[cells, 0]
            if with_margin:
                assert np.array_equal(margins.cpu().numpy(), want_m), (variant, n_chunks)
            assert np.array_equal(got_lab, lab), (variant, with_margin, n_chunks)
            # rows that are not listed keep their target code
            rest = np.setdiff1d(np.arange(n_rows), cells)
            assert np.array_equal(tile.cpu().numpy()[rest, 0], tile_np[rest, 0])
    assert n_chunks >= (1 if n_classes == 2 else n_classes)


def test_chunk_counts_cover_one_odd_even():
    """The parametrisation above really spans n_chunks == 1, odd > 1 and even (host-side check of the
    images so that a change of the chunk sizes cannot silently drop a regime)."""
    seen = set()
    for n_classes, n_iter, extra in CASES:
        rng = np.random.default_rng(1000 * n_classes + n_iter)
        doms = [7, 4, 30, 3, 9, 64, 2, 12] + extra
        dm = build_case(rng, doms, n_classes, n_iter, 64)[4]
        seen.add(dm.ranked.n_chunks)
    assert 1 in seen and any(c > 1 and c % 2 for c in seen) and any(c % 2 == 0 for c in seen), seen
