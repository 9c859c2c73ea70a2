"""Kernel-level parity tests: every C-ABI entry point against NumPy / the oracle on seeded inputs,
including ragged sizes (non-multiples of 32/128/1024), empty inputs and all-NULL columns.
Integer results must be bit-exact."""
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


def padded(codes, pad_to=128, fill=-1):
    n = len(codes)
    n_pad = (n + pad_to - 1) // pad_to * pad_to or pad_to
    out = np.full(n_pad, fill, dtype=codes.dtype)
    out[:n] = codes
    return out


def bits_of(t, n):
    w = t.cpu().numpy().view(np.uint32)
    b = np.unpackbits(w.view(np.uint8), bitorder="little")[:n]
    return b.astype(bool)


def rand_cols(rng, n, doms, null_p=0.05):
    cols = []
    for d in doms:
        c = rng.integers(0, max(d, 1), size=n).astype(np.int32) if d > 0 else np.full(n, -1, dtype=np.int32)
        c[rng.random(n) < null_p] = -1
        cols.append(c)
    return cols


@pytest.mark.parametrize("n", [0, 1, 31, 127, 128, 1000, 1024, 4097, 100003])
def test_scan_hist(ctx, n):
    rng = np.random.default_rng(n)
    doms = [2, 4, 8, 16, 32, 64, 48, 24, 80, 3, 1, 0, 13]
    cols = rand_cols(rng, n, doms)
    dcols = [dev(padded(c)) for c in cols]
    n_words = (len(padded(cols[0])) + 31) // 32
    bms = [torch.zeros(n_words, dtype=torch.int32, device="cuda") if i % 3 != 2 else None for i in range(len(doms))]
    hist = torch.zeros(sum(d + 1 for d in doms), dtype=torch.int64, device="cuda")
    ctx.scan_hist(dcols, doms, n, bms, hist)
    h = hist.cpu().numpy()
    off = 0
    for c, d, bm in zip(cols, doms, bms):
        want = np.bincount(c.astype(np.int64) + 1, minlength=d + 1)
        assert np.array_equal(h[off:off + d + 1], want)
        if bm is not None:
            assert np.array_equal(bits_of(bm, n), c < 0)
        off += d + 1


def test_scan_hist_many_columns_and_large_domain(ctx):
    rng = np.random.default_rng(5)
    n = 50000
    doms = [80] * 40 + [2000, 5000]          # several groups + the global-atomic fallback
    cols = rand_cols(rng, n, doms)
    dcols = [dev(padded(c)) for c in cols]
    n_words = (len(padded(cols[0])) + 31) // 32
    bms = [torch.zeros(n_words, dtype=torch.int32, device="cuda") for _ in doms]
    hist = torch.zeros(sum(d + 1 for d in doms), dtype=torch.int64, device="cuda")
    ctx.scan_hist(dcols, doms, n, bms, hist)
    h = hist.cpu().numpy()
    off = 0
    for c, d, bm in zip(cols, doms, bms):
        assert np.array_equal(h[off:off + d + 1], np.bincount(c.astype(np.int64) + 1, minlength=d + 1))
        assert np.array_equal(bits_of(bm, n), c < 0)
        off += d + 1


def test_scan_hist_unaligned_column(ctx):
    rng = np.random.default_rng(6)
    n = 5000
    base = rng.integers(-1, 7, size=n + 131).astype(np.int32)
    d = dev(base)
    view = d[3:3 + n]  # 12-byte offset: not 16-byte aligned -> scalar path
    hist = torch.zeros(8, dtype=torch.int64, device="cuda")
    bm = torch.zeros((n + 31) // 32 + 4, dtype=torch.int32, device="cuda")
    ctx.scan_hist([view], [7], n, [bm], hist)
    assert np.array_equal(hist.cpu().numpy(), np.bincount(base[3:3 + n].astype(np.int64) + 1, minlength=8))
    assert np.array_equal(bits_of(bm, n), base[3:3 + n] < 0)


@pytest.mark.parametrize("n", [1, 33, 1000, 70001])
def test_lut_scan_and_range_flag_and_discretize(ctx, n):
    rng = np.random.default_rng(n)
    c = rand_cols(rng, n, [50])[0]
    lut = (rng.random(50) < 0.3).astype(np.uint8)
    bm = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    ctx.lut_scan(dev(c), n, dev(lut), 50, bm)
    assert np.array_equal(bits_of(bm, n), (c < 0) | (lut[np.maximum(c, 0)] == 1))
    bm.zero_()
    ctx.lut_scan(dev(c), n, None, 0, bm)  # NULL-only
    assert np.array_equal(bits_of(bm, n), c < 0)
    v = rng.normal(size=n)
    v[rng.random(n) < 0.1] = np.nan
    bm.zero_()
    ctx.range_flag(dev(v), n, -0.5, 1.25, bm)
    with np.errstate(invalid="ignore"):
        assert np.array_equal(bits_of(bm, n), (v < -0.5) | (v > 1.25))
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    from oracle.stats import discretize_numeric
    vmin, vmax = float(np.nanmin(v)), float(np.nanmax(v)) if n > 1 else float(np.nanmin(v)) + 1
    if not np.isnan(vmin) and vmax > vmin:
        from repair.stats_host import discretize_params
        fmin, den = discretize_params("float", vmin, vmax)
        ctx.discretize(dev(v), n, fmin, den, 80, out)
        want = discretize_numeric(v, "float", vmin, vmax, 80)
        assert np.array_equal(out.cpu().numpy(), np.where(np.isnan(want), -1, want).astype(np.int32))


@pytest.mark.parametrize("n", [1, 2, 5, 506, 1001, 65537])
def test_quartiles(ctx, n):
    from oracle.detect import spark_percentile
    rng = np.random.default_rng(n)
    v = np.round(rng.normal(size=n) * 10, 1)
    if n > 4:
        v[rng.random(n) < 0.1] = np.nan
        v[0], v[1] = -0.0, 0.0
    q1, q3, cnt = ctx.quartiles(dev(v), n)
    s = np.sort(v[~np.isnan(v)])
    assert cnt == len(s)
    if cnt:
        assert q1 == spark_percentile(s, 0.25) and q3 == spark_percentile(s, 0.75)


def test_quartiles_kat(ctx):
    # ErrorDetectorSuite.scala:217-231 : 1000 x 100.0 and one 0.0
    v = np.array([100.0] * 1000 + [0.0])
    q1, q3, n = ctx.quartiles(dev(v), len(v))
    assert (q1, q3, n) == (100.0, 100.0, 1001)


@pytest.mark.parametrize("n", [1, 40, 3000, 50001])
def test_dc_const_and_fd(ctx, n):
    rng = np.random.default_rng(n)
    a, b, c = rand_cols(rng, n, [5, 3, 7], null_p=0.1)
    bm = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    ctx.dc_const([dev(a), dev(b), dev(c), dev(c)], [0, 1, 2, 3], [2, 1, 4, 2], n, bm)
    want = (a == 2) & (b != 1) & (c >= 0) & (c < 4) & (c >= 2)
    assert np.array_equal(bits_of(bm, n), want)
    # FD (a, b) -> c with NULL-safe keys, NULL counted as a value of c
    space = 6 * 4
    lo = torch.full((space,), 2 ** 31 - 1, dtype=torch.int32, device="cuda")
    hi = torch.full((space,), -2 ** 31, dtype=torch.int32, device="cuda")
    ctx.dc_fd_build([dev(a), dev(b)], [1, 6], dev(c), n, space, lo, hi)
    bm.zero_()
    ctx.dc_fd_flag([dev(a), dev(b)], [1, 6], n, space, lo, hi, bm)
    key = (a.astype(np.int64) + 1) + (b.astype(np.int64) + 1) * 6
    viol = np.zeros(n, dtype=bool)
    for k in np.unique(key):
        m = key == k
        viol[m] = len(np.unique(c[m])) > 1
    assert np.array_equal(bits_of(bm, n), viol)


@pytest.mark.parametrize("n", [0, 1, 31, 32, 33, 1000, 32768, 32769, 200001])
def test_bitmap_ops(ctx, n):
    rng = np.random.default_rng(n + 1)
    n_words = max((n + 31) // 32, 1)
    a = rng.integers(0, 2 ** 32, size=n_words, dtype=np.uint64).astype(np.uint32)
    b = rng.integers(0, 2 ** 32, size=n_words, dtype=np.uint64).astype(np.uint32)
    if n > 1000:
        a &= rng.integers(0, 2 ** 32, size=n_words, dtype=np.uint64).astype(np.uint32)  # sparser
    da, db = dev(a.view(np.int32)), dev(b.view(np.int32))
    bits_a = np.unpackbits(a.view(np.uint8), bitorder="little")[:n].astype(bool)
    assert ctx.bitmap_count(da, n) == int(bits_a.sum())
    rows = torch.empty(max(int(bits_a.sum()), 1), dtype=torch.int32, device="cuda")
    cnt = ctx.bitmap_to_rows(da, n, rows, int(bits_a.sum()))
    assert cnt == int(bits_a.sum())
    assert np.array_equal(rows[:cnt].cpu().numpy(), np.nonzero(bits_a)[0].astype(np.int32))
    # several popcounts with one round trip, then compaction with the count already known
    bits_b0 = np.unpackbits(b.view(np.uint8), bitorder="little")[:n].astype(bool)
    assert ctx.bitmap_count_many([da, db, da], n) == [int(bits_a.sum()), int(bits_b0.sum()), int(bits_a.sum())]
    rows2 = torch.full((max(cnt, 1),), -7, dtype=torch.int32, device="cuda")
    ctx.bitmap_to_rows_async(da, n, rows2, cnt)
    assert np.array_equal(rows2[:cnt].cpu().numpy(), np.nonzero(bits_a)[0].astype(np.int32))
    # ordered compaction of several bitmaps at once (an empty one among them is skipped)
    zero = torch.zeros_like(da)
    cb = int(bits_b0.sum())
    outs = [torch.full((max(c, 1),), -7, dtype=torch.int32, device="cuda") for c in (cnt, 0, cb)]
    ctx.bitmaps_to_rows_many([da, zero, db], n, outs, [cnt, 0, cb])
    assert np.array_equal(outs[0][:cnt].cpu().numpy(), np.nonzero(bits_a)[0].astype(np.int32))
    assert np.array_equal(outs[2][:cb].cpu().numpy(), np.nonzero(bits_b0)[0].astype(np.int32))
    assert int(outs[1][0]) == -7
    d2 = da.clone()
    ctx.bitmap_or(d2, db, n)
    if n:  # whole words are combined; n == 0 touches nothing
        assert np.array_equal(d2.cpu().numpy().view(np.uint32), a | b)
    ctx.bitmap_andnot(d2, db, n)
    if n:
        assert np.array_equal(d2.cpu().numpy().view(np.uint32), (a | b) & ~b)
    if cnt:
        out = torch.zeros((cnt + 31) // 32, dtype=torch.int32, device="cuda")
        ctx.bitmap_gather(db, rows, cnt, out)
        bits_b = np.unpackbits(b.view(np.uint8), bitorder="little")
        assert np.array_equal(bits_of(out, cnt), bits_b[np.nonzero(bits_a)[0]].astype(bool))
        flags = (rng.random(cnt) < 0.5).astype(np.uint8)
        d3 = da.clone()
        ctx.bitmap_clear_rows(d3, rows, dev(flags), cnt)
        want = bits_a.copy()
        want[np.nonzero(bits_a)[0][flags == 1]] = False
        assert np.array_equal(bits_of(d3, n), want)


@pytest.mark.parametrize("n", [300, 5000, 70000])
def test_pair_presence_and_cooc(ctx, n):
    rng = np.random.default_rng(n)
    doms = [2, 4, 8, 16, 33, 64, 80, 24]
    cols = rand_cols(rng, n, doms, null_p=0.03)
    cols[4] = (cols[5] * 7 % 33).astype(np.int32)  # a dependent column
    dcols = [dev(padded(c)) for c in cols]
    pairs = [(i, j) for i in range(len(doms)) for j in range(i + 1, len(doms))]
    pairs += [(5, 4), (7, 0)]
    px, py = [p[0] for p in pairs], [p[1] for p in pairs]
    tab_off, bit_off = [0], [0]
    for x, y in pairs:
        e = (doms[x] + 1) * (doms[y] + 1)
        tab_off.append(tab_off[-1] + e)
        bit_off.append(bit_off[-1] + (e + 31) // 32)
    out = torch.zeros(tab_off[-1], dtype=torch.int64, device="cuda")
    ctx.cooc(dcols, doms, px, py, tab_off, n, out)
    bits = torch.zeros(bit_off[-1], dtype=torch.int32, device="cuda")
    ctx.pair_presence(dcols, doms, px, py, bit_off, n, 256, (n + 255) // 256, bits)
    h, w = out.cpu().numpy(), bits.cpu().numpy().view(np.uint32)
    for q, (x, y) in enumerate(pairs):
        idx = (cols[x].astype(np.int64) + 1) * (doms[y] + 1) + cols[y] + 1
        want = np.bincount(idx, minlength=(doms[x] + 1) * (doms[y] + 1))
        assert np.array_equal(h[tab_off[q]:tab_off[q + 1]], want), (x, y)
        got_bits = np.unpackbits(w[bit_off[q]:bit_off[q + 1]].view(np.uint8), bitorder="little")[:len(want)]
        assert np.array_equal(got_bits.astype(bool), want > 0), (x, y)
    # a sample only yields lower bounds
    bits.zero_()
    ctx.pair_presence(dcols, doms, px, py, bit_off, n, 64, 2, bits)
    w2 = bits.cpu().numpy().view(np.uint32)
    assert np.all((w2 & ~w) == 0)


def test_pair_tables_split_over_launches(ctx):
    # 2016 pairs of 64 columns x 81 slots do not fit one launch: exercises the range splitting
    rng = np.random.default_rng(3)
    n, k = 4000, 20
    doms = [80] * k
    cols = rand_cols(rng, n, doms, null_p=0.02)
    dcols = [dev(padded(c)) for c in cols]
    pairs = [(i, j) for i in range(k) for j in range(i + 1, k)]
    off = [0]
    for _ in pairs:
        off.append(off[-1] + 81 * 81)
    out = torch.zeros(off[-1], dtype=torch.int64, device="cuda")
    ctx.cooc(dcols, doms, [p[0] for p in pairs], [p[1] for p in pairs], off, n, out)
    h = out.cpu().numpy()
    for q in (0, 57, len(pairs) - 1):
        x, y = pairs[q]
        want = np.bincount((cols[x].astype(np.int64) + 1) * 81 + cols[y] + 1, minlength=81 * 81)
        assert np.array_equal(h[off[q]:off[q + 1]], want)
    assert h.sum() == n * len(pairs)


@pytest.mark.parametrize("n,k", [(1, 3), (77, 7), (5000, 19), (40000, 64)])
def test_gather_rows_masked_and_tile_ops(ctx, n, k):
    rng = np.random.default_rng(n + k)
    N = n * 3 + 5
    cols = rand_cols(rng, N, [9] * k, null_p=0.1)
    vals = [rng.normal(size=N) for _ in range(2)]
    masks = [(rng.random(N) < 0.2) if i % 2 == 0 else None for i in range(k)]
    n_words = (N + 31) // 32

    def to_bm(m):
        w = np.zeros(n_words * 32, dtype=np.uint8)
        w[:N] = m
        return dev(np.packbits(w, bitorder="little").view(np.int32))
    dbms = [None if m is None else to_bm(m) for m in masks]
    rows = np.sort(rng.choice(N, size=n, replace=False)).astype(np.int32)
    tile = torch.empty((n, k), dtype=torch.int32, device="cuda")
    ctx.gather_rows_masked([dev(c) for c in cols], dbms, dev(rows), n, tile)
    want = np.stack([np.where(masks[i][rows], -1, cols[i][rows]) if masks[i] is not None else cols[i][rows]
                     for i in range(k)], axis=1)
    assert np.array_equal(tile.cpu().numpy(), want)
    # the same gather producing the NULL bitmap of every tile column on the way
    words = (n + 31) // 32 + 1
    tile2 = torch.empty((n, k), dtype=torch.int32, device="cuda")
    nulls = torch.zeros((k, words), dtype=torch.int32, device="cuda")
    ctx.gather_rows_masked([dev(c) for c in cols], dbms, dev(rows), n, tile2, null_out=nulls)
    assert np.array_equal(tile2.cpu().numpy(), want)
    for c in range(k):
        assert np.array_equal(bits_of(nulls[c], n), want[:, c] < 0)
    ctile = torch.empty((n, 2), dtype=torch.float64, device="cuda")
    ctx.gather_rows_masked([dev(v) for v in vals], [dbms[0], None], dev(rows), n, ctile, f64=True)
    w0 = np.where(masks[0][rows], np.nan, vals[0][rows])
    got = ctile.cpu().numpy()
    assert np.array_equal(np.isnan(got[:, 0]), np.isnan(w0)) and np.array_equal(got[:, 1], vals[1][rows])
    nb = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    ctx.tile_null_bitmap(tile, n, k, k - 1, nb)
    assert np.array_equal(bits_of(nb, n), want[:, k - 1] < 0)
    ctx.tile_null_bitmap(ctile, n, 2, 0, nb, f64=True)
    assert np.array_equal(bits_of(nb, n), np.isnan(w0))
    # lookup + gathers
    keys = rows[rng.integers(0, n, size=min(n, 50))]
    pos = torch.empty(len(keys), dtype=torch.int32, device="cuda")
    ctx.lookup_sorted(dev(rows), n, dev(keys), len(keys), pos)
    assert np.array_equal(rows[pos.cpu().numpy()], keys)
    miss = torch.empty(1, dtype=torch.int32, device="cuda")
    ctx.lookup_sorted(dev(rows), n, dev(np.array([N + 7], dtype=np.int32)), 1, miss)
    assert int(miss.item()) == -1
    g = torch.empty(len(keys), dtype=torch.int32, device="cuda")
    ctx.gather(dev(cols[1]), dev(keys), len(keys), g)
    assert np.array_equal(g.cpu().numpy(), cols[1][keys])
    ctx.tile_gather(tile, k, 2 % k, pos, len(keys), g)
    assert np.array_equal(g.cpu().numpy(), want[pos.cpu().numpy(), 2 % k])
    cells = dev(np.arange(0, n, 2, dtype=np.int32))
    ctx.tile_fill(tile, k, 0, cells, (n + 1) // 2, 5)
    want[::2, 0] = 5
    assert np.array_equal(tile.cpu().numpy(), want)


def _rand_forest(rng, n_feat, n_classes, n_iter):
    from tools.randforest import random_forest
    thr = [[-0.5, 0.5, 1.5, 2.5, 3.5] for _ in range(n_feat)]
    return random_forest(n_feat, n_classes, n_iter, thr, rng, leaf_scale=0.1)


@pytest.mark.parametrize("n_classes,n_iter,n", [(2, 7, 100), (3, 40, 1000), (64, 300, 700), (1, 25, 333)])
def test_forest_predict_matches_oracle(ctx, n_classes, n_iter, n):
    """Same flat forest in -> bit-identical margins and labels out (CUDA vs oracle)."""
    from oracle.forest import forest_margins, forest_predict
    from repair.forest import DeviceModel
    rng = np.random.default_rng(n_classes * 1000 + n)
    k = 6
    doms = [4, 5, 3, 6, 5, 4]
    tile_np = np.stack(rand_cols(rng, n, doms, null_p=0.1), axis=1).astype(np.int32)
    ctile_np = rng.normal(size=(n, 1)) * 2
    ctile_np[rng.random(n) < 0.1, 0] = np.nan
    regression = n_classes == 1
    feats = ["a%d" % i for i in range(1, k)]
    encoders = [{"attr": f, "type": "ordinal" if i % 2 else "sum", "categories": [-1] * (i == 1) +
                 list(range(doms[i + 1] - (i == 2)))} for i, f in enumerate(feats)]
    encoders.append({"attr": "x", "type": "cont"})
    from repair.forest import encoder_width
    n_feat = sum(encoder_width(e) for e in encoders)
    forest = _rand_forest(rng, n_feat, n_classes, n_iter)
    class_codes = None if regression else list(range(max(n_classes, 2)))
    spec = {"forest": forest, "encoders": encoders, "class_codes": class_codes, "integral": regression}
    tile_cols = {"a%d" % i: i for i in range(k)}
    dict_sizes = {"a%d" % i: doms[i] for i in range(k)}
    dm = DeviceModel(spec, tile_cols, dict_sizes, {"x": 0}, torch.device("cuda", 0))
    tile, ctile = dev(tile_np), dev(ctile_np)
    cells = np.sort(rng.choice(n, size=n // 2, replace=False)).astype(np.int32)
    margins = torch.empty((len(cells), dm.n_seq), dtype=torch.float64, device="cuda")
    ctx.forest_predict(dm.struct, tile, k, ctile, 1, dev(cells), len(cells), 0, margins)
    from repair.forest import encode_matrix
    X = encode_matrix(encoders, {f: tile_np[cells, tile_cols[f]] for f in feats}, {"x": ctile_np[cells, 0]}, dict_sizes)
    want_m = forest_margins(forest, X)
    assert np.array_equal(margins.cpu().numpy(), want_m)  # bit-exact float64
    want = forest_predict(forest, X)
    if regression:
        got = ctile.cpu().numpy()[cells, 0]
        assert np.array_equal(got, np.round(want))
    else:
        got = tile.cpu().numpy()[cells, 0]
        assert np.array_equal(got, want.astype(np.int32))
    untouched = np.setdiff1d(np.arange(n), cells)
    assert np.array_equal(tile.cpu().numpy()[untouched], tile_np[untouched])


def test_domain_score_matches_oracle(ctx):
    """Weak-label domain analysis vs oracle.domain on a random discrete table."""
    from oracle import domain as DM
    from oracle import stats as S
    from oracle.table import OTable
    rng = np.random.default_rng(11)
    n = 3000
    doms = [6, 5, 7]
    cols = rand_cols(rng, n, doms, null_p=0.05)
    cols[0] = np.where(rng.random(n) < 0.8, cols[1] % 6, cols[0]).astype(np.int32)
    names = ["t", "a", "b"]
    otbl = OTable(["tid"] + names, ["int", "str", "str", "str"],
                  [np.arange(n, dtype=np.float64)] + [c.astype(np.int64) for c in cols])
    ndv = {nm: len(np.unique(c[c >= 0])) for nm, c in zip(names, cols)}
    fs = S.compute_freq_stats(otbl, [["t"], ["a"], ["b"], ["t", "a"], ["t", "b"]], 0.0)
    cells_rows = np.sort(rng.choice(n, size=500, replace=False)).astype(np.int32)
    cells = [(int(r), "t", None if cols[0][r] < 0 else int(cols[0][r])) for r in cells_rows]
    for beta in (0.01, 0.5, 0.7):
        want = DM.compute_domain_in_error_cells(otbl, "tid", cells, [], ["t"], fs,
                                                {"t": [("a", 0.1), ("b", 0.2)]}, ndv, 2, 0.0, beta)
        weak_want = DM.weak_labeled_cells(otbl, want)
        hist_t = np.bincount(cols[0].astype(np.int64) + 1, minlength=doms[0] + 1)
        tabs = []
        for c, d in ((cols[1], doms[1]), (cols[2], doms[2])):
            idx = (c.astype(np.int64) + 1) * (doms[0] + 1) + cols[0] + 1
            tabs.append(dev(np.bincount(idx, minlength=(d + 1) * (doms[0] + 1)).astype(np.int64)))
        top1 = torch.empty(len(cells), dtype=torch.int32, device="cuda")
        prob = torch.empty(len(cells), dtype=torch.float64, device="cuda")
        weak = torch.empty(len(cells), dtype=torch.uint8, device="cuda")
        ctx.domain_score(dev(cells_rows), len(cells), dev(cols[0]), doms[0], [dev(cols[1]), dev(cols[2])],
                         [doms[1], doms[2]], tabs, dev(hist_t.astype(np.int64)), [0, 0], n, beta, top1, prob, weak)
        t1, pr, wk = top1.cpu().numpy(), prob.cpu().numpy(), weak.cpu().numpy()
        by_row = {r: d for r, _, _, d in want}
        for i, r in enumerate(cells_rows.tolist()):
            d = by_row[r]
            if d:
                assert t1[i] == d[0][0] and abs(pr[i] - d[0][1]) < 1e-12
            else:
                assert t1[i] == -1
            assert bool(wk[i]) == ((r, "t") in weak_want)


@pytest.mark.parametrize("n_classes,n_iter,n,extra", [(2, 9, 300, []), (5, 33, 2000, []), (64, 300, 1500, []),
                                                        (3, 40, 3000, [11, 10, 9, 8, 7])])
def test_forest_predict_ranked_matches_oracle_and_generic(ctx, n_classes, n_iter, n, extra):
    """All-discrete model: rank-coded kernel == generic kernel == oracle, margins bit-exact.
    `extra` attributes push the feature count past what the 512-cell wide tile holds, so that the
    256-cell wide kernels run."""
    from oracle.forest import forest_margins, forest_predict
    from repair.forest import DeviceModel, encode_matrix, encoder_width
    from tools.randforest import random_forest
    rng = np.random.default_rng(n_classes + n)
    doms = [7, 4, 30, 3, 9, 64, 2, 12] + extra
    k = len(doms)
    tile_np = np.stack(rand_cols(rng, n, doms, null_p=0.08), axis=1).astype(np.int32)
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
    assert dm.ranked is not None
    cells = np.sort(rng.choice(n, size=n * 2 // 3, replace=False)).astype(np.int32)
    X = encode_matrix(encoders, {nm: tile_np[cells, tile_cols[nm]] for nm in names[1:]}, {}, dict_sizes)
    want_m, want = forest_margins(forest, X), forest_predict(forest, X)
    assert (dm.ranked.n_slots > 72) == bool(extra) and dm.ranked.n_slots <= 148   # 512- / 256-cell wide tiles
    for variant in ("auto", "bytes", "wide8", "wide16", "generic"):
        dm.ranked.layout = {"auto": 0, "bytes": 1, "wide8": 2, "wide16": 3, "generic": 0}[variant]
        tile = dev(tile_np)
        margins = torch.empty((len(cells), dm.n_seq), dtype=torch.float64, device="cuda")
        dm.predict(ctx, tile, k, None, 0, dev(cells), len(cells), 0, margins, force_generic=variant == "generic")
        assert np.array_equal(margins.cpu().numpy(), want_m), variant
        assert np.array_equal(tile.cpu().numpy()[cells, 0], want.astype(np.int32)), variant


def test_tile_null_bitmaps_and_rows_after_count(ctx):
    rng = np.random.default_rng(9)
    n, k = 70001, 19
    tile_np = np.stack(rand_cols(rng, n, [5] * k, null_p=0.2), axis=1).astype(np.int32)
    words = (n + 31) // 32 + 1
    out = torch.zeros((k, words), dtype=torch.int32, device="cuda")
    ctx.tile_null_bitmaps(dev(tile_np), n, k, words, out)
    for c in range(k):
        assert np.array_equal(bits_of(out[c], n), tile_np[:, c] < 0)
        cnt = ctx.bitmap_count(out[c], n)
        rows = torch.empty(cnt, dtype=torch.int32, device="cuda")
        ctx.bitmap_rows_after_count(out[c], n, rows, cnt)
        assert np.array_equal(rows.cpu().numpy(), np.nonzero(tile_np[:, c] < 0)[0].astype(np.int32))


def test_dc_fd_large_key_space_uses_global_tables(ctx):
    rng = np.random.default_rng(12)
    n = 60000
    a, b, c = rand_cols(rng, n, [200, 150, 9], null_p=0.05)
    space = 201 * 151  # > 8192: global-atomic path
    lo = torch.full((space,), 2 ** 31 - 1, dtype=torch.int32, device="cuda")
    hi = torch.full((space,), -2 ** 31, dtype=torch.int32, device="cuda")
    ctx.dc_fd_build([dev(a), dev(b)], [1, 201], dev(c), n, space, lo, hi)
    bm = torch.zeros((n + 31) // 32, dtype=torch.int32, device="cuda")
    ctx.dc_fd_flag([dev(a), dev(b)], [1, 201], n, space, lo, hi, bm)
    key = (a.astype(np.int64) + 1) + (b.astype(np.int64) + 1) * 201
    order = np.argsort(key, kind="stable")
    ks, cs = key[order], c[order]
    starts = np.r_[0, np.nonzero(np.diff(ks))[0] + 1]
    mn, mx = np.minimum.reduceat(cs, starts), np.maximum.reduceat(cs, starts)
    viol_key = dict(zip(ks[starts].tolist(), (mn != mx).tolist()))
    want = np.array([viol_key[kk] for kk in key.tolist()])
    assert np.array_equal(bits_of(bm, n), want)


@pytest.mark.parametrize("n", [1, 15, 16, 1000, 128 * 4001])
def test_widen_u8(ctx, n):
    rng = np.random.default_rng(n)
    codes = rng.integers(-1, 254, size=n).astype(np.int32)
    n_pad = (n + 127) // 128 * 128
    src = np.full(n_pad, 255, dtype=np.uint8)
    src[:n] = codes.astype(np.uint8)
    dst = torch.empty(n_pad, dtype=torch.int32, device="cuda")
    ctx.widen_u8(dev(src), n_pad, dst)
    want = np.full(n_pad, -1, dtype=np.int32)
    want[:n] = codes
    assert np.array_equal(dst.cpu().numpy(), want)


@pytest.mark.parametrize("n_classes,n,n_iter", [(3, 1500, 12), (2, 900, 15), (1, 1200, 10), (17, 2500, 6)])
def test_gbdt_trainer_matches_oracle_bit_for_bit(ctx, n_classes, n, n_iter):
    """dr_gbdt_train == oracle/gbdt.py: identical tree structures, thresholds and float64 leaf values."""
    from oracle import gbdt as OG
    from repair import gbdt as PG
    rng = np.random.default_rng(n_classes * 31 + n)
    doms = [4, 9, 3, 6, 30, 2, 12]
    vals = [np.sort(rng.choice(np.arange(-3, 40), size=d, replace=False)).astype(np.float64) for d in doms]
    n_bins = np.array([d + 1 for d in doms], dtype=np.int32)
    bins = np.stack([rng.integers(0, d + 1, size=n) for d in doms], axis=1).astype(np.uint8)
    sig = (bins[:, 0].astype(int) * 3 + bins[:, 4] + (bins[:, 1] > 4) * 5)
    if n_classes == 1:
        y = sig * 0.37 + rng.normal(size=n)
        w = np.ones(n)
    else:
        y = ((sig + rng.integers(0, 2, size=n)) % n_classes).astype(np.int64)
        w = PG.class_weights(y, n_classes, balanced=True)
    lr, depth = 0.1, 5
    want = OG.to_flat_forest(OG.train(bins, n_bins, y, n_classes, w if n_classes > 1 else None, n_iter, lr, depth,
                                      num_leaves=15, min_data_in_leaf=10), vals, len(doms))
    got = PG.train_gpu(ctx, torch.device("cuda", 0), bins, n_bins, vals, y, n_classes, w, n_iter, lr, depth,
                       num_leaves=15, min_data_in_leaf=10)
    for k in ("tree_seq", "tree_offset", "feature", "missing_left", "left", "right"):
        assert np.array_equal(np.asarray(got[k]), np.asarray(want[k])), k
    assert np.array_equal(got["threshold"], want["threshold"])
    assert np.array_equal(got["value"], want["value"])          # bit-exact float64 leaves
    assert np.array_equal(got["baseline"], want["baseline"])
    assert (np.asarray(want["feature"]) >= 0).sum() > n_iter    # the trees actually split


@pytest.mark.parametrize("n_classes,kw", [
    (3, dict(reg_lambda=1.5)), (2, dict(colsample_bytree=0.4)), (4, dict(subsample=0.7, subsample_freq=2)),
    (1, dict(reg_lambda=0.3, colsample_bytree=0.6, subsample=0.8, subsample_freq=1, num_leaves=9, min_data_in_leaf=4,
             min_sum_hessian=0.5)),
    (5, dict(reg_lambda=4.0, colsample_bytree=0.15, subsample=0.55, subsample_freq=3, num_leaves=32))])
def test_gbdt_trainer_tuned_parameters_match_oracle_bit_for_bit(ctx, n_classes, kw):
    """The parameters of the reference's search space (train.py:148-156) in the trainer and in its
    specification: reg_lambda, colsample_bytree (hashed feature subsets), subsample / subsample_freq
    (hashed bags), num_leaves, min_child_samples, min_child_weight."""
    from oracle import gbdt as OG
    from repair import gbdt as PG
    rng = np.random.default_rng(n_classes * 17)
    n, n_iter = 900, 7
    doms = [4, 9, 3, 6, 30, 2, 12, 5]
    vals = [np.sort(rng.choice(np.arange(-3, 40), size=d, replace=False)).astype(np.float64) for d in doms]
    n_bins = np.array([d + 1 for d in doms], dtype=np.int32)
    bins = np.stack([rng.integers(0, d + 1, size=n) for d in doms], axis=1).astype(np.uint8)
    sig = (bins[:, 0].astype(int) * 3 + bins[:, 4] + (bins[:, 1] > 4) * 5)
    if n_classes == 1:
        y, w = sig * 0.37 + rng.normal(size=n), np.ones(n)
    else:
        y = ((sig + rng.integers(0, 2, size=n)) % n_classes).astype(np.int64)
        w = PG.class_weights(y, n_classes, balanced=True)
    args = dict(num_leaves=15, min_data_in_leaf=10)
    args.update(kw)
    want = OG.to_flat_forest(OG.train(bins, n_bins, y, n_classes, w if n_classes > 1 else None, n_iter, 0.1, 5, **args),
                             vals, len(doms))
    got = PG.train_gpu(ctx, torch.device("cuda", 0), bins, n_bins, vals, y, n_classes, w, n_iter, 0.1, 5, **args)
    for k in ("tree_seq", "tree_offset", "feature", "missing_left", "left", "right"):
        assert np.array_equal(np.asarray(got[k]), np.asarray(want[k])), k
    assert np.array_equal(got["threshold"], want["threshold"])
    assert np.array_equal(got["value"], want["value"])
    assert (np.asarray(want["feature"]) >= 0).sum() > n_iter


def test_byte_stager_double_buffering(ctx):
    """Successive byte batches reach the resident table in order, with the next copy in flight."""
    from repair.table import ByteStager, DeviceTable, EncodedTable
    rng = np.random.default_rng(5)
    n, k = 1000, 5
    table = EncodedTable.from_codes("tid", ["a%d" % i for i in range(k)], [np.zeros(0, np.int32)] * k, [200] * k,
                                    row_ids=np.arange(n, dtype=np.int64))
    table.n_rows = n
    dt = DeviceTable(table, torch.device("cuda", 0), codes=torch.zeros((k, 1024), dtype=torch.int32, device="cuda"))
    stager = ByteStager(dt, ctx)
    batches = []
    for _ in range(4):
        b = rng.integers(0, 256, size=(k, 1024)).astype(np.uint8)
        batches.append(torch.from_numpy(b).pin_memory())
    stager.prefetch(batches[0])
    for i in range(4):
        stager.next_batch()
        if i + 1 < 4:
            stager.prefetch(batches[i + 1])      # overlaps whatever is done with batch i
        got = dt.codes.cpu().numpy()
        want = batches[i].numpy().astype(np.int32)
        want[want == 255] = -1
        assert np.array_equal(got, want), i
    stager.next_batch(batches[2])                # nothing prefetched: copied on the spot
    want = batches[2].numpy().astype(np.int32)
    want[want == 255] = -1
    assert np.array_equal(dt.codes.cpu().numpy(), want)
