"""Device-side Arrow ingest / egress (EncodedTable.from_arrow_device, dr_h2d_copy, dr_index_presence,
dr_index_remap, dr_ids_unique_i64, _arrow_cells_frame) against the host path (EncodedTable.from_arrow,
which tests/test_host_cpu.py pins to from_pandas) and against the pandas-in / pandas-out run."""
import numpy as np
import pandas as pd
import pytest

import parity_utils as PU

pytestmark = pytest.mark.gpu


def _arrow_variants(df):
    """The same table as Arrow hands it over in practice: plain strings, dictionary columns with int8 /
    int32 indices, several chunks with DIFFERENT per-chunk dictionaries, a zero-copy slice (offsets that
    are not byte aligned in the validity bitmap) and dictionaries listing values no row uses."""
    import pyarrow as pa
    plain = pa.Table.from_pandas(df, preserve_index=False)
    out = {"plain": plain}
    names = [c for c in df.columns if c != "tid"]

    def dict_table(tbl, index_type, extra=False):
        cols = {"tid": tbl["tid"]}
        for c in names:
            arr = tbl[c].combine_chunks().dictionary_encode()
            it = index_type if len(arr.dictionary) < (1 << (index_type.bit_width - 1)) or \
                (pa.types.is_unsigned_integer(index_type) and len(arr.dictionary) < 256) else pa.int16()
            if extra:   # an entry nobody uses, placed first so that every index moves
                d = pa.concat_arrays([pa.array(["~unused~"], type=arr.dictionary.type), arr.dictionary])
                idx = pa.compute.add(arr.indices.cast(pa.int32()), 1)
                arr = pa.DictionaryArray.from_arrays(idx.cast(it), d)
            else:
                arr = pa.DictionaryArray.from_arrays(arr.indices.cast(it), arr.dictionary)
            cols[c] = arr
        return pa.table(cols)
    out["dict8"] = dict_table(plain, pa.int8())
    out["dict32_unused"] = dict_table(plain, pa.int32(), extra=True)
    out["dict_u8"] = dict_table(plain, pa.uint8())
    n = plain.num_rows
    cuts = [0, n // 3 + 1, n // 3 + 1, 2 * n // 3 + 5, n]      # includes an empty chunk
    parts = [dict_table(plain.slice(a, b - a), pa.int16()) for a, b in zip(cuts[:-1], cuts[1:])]
    out["chunked"] = pa.concat_tables(parts)                   # per-chunk dictionaries differ
    padded = pa.concat_tables([dict_table(plain.slice(0, 3), pa.int16()), dict_table(plain, pa.int16())])
    out["sliced"] = padded.slice(3)                            # first chunk sliced away to nothing
    one = dict_table(pa.concat_tables([plain.slice(0, 5), plain]).combine_chunks(), pa.int8())
    out["offset5"] = one.slice(5)                              # validity / index offset 5 inside ONE chunk
    return out


def _synthetic(n=3000, seed=5):
    rng = np.random.default_rng(seed)
    df = pd.DataFrame({"tid": np.arange(n) * 3 + 7,
                       "a": rng.integers(0, 9, n).astype(str), "b": rng.choice(list("pqrstuvw"), n),
                       "c": rng.integers(0, 200, n).astype(str), "d": rng.integers(0, 4, n).astype(str)})
    df["b"] = np.where(df["a"].astype(int) % 2 == 0, "p", df["b"])      # some structure for the models
    for c, p in (("a", 0.03), ("b", 0.05), ("d", 0.02)):
        df[c] = df[c].where(rng.random(n) > p, None)
    return df


@pytest.mark.parametrize("variant", ["plain", "dict8", "dict32_unused", "dict_u8", "chunked", "sliced", "offset5"])
def test_device_ingest_equals_host_ingest(variant, monkeypatch):
    import torch
    from repair import table as T
    from repair._native import Context
    from repair.table import EncodedTable
    if variant in ("chunked", "dict8"):
        monkeypatch.setattr(T, "INGEST_GROUP_COLS", 1 if variant == "chunked" else 3)   # several pipelined groups
    df = _synthetic()
    tbl = _arrow_variants(df)[variant]
    want = EncodedTable.from_arrow(tbl, "tid")
    ctx = Context(0)
    try:
        timings = {}
        got = EncodedTable.from_arrow_device(tbl, "tid", ctx, torch.device("cuda", 0), threads=3, timings=timings)
        assert got is not None
        t, dt = got
        assert t.names == want.names and t.n_rows == want.n_rows and t.row_id_kind == want.row_id_kind
        assert np.array_equal(np.asarray(t.row_ids), np.asarray(want.row_ids))
        assert np.array_equal(dt.ids.cpu().numpy(), np.asarray(want.row_ids).astype(np.int64))
        for a, b in zip(t.columns, want.columns):
            assert a.kind == b.kind and list(a.dictionary) == list(b.dictionary), a.name
            assert np.array_equal(a.codes, b.codes), a.name
        assert bool((dt.codes[:, t.n_rows:] == -1).all())
        assert timings["ingest_total_s"] > 0
    finally:
        ctx.close()


def test_device_ingest_rejects_and_defers_like_the_host_path():
    import pyarrow as pa
    import torch
    from repair._native import Context
    from repair.table import EncodedTable
    from repair.utils import AnalysisException
    ctx = Context(0)
    dev = torch.device("cuda", 0)
    try:
        dup = pa.table({"tid": [5, 1, 5, 2], "x": ["a", "b", "a", "c"], "y": ["a", "b", "a", "c"]})
        with pytest.raises(AnalysisException, match="Uniqueness does not hold in column 'tid'.*distinct 'tid': 3, # of rows: 4"):
            EncodedTable.from_arrow_device(dup, "tid", ctx, dev)
        # unique but unordered ids: the sort-based check accepts them
        ok = pa.table({"tid": [5, 1, 9, 2], "x": ["a", "b", "a", None], "y": ["a", "b", "a", "c"]})
        t, dt = EncodedTable.from_arrow_device(ok, "tid", ctx, dev)
        assert list(t.by_name["x"].codes) == [0, 1, 0, -1] and list(t.row_ids) == [5, 1, 9, 2]
        # numeric attribute / string row ids: the host path has to take them
        assert EncodedTable.from_arrow_device(pa.table({"tid": [1, 2], "x": ["a", "b"], "v": [1.0, 2.0]}),
                                              "tid", ctx, dev) is None
        assert EncodedTable.from_arrow_device(pa.table({"tid": ["1", "2"], "x": ["a", "b"], "y": ["a", "b"]}),
                                              "tid", ctx, dev) is None
        with pytest.raises(AnalysisException, match="unsupported ones found: boolean"):
            EncodedTable.from_arrow_device(pa.table({"tid": [1, 2], "x": ["a", "b"], "v": [True, False]}),
                                           "tid", ctx, dev)
        with pytest.raises(AnalysisException, match="Column 'nope' does not exist"):
            EncodedTable.from_arrow_device(ok, "nope", ctx, dev)
    finally:
        ctx.close()


def test_threaded_copies_round_trip():
    """dr_h2d_copy / dr_d2h_copy: many buffers of awkward sizes (below, at and above the 4 MB chunk)."""
    import torch
    from repair._native import Context
    ctx = Context(0)
    try:
        rng = np.random.default_rng(0)
        sizes = [1, 17, 4 << 20, (4 << 20) + 1, 9_000_001, 3, 12 << 20]
        host = [rng.integers(0, 255, s, dtype=np.uint8) for s in sizes]
        dev = [torch.zeros(s, dtype=torch.uint8, device="cuda") for s in sizes]
        ctx.h2d_copy([h.ctypes.data for h in host], dev, sizes, threads=5)
        for h, d in zip(host, dev):
            assert np.array_equal(d.cpu().numpy(), h)
        back = [np.zeros(s, dtype=np.uint8) for s in sizes]
        ctx.d2h_copy(dev, [b.ctypes.data for b in back], sizes, threads=4)
        for h, b in zip(host, back):
            assert np.array_equal(h, b)
        ctx.d2h_copy(dev[:2], [b.ctypes.data for b in back[:2]], sizes[:2], threads=1)
    finally:
        ctx.close()


def test_arrow_in_arrow_out_equals_pandas_run():
    """Arrow table in -> pyarrow.Table out (row ids, dictionary-encoded strings, NULLs as Arrow nulls),
    identical to the pandas-in / pandas-out run of the same model; a second run with the frozen models
    of the first gives the same frame again."""
    import pyarrow as pa
    from repair import ConstraintErrorDetector, NullErrorDetector, RepairModel
    df = _synthetic(4000, seed=9)
    dets = lambda: [NullErrorDetector(), ConstraintErrorDetector(constraints="t1&t2&EQ(t1.a,t2.a)&IQ(t1.d,t2.d)")]  # noqa: E731
    opt = lambda m: m.option("model.lgb.n_estimators", "20").option("model.hp.max_evals", "1")  # noqa: E731
    want = opt(RepairModel().setInput(df).setRowId("tid").setErrorDetectors(dets())).run()
    tbl = _arrow_variants(df)["chunked"]
    rm = opt(RepairModel().setArrowInput(tbl).setRowId("tid").setErrorDetectors(dets()))
    got = rm.run()
    assert isinstance(got, pa.Table)
    assert got.column_names == ["tid", "attribute", "current_value", "repaired"]
    assert PU.frame_tuples(got.to_pandas(), "tid") == PU.frame_tuples(want, "tid") and got.num_rows > 100
    assert rm.last_run["ingest_total_s"] > 0 and rm.last_run["egress_s"] > 0
    again = opt(RepairModel().setArrowInput(tbl).setRowId("tid").setErrorDetectors(dets())
                .setFrozenModels(rm.last_run["models"])).run()
    assert again.equals(got)
    cells = opt(RepairModel().setArrowInput(tbl).setRowId("tid").setErrorDetectors(dets())).run(detect_errors_only=True)
    want_cells = opt(RepairModel().setInput(df).setRowId("tid").setErrorDetectors(dets())).run(detect_errors_only=True)
    assert isinstance(cells, pa.Table)
    assert PU.frame_tuples(cells.to_pandas(), "tid") == PU.frame_tuples(want_cells, "tid")
