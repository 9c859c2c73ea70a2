"""Host-side table ingest: a DataFrame is collected ONCE into label-encoded int32 column-major
arrays (sorted dictionaries, NULL = -1) plus float64 side arrays for numeric columns, then copied
to the GPU through pinned staging buffers.  This replaces the Spark views the reference passes by
name between Python and Scala ("Data hand-off").
"""
import numpy as np

from .utils import AnalysisException, cell_to_string

_SUPPORTED_MSG = "tinyint,float,smallint,string,double,int,bigint"  # RepairBase.scala:41-44 order as Spark prints it
ROW_ALIGN = 128  # rows are padded so that every column starts 512-byte aligned (128-bit loads)


class Column:
    """kind: 'str' (discrete) | 'int' | 'float' (both continuous, RepairBase.scala:41-44).
    ``dictionary``: sorted distinct non-NULL values (str objects, or float64 for numerics);
    ``codes``: int32 index into it, -1 = NULL;  ``values``: float64 (NaN = NULL) for numerics."""

    def __init__(self, name, kind, dictionary, codes, values=None):
        self.name, self.kind, self.dictionary, self.codes, self.values = name, kind, dictionary, codes, values

    @property
    def dict_size(self):
        return len(self.dictionary)

    @property
    def continuous(self):
        return self.kind in ("int", "float")

    def code_of(self, value):
        """Dictionary code of a python value, or -2 if it does not occur."""
        if len(self.dictionary) == 0:
            return -2
        if self.kind == "str":
            if getattr(self, "_lut", None) is None:
                self._lut = {str(v): i for i, v in enumerate(self.dictionary)}
            return self._lut.get(str(value), -2)
        try:
            x = float(value)
        except (TypeError, ValueError):
            return -2
        i = int(np.searchsorted(self.dictionary, x))
        return i if i < len(self.dictionary) and self.dictionary[i] == x else -2

    def rank_bounds(self, value):
        """(lower_bound, upper_bound) ranks of a constant in the sorted dictionary: ``col < c`` is
        ``code < lower`` and ``col > c`` is ``code >= upper``."""
        if self.kind == "str":
            keys = [str(v) for v in self.dictionary]
            import bisect
            return bisect.bisect_left(keys, str(value)), bisect.bisect_right(keys, str(value))
        x = float(value)
        return int(np.searchsorted(self.dictionary, x, "left")), int(np.searchsorted(self.dictionary, x, "right"))

    def strings(self):
        """CAST(dictionary entry AS STRING) for every entry (regex LUTs, output frames)."""
        if self.kind == "str":
            return [str(v) for v in self.dictionary]
        return [cell_to_string(self.kind, float(v)) for v in self.dictionary]

    def decode(self, codes):
        """codes -> list of CAST(.. AS STRING) python values (None for NULL)."""
        strs = self.strings()
        return [None if c < 0 else strs[c] for c in np.asarray(codes).tolist()]


def _encode_numeric(name, kind, arr):
    vals = np.asarray(arr, dtype=np.float64)
    nul = np.isnan(vals)
    uniq = np.unique(vals[~nul])
    codes = np.full(len(vals), -1, dtype=np.int32)
    codes[~nul] = np.searchsorted(uniq, vals[~nul]).astype(np.int32)
    return Column(name, kind, uniq, codes, vals)


def _encode_strings(name, series):
    import pandas as pd
    obj = series.astype(object)
    isn = pd.isna(obj).to_numpy()
    strs = np.array([None if n else str(v) for v, n in zip(obj.tolist(), isn.tolist())], dtype=object)
    present = strs[~isn]
    uniq = np.array(sorted(set(present.tolist())), dtype=object)
    lut = {v: i for i, v in enumerate(uniq.tolist())}
    codes = np.full(len(strs), -1, dtype=np.int32)
    if len(present):
        codes[~isn] = np.fromiter((lut[v] for v in present.tolist()), dtype=np.int32, count=len(present))
    return Column(name, "str", uniq, codes, None)


class EncodedTable:
    """The collected input: row ids + encoded columns (row id excluded)."""

    def __init__(self, row_id, row_ids, row_id_kind, columns, name="input"):
        self.row_id, self.row_ids, self.row_id_kind = row_id, row_ids, row_id_kind
        self.columns = list(columns)
        self.name = name
        self.n_rows = len(row_ids)
        self.by_name = {c.name: c for c in self.columns}
        self.row_offset = 0      # first global row of this shard
        self.n_rows_global = self.n_rows

    @property
    def names(self):
        return [c.name for c in self.columns]

    @property
    def continuous_attrs(self):
        return [c.name for c in self.columns if c.continuous]

    # ---- constructors --------------------------------------------------------------------------
    @classmethod
    def from_pandas(cls, df, row_id, name="input"):
        """checkInputTable (RepairApi.scala:34-67): type gate, >= 3 columns, unique row id."""
        import pandas as pd
        if row_id not in df.columns:
            raise AnalysisException("Column '{}' does not exist in table '{}'".format(row_id, name))
        kinds, bad = {}, []
        for c in df.columns:
            k = df[c].dtype.kind
            if k == "b":
                bad.append("boolean")
            elif k in "iu":
                kinds[c] = "int"
            elif k == "f":
                kinds[c] = "float"
            elif k in "OUS" or str(df[c].dtype) in ("string", "str", "category"):
                kinds[c] = "str"
            elif k == "M":
                bad.append("timestamp")
            else:
                bad.append(str(df[c].dtype))
        if bad:
            raise AnalysisException("Supported types are {}, but unsupported ones found: {}".format(
                _SUPPORTED_MSG, ",".join(bad)))
        # object columns holding only numbers (e.g. nullable ints collected from Spark) are numeric
        for c, k in list(kinds.items()):
            if k == "str" and df[c].dtype == object:
                non_null = [v for v in df[c].tolist() if v is not None and not (isinstance(v, float) and v != v)]
                if non_null and all(isinstance(v, (int, np.integer)) and not isinstance(v, bool) for v in non_null):
                    kinds[c] = "int"
                elif non_null and all(isinstance(v, (int, float, np.integer, np.floating)) and
                                      not isinstance(v, bool) for v in non_null):
                    kinds[c] = "float"
        if not len(df.columns) >= 3:
            raise AnalysisException("A least three columns (`{}` columns + two more ones) in table '{}'".format(
                row_id, name))
        n_distinct = int(df[row_id].nunique(dropna=False))
        if n_distinct != len(df):
            raise AnalysisException(
                "Uniqueness does not hold in column '{}' of table '{}' (# of distinct '{}': {}, # of rows: {})".format(
                    row_id, name, row_id, n_distinct, len(df)))
        cols = []
        for c in df.columns:
            if c == row_id:
                continue
            if kinds[c] == "str":
                cols.append(_encode_strings(c, df[c]))
            else:
                arr = pd.to_numeric(df[c], errors="coerce").to_numpy(dtype=np.float64, na_value=np.nan)
                cols.append(_encode_numeric(c, kinds[c], arr))
        return cls(row_id, df[row_id].to_numpy(), kinds[row_id], cols, name)

    @classmethod
    def from_arrow(cls, tbl, row_id, name="input"):
        """Same gate and encoding as from_pandas for a ``pyarrow.Table`` (what ``spark_df.toArrow()`` or
        ``pyarrow.parquet.read_table(path, read_dictionary=[...])`` return), without a Python object
        per cell: strings are dictionary-encoded by Arrow's C++ kernels -- columns that already arrive
        dictionary-encoded (Parquet dictionary pages) are NOT decoded first -- and only the
        dictionaries (one entry per distinct value) are sorted and turned into Python strings."""
        import pyarrow as pa
        import pyarrow.compute as pc
        if row_id not in tbl.column_names:
            raise AnalysisException("Column '{}' does not exist in table '{}'".format(row_id, name))
        kinds, bad = {}, []
        for f in tbl.schema:
            t = f.type.value_type if pa.types.is_dictionary(f.type) else f.type
            if pa.types.is_boolean(t):
                bad.append("boolean")
            elif pa.types.is_integer(t):
                kinds[f.name] = "int"
            elif pa.types.is_floating(t):
                kinds[f.name] = "float"
            elif pa.types.is_string(t) or pa.types.is_large_string(t):
                kinds[f.name] = "str"
            elif pa.types.is_timestamp(t):
                bad.append("timestamp")
            elif pa.types.is_date(t):
                bad.append("date")
            else:
                bad.append(str(t))
        if bad:
            raise AnalysisException("Supported types are {}, but unsupported ones found: {}".format(
                _SUPPORTED_MSG, ",".join(bad)))
        if not tbl.num_columns >= 3:
            raise AnalysisException("A least three columns (`{}` columns + two more ones) in table '{}'".format(
                row_id, name))
        n_distinct = int(pc.count_distinct(tbl[row_id], mode="all").as_py())
        if n_distinct != tbl.num_rows:
            raise AnalysisException(
                "Uniqueness does not hold in column '{}' of table '{}' (# of distinct '{}': {}, # of rows: {})".format(
                    row_id, name, row_id, n_distinct, tbl.num_rows))

        def plain(col):
            col = col.unify_dictionaries() if pa.types.is_dictionary(col.type) else col
            return col.combine_chunks() if col.num_chunks != 1 else col.chunk(0)

        cols = []
        for f in tbl.schema:
            if f.name == row_id:
                continue
            arr = plain(tbl[f.name])
            if kinds[f.name] != "str":
                if pa.types.is_dictionary(arr.type):
                    arr = arr.dictionary_decode()
                vals = pc.cast(arr, pa.float64()).to_numpy(zero_copy_only=False)
                cols.append(_encode_numeric(f.name, kinds[f.name], np.asarray(vals, dtype=np.float64)))
                continue
            if not pa.types.is_dictionary(arr.type):
                arr = pc.dictionary_encode(arr)
            idx = arr.indices
            valid = np.asarray(idx.is_valid())
            raw = np.asarray(idx.fill_null(0).to_numpy(zero_copy_only=False), dtype=np.int64)
            entries = np.array(arr.dictionary.to_pylist(), dtype=object)
            # a Parquet dictionary page may list values the rows never use: keep what occurs, sorted
            used = np.zeros(len(entries), dtype=bool)
            used[raw[valid]] = True
            used &= np.array([e is not None for e in entries], dtype=bool) if len(entries) else used
            keep = np.nonzero(used)[0]
            order = keep[np.argsort(np.array([str(entries[i]) for i in keep], dtype=object), kind="stable")] \
                if len(keep) else keep
            rank = np.full(len(entries), -1, dtype=np.int32)
            rank[order] = np.arange(len(order), dtype=np.int32)
            codes = np.where(valid, rank[raw], -1).astype(np.int32) if len(entries) else \
                np.full(len(raw), -1, dtype=np.int32)
            cols.append(Column(f.name, "str", np.array([str(entries[i]) for i in order], dtype=object), codes, None))
        ids = plain(tbl[row_id])
        if pa.types.is_dictionary(ids.type):
            ids = ids.dictionary_decode()
        return cls(row_id, np.asarray(ids.to_numpy(zero_copy_only=False)), kinds[row_id], cols, name)

    @classmethod
    def from_codes(cls, row_id, names, codes, dict_sizes, name="input", row_ids=None, dictionaries=None):
        """Pre-encoded discrete table (already label-encoded upstream, e.g. Arrow dictionary pages
        or the synthetic generator): ``codes[k]`` int32, -1 = NULL; value c of column k prints as
        ``dictionaries[k][c]`` (default ``"v%03d" % c``)."""
        n = len(codes[0]) if len(codes) else 0
        cols = []
        for i, nm in enumerate(names):
            d = dictionaries[i] if dictionaries is not None else \
                np.array(["v%03d" % c for c in range(int(dict_sizes[i]))], dtype=object)
            cols.append(Column(nm, "str", d, np.ascontiguousarray(codes[i], dtype=np.int32), None))
        if row_ids is None:
            row_ids = np.arange(n, dtype=np.int64)
        if len(names) + 1 < 3:
            raise AnalysisException("A least three columns (`{}` columns + two more ones) in table '{}'".format(
                row_id, name))
        return cls(row_id, row_ids, "int", cols, name)

    def row_id_strings(self, positions):
        ids = self.row_ids[np.asarray(positions, dtype=np.int64)]
        return ids

    def unify(self, dist):
        """Row-sharded input (every rank collected its own rows): makes the dictionaries GLOBAL -- the
        sorted union of the shards' dictionaries, so that a code means the same value on every GPU and
        the count tensors of the shards add up -- and records the shard's place in the global table.
        Only dictionaries (one entry per distinct value) and row counts cross the wire."""
        import torch.distributed as td
        local = [(c.name, c.kind, list(c.dictionary) if c.kind == "str" else np.asarray(c.dictionary, dtype=np.float64))
                 for c in self.columns]
        gathered = [None] * dist.world
        td.all_gather_object(gathered, (self.n_rows, local), group=dist.group)
        names = [nm for nm, _, _ in local]
        for n_r, cols in gathered:
            if [nm for nm, _, _ in cols] != names:
                raise AnalysisException("the shards of table '{}' do not have the same columns".format(self.name))
        counts = [n_r for n_r, _ in gathered]
        cols = []
        for i, c in enumerate(self.columns):
            kinds = {g[1][i][1] for g in gathered}
            kind = c.kind if len(kinds) == 1 else ("float" if kinds <= {"int", "float"} else "str")
            if kind == "str":
                merged = np.array(sorted(set(str(v) for g in gathered for v in g[1][i][2])), dtype=object)
                mine = np.array([str(v) for v in c.dictionary], dtype=object)
                lut = np.searchsorted(merged.astype(str), mine.astype(str)).astype(np.int32) if len(mine) else \
                    np.zeros(0, dtype=np.int32)
                values = None
            else:
                merged = np.unique(np.concatenate([np.asarray(g[1][i][2], dtype=np.float64) for g in gathered]))
                lut = np.searchsorted(merged, np.asarray(c.dictionary, dtype=np.float64)).astype(np.int32)
                values = c.values
            codes = np.where(c.codes >= 0, np.r_[lut, np.int32(-1)][c.codes], -1).astype(np.int32) if len(lut) else \
                np.full(len(c.codes), -1, dtype=np.int32)
            cols.append(Column(c.name, kind, merged, codes, values))
        t = EncodedTable(self.row_id, self.row_ids, self.row_id_kind, cols, self.name)
        t.row_offset = int(sum(counts[:dist.rank]))
        t.n_rows_global = int(sum(counts))
        return t

    def shard(self, rank, world):
        """Contiguous row shard [lo, hi) for rank `rank` of `world` (global dictionaries kept)."""
        n = self.n_rows
        lo, hi = (n * rank) // world, (n * (rank + 1)) // world
        cols = [Column(c.name, c.kind, c.dictionary, c.codes[lo:hi], None if c.values is None else c.values[lo:hi])
                for c in self.columns]
        t = EncodedTable(self.row_id, self.row_ids[lo:hi], self.row_id_kind, cols, self.name)
        t.row_offset, t.n_rows_global = lo, n
        return t


class DeviceTable:
    """The encoded table resident in HBM: ``codes`` int32 [K][n_pad] (one contiguous column per
    attribute) and ``values`` float64 [Kc][n_pad] for the continuous attributes."""

    def __init__(self, table, device, codes=None, values=None, ctx=None):
        import torch
        self.table = table
        self.device = device
        self.n_rows = table.n_rows
        self.n_pad = (self.n_rows + ROW_ALIGN - 1) // ROW_ALIGN * ROW_ALIGN or ROW_ALIGN
        K = len(table.columns)
        self.cont_index = {c.name: i for i, c in enumerate([c for c in table.columns if c.continuous])}
        if codes is None:
            codes = torch.empty((K, self.n_pad), dtype=torch.int32, device=device)
            staging = torch.empty((self.n_pad,), dtype=torch.int32).pin_memory()
            narrow = narrow_dev = None
            for i, c in enumerate(table.columns):
                if c.dict_size <= 254 and ctx is not None:
                    # small dictionaries cross PCIe as one byte per cell and are widened on the device
                    if narrow is None:
                        narrow = torch.empty((self.n_pad,), dtype=torch.uint8).pin_memory()
                        narrow_dev = torch.empty((self.n_pad,), dtype=torch.uint8, device=device)
                    narrow[:self.n_rows].copy_(torch.from_numpy(np.ascontiguousarray(c.codes).astype(np.uint8)))
                    narrow[self.n_rows:].fill_(255)
                    narrow_dev.copy_(narrow, non_blocking=True)
                    ctx.widen_u8(narrow_dev, self.n_pad, codes[i])
                else:
                    staging[:self.n_rows].copy_(torch.from_numpy(np.ascontiguousarray(c.codes)))
                    staging[self.n_rows:].fill_(-1)
                    codes[i].copy_(staging, non_blocking=True)
                torch.cuda.current_stream().synchronize()
        self.codes = codes
        if values is None and self.cont_index:
            values = torch.empty((len(self.cont_index), self.n_pad), dtype=torch.float64, device=device)
            staging = torch.empty((self.n_pad,), dtype=torch.float64).pin_memory()
            for c in table.columns:
                if c.continuous:
                    staging[:self.n_rows].copy_(torch.from_numpy(np.array(c.values, dtype=np.float64)))  # (own, writable copy)
                    staging[self.n_rows:].fill_(float("nan"))
                    values[self.cont_index[c.name]].copy_(staging, non_blocking=True)
                    torch.cuda.current_stream().synchronize()
        self.values = values
        self.col_index = {c.name: i for i, c in enumerate(table.columns)}

    def col(self, name):
        return self.codes[self.col_index[name]]

    def val(self, name):
        return self.values[self.cont_index[name]]


class ByteStager:
    """Ingest of successive batches of an all-small-dictionary table (every dictionary <= 254 entries,
    one byte per cell in pinned host memory, 255 = NULL) into ONE resident DeviceTable: the host->device
    copy of batch i + 1 runs on its own stream into the other of two device staging buffers while
    batch i is being processed, so PCIe time leaves the critical path; what stays on it is the
    widening pass (dr_widen_u8: 1 B read + 4 B written per cell)."""

    def __init__(self, device_table, ctx):
        import torch
        self.torch, self.dt, self.ctx = torch, device_table, ctx
        shape = tuple(device_table.codes.shape)
        self.stage = [torch.empty(shape, dtype=torch.uint8, device=device_table.device) for _ in range(2)]
        self.copied = [torch.cuda.Event(), torch.cuda.Event()]     # H2D of the buffer finished
        self.consumed = [torch.cuda.Event(), torch.cuda.Event()]   # widening pass read the buffer
        self.stream = torch.cuda.Stream(device=device_table.device)
        self.pending = [False, False]
        self.n = 0                                                  # batches handed to the table so far

    def prefetch(self, host_bytes):
        """Start copying the NEXT batch (pinned uint8 [K][n_pad]) to the device."""
        j = (self.n + (1 if self.pending[self.n % 2] else 0)) % 2
        assert not self.pending[j], "both staging buffers are in flight"
        self.stream.wait_event(self.consumed[j])
        with self.torch.cuda.stream(self.stream):
            self.stage[j].copy_(host_bytes, non_blocking=True)
            self.copied[j].record(self.stream)
        self.pending[j] = True

    def next_batch(self, host_bytes=None):
        """Make the oldest prefetched batch (or `host_bytes`, copied now) the table's content."""
        j = self.n % 2
        if not self.pending[j]:
            assert host_bytes is not None, "nothing prefetched"
            self.prefetch(host_bytes)
        main = self.torch.cuda.current_stream()
        main.wait_event(self.copied[j])
        self.ctx.widen_u8(self.stage[j], self.stage[j].numel(), self.dt.codes)
        self.consumed[j].record(main)
        self.pending[j] = False
        self.n += 1


def cells_to_arrow(table, cells):
    """Egress without a Python string per cell: the (row id, attribute, current_value, repaired) frame of
    encoded repair output ``[(attr, row positions, current codes, repaired codes)]`` as a pyarrow.Table
    whose string columns are dictionary arrays over the column dictionaries (one chunk per attribute)."""
    import pyarrow as pa
    names = [a for a, _, _, _ in cells]
    ids, attr_idx, cur, rep = [], [], [], []
    for i, (a, rows, c, r) in enumerate(cells):
        col = table.by_name[a]
        if col.kind != "str":
            raise NotImplementedError("cells_to_arrow takes discrete attributes (dictionary codes)")
        strs = pa.array(col.strings(), type=pa.string())
        ids.append(pa.array(table.row_ids[np.asarray(rows, dtype=np.int64)]))
        attr_idx.append(np.full(len(rows), i, dtype=np.int32))
        for codes, out in ((c, cur), (r, rep)):
            codes = np.asarray(codes, dtype=np.int32)
            out.append(pa.DictionaryArray.from_arrays(pa.array(codes, mask=codes < 0), strs))
    if not cells:
        empty = pa.array([], type=pa.string())
        return pa.table({table.row_id: pa.array(table.row_ids[:0]), "attribute": empty, "current_value": empty,
                         "repaired": empty})
    attribute = pa.DictionaryArray.from_arrays(pa.array(np.concatenate(attr_idx)), pa.array(names, type=pa.string()))
    return pa.table({table.row_id: pa.chunked_array(ids), "attribute": attribute,
                     "current_value": pa.chunked_array(cur), "repaired": pa.chunked_array(rep)})

