"""Oracle table model (TEST INFRASTRUCTURE -- see oracle/__init__.py).

A table is an ordered list of columns; each column is ``kind`` in {"str","int","float"} plus a
NumPy array: object dtype (``None`` = NULL) for strings, float64 (``NaN`` = NULL) for numerics.
The kinds mirror the reference's type sets (``RepairBase.scala:41-44``): integral + fractional
types are "continuous", strings are discrete.
"""
import csv
import math
from collections import OrderedDict

import numpy as np


class AnalysisException(Exception):
    """Stand-in for Spark's AnalysisException (``RepairApi.scala:40-60``)."""


def spark_double_to_string(v):
    """``CAST(double AS STRING)`` = Java ``Double.toString`` (used by ``RepairApi.scala:90``)."""
    from decimal import Decimal
    v = float(v)
    if v != v:
        return "NaN"
    if v in (float("inf"), float("-inf")):
        return "Infinity" if v > 0 else "-Infinity"
    if v == 0:
        return "-0.0" if math.copysign(1.0, v) < 0 else "0.0"
    a = abs(v)
    _, digits, exp = Decimal(repr(a)).as_tuple()
    digits = list(digits)
    while len(digits) > 1 and digits[-1] == 0:
        digits.pop()
        exp += 1
    ds = "".join(map(str, digits))
    e10 = len(ds) + exp - 1  # decimal exponent of the leading digit
    sign = "-" if v < 0 else ""
    if 1e-3 <= a < 1e7:
        if e10 >= 0:
            ip = (ds[: e10 + 1]).ljust(e10 + 1, "0")
            fp = ds[e10 + 1:] or "0"
            return sign + ip + "." + fp
        return sign + "0." + "0" * (-e10 - 1) + ds
    return sign + ds[0] + "." + (ds[1:] or "0") + "E" + str(e10)


def cast_to_string(kind, v):
    """``CAST(cell AS STRING)`` for one cell; NULL stays ``None``."""
    if v is None:
        return None
    if kind == "str":
        return v
    if v != v:
        return None
    if kind == "int":
        return str(int(v))
    return spark_double_to_string(float(v))


class OTable:
    def __init__(self, names, kinds, arrays):
        assert len(names) == len(kinds) == len(arrays)
        self.names = list(names)
        self.kinds = OrderedDict(zip(names, kinds))
        self.cols = OrderedDict(zip(names, arrays))
        self.n_rows = len(arrays[0]) if arrays else 0

    def isnull(self, name):
        a = self.cols[name]
        if self.kinds[name] == "str":
            return np.array([x is None for x in a], dtype=bool) if a.dtype == object else (a < 0)
        return np.isnan(a)

    def strings(self, name):
        """Column as CAST(.. AS STRING) object array."""
        k = self.kinds[name]
        return np.array([cast_to_string(k, x) for x in self.cols[name]], dtype=object)

    def value(self, name, r):
        """One cell as a plain Python value (None = NULL); integral floats become ints."""
        a = self.cols[name]
        v = a[r]
        if a.dtype == object:
            return v
        if self.kinds[name] == "str":  # label-encoded discrete column
            return None if v < 0 else int(v)
        if v != v:
            return None
        return int(v) if float(v).is_integer() else float(v)

    def select(self, names):
        return OTable(names, [self.kinds[n] for n in names], [self.cols[n] for n in names])

    def copy(self):
        return OTable(self.names, list(self.kinds.values()), [a.copy() for a in self.cols.values()])

    def take(self, idx):
        return OTable(self.names, list(self.kinds.values()), [a[idx] for a in self.cols.values()])


def _infer_kind(py_values):
    kinds = set()
    for v in py_values:
        if v is None:
            continue
        if isinstance(v, bool):
            raise AnalysisException("unsupported ones found: boolean")
        if isinstance(v, (int, np.integer)):
            kinds.add("int")
        elif isinstance(v, (float, np.floating)):
            kinds.add("float")
        elif isinstance(v, str):
            kinds.add("str")
        else:
            raise AnalysisException("unsupported ones found: {}".format(type(v).__name__))
    if not kinds:
        return "str"
    if kinds == {"int"}:
        return "int"
    if kinds <= {"int", "float"}:
        return "float"
    return "str"


def _make_array(kind, py_values):
    if kind == "str":
        return np.array([None if v is None else str(v) for v in py_values], dtype=object)
    return np.array([np.nan if v is None else float(v) for v in py_values], dtype=np.float64)


def from_rows(names, rows, kinds=None):
    cols = list(zip(*rows)) if rows else [[] for _ in names]
    if kinds is None:
        kinds = [_infer_kind(c) for c in cols]
    return OTable(names, kinds, [_make_array(k, c) for k, c in zip(kinds, cols)])


def _csv_infer(col):
    """Spark CSV ``inferSchema``: all-int -> int, all-numeric -> double, else string."""
    is_int = is_float = True
    for v in col:
        if v is None:
            continue
        try:
            int(v)
        except ValueError:
            is_int = False
            try:
                float(v)
            except ValueError:
                is_float = False
                break
    if all(v is None for v in col):
        return "str"
    return "int" if is_int else ("float" if is_float else "str")


def from_csv(path, infer_schema=False, kinds=None):
    """Spark ``spark.read.option("header", True).csv``: empty field -> NULL."""
    with open(path, newline="") as f:
        rd = csv.reader(f)
        names = next(rd)
        rows = [[None if x == "" else x for x in r] for r in rd if r]
    cols = list(zip(*rows))
    if kinds is None:
        kinds = [_csv_infer(c) if infer_schema else "str" for c in cols]
    elif isinstance(kinds, dict):
        kinds = [kinds.get(n, _csv_infer(c)) for n, c in zip(names, cols)]
    return OTable(names, kinds, [_make_array(k, c) for k, c in zip(kinds, cols)])


def from_pandas(df):
    names = list(df.columns)
    kinds, arrays = [], []
    for n in names:
        s = df[n]
        if s.dtype.kind in "iu":
            kinds.append("int")
            arrays.append(s.to_numpy(dtype=np.float64))
        elif s.dtype.kind == "f":
            kinds.append("float")
            arrays.append(s.to_numpy(dtype=np.float64))
        else:
            vals = [None if (v is None or (isinstance(v, float) and v != v)) else v for v in s.tolist()]
            k = _infer_kind(vals)
            kinds.append(k)
            arrays.append(_make_array(k, vals))
    return OTable(names, kinds, arrays)


def from_codes(names, codes):
    """Label-encoded discrete table (``codes[k]`` int array, -1 = NULL); values are opaque ints.

    Used for the synthetic configs: the oracle then treats code ``c`` as the string ``"v%03d" % c``
    without materialising Python strings (equality, grouping and order coincide).
    """
    return OTable(names, ["str"] * len(names), [np.asarray(c, dtype=np.int64) for c in codes])


def factorize(tbl, name):
    """-> (codes int64 with -1 = NULL, sorted uniques).  Independent of the product's ingest."""
    a = tbl.cols[name]
    if tbl.kinds[name] == "str" and a.dtype != object:
        a = a.astype(np.int64)
        if a.size and a.max() >= 0:
            uniq = np.unique(a[a >= 0])
        else:
            uniq = np.array([], dtype=np.int64)
        codes = np.searchsorted(uniq, np.where(a >= 0, a, 0)).astype(np.int64)
        return np.where(a >= 0, codes, -1), uniq
    nul = tbl.isnull(name)
    if tbl.kinds[name] == "str":
        vals = np.array([x for x in a[~nul]], dtype=object)
        uniq = np.array(sorted(set(vals.tolist())), dtype=object)
        lut = {v: i for i, v in enumerate(uniq.tolist())}
        codes = np.full(len(a), -1, dtype=np.int64)
        codes[~nul] = [lut[v] for v in vals.tolist()]
        return codes, uniq
    uniq = np.unique(a[~nul])
    codes = np.full(len(a), -1, dtype=np.int64)
    codes[~nul] = np.searchsorted(uniq, a[~nul])
    return codes, uniq
