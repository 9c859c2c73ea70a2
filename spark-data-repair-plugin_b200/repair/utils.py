"""Small host-side helpers shared by the API mirror.

The argument checker and the option reader keep the reference's user-visible behaviour (the exact
``TypeError`` / ``ValueError`` texts its tests assert, ``python/repair/tests/test_model.py:98-229``
against ``python/repair/utils.py:50-75,149-216``) but are written from scratch.
"""
import functools
import inspect
import logging
import math
import os
import time
import typing
from decimal import Decimal


def setup_logger():
    logger = logging.getLogger("repair")
    logger.addHandler(logging.NullHandler())
    return logger


_logger = setup_logger()


class AnalysisException(Exception):
    """Raised where the reference surfaces Spark's ``AnalysisException`` through Py4J
    (``RepairApi.scala:40-60``).  If pyspark is importable its class is used instead."""


try:  # pragma: no cover - pyspark is not installed in the build image
    from pyspark.sql.utils import AnalysisException as _SparkAnalysisException  # type: ignore

    AnalysisException = _SparkAnalysisException  # noqa: F811
except Exception:
    pass


def is_testing():
    return os.environ.get("SPARK_TESTING") is not None


def to_list_str(values, sep=",", quote=False):
    return sep.join("'{}'".format(v) if quote else str(v) for v in values)


def get_option_value(opts, key, default_value, type_class=str, validator=None, err_msg=None):
    """Typed read of a string option: cast/validation failures warn and fall back to the default,
    or raise under SPARK_TESTING (utils.py:50-75 semantics)."""
    assert type(default_value) is type_class, "key={}".format(key)
    if key not in opts:
        return default_value
    raw = opts[key]
    problem = None
    try:
        value = type_class(raw)
    except Exception:
        problem = 'Failed to cast "{}" into {} data: key={}'.format(raw, type_class.__name__, key)
    else:
        if validator is not None and not validator(value):
            problem = "{}, got {}".format(str(err_msg).format(key), value)
    if problem is None:
        return value
    if is_testing():
        raise ValueError(problem)
    _logger.warning(problem)
    return default_value


def _type_name(tp):
    origin = getattr(tp, "__origin__", None)
    if origin is list:
        return "list[{}]".format(_type_name(tp.__args__[0]))
    if origin is dict:
        return "dict[{},{}]".format(_type_name(tp.__args__[0]), _type_name(tp.__args__[1]))
    return getattr(tp, "__name__", str(tp))


def _matches(value, tp):
    origin = getattr(tp, "__origin__", None)
    if origin is list:
        return type(value) is list and all(_matches(v, tp.__args__[0]) for v in value)
    if origin is dict:
        return type(value) is dict and all(_matches(k, tp.__args__[0]) for k in value) \
            and all(_matches(v, tp.__args__[1]) for v in value.values())
    if origin is typing.Union:
        return any(_matches(value, a) for a in tp.__args__)
    return isinstance(value, tp)


def _first_mismatch(values, tp):
    for v in values:
        if not _matches(v, tp):
            return type(v).__name__
    return None


def argtype_check(f):
    """Validates call arguments against the annotations of ``f`` (Union / list / dict / plain)."""
    sig = inspect.signature(f)

    @functools.wraps(f)
    def wrapper(self, *args, **kwargs):
        for name, value in sig.bind(self, *args, **kwargs).arguments.items():
            tp = sig.parameters[name].annotation
            if tp is inspect.Parameter.empty or name == "self":
                continue
            origin = getattr(tp, "__origin__", None)
            got = type(value).__name__
            if origin is typing.Union:
                if not _matches(value, tp):
                    want = "/".join(_type_name(a) for a in tp.__args__)
                    raise TypeError("`{}` should be provided as {}, got {}".format(name, want, got))
            elif origin is list:
                want = _type_name(tp)
                if type(value) is not list:
                    raise TypeError("`{}` should be provided as {}, got {}".format(name, want, got))
                bad = _first_mismatch(value, tp.__args__[0])
                if bad is not None:
                    raise TypeError("`{}` should be provided as {}, got {} in elements".format(name, want, bad))
            elif origin is dict:
                want = _type_name(tp)
                if type(value) is not dict:
                    raise TypeError("`{}` should be provided as {}, got {}".format(name, want, got))
                bad = _first_mismatch(value.keys(), tp.__args__[0])
                if bad is not None:
                    raise TypeError("`{}` should be provided as {}, got {} in keys".format(name, want, bad))
                bad = _first_mismatch(value.values(), tp.__args__[1])
                if bad is not None:
                    raise TypeError("`{}` should be provided as {}, got {} in values".format(name, want, bad))
            elif not _matches(value, tp):
                raise TypeError("`{}` should be provided as {}, got {}".format(name, _type_name(tp), got))
        return f(self, *args, **kwargs)

    return wrapper


def elapsed_time(f):
    @functools.wraps(f)
    def wrapper(self, *args, **kwargs):
        t0 = time.time()
        ret = f(self, *args, **kwargs)
        return ret, time.time() - t0

    return wrapper


def phase(name):
    """Logs "Elapsed time (name: ...)" like the reference's ``spark_job_group`` (utils.py:130-146)."""
    def decorator(f):
        @functools.wraps(f)
        def wrapper(self, *args, **kwargs):
            t0 = time.time()
            ret = f(self, *args, **kwargs)
            _logger.info("Elapsed time (name: {}) is {}(s)".format(name, time.time() - t0))
            return ret
        return wrapper
    return decorator


def double_to_string(v):
    """``CAST(double AS STRING)`` (Java ``Double.toString``): plain notation in [1e-3, 1e7),
    computerised scientific notation outside, shortest round-trip digits."""
    v = float(v)
    if v != v:
        return "NaN"
    if math.isinf(v):
        return "Infinity" if v > 0 else "-Infinity"
    if v == 0.0:
        return "-0.0" if math.copysign(1.0, v) < 0 else "0.0"
    sign = "-" if v < 0 else ""
    mag = abs(v)
    digits, exp = Decimal(repr(mag)).as_tuple()[1:]
    ds = "".join(str(d) for d in digits).rstrip("0") or "0"
    exp += len(digits) - len(ds)
    lead = len(ds) + exp - 1
    if 1e-3 <= mag < 1e7:
        if lead < 0:
            return sign + "0." + "0" * (-lead - 1) + ds
        whole = ds[:lead + 1].ljust(lead + 1, "0")
        return sign + whole + "." + (ds[lead + 1:] or "0")
    return "{}{}.{}E{}".format(sign, ds[0], ds[1:] or "0", lead)


def cell_to_string(kind, v):
    """``CAST(cell AS STRING)`` by column kind; ``None`` stays ``None``."""
    if v is None:
        return None
    if kind == "str":
        return v
    if v != v:
        return None
    return str(int(v)) if kind == "int" else double_to_string(v)


def row_positions(row_ids, wanted):
    """Vectorised join of `wanted` row ids against the table's `row_ids` (compared as strings, like the
    reference's CAST(.. AS STRING) joins): -> (positions int64, found bool) per wanted id.  No Python dict
    over the table: integer ids take a binary search (directly when the ids are increasing, else through one
    argsort), everything else pandas' hashed ``Index.get_indexer``."""
    import numpy as np
    import pandas as pd
    ids = np.asarray(row_ids)
    want = np.asarray(list(wanted) if not isinstance(wanted, np.ndarray) else wanted)
    n = len(ids)
    if n == 0 or len(want) == 0:
        return np.zeros(len(want), dtype=np.int64), np.zeros(len(want), dtype=bool)
    if ids.dtype.kind in "iu":
        as_int = pd.to_numeric(pd.Series(want), errors="coerce")
        ok = as_int.notna().to_numpy() & (as_int.fillna(0) % 1 == 0).to_numpy()
        w = as_int.fillna(0).to_numpy().astype(np.int64)
        if bool(np.all(ids[1:] > ids[:-1])):
            pos = np.searchsorted(ids, w)
            pos = np.minimum(pos, n - 1)
            found = ok & (ids[pos] == w)
        else:
            order = np.argsort(ids, kind="stable")
            p = np.minimum(np.searchsorted(ids[order], w), n - 1)
            pos = order[p]
            found = ok & (ids[pos] == w)
        return pos.astype(np.int64), found
    idx = pd.Index(pd.Series(ids).astype(str))
    pos = idx.get_indexer(pd.Index(pd.Series(want).astype(str)))
    return np.maximum(pos, 0).astype(np.int64), pos >= 0
