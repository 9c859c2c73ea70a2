"""ctypes wrapper of oracle/c/liboracle_kernels.so (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Compiled, OpenMP-threaded versions of the oracle's three heavy loops so that the CPU baseline in
bench.py is timed on all host cores; each is checked against the NumPy oracle in
tests/test_oracle_c.py."""
import ctypes
import os

import numpy as np

_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c", "liboracle_kernels.so")
_lib = None


def available():
    return os.path.exists(_PATH)


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_PATH)
        _lib.o_num_threads.restype = ctypes.c_int
    return _lib


def num_threads():
    return int(lib().o_num_threads())


def use_all_cores():
    """OpenMP threads = the cores this process may run on, whatever OMP_NUM_THREADS says (torchrun
    exports OMP_NUM_THREADS=1 to every rank).  -> the thread count now in force."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    lib().o_set_num_threads(ctypes.c_int(n))
    return num_threads()


def _pp(arrays):
    arr = (ctypes.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])
    return arr


def scan_hist(cols, dom):
    cols = [np.ascontiguousarray(c, dtype=np.int32) for c in cols]
    dom = np.ascontiguousarray(dom, dtype=np.int32)
    hist = np.zeros(int((dom + 1).sum()), dtype=np.int64)
    lib().o_scan_hist(_pp(cols), dom.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(cols)),
                      ctypes.c_int64(len(cols[0])), hist.ctypes.data_as(ctypes.c_void_p))
    return hist


def cooc(cols, dom, px, py):
    cols = [np.ascontiguousarray(c, dtype=np.int32) for c in cols]
    dom = np.ascontiguousarray(dom, dtype=np.int32)
    px, py = np.ascontiguousarray(px, dtype=np.int32), np.ascontiguousarray(py, dtype=np.int32)
    off = np.zeros(len(px) + 1, dtype=np.int64)
    off[1:] = np.cumsum((dom[px].astype(np.int64) + 1) * (dom[py] + 1))
    out = np.zeros(int(off[-1]), dtype=np.int64)
    lib().o_cooc(_pp(cols), dom.ctypes.data_as(ctypes.c_void_p), px.ctypes.data_as(ctypes.c_void_p),
                 py.ctypes.data_as(ctypes.c_void_p), ctypes.c_int(len(px)), off.ctypes.data_as(ctypes.c_void_p),
                 ctypes.c_int64(len(cols[0])), out.ctypes.data_as(ctypes.c_void_p))
    return out, off


def forest_margins(forest, X):
    X = np.ascontiguousarray(X, dtype=np.float64)
    n, f = X.shape
    S = len(forest["baseline"])
    raw = np.empty((n, S), dtype=np.float64)
    a = {k: np.ascontiguousarray(forest[k], dtype=dt) for k, dt in (
        ("baseline", np.float64), ("tree_seq", np.int32), ("tree_offset", np.int64), ("feature", np.int32),
        ("threshold", np.float64), ("missing_left", np.uint8), ("left", np.int32), ("right", np.int32),
        ("value", np.float64))}
    p = lambda x: x.ctypes.data_as(ctypes.c_void_p)  # noqa: E731
    lib().o_forest_margins(ctypes.c_int64(n), ctypes.c_int(f), p(X), ctypes.c_int(S), p(a["baseline"]),
                           ctypes.c_int(len(a["tree_seq"])), p(a["tree_seq"]), p(a["tree_offset"]), p(a["feature"]),
                           p(a["threshold"]), p(a["missing_left"]), p(a["left"]), p(a["right"]), p(a["value"]), p(raw))
    return raw
