"""Oracle flat-forest evaluator and feature encoders (TEST INFRASTRUCTURE -- see oracle/__init__.py).

The reference only ever calls ``model.predict(X)`` / ``predict_proba(X)`` / ``classes_`` on a
LightGBM 3.3.1 model (``model.py:1118-1133``) after ``category_encoders`` 2.2.2 transforms
(``model.py:701-729,1111-1116``).  Neither package is vendored or installed: both are restated
from their published behaviour and are PARITY-UNPINNED against the originals; the evaluator is
pinned against scikit-learn's ``HistGradientBoosting*.predict`` (tests/test_oracle_forest.py).

Flat forest (dict of NumPy arrays), one tree sequence per class:
  n_classes      1 = regression, 2 = binary (ONE sequence), C > 2 = C sequences
  baseline       float64[S]        S = number of sequences (1 for regression/binary)
  tree_seq       int32[T]          sequence (class) of tree t
  tree_offset    int64[T+1]        node range of tree t in the node arrays
  feature        int32[M]          split feature of node m, -1 = leaf
  threshold      float64[M]        go left iff x <= threshold
  missing_left   uint8[M]          NaN goes left iff 1
  left, right    int32[M]          child node index RELATIVE to the tree's first node
  value          float64[M]        leaf value (0 on internal nodes)
Margins: raw[s] = baseline[s] + sum over the trees of sequence s IN TREE ORDER (sequential
float64 adds -- the CUDA kernel uses the same order so margins are bit-identical).
Label: regression raw[0]; binary 1 iff raw[0] > 0; multiclass argmax_s raw[s], ties -> lowest s
(= ``classes_[argmax(predict_proba)]`` barring floating-point collisions in softmax).
"""
import numpy as np


def forest_margins(forest, X):
    """X: float64 [n, F'] (NaN = missing) -> raw float64 [n, S]."""
    n = X.shape[0]
    S = len(forest["baseline"])
    raw = np.tile(np.asarray(forest["baseline"], dtype=np.float64), (n, 1))
    off = forest["tree_offset"]
    feat, thr = forest["feature"], forest["threshold"]
    ml, left, right, val = forest["missing_left"], forest["left"], forest["right"], forest["value"]
    rows = np.arange(n)
    for t in range(len(forest["tree_seq"])):
        base = int(off[t])
        node = np.zeros(n, dtype=np.int64)
        active = feat[base + node] >= 0
        while active.any():
            idx = base + node[active]
            f = feat[idx]
            x = X[rows[active], f]
            isn = np.isnan(x)
            with np.errstate(invalid="ignore"):
                go_left = np.where(isn, ml[idx] == 1, x <= thr[idx])
            node[active] = np.where(go_left, left[idx], right[idx])
            active = feat[base + node] >= 0
        raw[:, int(forest["tree_seq"][t])] += val[base + node]
    return raw


def forest_predict(forest, X):
    """-> class index int64[n] (classification) or float64[n] (regression)."""
    raw = forest_margins(forest, X)
    C = int(forest["n_classes"])
    if C == 1:
        return raw[:, 0]
    if C == 2:
        return (raw[:, 0] > 0).astype(np.int64)
    return np.argmax(raw, axis=1).astype(np.int64)


def forest_proba(forest, X):
    raw = forest_margins(forest, X)
    C = int(forest["n_classes"])
    if C == 2:
        p = 1.0 / (1.0 + np.exp(-raw[:, 0]))
        return np.stack([1.0 - p, p], axis=1)
    e = np.exp(raw - raw.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


# --------------------------------------------------------------------------------------
# Encoders: restatement of category_encoders 2.2.2 as used at model.py:701-729
#   ce.SumEncoder(cols=small_domain_columns, handle_unknown='impute')
#   ce.OrdinalEncoder(cols=other_discrete_columns, handle_unknown='impute')
# 'impute' is not a 2.x option name, so no unknown-value branch fires: unknown -> NaN.
# handle_missing keeps its default 'value'.
# --------------------------------------------------------------------------------------
def first_seen_categories(values):
    """pandas ``Series.unique()`` order; NULL (None) is a category where it first appears."""
    seen, out = set(), []
    for v in values:
        k = ("__null__",) if v is None else v
        if k not in seen:
            seen.add(k)
            out.append(v)
    return out


def encoder_kind(attr, continuous, domain_stats, small_domain_threshold):
    if attr in continuous:
        return "cont"
    return "sum" if int(domain_stats[attr]) < small_domain_threshold else "ordinal"


def encoder_width(enc):
    if enc["type"] == "sum":
        k = len(enc["categories"])
        return k - 1 if k >= 2 else 0
    return 1


def encode_value(enc, v):
    """One cell -> list of floats."""
    t = enc["type"]
    if t == "cont":
        return [np.nan if v is None else float(v)]
    cats = enc["categories"]
    null_seen = any(c is None for c in cats)
    if t == "ordinal":
        if v is None:
            return [float(cats.index(None) + 1)] if null_seen else [-2.0]
        return [float(cats.index(v) + 1)] if v in cats else [np.nan]
    # sum (deviation) contrast: level i < k-1 -> e_i ; last level -> all -1
    k = len(cats)
    if k < 2:
        return []
    if v is None and not null_seen:
        return [0.0] * (k - 1)  # ordinal -2 -> zeros row (handle_missing='value')
    if v not in cats:
        return [np.nan] * (k - 1)  # ordinal -1 has no row -> reindex gives NaN
    i = cats.index(v)
    if i == k - 1:
        return [-1.0] * (k - 1)
    row = [0.0] * (k - 1)
    row[i] = 1.0
    return row


def encode_rows(encoders, columns):
    """columns: {attr: list of python values (None = NULL)} -> float64 [n, F']"""
    n = len(next(iter(columns.values()))) if columns else 0
    width = sum(encoder_width(e) for e in encoders)
    X = np.empty((n, width), dtype=np.float64)
    j = 0
    for e in encoders:
        w = encoder_width(e)
        if w == 0:
            continue
        col = columns[e["attr"]]
        # one encode_value per DISTINCT cell value (hashed in C), then a table look-up for the whole column
        if n:
            import pandas as pd
            where, uniques = pd.factorize(np.asarray(col, dtype=object), use_na_sentinel=True)   # None -> -1
            table = [encode_value(e, v) for v in uniques.tolist()] + [encode_value(e, None)]
            where = np.where(where < 0, len(table) - 1, where)
            X[:, j:j + w] = np.asarray(table, dtype=np.float64).reshape(len(table), w)[where]
        j += w
    return X
