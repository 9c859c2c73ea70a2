"""Repair-model producer.

The reference trains LightGBM 3.3.1 under a hyperopt TPE search (``train.py:89-229``); neither
package is available offline, and SURVEY.md section 8 keeps training out of the first bar (the model
is *consumed* by the hot path).  This module produces the same kind of object -- a gradient-boosted
forest with the reference's fixed parameters -- with scikit-learn's histogram GBDT and flattens
it into the exchange format of ``forest.py``.  Training parity with LightGBM is therefore
UNPINNED; inference parity is exact for whatever forest is handed over.
"""
import logging

import numpy as np

from .utils import get_option_value

_logger = logging.getLogger("repair")

# option keys accepted for API compatibility with train.py:36-86 (key, default, type, validator, message)
_OPTS = [
    ("model.lgb.boosting_type", "gbdt", str, lambda v: v in ["gbdt", "dart", "goss", "rf"],
     "`{}` should be in ['gbdt', 'dart', 'goss', 'rf']"),
    ("model.lgb.class_weight", "balanced", str, None, None),
    ("model.lgb.learning_rate", 0.01, float, lambda v: v > 0.0, "`{}` should be positive"),
    ("model.lgb.max_depth", 7, int, None, None),
    ("model.lgb.max_bin", 255, int, None, None),
    ("model.lgb.reg_alpha", 0.0, float, lambda v: v >= 0.0, "`{}` should be greater than or equal to 0.0"),
    ("model.lgb.min_split_gain", 0.0, float, lambda v: v >= 0.0, "`{}` should be greater than or equal to 0.0"),
    ("model.lgb.n_estimators", 300, int, lambda v: v > 0, "`{}` should be positive"),
    ("model.lgb.importance_type", "gain", str, lambda v: v in ["split", "gain"], "`{}` should be in ['split', 'gain']"),
    ("model.cv.n_splits", 3, int, lambda v: v >= 3, "`{}` should be greater than 2"),
    ("model.hp.timeout", 0, int, None, None),
    ("model.hp.max_evals", 100000000, int, lambda v: v > 0, "`{}` should be positive"),
    ("model.hp.no_progress_loss", 50, int, lambda v: v > 0, "`{}` should be positive"),
]
train_option_keys = [o[0] for o in _OPTS]
_OPT = {o[0]: o for o in _OPTS}


def _get(opts, key):
    return get_option_value(opts, *_OPT[key])


def validate_options(opts):
    for key in train_option_keys:
        _get(opts, key)


def flatten_sklearn(est, n_features, n_classes):
    """HistGradientBoosting{Classifier,Regressor} -> flat forest (numerical splits only)."""
    feature, threshold, missing_left, left, right, value = [], [], [], [], [], []
    tree_seq, tree_offset = [], [0]
    for per_iter in est._predictors:
        for s, pred in enumerate(per_iter):
            nd = pred.nodes
            if nd["is_categorical"].any():
                raise ValueError("categorical splits are not part of the exchange format")
            leaf = nd["is_leaf"].astype(bool)
            feature.append(np.where(leaf, -1, nd["feature_idx"]).astype(np.int32))
            threshold.append(np.where(leaf, 0.0, nd["num_threshold"]).astype(np.float64))
            missing_left.append(nd["missing_go_to_left"].astype(np.uint8))
            left.append(np.where(leaf, 0, nd["left"]).astype(np.int32))
            right.append(np.where(leaf, 0, nd["right"]).astype(np.int32))
            value.append(np.where(leaf, nd["value"], 0.0).astype(np.float64))
            tree_seq.append(s)
            tree_offset.append(tree_offset[-1] + len(nd))
    cat = (lambda xs, dt: np.concatenate(xs).astype(dt) if xs else np.zeros(0, dtype=dt))
    return {
        "n_features": int(n_features), "n_classes": int(n_classes),
        "baseline": np.asarray(est._baseline_prediction, dtype=np.float64).reshape(-1).copy(),
        "tree_seq": np.asarray(tree_seq, dtype=np.int32), "tree_offset": np.asarray(tree_offset, dtype=np.int64),
        "feature": cat(feature, np.int32), "threshold": cat(threshold, np.float64),
        "missing_left": cat(missing_left, np.uint8), "left": cat(left, np.int32), "right": cat(right, np.int32),
        "value": cat(value, np.float64),
    }


def build_model(X, y, is_discrete, num_class, opts):
    """-> (flat forest, class labels ascending or None) or (None, None) when training fails
    (the reference swallows failures into PoorModel(None), train.py:227-229).

    Fixed parameters follow train.py:102-115; the seven parameters the reference tunes with
    hyperopt stay at LightGBM's defaults (num_leaves 31, min_child_samples 20, no subsampling,
    reg_lambda 0)."""
    from sklearn.ensemble import HistGradientBoostingClassifier, HistGradientBoostingRegressor
    try:
        common = dict(
            learning_rate=_get(opts, "model.lgb.learning_rate"),
            max_iter=_get(opts, "model.lgb.n_estimators"),
            max_depth=_get(opts, "model.lgb.max_depth") if _get(opts, "model.lgb.max_depth") > 0 else None,
            max_leaf_nodes=31, min_samples_leaf=20, max_bins=min(255, max(2, _get(opts, "model.lgb.max_bin"))),
            l2_regularization=0.0, early_stopping=False, random_state=42)
        X = np.asarray(X, dtype=np.float64)
        if X.shape[1] == 0:
            X = np.zeros((X.shape[0], 1))
        if is_discrete:
            cw = _get(opts, "model.lgb.class_weight")
            est = HistGradientBoostingClassifier(class_weight="balanced" if cw == "balanced" else None, **common)
            est.fit(X, np.asarray(y))
            classes = [c.item() if hasattr(c, "item") else c for c in est.classes_]
            return flatten_sklearn(est, X.shape[1], len(classes)), classes
        est = HistGradientBoostingRegressor(**common)
        est.fit(X, np.asarray(y, dtype=np.float64))
        return flatten_sklearn(est, X.shape[1], 1), None
    except Exception as e:  # noqa: BLE001
        _logger.warning("Failed to build a stat model because: {}".format(e))
        return None, None


def _random_shape(rng, max_depth, max_leaves):
    """One random tree topology: arrays (left, right, depth) with -1 children on leaves."""
    left, right, depth = [-1], [-1], [0]
    leaves = [0]
    while len(leaves) < max_leaves:
        cand = [n for n in leaves if depth[n] < max_depth]
        if not cand:
            break
        n = cand[int(rng.integers(len(cand)))]
        left[n], right[n] = len(left), len(left) + 1
        for _ in range(2):
            left.append(-1), right.append(-1), depth.append(depth[n] + 1)
        leaves.remove(n)
        leaves += [len(left) - 2, len(left) - 1]
    return np.array(left), np.array(right)


def random_forest(n_features, n_classes, n_iter, feature_thresholds, rng, max_depth=7, max_leaves=31,
                  leaf_scale=0.01, n_shapes=16):
    """Random-init forest of the reference's architecture (n_iter boosting rounds x one tree per
    class, depth <= max_depth, <= max_leaves leaves): used by bench.py where no training data of
    the named size can be fitted in the time budget.  ``feature_thresholds[f]`` lists the split
    points that make sense for encoded feature f.  Trees draw their topology from a pool of
    `n_shapes` random shapes; split features, thresholds, NaN directions and leaf values are
    independent per tree."""
    S = 1 if n_classes <= 2 else n_classes
    T = n_iter * S
    usable = np.array([f for f in range(n_features) if len(feature_thresholds[f])], dtype=np.int64)
    shapes = [_random_shape(rng, max_depth if len(usable) else 0, max_leaves) for _ in range(n_shapes)]
    sizes = np.array([len(sh[0]) for sh in shapes])
    shape_of = rng.integers(0, n_shapes, size=T)
    tree_sizes = sizes[shape_of]
    tree_offset = np.zeros(T + 1, dtype=np.int64)
    tree_offset[1:] = np.cumsum(tree_sizes)
    M = int(tree_offset[-1])
    feature = np.full(M, -1, dtype=np.int32)
    threshold = np.zeros(M)
    missing_left = np.zeros(M, dtype=np.uint8)
    left = np.zeros(M, dtype=np.int32)
    right = np.zeros(M, dtype=np.int32)
    value = np.zeros(M)
    max_thr = max([len(t) for t in feature_thresholds] + [1])
    thr_tab = np.zeros((max(n_features, 1), max_thr))
    thr_len = np.ones(max(n_features, 1), dtype=np.int64)
    for f, t in enumerate(feature_thresholds):
        if len(t):
            thr_tab[f, :len(t)] = t
            thr_len[f] = len(t)
    for si, (l, r) in enumerate(shapes):
        trees = np.nonzero(shape_of == si)[0]
        if len(trees) == 0:
            continue
        idx = tree_offset[trees][:, None] + np.arange(len(l))[None, :]        # [n_trees, nodes]
        internal = l >= 0
        n_int = int(internal.sum())
        left[idx] = np.where(internal, l, 0)[None, :]
        right[idx] = np.where(internal, r, 0)[None, :]
        if n_int:
            f = usable[rng.integers(0, len(usable), size=(len(trees), n_int))]
            u = rng.random(size=f.shape)
            th = thr_tab[f, np.minimum((u * thr_len[f]).astype(np.int64), thr_len[f] - 1)]
            feature[idx[:, internal]] = f
            threshold[idx[:, internal]] = th
            missing_left[idx[:, internal]] = rng.integers(0, 2, size=f.shape)
        value[idx[:, ~internal]] = rng.normal(0.0, leaf_scale, size=(len(trees), int((~internal).sum())))
    return {
        "n_features": int(n_features), "n_classes": int(n_classes), "baseline": np.zeros(S),
        "tree_seq": (np.arange(T) % S).astype(np.int32), "tree_offset": tree_offset,
        "feature": feature, "threshold": threshold, "missing_left": missing_left, "left": left, "right": right,
        "value": value,
    }
