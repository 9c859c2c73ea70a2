"""Training-data rebalancing (``setTrainingDataRebalancingEnabled(True)``, reference ``train.py:242-293``).

The reference brings every class of a discrete target to the MEDIAN class size: classes below it are
over-sampled with imbalanced-learn's SMOTEN (SMOTE for nominal features), classes above it are cut by
``RandomUnderSampler``.  imbalanced-learn is a third-party dependency that is absent from
``/root/reference`` and from this image, so its algorithms are restated here from their published
behaviour (imblearn 0.8: ``SMOTEN._generate_samples``, ``ValueDifferenceMetric``,
``RandomUnderSampler``); the random draws are seeded (42, like the reference) but come from NumPy's
``default_rng`` rather than imblearn's ``RandomState`` stream -- parity with imblearn's exact samples is
UNPINNED, what is reproduced is the procedure:

* median class size as the target (``train.py:247``); a class below the median is only over-sampled
  when it has more than k = 5 rows (``train.py:262-271``), to ``median`` rows;
* SMOTEN: distances between two rows = sum over features of the Value Difference Metric of their
  categories, VDM(a, b) = sum_classes |P(class | a) - P(class | b)|; a synthetic row starts from a random
  row of the class and takes, per feature, the most frequent category among that row's k nearest
  neighbours inside the class (ties: smallest category);
* classes above the median keep ``median`` rows drawn without replacement.

Everything works on the label-encoded sample (one int code per cell; NULL = -1 is a category like any
other -- the reference's encoders map it to a value as well, so its ``dropna`` split is empty in practice).
"""
import logging

import numpy as np

_logger = logging.getLogger("repair")

K_NEIGHBORS = 5  # SMOTEN default, train.py:260


def _vdm_tables(codes, y_idx, n_classes):
    """Per feature: (distinct values, P(class | value) matrix [n_values, n_classes])."""
    out = []
    for j in range(codes.shape[1]):
        vals, inv = np.unique(codes[:, j], return_inverse=True)
        cnt = np.zeros((len(vals), n_classes), dtype=np.float64)
        np.add.at(cnt, (inv, y_idx), 1.0)
        out.append((vals, cnt / cnt.sum(axis=1, keepdims=True), inv))
    return out


def rebalance(codes, y_values, seed=42):
    """codes: int [n, F] label-encoded features of the training sample, y_values: int [n] class codes.
    -> (row sources int64 [m] into the original sample or -1 for synthetic rows, codes [m, F], y [m])."""
    codes = np.asarray(codes)
    y_values = np.asarray(y_values)
    classes, y_idx = np.unique(y_values, return_inverse=True)
    counts = np.bincount(y_idx, minlength=len(classes))
    median = int(np.median(counts))
    rng = np.random.default_rng(seed)
    keep_rows, new_codes, new_y = [], [], []
    tables = None
    for c, cnt in enumerate(counts.tolist()):
        rows = np.nonzero(y_idx == c)[0]
        if cnt > median:                                       # RandomUnderSampler
            rows = np.sort(rng.choice(rows, size=median, replace=False))
        keep_rows.append(rows)
        if cnt < median:
            if cnt <= K_NEIGHBORS:
                _logger.warning("Over-sampling of '{}' failed because the number of the clean rows is too small: {}"
                                .format(classes[c], cnt))
                continue
            if tables is None:
                tables = _vdm_tables(codes, y_idx, len(classes))
            n_new = median - cnt
            # chunked pairwise distances keep the memory of a large class bounded
            base = rng.integers(0, cnt, size=n_new)
            uniq_base = np.unique(base)
            nn_of = {}
            for s in range(0, len(uniq_base), 256):
                part = uniq_base[s:s + 256]
                dist = np.zeros((len(part), cnt), dtype=np.float64)
                for vals, prob, inv in tables:
                    pa, pb = prob[inv[rows[part]]], prob[inv[rows]]
                    dist += np.abs(pa[:, None, :] - pb[None, :, :]).sum(axis=2)
                dist[np.arange(len(part)), part] = -1.0         # the row itself sorts first and is dropped
                order = np.argsort(dist, axis=1, kind="stable")[:, 1:K_NEIGHBORS + 1]
                for b, nn in zip(part.tolist(), order):
                    nn_of[b] = nn
            synth = np.empty((n_new, codes.shape[1]), dtype=codes.dtype)
            for i, b in enumerate(base.tolist()):
                neigh = codes[rows[nn_of[b]]]                     # [k, F]
                for j in range(codes.shape[1]):
                    v, n_v = np.unique(neigh[:, j], return_counts=True)
                    synth[i, j] = v[np.argmax(n_v)]               # mode; ties -> smallest category
            new_codes.append(synth)
            new_y.append(np.full(n_new, classes[c], dtype=y_values.dtype))
    src = np.concatenate(keep_rows) if keep_rows else np.zeros(0, dtype=np.int64)
    order = np.argsort(src, kind="stable")
    src = src[order]
    out_codes, out_y = codes[src], y_values[src]
    if new_codes:
        out_codes = np.concatenate([out_codes] + new_codes)
        out_y = np.concatenate([out_y] + new_y)
        src = np.concatenate([src, np.full(sum(len(s) for s in new_codes), -1, dtype=np.int64)])
    _logger.info("Rebalanced training data (median={}): #rows={} -> #rows={}".format(median, len(codes), len(out_codes)))
    return src, out_codes, out_y
