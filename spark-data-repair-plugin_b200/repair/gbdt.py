"""GPU histogram-GBDT trainer (host side of ``dr_gbdt_train``): pre-bins the encoded training
sample, computes class weights / initial scores, runs the whole boosting loop on the device and
flattens the result into the exchange format of ``forest.py``.

Eligibility: every feature discrete (label-encoded attribute behind a sum / ordinal encoder) and at
most 128 encoded features; a feature with more than 254 distinct encoded values is binned like
LightGBM's max_bin does (adjacent values share a bin).  Other models (continuous features) use
``train.build_model`` (scikit-learn)."""
import numpy as np

from .forest import encoder_lut

NODE_DTYPE = np.dtype([("feature", np.int16), ("thr_bin", np.uint8), ("missing_left", np.uint8), ("left", np.uint8),
                       ("right", np.uint8), ("pad", np.uint8, (2,)), ("value", np.float64)])
MAX_NODES = 64
MAX_BINS = 254   # real bins per feature (one more is the missing bin; bins travel as bytes)


def quant_bits(n_rows):
    """Quantisation width: every histogram bin (a sum over <= n_rows rows) must fit a signed 32-bit
    integer, so that the device can use native 32-bit shared-memory atomics."""
    return int(min(24, 30 - int(np.ceil(np.log2(max(n_rows, 2))))))


def bin_sample(encoders, sample_codes, dict_sizes):
    """-> (bins uint8 [n, F'], n_bins int32 [F'], bin_values list of float arrays) or None."""
    cols, n_bins, values = [], [], []
    for e in encoders:
        if e["type"] == "cont":
            return None
        lut = encoder_lut(e, dict_sizes[e["attr"]])
        codes = np.asarray(sample_codes[e["attr"]], dtype=np.int64)
        for j in range(lut.shape[1]):
            col = lut[:, j]
            vals = np.unique(col[~np.isnan(col)])
            ok = ~np.isnan(col)
            if len(vals) > MAX_BINS:
                # more distinct values than bins (LightGBM: max_bin = 255, train.py:106): adjacent values
                # share a bin, bins of about equal sample counts; thresholds only fall between bins
                enc = col[codes + 1]
                cnt = np.bincount(np.searchsorted(vals, enc[~np.isnan(enc)]), minlength=len(vals)).astype(np.float64)
                edge = np.floor(np.cumsum(cnt) / max(cnt.sum(), 1.0) * MAX_BINS - 1e-9).astype(np.int64)
                group = np.minimum(np.maximum.accumulate(np.clip(edge, 0, MAX_BINS - 1)), MAX_BINS - 1)
                _, group = np.unique(group, return_inverse=True)      # dense bin ids, in value order
                n_real = int(group.max()) + 1
                upper = np.array([vals[group == b].max() for b in range(n_real)])
                lower = np.array([vals[group == b].min() for b in range(n_real)])
                bin_of = np.full(len(col), n_real, dtype=np.uint8)           # missing bin
                bin_of[ok] = group[np.searchsorted(vals, col[ok])].astype(np.uint8)
                cols.append(bin_of[codes + 1])
                n_bins.append(n_real + 1)
                values.append(np.stack([upper, lower]))
                continue
            bin_of = np.full(len(col), len(vals), dtype=np.uint8)           # missing bin
            bin_of[ok] = np.searchsorted(vals, col[ok]).astype(np.uint8)
            cols.append(bin_of[codes + 1])
            n_bins.append(len(vals) + 1)
            values.append(vals)
    if not cols or len(cols) > 128:
        return None
    return np.stack(cols, axis=1).astype(np.uint8), np.asarray(n_bins, dtype=np.int32), values


def class_weights(y_idx, n_classes, balanced):
    """class_weight='balanced' (train.py:105): n / (C * count[class])."""
    n = len(y_idx)
    if not balanced:
        return np.ones(n)
    cnt = np.bincount(y_idx, minlength=n_classes).astype(np.float64)
    return (float(n) / (float(n_classes) * cnt))[y_idx]


def initial_scores(y, n_classes, weight):
    if n_classes == 1:
        yv = np.asarray(y, dtype=np.float64)
        return np.array([np.cumsum(yv)[-1] / len(yv)])
    if n_classes == 2:
        yv = np.asarray(y, dtype=np.float64)
        sw, swy = np.cumsum(weight)[-1], np.cumsum(weight * yv)[-1]
        pavg = min(max(swy / sw, 1e-15), 1.0 - 1e-15)
        return np.array([np.log(pavg / (1.0 - pavg))])
    return np.zeros(n_classes)


def train_gpu(ctx, device, bins, n_bins, bin_values, y, n_classes, weight, n_iter, learning_rate, max_depth,
              num_leaves=31, min_data_in_leaf=20, min_sum_hessian=1e-3, reg_lambda=0.0, colsample_bytree=1.0,
              subsample=1.0, subsample_freq=0, seed=42):
    """-> flat forest (forest.py layout)."""
    import torch
    from ._native import dr_gbdt_params
    n, F = bins.shape
    S = 1 if n_classes <= 2 else n_classes
    init = initial_scores(y, n_classes, weight)
    if n_classes == 1:
        yv = np.asarray(y, dtype=np.float64)
        qscale = float(2 ** quant_bits(n)) / max(float(np.abs(yv - init[0]).max()), 1e-300)
    else:
        qscale = float(2 ** quant_bits(n)) / float(np.max(weight))
    prm = dr_gbdt_params(n, F, n_classes, n_iter, max_depth, num_leaves, min_data_in_leaf, learning_rate,
                         min_sum_hessian, qscale, float(reg_lambda), float(colsample_bytree), float(subsample),
                         int(subsample_freq), int(seed))
    d_bins = torch.from_numpy(np.ascontiguousarray(bins)).to(device)
    d_yc = torch.from_numpy(np.ascontiguousarray(y, dtype=np.int32)).to(device) if n_classes >= 2 else None
    d_yv = torch.from_numpy(np.ascontiguousarray(y, dtype=np.float64)).to(device) if n_classes == 1 else None
    d_w = torch.from_numpy(np.ascontiguousarray(weight, dtype=np.float64)).to(device)
    ws = torch.empty(ctx.gbdt_workspace_bytes(n, S), dtype=torch.uint8, device=device)
    out_nodes = torch.zeros(n_iter * S * MAX_NODES * NODE_DTYPE.itemsize, dtype=torch.uint8, device=device)
    out_counts = torch.zeros(n_iter * S, dtype=torch.int32, device=device)
    ctx.gbdt_train(prm, d_bins, n_bins, d_yc, d_yv, d_w, init, ws, out_nodes, out_counts)
    nodes = out_nodes.cpu().numpy().view(NODE_DTYPE).reshape(n_iter, S, MAX_NODES)
    counts = out_counts.cpu().numpy().reshape(n_iter, S)
    return flatten(nodes, counts, init, bin_values, F, n_classes)


def flatten(nodes, counts, init, bin_values, n_features, n_classes):
    """device node records -> flat forest; threshold = midpoint between the split bin's value and the
    next one in the feature's encoded value space."""
    n_iter, S, _ = nodes.shape
    sizes = counts.reshape(-1).astype(np.int64)
    tree_offset = np.zeros(len(sizes) + 1, dtype=np.int64)
    tree_offset[1:] = np.cumsum(sizes)
    keep = (np.arange(MAX_NODES)[None, :] < sizes[:, None]).reshape(-1)
    flat = nodes.reshape(-1)[keep]
    feat = flat["feature"].astype(np.int32)
    leaf = feat < 0
    thr = np.zeros(len(flat))
    # a bin holds one encoded value (1-D entry) or a run of adjacent values (2 x n entry: largest / smallest
    # value of every bin): the threshold sits midway between the split bin's largest and the next bin's smallest
    hi_lo = [(v, v) if np.ndim(v) == 1 else (v[0], v[1]) for v in bin_values]
    max_bins = max(len(h) for h, _ in hi_lo)
    tab_hi, tab_lo = np.zeros((n_features, max_bins + 1)), np.zeros((n_features, max_bins + 1))
    for f, (h, l) in enumerate(hi_lo):
        tab_hi[f, :len(h)] = h
        tab_lo[f, :len(l)] = l
    fi, ti = np.where(leaf, 0, feat), flat["thr_bin"].astype(np.int64)
    thr = np.where(leaf, 0.0, (tab_hi[fi, ti] + tab_lo[fi, np.minimum(ti + 1, max_bins)]) / 2.0)
    return {
        "n_features": int(n_features), "n_classes": int(n_classes), "baseline": np.asarray(init, dtype=np.float64),
        "tree_seq": np.tile(np.arange(S, dtype=np.int32), n_iter), "tree_offset": tree_offset,
        "feature": feat, "threshold": thr, "missing_left": np.where(leaf, 0, flat["missing_left"]).astype(np.uint8),
        "left": np.where(leaf, 0, flat["left"]).astype(np.int32),
        "right": np.where(leaf, 0, flat["right"]).astype(np.int32),
        "value": np.where(leaf, flat["value"], 0.0).astype(np.float64),
    }
