"""Oracle histogram-GBDT trainer (TEST INFRASTRUCTURE -- see oracle/__init__.py).

The reference trains its repair models with LightGBM 3.3.1 (``train.py:89-229``), which is neither
vendored under /root/reference nor installed: PARITY WITH LIGHTGBM IS UNPINNED.  What is restated
here is the published algorithm with the reference's fixed parameters (``train.py:102-115``:
gbdt, learning_rate 0.01, max_depth 7, n_estimators 300, class_weight balanced; LightGBM defaults
num_leaves 31, min_child_samples 20, min_child_weight 1e-3, reg_lambda 0):

* objectives: softmax (C > 2, hessian factor C/(C-1)), logistic (C == 2), L2 (regression);
  ``boost_from_average`` initial score for binary / regression, 0 for multiclass;
* features are pre-binned (one bin per distinct encoded value + one "missing" bin);
* gradients / hessians are QUANTISED to integers (scale 2^24 / max sample weight) so that every
  histogram is an exact integer sum: split decisions do not depend on summation order, which is
  what lets the CUDA trainer reproduce this trainer bit for bit;
* exp() is evaluated by ``exp_det`` (range reduction + degree-13 Horner polynomial, separate
  multiplies and adds) on both sides for the same reason;
* growth: level by level up to ``max_depth``; at every level each leaf proposes its best split
  (gain = GL^2/HL + GR^2/HR - G^2/H, both directions tried for the missing bin, constraints
  min_data_in_leaf / min_sum_hessian), proposals are applied in order of decreasing gain while the
  tree has fewer than ``num_leaves`` leaves;
* leaf value = -G/(H + lambda) * learning_rate.

The seven parameters the reference tunes with hyperopt (``train.py:148-156``) are all honoured:
``num_leaves``, ``min_child_samples`` (min_data_in_leaf), ``min_child_weight`` (min_sum_hessian),
``reg_lambda`` (added to every hessian sum of the gain and of the leaf value), ``colsample_bytree``
(feature f takes part in tree (iteration, sequence) iff a counter-based hash of (seed, iteration,
sequence, f) falls below the fraction; the feature with the smallest hash always takes part) and
``subsample`` / ``subsample_freq`` (every ``subsample_freq`` iterations a new bag: row i is in the bag
iff a hash of (seed, bag number, i) falls below the fraction; rows outside the bag contribute neither
gradients nor counts to the trees of those iterations, their scores are still updated).
"""
import numpy as np

LN2_HI = 6.93147180369123816490e-01
LN2_LO = 1.90821492927058770002e-10
INV_LN2 = 1.44269504088896338700e+00
_COEF = [1.0 / float(np.prod(np.arange(1, k + 1, dtype=np.float64))) if k else 1.0 for k in range(14)]


_M64 = (1 << 64) - 1


def mix64(z):
    """splitmix64 finaliser on Python ints (the CUDA trainer evaluates the same function)."""
    z &= _M64
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
    return z ^ (z >> 31)


def feature_used(seed, it, s, n_features, colsample):
    """bool[F]: the features tree (it, s) may split on."""
    if colsample >= 1.0:
        return np.ones(n_features, dtype=bool)
    thr = int(colsample * float(1 << 24))
    h = [mix64(seed * 0x9E3779B97F4A7C15 + (it << 40) + (s << 20) + f + 1) >> 40 for f in range(n_features)]
    used = np.array([x < thr for x in h], dtype=bool)
    used[int(np.argmin(h))] = True
    return used


def rows_in_bag(seed, bag, n_rows, subsample):
    """bool[n]: the rows of bag number `bag`."""
    thr = int(subsample * float(1 << 24))
    return np.array([(mix64((seed + 1) * 0x9E3779B97F4A7C15 + (bag << 32) + i) >> 40) < thr for i in range(n_rows)],
                    dtype=bool)


def quant_bits(n_rows):
    """Quantisation width: every histogram bin (a sum over <= n_rows rows) must fit a signed 32-bit
    integer, so that the device can use native 32-bit shared-memory atomics."""
    return int(min(24, 30 - int(np.ceil(np.log2(max(n_rows, 2))))))


def exp_det(x):
    """exp(x) with a fixed sequence of IEEE operations (no fused multiply-add)."""
    x = np.clip(np.asarray(x, dtype=np.float64), -700.0, 700.0)
    k = np.rint(x * INV_LN2)
    r = (x - k * LN2_HI) - k * LN2_LO
    p = np.full_like(r, _COEF[13])
    for c in _COEF[12::-1]:
        p = p * r + c
    return np.ldexp(p, k.astype(np.int32))


def softmax_det(scores):
    m = scores.max(axis=1, keepdims=True)
    e = exp_det(scores - m)
    tot = np.zeros(len(scores))
    for k in range(scores.shape[1]):  # sequential, ascending class
        tot = tot + e[:, k]
    return e / tot[:, None]


def sigmoid_det(s):
    e = exp_det(-np.abs(s))
    return np.where(s >= 0, 1.0 / (1.0 + e), e / (1.0 + e))


def train(bins, n_bins, y, n_classes, sample_weight=None, n_iter=300, learning_rate=0.01, max_depth=7,
          num_leaves=31, min_data_in_leaf=20, min_sum_hessian=1e-3, reg_lambda=0.0, colsample_bytree=1.0,
          subsample=1.0, subsample_freq=0, seed=42):
    """bins: uint8 [n, F], value bins 0..n_bins[f]-2, missing bin = n_bins[f]-1.
    y: class index (n_classes >= 2) or float target (n_classes == 1).
    -> dict(init float64[S], trees = list over iterations of list over sequences of node lists
            [(feature, thr_bin, missing_left, left, right, value)], feature -1 = leaf)"""
    bins = np.asarray(bins)
    n, F = bins.shape
    S = 1 if n_classes <= 2 else n_classes
    w = np.ones(n) if sample_weight is None else np.asarray(sample_weight, dtype=np.float64)
    qscale = float(2 ** quant_bits(n)) / float(w.max())
    scores = np.zeros((n, S))
    if n_classes == 1:
        yv = np.asarray(y, dtype=np.float64)
        init = np.array([np.cumsum(yv)[-1] / n])  # sequential sum
        qscale = float(2 ** quant_bits(n)) / max(float(np.abs(yv - init[0]).max()), 1e-300)
    elif n_classes == 2:
        yv = np.asarray(y, dtype=np.float64)
        sw, swy = np.cumsum(w)[-1], np.cumsum(w * yv)[-1]  # sequential sums
        pavg = min(max(swy / sw, 1e-15), 1.0 - 1e-15)
        init = np.array([np.log(pavg / (1.0 - pavg))])
    else:
        init = np.zeros(S)
        onehot = np.zeros((n, S))
        onehot[np.arange(n), np.asarray(y, dtype=np.int64)] = 1.0
    scores += init[None, :]
    offs = np.zeros(F + 1, dtype=np.int64)
    offs[1:] = np.cumsum(n_bins)
    flat = bins.astype(np.int64) + offs[:-1][None, :]      # [n, F] global bin ids
    trees = []
    lam_q = float(reg_lambda) * qscale
    bagging = subsample < 1.0 and subsample_freq > 0
    in_bag = np.ones(n, dtype=bool)
    for it in range(n_iter):
        if bagging and it % subsample_freq == 0:
            in_bag = rows_in_bag(seed, it // subsample_freq, n, subsample)
        if n_classes == 1:
            g, h = scores - yv[:, None], np.ones((n, 1))
        elif n_classes == 2:
            p = sigmoid_det(scores[:, 0])
            g, h = ((p - yv) * w)[:, None], (p * (1.0 - p) * w)[:, None]
        else:
            p = softmax_det(scores)
            factor = float(S) / float(S - 1)
            g, h = (p - onehot) * w[:, None], factor * p * (1.0 - p) * w[:, None]
        gq, hq = np.rint(g * qscale).astype(np.int64), np.rint(h * qscale).astype(np.int64)
        it_trees = []
        for s in range(S):
            nodes = [[-1, 0, 0, 0, 0, 0.0]]
            node_of = np.zeros(n, dtype=np.int64)
            used = feature_used(seed, it, s, F, colsample_bytree)
            sums = {0: (int(gq[in_bag, s].sum()), int(hq[in_bag, s].sum()), int(in_bag.sum()))}
            active, n_leaves = [0], 1
            for depth in range(max_depth):
                props = []
                for leaf in active:
                    rows = np.nonzero((node_of == leaf) & in_bag)[0]
                    G, H, cnt = sums[leaf]
                    if cnt < 2 * min_data_in_leaf or H <= 0:
                        continue
                    idx = flat[rows].reshape(-1)
                    hg = np.zeros(offs[-1], dtype=np.int64)
                    hh = np.zeros(offs[-1], dtype=np.int64)
                    hc = np.zeros(offs[-1], dtype=np.int64)
                    np.add.at(hg, idx, np.repeat(gq[rows, s], F))
                    np.add.at(hh, idx, np.repeat(hq[rows, s], F))
                    np.add.at(hc, idx, 1)
                    best = None
                    parent = (float(G) * float(G)) / (float(H) + lam_q)
                    for f in range(F):
                        nb = int(n_bins[f])
                        if nb < 3 or not used[f]:
                            continue
                        o = int(offs[f])
                        mg, mh, mc = int(hg[o + nb - 1]), int(hh[o + nb - 1]), int(hc[o + nb - 1])
                        cg = ch = cc = 0
                        for t in range(nb - 2):  # split after value bin t
                            cg += int(hg[o + t]); ch += int(hh[o + t]); cc += int(hc[o + t])
                            for ml in (0, 1):
                                GL, HL, CL = (cg + mg, ch + mh, cc + mc) if ml else (cg, ch, cc)
                                GR, HR, CR = G - GL, H - HL, cnt - CL
                                if CL < min_data_in_leaf or CR < min_data_in_leaf:
                                    continue
                                if HL < min_sum_hessian * qscale or HR < min_sum_hessian * qscale:
                                    continue
                                gain = ((float(GL) * float(GL)) / (float(HL) + lam_q) +
                                        (float(GR) * float(GR)) / (float(HR) + lam_q)) - parent
                                if gain > 0.0 and (best is None or gain > best[0]):
                                    best = (gain, f, t, ml, GL, HL, CL)
                    if best is not None:
                        props.append((best, leaf))
                props.sort(key=lambda pr: (-pr[0][0], pr[1]))
                new_active = []
                for (gain, f, t, ml, GL, HL, CL), leaf in props:
                    if n_leaves >= num_leaves:
                        break
                    G, H, cnt = sums[leaf]
                    li, ri = len(nodes), len(nodes) + 1
                    nodes[leaf][0:5] = [f, t, ml, li, ri]
                    nodes += [[-1, 0, 0, 0, 0, 0.0], [-1, 0, 0, 0, 0, 0.0]]
                    sums[li], sums[ri] = (GL, HL, CL), (G - GL, H - HL, cnt - CL)
                    rows = np.nonzero(node_of == leaf)[0]
                    b = bins[rows, f].astype(np.int64)
                    go_left = np.where(b == n_bins[f] - 1, ml == 1, b <= t)
                    node_of[rows] = np.where(go_left, li, ri)
                    new_active += [li, ri]
                    n_leaves += 1
                active = new_active
                if not active:
                    break
            for i, nd in enumerate(nodes):
                if nd[0] < 0:
                    G, H, _ = sums[i]
                    nd[5] = (-(float(G) / (float(H) + lam_q)) * learning_rate) if H > 0 else 0.0
            vals = np.array([nd[5] for nd in nodes])
            scores[:, s] = scores[:, s] + vals[node_of]
            it_trees.append([tuple(nd) for nd in nodes])
        trees.append(it_trees)
    return {"init": init, "trees": trees, "n_classes": n_classes}


def to_flat_forest(model, bin_values, n_features):
    """-> flat forest (oracle/forest.py layout); threshold = midpoint between the split bin's value
    and the next one in the feature's encoded value space."""
    feature, threshold, missing_left, left, right, value, tree_seq, tree_offset = [], [], [], [], [], [], [], [0]
    for it in model["trees"]:
        for s, nodes in enumerate(it):
            for (f, t, ml, l, r, v) in nodes:
                feature.append(f)
                threshold.append(0.0 if f < 0 else (bin_values[f][t] + bin_values[f][t + 1]) / 2.0)
                missing_left.append(ml if f >= 0 else 0)
                left.append(l if f >= 0 else 0)
                right.append(r if f >= 0 else 0)
                value.append(v if f < 0 else 0.0)
            tree_seq.append(s)
            tree_offset.append(tree_offset[-1] + len(nodes))
    return {"n_features": n_features, "n_classes": model["n_classes"], "baseline": np.asarray(model["init"]),
            "tree_seq": np.asarray(tree_seq, dtype=np.int32), "tree_offset": np.asarray(tree_offset, dtype=np.int64),
            "feature": np.asarray(feature, dtype=np.int32), "threshold": np.asarray(threshold, dtype=np.float64),
            "missing_left": np.asarray(missing_left, dtype=np.uint8), "left": np.asarray(left, dtype=np.int32),
            "right": np.asarray(right, dtype=np.int32), "value": np.asarray(value, dtype=np.float64)}
