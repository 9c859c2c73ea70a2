"""Probability-mass-function modes: ``compute_repair_candidate_prob``, ``compute_repair_prob``,
``compute_repair_score`` and ``maximal_likelihood_repair`` (reference ``model.py:1145-1277``).

The device returns float64 class margins for every predicted cell (``dr_forest_predict*`` with
``out_margin``); everything after that is small per-cell arithmetic on the host."""
import math

import numpy as np


def probabilities(margins):
    """LightGBM's predict_proba: logistic for one sequence (binary), softmax otherwise."""
    m = np.asarray(margins, dtype=np.float64)
    if m.shape[1] == 1:
        p = 1.0 / (1.0 + np.exp(-m[:, 0]))
        return np.stack([1.0 - p, p], axis=1)
    e = np.exp(m - m.max(axis=1, keepdims=True))
    return e / e.sum(axis=1, keepdims=True)


def shape_pmf(cells, opts, cost_function=None):
    """-> [(row id, attribute, (current, its prob), [(class, prob)])]: candidates re-weighted by the
    update cost (if any), sorted by probability (stable), thresholded and cut to top-k
    (model.py:1174-1225)."""
    thr, top_k, weight = opts["repair.pmf.prob_threshold"], opts["repair.pmf.prob_top_k"], opts["repair.pmf.cost_weight"]
    restrict = list(cost_function.targets) if cost_function is not None else []
    shaped = []
    for rid, attr, cur, classes, probs in cells:
        if classes is None:  # continuous attribute: a single certain candidate
            shaped.append((rid, attr, (cur, 0.0), [(probs, 1.0)]))
            continue
        probs = list(probs)[:len(classes)]
        if cost_function is not None:
            if cur and (not restrict or attr in restrict):
                probs = [p * (1.0 / (1.0 + weight * cost_function.compute(cur, c))) for p, c in zip(probs, classes)]
            total = 0.0
            for p in probs:
                total += p
            probs = [p / total for p in probs]
        cur_prob = probs[classes.index(cur)] if cur in classes else 0.0
        order = sorted(range(len(classes)), key=lambda i: -probs[i])
        ranked = [(classes[i], probs[i]) for i in order if probs[i] > thr][:top_k]
        shaped.append((rid, attr, (cur, cur_prob), ranked))
    return shaped


def compute_score(shaped, cost_function):
    """ln(p_repaired / p_current) / (1 + cost)  (model.py:1238-1246)."""
    scored = []
    for rid, attr, (cur, cur_prob), pmf in shaped:
        rep, rep_prob = pmf[0]
        cost = cost_function.compute(cur if cur is not None else rep, rep)
        cost = 256.0 if cost is None else cost
        score = math.log(rep_prob / (cur_prob if cur_prob > 0.0 else 1e-6)) * (1.0 / (1.0 + cost))
        scored.append((rid, attr, cur, rep, score))
    return scored


def _percentile(sorted_vals, p):
    pos = (len(sorted_vals) - 1) * p
    lo, hi = int(math.floor(pos)), int(math.ceil(pos))
    if lo == hi:
        return float(sorted_vals[lo])
    return (hi - pos) * float(sorted_vals[lo]) + (pos - lo) * float(sorted_vals[hi])


def maximal_likelihood_repair(scored, delta):
    """Keeps the repairs whose score reaches the (1 - delta/n) percentile (model.py:1259-1277)."""
    n = len(scored)
    thres = _percentile(sorted(s[4] for s in scored), min(1.0, 1.0 - delta / n))
    return [s[:4] for s in scored if s[4] >= thres]
