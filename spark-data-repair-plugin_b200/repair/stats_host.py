"""Host-side arithmetic over the (tiny) count tensors the kernels produce.

After the scans, everything the reference computes with dozens of small Spark jobs
(``RepairApi.scala:280-477``) is a function of per-column histograms (<= 81 bins) and pair tables
(<= 81 x 81): candidate-pair scoring, conditional entropies and their ordering, the thresholds of
the domain analysis.  These run in float64 on the host, identically on every rank after the count
all-reduce.  Layout conventions: ``hist[a]`` is int64[dom+1] with slot 0 = NULL;
``tables[(x, y)]`` is int64[dom_x+1, dom_y+1] in the orientation the pair was first requested.
"""
import math

import numpy as np

_LOG2 = math.log(2.0)


def log2(v):  # RepairApi.scala:280-282
    return math.log(v) / _LOG2


def having_threshold(n_rows, attr_freq_ratio_threshold):
    """HAVING cnt > (rowCount * threshold).toInt  (RepairApi.scala:255-262); None = no filter."""
    if attr_freq_ratio_threshold > 0.0:
        return int(n_rows * attr_freq_ratio_threshold)
    return None


def apply_having(counts, having):
    if having is None:
        return counts
    return np.where(counts > having, counts, 0)


def _entropy_with_correction(flat_counts, n_rows, ub_domain):
    """-sum p log2 p over the visible groups + the missing-mass term (RepairApi.scala:306-339)."""
    nz = flat_counts[flat_counts > 0]
    h = 0.0
    for c in nz.tolist():
        p = c / n_rows
        h += p * log2(p)
    h = -h
    total = int(nz.sum())
    if n_rows > total:
        ub = max(ub_domain - len(nz), 1)
        avg = max((n_rows - total + 0.0) / ub, 1.0)
        h += -ub * (avg / n_rows) * log2(avg / n_rows)
    return h


def pairwise_entropies(n_rows, hist, tables, pairs, domain_stats, having=None):
    """computePairwiseStats: {x: [(y, H(x|y) = H(x,y) - H(y))] ascending}."""
    if not pairs:
        return {}
    hxy, hy = {}, {}
    for x, y in pairs:
        key = frozenset((x, y))
        if key not in hxy:
            tab = tables[(x, y)] if (x, y) in tables else tables[(y, x)]
            flat = apply_having(np.asarray(tab).reshape(-1), having)
            hxy[key] = _entropy_with_correction(flat, n_rows, domain_stats[x] * domain_stats[y])
        for a in (x, y):
            if a not in hy:
                hy[a] = _entropy_with_correction(apply_having(np.asarray(hist[a]), having), n_rows, domain_stats[a])
    out = {}
    for x, y in pairs:
        out.setdefault(x, []).append((y, hxy[frozenset((x, y))] - hy[y]))
    return {k: sorted(v, key=lambda t: t[1]) for k, v in out.items()}


def candidate_pairs(targets, disc_attrs):
    """{target: [(target, other)]} in discretised-column order (RepairApi.scala:430-431)."""
    return {t: [(t, a) for a in disc_attrs if a != t] for t in targets}


def select_scored(cands, nnz, domain_stats, pairwise_thr, max_attrs):
    """filter(ratio < thr).sortBy(ratio).take(max) with ratio = nnz / (ndv_x * ndv_y); `nnz` maps
    frozenset({x, y}) -> distinct pair count (exact, or a lower bound that already disqualifies)."""
    scored = []
    for (x, y) in cands:
        den = domain_stats[x] * domain_stats[y]
        ratio = (nnz[frozenset((x, y))] + 0.0) / den if den != 0 else float("inf")
        if ratio < pairwise_thr:
            scored.append((ratio, (x, y)))
    scored.sort(key=lambda s: s[0])
    return [p for _, p in scored[:max_attrs]]


def undecided(cands, nnz_lower, domain_stats, pairwise_thr):
    """Candidates a lower bound of nnz cannot yet exclude."""
    out = []
    for (x, y) in cands:
        den = domain_stats[x] * domain_stats[y]
        if den != 0 and (nnz_lower[frozenset((x, y))] + 0.0) / den < pairwise_thr:
            out.append((x, y))
    return out


def tau_for(alpha, n_rows, ndv_attr, ndv_target):
    """(alpha * (N / (ndv_attr * ndv_target))).toLong with integer division (RepairApi.scala:572-576)."""
    return int(alpha * (n_rows // (ndv_attr * ndv_target)))


def discretize_params(kind, vmin, vmax):
    """(min, denominator) doubles for ``int((v - min) / (max - min) * thres)`` exactly as Spark
    folds the spliced literals: plain-notation literals are DECIMAL, so ``max - min`` is an exact
    decimal difference rounded once; E-notation literals are DOUBLE."""
    from decimal import Decimal

    from .utils import double_to_string
    if kind == "int":
        smin, smax = str(int(vmin)), str(int(vmax))
    else:
        smin, smax = double_to_string(vmin), double_to_string(vmax)
    if "E" in smin or "E" in smax:
        den = float(smax) - float(smin)
    else:
        den = float(Decimal(smax) - Decimal(smin))
    return float(smin), den
