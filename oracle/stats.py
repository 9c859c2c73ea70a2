"""Oracle attribute statistics (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Restates ``RepairApi.scala``: checkInputTable (:34-67), column stats + discretisation
(:106-169), computeFreqStats (:231-273), computePairwiseStats (:280-394), computeAttrStats
(:396-477).  Distinct counts are exact (the reference's are HLL++ estimates).
"""
import math
from decimal import Decimal

import numpy as np

from .table import AnalysisException, OTable, factorize, spark_double_to_string

FREQ_GROUP_PREFIX = "__generated_freq_group_"  # RepairApi.scala:213


# --------------------------------------------------------------------------------------
# a1  checkInputTable  (RepairApi.scala:34-67)
# --------------------------------------------------------------------------------------
def check_input_table(tbl, row_id, table_name="input"):
    if row_id not in tbl.cols:
        raise AnalysisException("Column '{}' does not exist in '{}'".format(row_id, table_name))
    if not len(tbl.names) >= 3:
        raise AnalysisException("A least three columns (`{}` columns + two more ones) in table '{}'".format(
            row_id, table_name))
    ids = tbl.cols[row_id]
    n_distinct = len(set(ids.tolist()))
    if n_distinct != tbl.n_rows:
        raise AnalysisException(
            "Uniqueness does not hold in column '{}' of table '{}' (# of distinct '{}': {}, # of rows: {})".format(
                row_id, table_name, row_id, n_distinct, tbl.n_rows))
    return [n for n in tbl.names if n != row_id and tbl.kinds[n] in ("int", "float")]


# --------------------------------------------------------------------------------------
# a7  computeAndGetTableStats / discretizeTable / convertToDiscretizedTable (:106-169)
# --------------------------------------------------------------------------------------
def column_stats(tbl, name):
    """-> (distinct count of non-NULL values, min, max)"""
    codes, uniq = factorize(tbl, name)
    if len(uniq) == 0:
        return 0, None, None
    return len(uniq), uniq[0], uniq[-1]


def _literal(kind, v):
    """How the Scala side splices a stat value into SQL (:139): ``${min.get}``."""
    if kind == "int":
        return str(int(v))
    return spark_double_to_string(float(v))


def discretize_params(kind, vmin, vmax):
    """-> (min as double, denominator as double) exactly as Spark evaluates
    ``(attr - <min>) / (<max> - <min>)``: plain-notation literals are DECIMALs, so the
    denominator is an exact decimal difference rounded once to double; E-notation literals are
    doubles.  ``attr - <min>`` is evaluated in double (int columns: exact integer difference)."""
    smin, smax = _literal(kind, vmin), _literal(kind, vmax)
    if "E" in smin or "E" in smax:
        den = float(smax) - float(smin)
    else:
        den = float(Decimal(smax) - Decimal(smin))
    return float(smin), den


def discretize_numeric(vals, kind, vmin, vmax, thres):
    """``int((v - min) / (max - min) * thres)``; NULL stays NULL; max == min -> NULL (x/0)."""
    fmin, den = discretize_params(kind, vmin, vmax)
    out = np.full(len(vals), np.nan)
    if den == 0.0:
        return out
    ok = ~np.isnan(vals)
    out[ok] = np.trunc((vals[ok] - fmin) / den * thres)
    return out


def convert_to_discretized_table(tbl, row_id, thres):
    """-> (discretized OTable incl. row id, domain_stats {attr: ndv} for every non-row-id attr)"""
    assert 2 <= thres < 65536
    names, kinds, arrays = [row_id], [tbl.kinds[row_id]], [tbl.cols[row_id]]
    domain_stats = {}
    for a in tbl.names:
        if a == row_id:
            continue
        ndv, vmin, vmax = column_stats(tbl, a)
        domain_stats[a] = ndv
        k = tbl.kinds[a]
        if k in ("int", "float"):  # continuous types are always kept and bucketed (:136-139)
            if ndv == 0:
                # min/max are None -> Spark's `.get` throws; treated as all-NULL column
                names.append(a), kinds.append("int"), arrays.append(np.full(tbl.n_rows, np.nan))
            else:
                names.append(a), kinds.append("int")
                arrays.append(discretize_numeric(tbl.cols[a], k, vmin, vmax, thres))
        elif 1 < ndv <= thres:  # (:140-142)
            names.append(a), kinds.append("str"), arrays.append(tbl.cols[a])
        # else dropped with a warning (:143-145)
    return OTable(names, kinds, arrays), domain_stats


# --------------------------------------------------------------------------------------
# a8  computeFreqStats  (RepairApi.scala:231-273)
# --------------------------------------------------------------------------------------
class FreqStats:
    """The ``attr_freq_stats`` view: columns ``attrs`` (+ grouping flags) and ``cnt``.
    ``rows`` = list of (values tuple, flags tuple, cnt); value None = NULL."""

    def __init__(self, attrs, rows):
        self.attrs = list(attrs)
        self.rows = rows
        self.pos = {a: i for i, a in enumerate(self.attrs)}

    def as_tuples(self):
        out = []
        for vals, flags, cnt in self.rows:
            t = []
            for v, f in zip(vals, flags):
                t += [v, f]
            out.append(tuple(t + [cnt]))
        return out


def _py(v):
    if isinstance(v, (np.integer,)):
        return int(v)
    if isinstance(v, (np.floating,)):
        return int(v) if float(v).is_integer() else float(v)
    return v


def compute_freq_stats(disc, attr_sets, attr_freq_ratio_threshold):
    assert attr_sets
    assert 0.0 <= attr_freq_ratio_threshold <= 1.0
    target_attrs = []
    for s in attr_sets:
        for a in s:
            if a not in target_attrs:
                target_attrs.append(a)
    distinct_sets, seen = [], set()
    for s in attr_sets:
        key = frozenset(s)
        if key not in seen:
            seen.add(key)
            distinct_sets.append(list(s))
    if len(target_attrs) > 64:
        raise AnalysisException("Cannot handle the target attributes whose length is more than 64, "
                                "but got: {}".format(",".join(target_attrs)))
    for s in distinct_sets:
        if len(set(s)) > 2:
            raise RuntimeError("Cannot handle more than two entries: {}".format(",".join(s)))
    having = None
    if attr_freq_ratio_threshold > 0.0:
        having = int(disc.n_rows * attr_freq_ratio_threshold)
    fact = {a: factorize(disc, a) for a in target_attrs}
    rows = []
    for s in distinct_sets:
        s = list(dict.fromkeys(s))
        if len(s) == 1:
            codes, uniq = fact[s[0]]
            cnts = np.bincount(codes + 1, minlength=len(uniq) + 1)
            for c in np.nonzero(cnts)[0]:
                cnt = int(cnts[c])
                if having is not None and not cnt > having:
                    continue
                vals = [None] * len(target_attrs)
                flags = [1] * len(target_attrs)
                i = target_attrs.index(s[0])
                vals[i] = None if c == 0 else _py(uniq[c - 1])
                flags[i] = 0
                rows.append((tuple(vals), tuple(flags), cnt))
        else:
            (cx, ux), (cy, uy) = fact[s[0]], fact[s[1]]
            ny = len(uy) + 1
            cnts = np.bincount((cx + 1) * ny + (cy + 1), minlength=(len(ux) + 1) * ny)
            ix, iy = target_attrs.index(s[0]), target_attrs.index(s[1])
            for c in np.nonzero(cnts)[0]:
                cnt = int(cnts[c])
                if having is not None and not cnt > having:
                    continue
                vx, vy = divmod(int(c), ny)
                vals = [None] * len(target_attrs)
                flags = [1] * len(target_attrs)
                vals[ix] = None if vx == 0 else _py(ux[vx - 1])
                vals[iy] = None if vy == 0 else _py(uy[vy - 1])
                flags[ix] = flags[iy] = 0
                rows.append((tuple(vals), tuple(flags), cnt))
    return FreqStats(target_attrs, rows)


# --------------------------------------------------------------------------------------
# a8  computePairwiseStats  (RepairApi.scala:280-394)
# --------------------------------------------------------------------------------------
def log2(v):  # :280-282
    return math.log(v) / math.log(2.0)


def _entropy(cnts, row_count):
    # -COALESCE(SUM((cnt / N) * log2(cnt / N)), 0.0)
    if not cnts:
        return -0.0
    s = 0.0
    for c in cnts:
        p = c / row_count
        s += p * log2(p)
    return -s


def _corr_term(row_count, n_groups, total, ub_domain):
    if row_count > total:  # :318-324, :358-364
        ub = max(ub_domain - n_groups, 1)
        avg = max((row_count - total + 0.0) / ub, 1.0)
        return -ub * (avg / row_count) * log2(avg / row_count)
    return 0.0


def compute_pairwise_stats(row_count, fs, pairs, domain_stats):
    """-> {x: [(y, H(x|y))] ascending}  with H(x|y) = H(x,y) - H(y)."""
    if not pairs:
        return {}
    target_attrs = []
    for x, y in pairs:
        for a in (x, y):
            if a not in target_attrs:
                target_attrs.append(a)
    assert row_count > 0
    assert set(target_attrs) <= set(fs.attrs)
    hxy = {}
    for x, y in pairs:
        key = frozenset((x, y))
        if key in hxy:
            continue
        ix, iy = fs.pos[x], fs.pos[y]
        cnts = [cnt for vals, flags, cnt in fs.rows if flags[ix] == 0 and flags[iy] == 0]
        corr = _corr_term(row_count, len(cnts), sum(cnts), domain_stats[x] * domain_stats[y])
        hxy[key] = _entropy(cnts, row_count) + corr
    hy = {}
    for a in target_attrs:
        ia = fs.pos[a]
        cnts = [cnt for vals, flags, cnt in fs.rows
                if flags[ia] == 0 and all(f == 1 for j, f in enumerate(flags) if j != ia)]
        corr = _corr_term(row_count, len(cnts), sum(cnts), domain_stats[a])
        hy[a] = _entropy(cnts, row_count) + corr
    out = {}
    for x, y in pairs:
        out.setdefault(x, []).append((y, hxy[frozenset((x, y))] - hy[y]))
    return {k: sorted(v, key=lambda t: t[1]) for k, v in out.items()}


# --------------------------------------------------------------------------------------
# a8  computeAttrStats  (RepairApi.scala:396-477)
# --------------------------------------------------------------------------------------
def distinct_pair_count(disc, x, y):
    """Exact stand-in for ``approx_count_distinct(struct(x, y))`` (:435): a struct is never NULL,
    so combinations with NULL members count."""
    (cx, ux), (cy, uy) = factorize(disc, x), factorize(disc, y)
    ny = len(uy) + 1
    return int(len(np.unique((cx + 1) * ny + (cy + 1))))


def select_candidate_pairs(disc, row_id, targets, domain_stats, pairwise_thr, max_attrs):
    disc_attrs = [a for a in disc.names if a != row_id]
    pairs = []
    for t in targets:
        cands = [(t, a) for a in disc_attrs if a != t]
        if len(cands) > max_attrs:
            scored = []
            for (x, y) in cands:
                co = distinct_pair_count(disc, x, y)
                den = domain_stats[x] * domain_stats[y]
                ratio = (co + 0.0) / den if den != 0 else float("inf")
                scored.append((ratio, (x, y)))
            scored = [s for s in scored if s[0] < pairwise_thr]
            scored.sort(key=lambda s: s[0])  # stable, like Scala's sortBy
            pairs += [p for _, p in scored[:max_attrs]]
        else:
            pairs += cands
    return pairs


def compute_attr_stats(disc, row_id, targets, domain_stats, attr_freq_thr, pairwise_thr, max_attrs):
    """-> (FreqStats, {target: [(attr, H)]}, candidate pairs)"""
    assert 0.0 <= attr_freq_thr <= 1.0 and 0.0 <= pairwise_thr <= 1.0 and max_attrs > 0
    assert targets
    disc_attrs = [a for a in disc.names if a != row_id]
    pairs = select_candidate_pairs(disc, row_id, targets, domain_stats, pairwise_thr, max_attrs)
    sets = [[a] for a in disc_attrs] + [[x, y] for x, y in pairs]
    fs = compute_freq_stats(disc, sets, attr_freq_thr)
    stat_map = compute_pairwise_stats(disc.n_rows, fs, pairs, domain_stats)
    for t in targets:
        stat_map.setdefault(t, [])
    return fs, stat_map, pairs
