"""Oracle error detectors (TEST INFRASTRUCTURE -- see oracle/__init__.py).

Every detector returns a set of ``(row_position, attribute)`` pairs; callers map positions to
row ids.  Restates ``ErrorDetectorApi.scala`` / ``DenialConstraints.scala`` / ``errors.py``.
"""
import math
import re
import warnings

import numpy as np

OP_SIGNS = ("EQ", "IQ", "LT", "GT")  # DenialConstraints.scala:74


# --------------------------------------------------------------------------------------
# a2  NullErrorDetector.detect  (ErrorDetectorApi.scala:128-157)
# --------------------------------------------------------------------------------------
def null_cells(tbl, row_id, targets):
    out = set()
    for attr in tbl.names:  # inputDf.columns order, filtered (:140)
        if attr == row_id or attr not in targets:
            continue
        for r in np.nonzero(tbl.isnull(attr))[0]:
            out.add((int(r), attr))
    return out


# --------------------------------------------------------------------------------------
# a5  RegExErrorDetector.detect  (ErrorDetectorApi.scala:159-187)
#     flag  CAST(v AS STRING) NOT RLIKE regex  OR  v IS NULL ; RLIKE = regex *find*
# --------------------------------------------------------------------------------------
def regex_cells(tbl, row_id, targets, attr, regex):
    if attr not in targets or regex is None or regex.strip() == "":  # :171
        return set()
    if attr not in tbl.cols:
        return set()
    pat = re.compile(regex)
    out = set()
    for r, s in enumerate(tbl.strings(attr)):
        if s is None or pat.search(str(s)) is None:
            out.add((r, attr))
    return out


def domain_values_regex(tbl, attr, values, autofill, min_count_thres, continuous):
    """DomainValues._detect_impl (errors.py:112-129) -> regex string, or None when the detector
    yields an empty frame (continuous attribute)."""
    if attr in continuous:
        return None
    domain_values = list(values) if not autofill else []
    if autofill:
        strs = tbl.strings(attr)
        cnt = {}
        for s in strs:
            if s is not None:
                cnt[s] = cnt.get(s, 0) + 1
        filled = [k for k, c in cnt.items() if c > min_count_thres]
        if filled:
            domain_values = [str(v) for v in filled]
    return "({})".format("|".join(domain_values)) if domain_values else "$^"


# --------------------------------------------------------------------------------------
# a3  DenialConstraints.parse / parseAlt / parseAndVerifyConstraints
#     (DenialConstraints.scala:82-225)
# --------------------------------------------------------------------------------------
class Pred:
    """sign in OP_SIGNS; left is always an attribute of t1; right is ('attr', name) of t2 or
    ('const', text)."""

    def __init__(self, sign, left, right_kind, right):
        self.sign, self.left, self.right_kind, self.right = sign, left, right_kind, right

    @property
    def references(self):  # :44-46  (distinct, left first)
        refs = [self.left]
        if self.right_kind == "attr" and self.right not in refs:
            refs.append(self.right)
        return refs

    def __repr__(self):
        return "{}({},{}:{})".format(self.sign, self.left, self.right_kind, self.right)

    def __eq__(self, o):
        return (self.sign, self.left, self.right_kind, self.right) == (o.sign, o.left, o.right_kind, o.right)


def _is_identifier(s):  # :123-125
    return re.fullmatch(r"[a-zA-Z]+[a-zA-Z0-9]*", s) is not None


def parse(c):  # :128-182
    raw = c.split("&")
    if "&" in c:  # Java's String.split drops trailing empty strings (not the no-delimiter case)
        while raw and raw[-1] == "":
            raw.pop()
    parts = [p.strip() for p in raw]
    if len(parts) >= 2 and _is_identifier(parts[0]) and _is_identifier(parts[1]):
        t1, t2, cons = parts[0], parts[1], parts[2:]
        if len(cons) < 2:
            raise ValueError("At least two predicate candidates should be given, "
                             "but {} candidates found: {}".format(len(cons), c))
        pat = re.compile(r"({})\s*\(\s*{}\.(.*)\s*,\s*{}\.(.*)\s*\)".format(
            "|".join(OP_SIGNS), re.escape(t1), re.escape(t2)))
        preds, bad = [], []
        for s in cons:
            m = pat.fullmatch(s)
            if m:
                preds.append(Pred(m.group(1), m.group(2).strip(), "attr", m.group(3).strip()))
            else:
                bad.append(s)
        if bad:
            raise ValueError("Illegal predicates found: " + ", ".join(bad))
        return preds
    if len(parts) >= 1 and _is_identifier(parts[0]):
        t1, cons = parts[0], parts[1:]
        if len(cons) < 2:
            raise ValueError("At least two predicate candidates should be given, "
                             "but {} candidates found: {}".format(len(cons), c))
        pat = re.compile(r"({})\s*\(\s*{}\.(.*)\s*,\s*(.*)\)".format("|".join(OP_SIGNS), re.escape(t1)))
        preds, bad = [], []
        for s in cons:
            m = pat.fullmatch(s)
            if m:
                preds.append(Pred(m.group(1), m.group(2).strip(), "const", m.group(3).strip()))
            else:
                bad.append(s)
        if bad:
            raise ValueError("Illegal predicates found: " + ", ".join(bad))
        return preds
    if parts:
        raise ValueError("Failed to parse an input string: '{}'".format(c))
    return []


def parse_alt(c):  # :185-195   "X->Y"
    parts = [p.strip() for p in c.split("->")]
    parts = [p for p in parts if p]
    if len(parts) == 2:
        x, y = parts
        return [Pred("EQ", x, "attr", x), Pred("IQ", y, "attr", y)]
    if parts:
        raise ValueError("Failed to parse an input string: '{}'".format(c))
    return []


def load_constraint_stmts(path, constraints):  # :198-224
    lines = []
    if path is not None and path.strip() != "":
        try:
            p = path[len("file://"):] if path.startswith("file://") else path
            with open(p) as f:
                lines += [ln.rstrip("\n").rstrip("\r") for ln in f.readlines()]
        except Exception:
            warnings.warn("Failed to load constrains from '{}'".format(path))
    if constraints is not None:
        lines += [s.strip() for s in constraints.split(";") if s.strip()]
    return lines


def parse_and_verify_constraints(lines, table_attrs):  # :82-119
    """-> (list of predicate lists, references)"""
    predicates = []
    for c in lines:
        try:
            try:
                ps = parse(c)
            except Exception:
                ps = parse_alt(c)
            predicates.append(ps)
        except Exception:
            warnings.warn("Illegal constraint format found: {}".format(c))
    refs = []
    for ps in predicates:
        for p in ps:
            for r in p.references:
                if r not in refs:
                    refs.append(r)
    predicates = [ps for ps in predicates if ps]
    attr_set = set(table_attrs)
    absent = [a for a in refs if a not in attr_set]
    if absent:
        kept = [ps for ps in predicates if all(r in attr_set for p in ps for r in p.references)]
        if kept:
            return kept, [r for r in refs if r in attr_set]
        return [], []
    return predicates, refs


# --------------------------------------------------------------------------------------
# a3  ConstraintErrorDetector.detect (ErrorDetectorApi.scala:189-244): for every DC, rows t1
#     with EXISTS t2 (any row, t1 included) such that all predicates hold; flag every
#     referenced attribute that is a target.
# --------------------------------------------------------------------------------------
def _const_value(text, kind):
    t = text.strip()
    if len(t) >= 2 and t[0] == t[-1] and t[0] in "\"'":
        return t[1:-1]
    if kind != "str":
        try:
            return float(t)
        except ValueError:
            return t
    return t


def _cmp_matrix(sign, lv, lnull, rv, rnull):
    """lv: [n1] values, rv: [n2] values -> bool [n1, n2] under SQL semantics."""
    eqm = (lv[:, None] == rv[None, :])
    both = (~lnull)[:, None] & (~rnull)[None, :]
    if sign in ("EQ", "IQ"):  # <=>  /  NOT(<=>)   (DenialConstraints.scala:76-77)
        nse = (eqm & both) | (lnull[:, None] & rnull[None, :])
        return nse if sign == "EQ" else ~nse
    if sign == "LT":
        return (lv[:, None] < rv[None, :]) & both
    return (lv[:, None] > rv[None, :]) & both


def _col_cmp_arrays(tbl, name):
    """values usable with ==,<,> (object str or float) with NULLs replaced by a placeholder."""
    nul = tbl.isnull(name)
    a = tbl.cols[name]
    if tbl.kinds[name] == "str" and a.dtype == object:
        v = np.array(["" if x is None else x for x in a], dtype=object)
    else:
        v = np.where(nul, 0, a)
    return v, nul


def constraint_rows_bruteforce(tbl, preds):
    """Rows t1 for which EXISTS t2 with all predicates true.  O(N^2) -- small inputs only."""
    n = tbl.n_rows
    if n == 0:
        return np.zeros(0, dtype=bool)
    ok = np.ones((n, n), dtype=bool)
    for p in preds:
        lv, lnull = _col_cmp_arrays(tbl, p.left)
        if p.right_kind == "attr":
            rv, rnull = _col_cmp_arrays(tbl, p.right)
            m = _cmp_matrix(p.sign, lv, lnull, rv, rnull)
        else:
            c = _const_value(p.right, tbl.kinds[p.left])
            cv = np.array([c], dtype=object if isinstance(c, str) else np.float64)
            if isinstance(c, str) and lv.dtype != object:
                lv = np.array([str(x) for x in lv], dtype=object)
            m = _cmp_matrix(p.sign, lv, lnull, cv, np.zeros(1, dtype=bool))
            m = np.repeat(m, n, axis=1)
        ok &= m
    return ok.any(axis=1)


def constraint_rows_grouped(tbl, preds):
    """Same result for the decomposable shapes, O(N log N):
    * single-tuple constant DCs (row predicate);
    * EQ(a,a)* & IQ(b,b): a row violates iff its NULL-safe key group holds >= 2 distinct b
      (NULL counted as a value) -- SURVEY.md section 8(a) a3 note;
    * EQ-only two-tuple DCs: every row matches itself.
    Returns None for shapes that need the brute-force path."""
    from .table import factorize
    n = tbl.n_rows
    if all(p.right_kind == "const" for p in preds):
        ok = np.ones(n, dtype=bool)
        for p in preds:
            lv, lnull = _col_cmp_arrays(tbl, p.left)
            c = _const_value(p.right, tbl.kinds[p.left])
            if isinstance(c, str) and lv.dtype != object:
                lv = np.array([str(x) for x in lv], dtype=object)
            cv = np.array([c], dtype=object if isinstance(c, str) else np.float64)
            ok &= _cmp_matrix(p.sign, lv, lnull, cv, np.zeros(1, dtype=bool))[:, 0]
        return ok
    if any(p.right_kind != "attr" or p.left != p.right for p in preds):
        return None
    eqs = [p.left for p in preds if p.sign == "EQ"]
    iqs = [p.left for p in preds if p.sign == "IQ"]
    if len(eqs) + len(iqs) != len(preds):
        return None
    if not iqs:
        return np.ones(n, dtype=bool)
    if len(iqs) != 1:
        return None
    key = np.zeros(n, dtype=np.int64)
    for a in eqs:
        codes, uniq = factorize(tbl, a)
        key = key * (len(uniq) + 1) + (codes + 1)
    _, key = np.unique(key, return_inverse=True)
    b, ub = factorize(tbl, iqs[0])
    b = b + 1
    nk = int(key.max()) + 1 if n else 0
    lo = np.full(nk, np.iinfo(np.int64).max)
    hi = np.full(nk, -1)
    np.minimum.at(lo, key, b)
    np.maximum.at(hi, key, b)
    return (lo != hi)[key]


def constraint_cells(tbl, row_id, targets, constraint_path, constraints, force_bruteforce=False):
    stmts = load_constraint_stmts(constraint_path, constraints)
    if not stmts:
        return set()
    pred_lists, _ = parse_and_verify_constraints(stmts, tbl.names)
    out = set()
    for preds in pred_lists:
        attrs = []
        for p in preds:
            for r in p.references:
                if r in targets and r not in attrs:
                    attrs.append(r)
        if not attrs:
            continue
        rows = None if force_bruteforce else constraint_rows_grouped(tbl, preds)
        if rows is None:
            rows = constraint_rows_bruteforce(tbl, preds)
        for r in np.nonzero(rows)[0]:
            for a in attrs:
                out.add((int(r), a))
    return out


# --------------------------------------------------------------------------------------
# a4  GaussianOutlierErrorDetector.detect (ErrorDetectorApi.scala:249-300)
# --------------------------------------------------------------------------------------
def spark_percentile(sorted_vals, p):
    """Spark's exact ``percentile``: position p*(n-1), linear interpolation
    ``(higher - pos) * v[lower] + (pos - lower) * v[higher]``."""
    n = len(sorted_vals)
    pos = (n - 1) * p
    lo, hi = int(math.floor(pos)), int(math.ceil(pos))
    if lo == hi:
        return float(sorted_vals[lo])
    return (hi - pos) * float(sorted_vals[lo]) + (pos - lo) * float(sorted_vals[hi])


def outlier_bounds(vals):
    v = np.sort(vals[~np.isnan(vals)])
    if len(v) == 0:
        return None
    q1, q3 = spark_percentile(v, 0.25), spark_percentile(v, 0.75)
    return q1 - 1.5 * (q3 - q1), q3 + 1.5 * (q3 - q1)  # :286


def outlier_cells(tbl, row_id, continuous, targets, approx_enabled=False):
    """``approx_enabled`` (percentile_approx with accuracy 1000) is answered with the exact
    percentile: Spark's sketch is not restated (pinned only loosely)."""
    out = set()
    for attr in [a for a in continuous if a in targets]:
        if attr not in tbl.cols:
            continue
        vals = tbl.cols[attr]
        b = outlier_bounds(vals)
        if b is None:
            continue
        lower, upper = b
        with np.errstate(invalid="ignore"):
            bad = (vals < lower) | (vals > upper)
        for r in np.nonzero(bad)[0]:
            out.add((int(r), attr))
    return out
