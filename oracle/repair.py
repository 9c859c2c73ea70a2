"""Oracle pipeline: ErrorModel.detect + RepairModel._run (TEST INFRASTRUCTURE -- see
oracle/__init__.py).

Restates ``errors.py:389-461,545-582`` (detection pipeline), ``RepairApi.scala:69-104``
(current values), ``:171-211`` (NULL masking), ``model.py:533-555`` (split), ``:677-729``
(features / encoders), ``:955-1052`` (model bookkeeping), ``:1062-1143`` (inference chain) and
``:1398-1401`` (output shaping).  Model *training* is delegated to a ``model_provider`` callback
(the reference delegates it to LightGBM, ``train.py:89-229``).
"""
import re

import numpy as np

from . import detect as D
from . import domain as DM
from . import stats as S
from .forest import (encode_rows, encoder_kind, first_seen_categories, forest_predict)
from .table import cast_to_string

DEFAULT_OPTS = {  # errors.py:321-338, model.py:115-123
    "error.attr_freq_ratio_threshold": 0.0,
    "error.pairwise_freq_ratio_threshold": 0.05,
    "error.max_attrs_to_compute_pairwise_stats": 3,
    "error.max_attrs_to_compute_domains": 2,
    "error.domain_threshold_alpha": 0.0,
    "error.domain_threshold_beta": 0.70,
    "model.max_training_row_num": 10000,
    "model.max_training_column_num": 65536,
    "model.small_domain_threshold": 12,
}


def _opt(opts, key):
    v = (opts or {}).get(key, DEFAULT_OPTS[key])
    return type(DEFAULT_OPTS[key])(v)


def _detector_targets(det, pipeline_targets):
    # ErrorDetector.setUp (errors.py:49-59)
    own = det.get("targets") or []
    if own:
        return [t for t in pipeline_targets if t in set(own)]
    return list(pipeline_targets)


def run_detectors(tbl, row_id, targets, detectors, continuous):
    """_detect_error_cells (errors.py:405-421): union + distinct of all detector outputs."""
    target_attrs = [c for c in tbl.names if c != row_id]
    if targets:
        target_attrs = [c for c in target_attrs if c in set(targets)]  # errors.py:398-402
    if not detectors:  # errors.py:389-396
        detectors = [{"type": "null"}]
        for c in (targets if targets else [c for c in tbl.names if c != row_id]):
            detectors.append({"type": "domain", "attr": c, "values": [], "autofill": True, "min_count_thres": 4})
    cells = set()
    for det in detectors:
        tg = _detector_targets(det, target_attrs)
        t = det["type"]
        if t == "null":
            cells |= D.null_cells(tbl, row_id, tg)
        elif t == "regex":
            cells |= D.regex_cells(tbl, row_id, tg, det["attr"], det["regex"])
        elif t == "domain":
            rx = D.domain_values_regex(tbl, det["attr"], det.get("values", []), det.get("autofill", False),
                                       det.get("min_count_thres", 12), continuous)
            if rx is not None and det["attr"] in tbl.cols:
                cells |= D.regex_cells(tbl, row_id, tg, det["attr"], rx)
        elif t == "constraint":
            cells |= D.constraint_cells(tbl, row_id, tg, det.get("path", ""), det.get("constraints", ""))
        elif t == "outlier":
            cells |= D.outlier_cells(tbl, row_id, continuous, tg, det.get("approx", False))
        else:
            raise ValueError(t)
    return cells


def with_current_values(tbl, cells):
    """RepairApi.withCurrentValues (:69-104): current_value = CAST(cell AS STRING)."""
    col_pos = {n: i for i, n in enumerate(tbl.names)}
    out = []
    for (r, a) in sorted(cells, key=lambda c: (c[0], col_pos[c[1]])):
        out.append((r, a, cast_to_string(tbl.kinds[a], tbl.value(a, r))))
    return out


def error_model_detect(tbl, row_id, targets, discrete_thres, detectors, given_error_cells, continuous, opts=None):
    """ErrorModel.detect (errors.py:545-582).

    -> (error cells [(row_pos, attr, current_value)], target_columns, pairwise_stats, domain_stats)
    ``given_error_cells``: optional list of (row_id_value, attribute)."""
    if given_error_cells is not None:  # errors.py:434-446
        idpos = {cast_to_string(tbl.kinds[row_id], tbl.value(row_id, r)): r for r in range(tbl.n_rows)}
        keep = set(targets) if targets else set(tbl.names)
        cells = set()
        for rid, a in given_error_cells:
            if a in keep and a in tbl.cols and str(rid) in idpos:
                cells.add((idpos[str(rid)], a))
    else:
        cells = run_detectors(tbl, row_id, targets, detectors, continuous)
    if not cells:
        return [], [], {}, {}
    noisy_columns = [c for c in tbl.names if c in {a for _, a in cells}]
    noisy = with_current_values(tbl, cells)
    disc, domain_stats = S.convert_to_discretized_table(tbl, row_id, discrete_thres)
    disc_cols = disc.names
    if len(disc_cols) == 0:
        return noisy, [], {}, {}
    target_columns = [c for c in noisy_columns if c in disc_cols]
    if len(target_columns) == 0 or len(disc_cols) <= 1:
        return noisy, target_columns, {}, domain_stats
    fs, pairwise, _ = S.compute_attr_stats(
        disc, row_id, target_columns, domain_stats,
        _opt(opts, "error.attr_freq_ratio_threshold"),
        _opt(opts, "error.pairwise_freq_ratio_threshold"),
        _opt(opts, "error.max_attrs_to_compute_pairwise_stats"))
    error_cells = noisy
    if given_error_cells is None:  # errors.py:569-576
        doms = DM.compute_domain_in_error_cells(
            disc, row_id, noisy, continuous, target_columns, fs, pairwise, domain_stats,
            _opt(opts, "error.max_attrs_to_compute_domains"),
            _opt(opts, "error.domain_threshold_alpha"),
            _opt(opts, "error.domain_threshold_beta"))
        weak = DM.weak_labeled_cells(disc, doms)
        error_cells = [c for c in noisy if (c[0], c[1]) not in weak]
    return error_cells, target_columns, pairwise, domain_stats


def convert_error_cells_to_null(tbl, cells, target_columns):
    """RepairApi.convertErrorCellsToNull (:171-211) on the ORIGINAL table (model.py:1323)."""
    out = tbl.copy()
    for (r, a, _) in cells:
        if a in target_columns:
            if out.cols[a].dtype == object:
                out.cols[a][r] = None
            elif out.kinds[a] == "str":
                out.cols[a][r] = -1
            else:
                out.cols[a][r] = np.nan
    return out


def sample_training_rows(candidates, max_rows, seed=42):
    """Deterministic stand-in for the unseeded ``df.sample(ratio)`` (model.py:755-766): exactly
    ``max_rows`` rows without replacement when there are more, in table order."""
    candidates = np.asarray(candidates, dtype=np.int64)
    if len(candidates) <= max_rows:
        return candidates
    rng = np.random.default_rng(seed)
    idx = np.sort(rng.choice(len(candidates), size=max_rows, replace=False))
    return candidates[idx]


def select_features(pairwise_stats, y, features, max_training_column_num):
    """RepairModel._select_features (model.py:677-699)."""
    if max_training_column_num < len(features) and y in pairwise_stats:
        fts = sorted([(float(corr), f) for f, corr in pairwise_stats[y] if f in features])
        top = []
        for corr, f in fts:
            if len(top) <= 1 or (corr >= 0.0 and len(top) < max_training_column_num):
                top.append((corr, f))
        features = [f for _, f in top]
    return features


def column_values(tbl, name, rows=None):
    """Column (or the listed rows of it) as plain Python values, cell for cell what ``tbl.value`` returns
    (None = NULL; label-encoded discrete columns as ints; integral floats as ints) -- converted array-wise:
    the per-cell accessor made this the most expensive line of the CPU baseline."""
    a = tbl.cols[name]
    if rows is not None:
        a = a[np.asarray([int(r) for r in rows], dtype=np.int64)] if len(rows) else a[:0]
    if a.dtype == object:
        return a.tolist()
    if tbl.kinds[name] == "str":  # label-encoded discrete column
        return [None if v < 0 else v for v in a.astype(np.int64).tolist()]
    return [None if v != v else (int(v) if v.is_integer() else v) for v in a.astype(np.float64).tolist()]


def levenshtein(a, b):
    """python-Levenshtein distance as used by costs.Levenshtein (costs.py:38-49)."""
    a, b = str(a), str(b)
    prev = list(range(len(b) + 1))
    for i, ca in enumerate(a, 1):
        cur = [i]
        for j, cb in enumerate(b, 1):
            cur.append(min(prev[j] + 1, cur[j - 1] + 1, prev[j - 1] + (ca != cb)))
        prev = cur
    return float(prev[-1])


class _Unknown:
    """Stands for the JSON pmf string the reference writes into a repaired cell in pmf mode
    (model.py:1118-1128): later models see a value that is neither NULL nor a known category."""

    def __repr__(self):
        return "<pmf>"


UNKNOWN = _Unknown()


def shape_pmf(cells_pmf, opts=None, cost_fn=None, cost_targets=None):
    """_compute_repair_pmf (model.py:1174-1225).  cells_pmf: [(rid, attr, cur, classes, probs)]
    (classes None for continuous: pmf = [(value, 1.0)]).
    -> [(rid, attr, (cur, cur_prob), [(class, prob)] sorted by prob desc, filtered, top-k)]"""
    thr = float((opts or {}).get("repair.pmf.prob_threshold", 0.0))
    top_k = int((opts or {}).get("repair.pmf.prob_top_k", 32))
    weight = float((opts or {}).get("repair.pmf.cost_weight", 0.1))
    out = []
    for rid, attr, cur, classes, probs in cells_pmf:
        if classes is None:
            out.append((rid, attr, (cur, 0.0), [(probs, 1.0)]))
            continue
        classes = [cast_to_string("str", c) if isinstance(c, str) else
                   (str(c) if not isinstance(c, float) else cast_to_string("float", c)) for c in classes]
        probs = [float(p) for p in probs][:len(classes)]
        if cost_fn is not None and (not cost_targets or attr in cost_targets) and cur:
            probs = [p * (1.0 / (1.0 + weight * cost_fn(cur, c))) for p, c in zip(probs, classes)]
        if cost_fn is not None:
            norm = 0.0
            for p in probs:
                norm += p
            probs = [p / norm for p in probs]
        cur_prob = probs[classes.index(cur)] if cur in classes else 0.0
        pmf = sorted(zip(classes, probs), key=lambda t: -t[1])  # stable: ties keep class order
        pmf = [t for t in pmf if t[1] > thr][:top_k]
        out.append((rid, attr, (cur, cur_prob), pmf))
    return out


def compute_score(shaped, cost_fn):
    """_compute_score (model.py:1227-1248) -> [(rid, attr, cur, repaired, score)]"""
    import math
    out = []
    for rid, attr, (cur, cur_prob), pmf in shaped:
        rep, rep_prob = pmf[0]
        base = cur if cur is not None else rep
        cost = cost_fn(base, rep)
        cost = 256.0 if cost is None else cost
        score = math.log(rep_prob / (cur_prob if cur_prob > 0.0 else 1e-6)) * (1.0 / (1.0 + cost))
        out.append((rid, attr, cur, rep, score))
    return out


def maximal_likelihood_repair(scored, delta):
    """_maximal_likelihood_repair (model.py:1259-1277): keep score >= percentile(score, 1 - delta/n)."""
    n = len(scored)
    percent = min(1.0, 1.0 - delta / n)
    thres = D.spark_percentile(sorted(s[4] for s in scored), percent)
    return [s[:4] for s in scored if s[4] >= thres]


# --------------------------------------------------------------------------------------
# Rule-based repairs (model.py:583-673, 731-753, 928-953; DepGraph.scala:257-317)
# --------------------------------------------------------------------------------------
def cost_of(cost_fn, x, y):
    """UpdateCostFunction.compute (costs.py:33-34): NULL unless both operands are truthy."""
    return cost_fn(str(x), str(y)) if x and y else None


def repair_by_nearest_values(base, error_cells, target_columns, cost_fn, cost_targets, merge_threshold):
    """_repair_by_nearest_values (model.py:583-626): an error cell whose current value is within
    `merge_threshold` of exactly one nearest value of the attribute's (masked) domain takes that value.
    Candidates whose cost is NULL can never win and never count as the runner-up (Spark's comparator
    treats them as equal to everything, which leaves their position unspecified; `cost < NULL` is
    never true either way).
    -> (remaining error cells, [(row, attr, current_value, repaired typed value)])"""
    targets = [c for c in target_columns if c in cost_targets] if cost_targets else list(target_columns)
    domains = {}
    for a in targets:
        seen = []
        for v in column_values(base, a):
            if v is not None and v not in seen:
                seen.append(v)
        domains[a] = seen
    remaining, repaired, memo = [], [], {}
    for (r, a, cur) in error_cells:
        best = None
        if a in domains and domains[a] and cur:
            if (a, cur) not in memo:  # same current value, same answer
                costs = [(cost_of(cost_fn, cur, v), v) for v in domains[a]]
                known = sorted([c for c in costs if c[0] is not None], key=lambda c: c[0])
                ok = len(known) >= 2 and known[0][0] <= merge_threshold and known[0][0] < known[1][0]
                memo[(a, cur)] = known[0][1] if ok else None
            best = memo[(a, cur)]
        if best is None:
            remaining.append((r, a, cur))
        else:
            repaired.append((r, a, cur, best))
    return remaining, repaired


def repair_by_regexs(error_cells, regexs):
    """_repair_by_regexs (model.py:633-651) over RepairApi.repairByRegularExpression (:677-706): the
    error cells of `attr` whose current value can be rebuilt from the detector's regex take the
    rebuilt value; a regex the lexer rejects repairs nothing.
    -> (remaining error cells, [(row, attr, current_value, repaired string)])"""
    from .regex_repair import RegexStructureRepair
    done = {}
    for attr, regex in regexs:
        try:
            fix = RegexStructureRepair(regex)
        except (ValueError, re.error):  # lexer error / uncompilable pattern: logged, nothing repaired
            continue
        for (r, a, cur) in error_cells:
            if a == attr and (r, a) not in done:
                rep = fix(cur)
                if rep is not None:
                    done[(r, a)] = (r, a, cur, rep)
    remaining = [c for c in error_cells if (c[0], c[1]) not in done]
    return remaining, list(done.values())


def functional_deps(table_attrs, constraint_path, constraints, target_attrs):
    """DepGraph.computeFunctionalDeps (:257-298): {y: sorted [x]} for every constraint made of exactly
    one EQ and one IQ predicate over a single attribute each, in statement order, skipping a
    dependency whose reverse is already there."""
    stmts = D.load_constraint_stmts(constraint_path, constraints)
    pred_lists, _ = D.parse_and_verify_constraints(stmts, table_attrs) if stmts else ([], [])
    fd = {}
    for preds in pred_lists:
        if len(preds) != 2 or {p.sign for p in preds} != {"EQ", "IQ"}:
            continue
        if any(len(p.references) != 1 for p in preds):
            continue
        x = [p for p in preds if p.sign == "EQ"][0].references[0]
        y = [p for p in preds if p.sign == "IQ"][0].references[0]
        no_cycle = y not in fd.get(x, set()) and x not in fd.get(y, set())
        if y in target_attrs and no_cycle:
            fd.setdefault(y, set()).add(x)
    return {y: sorted(xs) for y, xs in fd.items()}


def functional_dep_map(tbl, x, y):
    """DepGraph.computeFunctionalDepMap (:300-316): x value -> the single y value it occurs with
    (NULL y ignored by collect_set; the NULL x group is dropped)."""
    groups = {}
    for xv, yv in zip(column_values(tbl, x), column_values(tbl, y)):
        if xv is None:
            continue
        g = groups.setdefault(xv, [])
        if yv is not None and yv not in g:
            g.append(yv)
    return {xv: g[0] for xv, g in groups.items() if len(g) == 1}


def resolve_prediction_order(models, target_columns):
    """_resolve_prediction_order (model.py:928-953): statistical models first, then every
    functional-dependency model once its determinant is no longer waiting for a repair."""
    by_y = {m[0]: m for m in models}
    ordered, waiting = [], list(target_columns)
    for y in target_columns:
        if "fd" not in by_y[y][1]:
            ordered.append(by_y[y])
            waiting.remove(y)
    while waiting:
        before = len(waiting)
        for y in list(waiting):
            if by_y[y][1]["fd"]["x"] not in waiting:
                ordered.append(by_y[y])
                waiting.remove(y)
        assert len(waiting) < before
    return ordered


def repair(tbl, row_id, error_cells, target_columns, pairwise_stats, domain_stats, continuous, opts,
           model_provider, repair_data=False, pmf_mode=False, rules=None):
    """RepairModel._run phases 2-3 (model.py:1311-1408), default mode.

    ``model_provider(ctx) -> spec`` is called once per target that needs a statistical model with
    ctx = {y, features, encoders (fitted categories), X (encoded training matrix), y_values,
    is_discrete, num_class, train_rows};  spec = {"forest": flat forest, "classes": [labels
    ascending] or None}  or  {"const": value}.
    ``rules`` (setRepairByRules(True)): {"cost_fn", "cost_targets", "nearest_values": bool,
    "merge_threshold", "functional_deps": bool, "max_domain_size", "constraints": [detector specs],
    "regexs": [(attr, regex)] when repair-by-regex is enabled}.
    -> list of (row_id_string, attribute, current_value, repaired)."""
    if not error_cells:
        return []
    if len(target_columns) == 0:
        raise ValueError("At least one valid discretizable feature is needed to repair error cells, "
                         "but no such feature found")
    error_cells = [c for c in error_cells if c[1] in target_columns]  # model.py:1316
    base = convert_error_cells_to_null(tbl, error_cells, target_columns)
    by_rules = []
    if rules and rules.get("regexs"):  # model.py:662-665 (regex repairs come first)
        error_cells, by_regex = repair_by_regexs(error_cells, rules["regexs"])
        for (r, a, cur, v) in by_regex:
            if base.kinds[a] != "str":
                raise NotImplementedError("regex structure repair of a numeric attribute")
            if base.cols[a].dtype != object:
                raise NotImplementedError("regex structure repair needs string values, not codes")
            base.cols[a][r] = v
        by_rules += by_regex
    if rules and rules.get("nearest_values") and rules.get("cost_fn") is not None:  # model.py:1326-1328
        error_cells, by_nv = repair_by_nearest_values(
            base, error_cells, target_columns, rules["cost_fn"], rules.get("cost_targets") or [],
            rules.get("merge_threshold", 2.0))
        by_rules += by_nv
        for (r, a, _, v) in by_nv:  # _repair_attrs: the repaired cells join the repair base
            if base.cols[a].dtype == object or base.kinds[a] == "str":
                base.cols[a][r] = v
            else:
                base.cols[a][r] = float(v)
    if by_rules and not error_cells:
        # nothing left for the statistical models (the reference still trains them; their output
        # would be empty)
        if repair_data:
            return base
        return [(cast_to_string(tbl.kinds[row_id], tbl.value(row_id, r)), a, cur, cast_to_string(tbl.kinds[a], v))
                for (r, a, cur, v) in by_rules]
    dirty_pos = sorted({c[0] for c in error_cells})
    columns = [c for c in base.names if c != row_id]  # train_df.drop(row_id) (model.py:985)
    integral = {c for c in base.names if base.kinds[c] == "int"}
    fdeps = None
    if rules and rules.get("functional_deps"):  # _get_functional_deps (model.py:735-753)
        ceds = rules.get("constraints") or []
        if len(ceds) == 1:
            ct = [c for c in target_columns if c in ceds[0]["targets"]] if ceds[0].get("targets") else target_columns
            fdeps = functional_deps(columns, ceds[0].get("path", ""), ceds[0].get("constraints", ""), ct)
    models = []
    for y in target_columns:  # model.py:1001-1052
        is_discrete = y not in continuous
        y_all = column_values(base, y)
        input_columns = [c for c in columns if c != y]
        num_class = len({v for v in y_all if v is not None}) if is_discrete else 0
        if is_discrete and num_class <= 1:
            v = next((v for v in y_all if v is not None), None) if num_class == 1 else None
            models.append((y, {"const": v}, input_columns, None))
            continue
        if fdeps is not None and y in fdeps:  # model.py:1018-1029
            fx = [x for x in fdeps[y] if int(domain_stats[x]) < int(rules.get("max_domain_size", 1000))]
            if fx:
                if base.kinds[y] != "str":
                    raise NotImplementedError("functional-dependency models only repair string attributes "
                                              "(the reference's FunctionalDepModel predicts strings)")
                fmap = functional_dep_map(base, fx[0], y) if base.kinds[fx[0]] == "str" else {}
                models.append((y, {"fd": {"x": fx[0], "map": fmap}}, [fx[0]], None))
                continue
        features = select_features(pairwise_stats, y, input_columns, _opt(opts, "model.max_training_column_num"))
        cand = [r for r in range(base.n_rows) if y_all[r] is not None]
        if len(cand) == 0:
            models.append((y, {"const": None}, features, None))
            continue
        rows = sample_training_rows(cand, _opt(opts, "model.max_training_row_num"))
        train_cols = {f: column_values(base, f, rows) for f in features}
        encoders = []
        for f in features:
            kind = encoder_kind(f, continuous, domain_stats, _opt(opts, "model.small_domain_threshold"))
            e = {"attr": f, "type": kind}
            if kind != "cont":
                e["categories"] = first_seen_categories(train_cols[f])
            encoders.append(e)
        X = encode_rows(encoders, train_cols)
        ctx = {"y": y, "features": features, "encoders": encoders, "X": X,
               "y_values": [y_all[int(r)] for r in rows], "is_discrete": is_discrete,
               "num_class": num_class, "train_rows": rows}
        spec = model_provider(ctx)
        models.append((y, spec, features, encoders))
    if any("fd" in m[1] for m in models):
        models = resolve_prediction_order(models, target_columns)
    # ---- repair UDF (model.py:1096-1135): sequential chain over targets on the dirty rows ----
    dirty = base.take(np.array(dirty_pos, dtype=np.int64))
    for c in dirty.names:  # pmf mode parks JSON strings in repaired cells: needs object columns
        if pmf_mode and dirty.kinds[c] == "str" and dirty.cols[c].dtype != object:
            dirty.cols[c] = np.array([None if v < 0 else int(v) for v in dirty.cols[c]], dtype=object)
    pmf_of = {}
    for (y, spec, features, encoders) in models:
        ycol = dirty.cols[y]
        n = dirty.n_rows
        if pmf_mode and y not in continuous:  # model.py:1118-1128
            from .forest import forest_proba
            if "const" in spec:
                classes, probs = [spec["const"]], [[1.0]] * n
            elif "fd" in spec:  # FunctionalDepModel.predict_proba (model.py:89-100): one-hot or nothing
                classes = sorted(set(spec["fd"]["map"].values()))
                xs = column_values(dirty, spec["fd"]["x"])
                probs = [[1.0 if c == spec["fd"]["map"][xv] else 0.0 for c in classes]
                         if xv in spec["fd"]["map"] else None for xv in xs]
            else:
                X = encode_rows(encoders, {f: column_values(dirty, f) for f in features})
                classes, probs = spec["classes"], forest_proba(spec["forest"], X)
            for i in range(n):
                if dirty.value(y, i) is None:
                    pmf_of[(dirty_pos[i], y)] = (classes, list(probs[i])) if probs[i] is not None else ([], [])
                    ycol[i] = UNKNOWN
            continue
        if "const" in spec:
            pred = [spec["const"]] * n
        elif "fd" in spec:  # FunctionalDepModel.predict (model.py:86-87)
            pred = [spec["fd"]["map"].get(xv) if xv is not UNKNOWN else None
                    for xv in column_values(dirty, spec["fd"]["x"])]
        else:
            X = encode_rows(encoders, {f: column_values(dirty, f) for f in features})
            p = forest_predict(spec["forest"], X)
            if spec.get("classes") is not None:
                pred = [spec["classes"][int(i)] for i in p]
            elif y in integral:
                pred = [int(v) for v in np.round(p)]  # model.py:1131-1132
            else:
                pred = [float(v) for v in p]
        for i in range(n):  # pdf[y].where(pdf[y].notna(), predicted)
            if dirty.value(y, i) is None:
                v = pred[i]
                if ycol.dtype == object:
                    ycol[i] = v
                elif dirty.kinds[y] == "str":
                    ycol[i] = -1 if v is None else int(v)
                else:
                    ycol[i] = np.nan if v is None else float(v)
    if pmf_mode:
        dpos = {r: i for i, r in enumerate(dirty_pos)}
        out = []
        for (r, a, cur) in error_cells:
            rid = cast_to_string(tbl.kinds[row_id], tbl.value(row_id, r))
            if a in continuous:
                out.append((rid, a, cur, None, cast_to_string(dirty.kinds[a], dirty.value(a, dpos[r]))))
            else:
                classes, probs = pmf_of[(r, a)]
                out.append((rid, a, cur, classes, probs))
        return out
    if repair_data:
        out = base.copy()
        for i, r in enumerate(dirty_pos):
            for c in columns:
                out.cols[c][r] = dirty.cols[c][i]
        return out
    # ---- flatten + join + filter (model.py:1398-1401) ----
    dpos = {r: i for i, r in enumerate(dirty_pos)}
    out = []
    for (r, a, cur) in error_cells:
        rep = cast_to_string(dirty.kinds[a], dirty.value(a, dpos[r]))
        if rep is None or cur is None or str(cur) != str(rep):  # repaired IS NULL OR NOT(cur <=> repaired)
            rid = cast_to_string(tbl.kinds[row_id], tbl.value(row_id, r))
            out.append((rid, a, cur, rep))
    for (r, a, cur, v) in by_rules:  # model.py:1403-1404: rule repairs are appended unfiltered
        out.append((cast_to_string(tbl.kinds[row_id], tbl.value(row_id, r)), a, cur, cast_to_string(tbl.kinds[a], v)))
    return out


def run(tbl, row_id, detectors=None, targets=None, discrete_thres=80, given_error_cells=None, opts=None,
        model_provider=None, detect_errors_only=False, repair_data=False, pmf_mode=False, rules=None):
    """RepairModel.run, default / detect_errors_only / repair_data modes."""
    continuous = S.check_input_table(tbl, row_id)
    targets = targets or []
    if targets and not (set(targets) & set(tbl.names)):
        raise ValueError("Target attributes not found in input: {}".format(",".join(targets)))
    cells, target_columns, pairwise, domain_stats = error_model_detect(
        tbl, row_id, targets, discrete_thres, detectors or [], given_error_cells, continuous, opts)
    if detect_errors_only:
        return [(cast_to_string(tbl.kinds[row_id], tbl.value(row_id, r)), a, cur) for (r, a, cur) in cells]
    if not cells:
        return tbl if repair_data else []
    if rules is not None:
        rules = dict(rules)
        rules.setdefault("constraints", [d for d in (detectors or []) if d["type"] == "constraint"])
    return repair(tbl, row_id, cells, target_columns, pairwise, domain_stats, continuous, opts,
                  model_provider, repair_data=repair_data, pmf_mode=pmf_mode, rules=rules)
