"""Shared helpers for the parity tests: run the CUDA path (through the public API and the C ABI) and
the CPU oracle on the same input, feeding the oracle the very forests the product trained."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for _p in (ROOT, os.path.join(ROOT, "spark-data-repair-plugin_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)


def make_detectors(specs):
    from repair.errors import (ConstraintErrorDetector, DomainValues, GaussianOutlierErrorDetector,
                               NullErrorDetector, RegExErrorDetector)
    out = []
    for s in specs:
        t = s["type"]
        if t == "null":
            out.append(NullErrorDetector())
        elif t == "regex":
            out.append(RegExErrorDetector(s["attr"], s["regex"]))
        elif t == "domain":
            out.append(DomainValues(s["attr"], s.get("values", []), s.get("autofill", False),
                                    s.get("min_count_thres", 12)))
        elif t == "constraint":
            out.append(ConstraintErrorDetector(s.get("path", ""), s.get("constraints", ""), s.get("targets", [])))
        elif t == "outlier":
            out.append(GaussianOutlierErrorDetector(s.get("approx", False)))
    return out


def _nan_equal(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return a.shape == b.shape and bool(np.all((a == b) | (np.isnan(a) & np.isnan(b))))


def provider_from_product(rm, table, value_of):
    """Oracle model_provider that hands back the product's forests and cross-checks the training
    bookkeeping both sides derived independently (features, encoders, encoded sample)."""
    models = dict(rm.last_run.get("models", []))  # no models when the rules decided every cell

    def provider(ctx):
        m = models[ctx["y"]]
        if m[0] == "const":
            return {"const": None if m[1] is None else value_of(ctx["y"], m[1])}
        info = m[2]
        pctx, spec = info["ctx"], info["spec"]
        assert pctx["features"] == ctx["features"], (pctx["features"], ctx["features"])
        for pe, oe in zip(spec["encoders"], ctx["encoders"]):
            assert pe["attr"] == oe["attr"] and pe["type"] == oe["type"]
            if pe["type"] != "cont":
                want = [None if c < 0 else value_of(pe["attr"], c) for c in pe["categories"]]
                assert want == oe["categories"], (pe["attr"], want, oe["categories"])
        assert _nan_equal(pctx["X"], ctx["X"]), "encoded training matrices differ for " + ctx["y"]
        assert list(np.asarray(pctx["train_rows"])) == list(np.asarray(ctx["train_rows"]))
        classes = None
        if spec["class_codes"] is not None:
            classes = [value_of(ctx["y"], c) for c in spec["class_codes"]]
        return {"forest": spec["forest"], "classes": classes}

    return provider


def oracle_rules(rules, specs=()):
    """parity-test rule settings -> the oracle's `rules` argument."""
    if not rules:
        return None
    from oracle import repair as OR
    return {"regexs": [(s["attr"], s["regex"]) for s in specs if s["type"] == "regex"] if rules.get("regex") else [],
            "nearest_values": bool(rules.get("nearest")), "cost_targets": list(rules.get("cf_targets", [])),
            "cost_fn": (lambda a, b: float(OR.levenshtein(a, b))) if rules.get("nearest") else None,
            "merge_threshold": float(rules.get("threshold", 2.0)), "functional_deps": rules.get("fd", True),
            "max_domain_size": int(rules.get("max_domain_size", 1000))}


def run_product(inp, row_id, specs, targets=None, thres=80, opts=None, given=None, mode="repair", encoded=False,
                rules=None):
    from repair import RepairModel
    rm = RepairModel()
    if rules:
        from repair.costs import Levenshtein
        rm.setRepairByRules(True)
        if rules.get("nearest"):
            rm.setUpdateCostFunction(Levenshtein(targets=list(rules.get("cf_targets", []))))
            rm.option("model.rule.repair_by_nearest_values.disabled", "")
            rm.option("model.rule.merge_threshold", str(rules.get("threshold", 2.0)))
        if rules.get("regex"):
            rm.option("model.rule.repair_by_regex.disabled", "")
        if not rules.get("fd", True):
            rm.option("model.rule.repair_by_functional_deps.disabled", "1")
        if "max_domain_size" in rules:
            rm.option("model.rule.max_domain_size", str(rules["max_domain_size"]))
    if encoded:
        rm.setEncodedInput(inp)
    else:
        rm.setInput(inp).setRowId(row_id)
    if specs:
        rm.setErrorDetectors(make_detectors(specs))
    if targets:
        rm.setTargets(targets)
    if thres != 80:
        rm.setDiscreteThreshold(thres)
    if given is not None:
        rm.setErrorCells(given)
    # like the reference's own unit tests (tests/test_model.py:269-273), one evaluation = no search,
    # unless the test asks for one
    rm.option("model.hp.max_evals", "1")
    for k, v in (opts or {}).items():
        if k.startswith("_"):
            rm.opts[k] = v
        else:
            rm.option(k, str(v))
    if mode in ("pmf", "prob", "score", "mlr"):
        from repair.costs import Levenshtein
        if mode in ("score", "mlr") or (opts or {}).get("_cost"):
            rm.setUpdateCostFunction(Levenshtein())
        if mode in ("score", "mlr"):
            rm.setRepairDelta(int((opts or {}).get("_delta", 3)))
        rm.opts.pop("_cost", None), rm.opts.pop("_delta", None)
        kw = {"pmf": {"compute_repair_candidate_prob": True}, "prob": {"compute_repair_prob": True},
              "score": {"compute_repair_score": True}, "mlr": {"maximal_likelihood_repair": True}}[mode]
        out = rm.run(**kw)
    elif mode == "detect":
        out = rm.run(detect_errors_only=True)
    elif mode == "repair_data":
        out = rm.run(repair_data=True)
    else:
        out = rm.run()
    return rm, out


def frame_tuples(df, row_id):
    cols = [c for c in df.columns]
    out = []
    for rec in df.itertuples(index=False):
        d = dict(zip(cols, rec))
        t = [str(d[row_id]), d["attribute"], d.get("current_value")]
        if "repaired" in d:
            t.append(d["repaired"])
        out.append(tuple(None if (isinstance(x, float) and x != x) else x for x in t))
    return sorted(out, key=lambda t: (t[0], t[1]))


def run_both_frame(df, row_id, specs, targets=None, thres=80, opts=None, mode="repair", kinds=None, rules=None,
                   given=None):
    """pandas input: product through RepairModel, oracle through oracle.repair.run."""
    from oracle import repair as OR
    from oracle.table import from_pandas
    from repair.table import EncodedTable
    rm, out = run_product(df, row_id, specs, targets, thres, opts, given=given, mode=mode, rules=rules)
    got = frame_tuples(out, row_id)
    otbl = from_pandas(df)
    enc = EncodedTable.from_pandas(df, row_id)

    def value_of(attr, code):
        col = enc.by_name[attr]
        v = col.dictionary[code]
        if col.kind == "str":
            return str(v)
        return int(v) if col.kind == "int" else float(v)

    provider = provider_from_product(rm, enc, value_of) if mode == "repair" else None
    o_opts = {k: v for k, v in (opts or {}).items() if k.startswith("error.") or k in OR.DEFAULT_OPTS}
    o_given = None if given is None else [tuple(x) for x in given[[row_id, "attribute"]].itertuples(index=False)]
    want = OR.run(otbl, row_id, specs, targets, thres, o_given, o_opts, provider,
                  detect_errors_only=(mode == "detect"), rules=oracle_rules(rules, specs))
    want = sorted([tuple(w) for w in want], key=lambda t: (t[0], t[1]))
    return got, want, {"gpu_launches": rm.last_run.get("gpu_launches", 0), "rm": rm}


def synth_inputs(n_rows, n_cols, seed=0, c4=True):
    from repair import synth
    from repair.table import EncodedTable
    spec = synth.SynthSpec.c4(n_rows, n_cols, seed) if c4 else synth.SynthSpec.c3(n_rows, n_cols, seed)
    codes = synth.generate_numpy(spec)
    names = synth.column_names(n_cols)
    enc = EncodedTable.from_codes("tid", names, codes, spec.dom)
    specs = [{"type": "null"}]
    if c4 and synth.fd_constraints(n_cols):
        specs.append({"type": "constraint", "constraints": synth.fd_constraints(n_cols)})
    return spec, names, codes, enc, specs


def run_both_synth(n_rows, n_cols, seed=0, n_estimators=20, c4=True, mode="repair", max_training_rows=2000):
    """Label-encoded synthetic table (configs C3 / C4 at test size)."""
    from oracle import repair as OR
    from oracle.table import OTable
    spec, names, codes, enc, specs = synth_inputs(n_rows, n_cols, seed, c4)
    opts = {"error.pairwise_freq_ratio_threshold": 1.0, "model.lgb.n_estimators": n_estimators,
            "model.max_training_row_num": max_training_rows}
    rm, out = run_product(enc, "tid", specs, None, 80, opts, mode=mode, encoded=True)
    got = frame_tuples(out, "tid")
    otbl = OTable(["tid"] + names, ["int"] + ["str"] * n_cols,
                  [np.arange(n_rows, dtype=np.float64)] + [c.astype(np.int64) for c in codes])
    provider = provider_from_product(rm, enc, lambda attr, code: int(code)) if mode == "repair" else None
    o_opts = {"error.pairwise_freq_ratio_threshold": 1.0, "model.max_training_row_num": max_training_rows}
    want = OR.run(otbl, "tid", specs, None, 80, None, o_opts, provider, detect_errors_only=(mode == "detect"))
    fmt = (lambda v: None if v is None else "v%03d" % int(v))
    want = sorted([tuple([w[0], w[1]] + [fmt(x) for x in w[2:]]) for w in want], key=lambda t: (t[0], t[1]))
    return got, want, {"gpu_launches": rm.last_run.get("gpu_launches", 0), "rm": rm}


def run_both_pmf(df, row_id, specs, mode, targets=None, thres=80, opts=None):
    """pmf / prob / score / maximal-likelihood modes: product frame vs the oracle's restatement fed
    with the product's forests.  -> (got rows, want rows) as comparable tuples."""
    from oracle import repair as OR
    from oracle.table import from_pandas
    from repair.table import EncodedTable
    opts = dict(opts or {})
    rm, out = run_product(df, row_id, specs, targets, thres, opts, mode=mode)
    enc = EncodedTable.from_pandas(df, row_id)

    def value_of(attr, code):
        col = enc.by_name[attr]
        v = col.dictionary[code]
        return str(v) if col.kind == "str" else (int(v) if col.kind == "int" else float(v))

    provider = provider_from_product(rm, enc, value_of)
    o_opts = {k: v for k, v in opts.items() if k in OR.DEFAULT_OPTS}
    cells = OR.run(from_pandas(df), row_id, specs, targets, thres, None, o_opts, provider, pmf_mode=True)
    cost = OR.levenshtein if (mode in ("score", "mlr") or opts.get("_cost")) else None
    pmf_opts = {k: v for k, v in opts.items() if k.startswith("repair.pmf.")}
    shaped = OR.shape_pmf(cells, pmf_opts, cost)
    if mode == "pmf":
        want = [(s[0], s[1], s[2][0], tuple(s[3])) for s in shaped]
        got = [(str(r[row_id]), r["attribute"], r["current_value"], tuple((d["class"], d["prob"]) for d in r["pmf"]))
               for r in out.to_dict("records")]
    elif mode == "prob":
        want = [(s[0], s[1], s[2][0], s[3][0][0], s[3][0][1]) for s in shaped]
        got = [(str(r[row_id]), r["attribute"], r["current_value"], r["repaired"], r["prob"])
               for r in out.to_dict("records")]
    else:
        scored = OR.compute_score(shaped, OR.levenshtein)
        if mode == "score":
            want = [tuple(s) for s in scored]
            got = [(str(r[row_id]), r["attribute"], r["current_value"], r["repaired"], r["score"])
                   for r in out.to_dict("records")]
        else:
            want = [tuple(s) for s in OR.maximal_likelihood_repair(scored, int(opts.get("_delta", 3)))]
            got = [(str(r[row_id]), r["attribute"], r["current_value"], r["repaired"]) for r in out.to_dict("records")]
    key = lambda t: (t[0], t[1])  # noqa: E731
    got = [tuple(None if (isinstance(x, float) and x != x) else x for x in g) for g in got]
    return sorted(got, key=key), sorted(want, key=key)
