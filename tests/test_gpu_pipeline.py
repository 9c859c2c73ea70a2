"""End-to-end parity: the public API (RepairModel -> C ABI -> CUDA) against the CPU oracle on the
reference's fixtures and on seeded synthetic tables, plus the reference's own goldens."""
import os

import numpy as np
import pandas as pd
import pytest

import parity_utils as PU
from conftest import GOLDEN

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")

FAST = {"model.lgb.n_estimators": 30}


def read_csv(name, infer=True):
    df = pd.read_csv(os.path.join(GOLDEN, name), dtype=None if infer else str, keep_default_na=True)
    return df


def adult():
    df = pd.read_csv(os.path.join(GOLDEN, "adult.csv"))
    return df


def test_adult_detect_errors_only_golden():
    # tests/test_model.py:510-518 + bin/testdata/adult_repair.csv (error-cell set)
    rm, out = PU.run_product(adult(), "tid", [{"type": "null"}], mode="detect")
    got = PU.frame_tuples(out, "tid")
    exp = sorted([("3", "Sex", None), ("5", "Age", None), ("5", "Income", None), ("7", "Sex", None),
                  ("12", "Age", None), ("12", "Sex", None), ("16", "Income", None)], key=lambda t: (t[0], t[1]))
    assert got == exp


@pytest.mark.parametrize("specs,targets", [
    ([{"type": "null"}], None),
    ([{"type": "null"}], ["Sex"]),
    ([{"type": "domain", "attr": "Country", "values": ["United-States"]},
      {"type": "domain", "attr": "Income", "values": ["LessThan50K", "MoreThan50K"]}], None),
    ([{"type": "regex", "attr": "Country", "regex": "United-States"},
      {"type": "regex", "attr": "Relationship", "regex": "(Husband|Own-child|Not-in-family)"}],
     ["Country", "Relationship"]),
    ([{"type": "constraint", "path": os.path.join(GOLDEN, "adult_constraints.txt")}], ["Sex", "Relationship"]),
    ([{"type": "constraint", "path": os.path.join(GOLDEN, "adult_constraints.txt"), "targets": ["Sex"]}],
     ["Sex", "Relationship"]),
    ([], None),  # default detectors: NULL + DomainValues(autofill, 4) per attribute
])
def test_adult_detect_parity(specs, targets):
    got, want, _ = PU.run_both_frame(adult(), "tid", specs, targets, mode="detect")
    assert got == want
    assert len(got) > 0


def test_adult_repair_parity_and_golden_cells():
    got, want, info = PU.run_both_frame(adult(), "tid", [{"type": "null"}], opts=FAST)
    assert got == want
    # bin/testdata/adult_repair.csv: the 7 repaired cells (values depend on LightGBM, not pinned)
    golden = pd.read_csv(os.path.join(GOLDEN, "adult_repair.csv"))
    assert sorted((str(t), a) for t, a in zip(golden.tid, golden.attribute)) == sorted((g[0], g[1]) for g in got)
    assert all(g[3] is not None for g in got)
    assert info["gpu_launches"] > 0


def test_adult_repair_with_pair_stats():
    opts = dict(FAST)
    opts["error.pairwise_freq_ratio_threshold"] = 1.0
    got, want, _ = PU.run_both_frame(adult(), "tid", [], opts=opts)
    assert got == want


def test_adult_given_error_cells():
    dirty = pd.read_csv(os.path.join(GOLDEN, "adult_dirty.csv"))
    rm, out = PU.run_product(adult(), "tid", [{"type": "null"}], opts=FAST, given=dirty)
    got = sorted((g[0], g[1]) for g in PU.frame_tuples(out, "tid"))
    assert got == sorted((str(t), a) for t, a in zip(dirty.tid, dirty.attribute))


def test_adult_repair_data():
    rm, out = PU.run_product(adult(), "tid", [{"type": "null"}], opts=FAST, mode="repair_data")
    assert len(out) == 20 and out.isna().sum().sum() == 0
    src = adult()
    same = (out.fillna("<n>") == src.fillna("<n>"))
    assert int((~same).sum().sum()) == 7  # exactly the 7 NULL cells changed


def hospital():
    return pd.read_csv(os.path.join(GOLDEN, "hospital.csv"), dtype=str).astype({"tid": int})


def test_hospital_detect_parity_null_and_constraints():
    # config C2: Null + Constraint detectors, pair stats enabled like tests/test_model_perf.py:204-209
    specs = [{"type": "null"}, {"type": "constraint", "path": os.path.join(GOLDEN, "hospital_constraints.txt")}]
    opts = {"error.pairwise_freq_ratio_threshold": 1.0}
    got, want, _ = PU.run_both_frame(hospital(), "tid", specs, opts=opts, mode="detect")
    assert got == want
    assert len(got) > 2000


def test_hospital_repair_parity():
    specs = [{"type": "null"}, {"type": "constraint", "path": os.path.join(GOLDEN, "hospital_constraints.txt")}]
    opts = {"error.pairwise_freq_ratio_threshold": 1.0, "model.lgb.n_estimators": 5}
    got, want, _ = PU.run_both_frame(hospital(), "tid", specs, opts=opts,
                                     targets=["City", "State", "ZipCode", "EmergencyService", "HospitalType"])
    assert got == want
    assert len(got) > 0


def boston():
    df = pd.read_csv(os.path.join(GOLDEN, "boston.csv"))
    # tests/test_model_perf.py:74-76: CHAS and RAD are strings, ZN / TAX ints, the rest doubles
    df["CHAS"] = df["CHAS"].map(lambda v: None if v != v else str(v))
    df["RAD"] = df["RAD"].map(lambda v: None if v != v else str(int(v)) if float(v).is_integer() else str(v))
    return df


def test_boston_outlier_detect_parity():
    got, want, _ = PU.run_both_frame(boston(), "tid", [{"type": "null"}, {"type": "outlier"}], mode="detect")
    assert got == want
    assert any(g[1] == "CRIM" and g[2] is not None for g in got)


def test_boston_regression_repair_parity():
    opts = {"model.lgb.n_estimators": 20}
    got, want, _ = PU.run_both_frame(boston(), "tid", [{"type": "null"}], opts=opts)
    assert got == want
    assert {g[1] for g in got} >= {"CRIM", "TAX", "LSTAT"}


@pytest.mark.parametrize("n_rows,n_cols,c4", [(3000, 8, True), (20000, 16, True), (20000, 16, False)])
def test_synthetic_parity(n_rows, n_cols, c4):
    got, want, info = PU.run_both_synth(n_rows, n_cols, seed=n_rows, n_estimators=10, c4=c4)
    assert got == want
    assert len(got) > 0.005 * n_rows


def test_synthetic_detect_parity_larger():
    got, want, _ = PU.run_both_synth(200000, 16, seed=1, c4=True, mode="detect")
    assert got == want


def test_escaped_column_names_and_mixed_types():
    # tests/test_model.py:687-707 shape: spaces in names, string + double columns
    rows = [(1, "1", None, 1.0), (2, None, "test-2", 2.0), (3, "1", "test-1", 1.0), (4, "2", "test-2", 2.0),
            (5, "2", "test-2", 1.0), (6, "1", "test-1", 1.0)]
    df = pd.DataFrame(rows, columns=["t i d", "x x", "y y", "z z"])
    got, want, _ = PU.run_both_frame(df, "t i d", [{"type": "null"}], thres=10, opts=FAST)
    assert got == want
    assert [(g[0], g[1]) for g in got] == [("1", "y y"), ("2", "x x")]


def test_error_paths():
    from repair import RepairModel
    from repair.utils import AnalysisException
    df = pd.DataFrame([(1, 1, None), (1, 1, "test-1"), (1, 2, "test-1")], columns=["tid", "x", "y"])
    with pytest.raises(AnalysisException, match="Uniqueness does not hold in column 'tid'"):
        RepairModel().setInput(df).setRowId("tid").run()
    df = pd.DataFrame([(1, None), (2, "test-1")], columns=["tid", "x"])
    with pytest.raises(AnalysisException, match="A least three columns"):
        RepairModel().setInput(df).setRowId("tid").run()
    # tests/test_model.py:797-815: no discretizable feature
    df = pd.DataFrame([(1, "1", None), (2, "1", None), (3, "1", "test-1"), (4, "1", "test-1"), (5, "1", "test-1"),
                       (6, "1", None)], columns=["tid", "x", "y"])
    from repair.errors import NullErrorDetector
    with pytest.raises(ValueError, match="At least one valid discretizable feature is needed"):
        RepairModel().setInput(df).setRowId("tid").setErrorDetectors([NullErrorDetector()]).run()


def test_standalone_detector_protocol():
    # ErrorDetector.setUp(...).detect() as tests/test_errors.py drives it
    from repair.errors import ConstraintErrorDetector, NullErrorDetector, RegExErrorDetector
    df = adult()
    out = NullErrorDetector().setUp("tid", df, [], ["Sex", "Age", "Income"]).detect()
    assert sorted(zip(out.tid.tolist(), out.attribute.tolist())) == \
        [(3, "Sex"), (5, "Age"), (5, "Income"), (7, "Sex"), (12, "Age"), (12, "Sex"), (16, "Income")]
    out = RegExErrorDetector("Country", "United-States").setUp("tid", df, [], ["Unknown", "Country"]).detect()
    assert sorted(out.tid.tolist()) == [7, 19]
    path = os.path.join(GOLDEN, "adult_constraints.txt")
    out = ConstraintErrorDetector(path, targets=["Relationship"]).setUp("tid", df, [], ["Relationship", "Sex"]).detect()
    assert sorted(zip(out.tid.tolist(), out.attribute.tolist())) == [(4, "Relationship"), (11, "Relationship")]
    out = NullErrorDetector().setUp("tid", df, [], ["Non-existent"]).detect()
    assert len(out) == 0


def test_encoded_output_with_frozen_models_matches_string_frame():
    """bench.py's path (frozen models, dictionary-encoded frame assembled on the device) == public API."""
    from repair.engine import Engine
    from repair.errors import ErrorModelOptions
    from repair.model import repair_cells
    spec, names, codes, enc, specs = PU.synth_inputs(30000, 16, seed=5, c4=True)
    opts = {"error.pairwise_freq_ratio_threshold": 1.0, "model.lgb.n_estimators": 8, "model.max_training_row_num": 1500}
    rm, out = PU.run_product(enc, "tid", specs, None, 80, opts, encoded=True)
    want = PU.frame_tuples(out, "tid")
    engine = Engine(enc, 0)
    try:
        res = engine.detect(specs, [], 80, ErrorModelOptions.resolve({k: str(v) for k, v in opts.items()}))
        cells = repair_cells(rm, engine, enc, res, [], models=rm.last_run["models"], encoded_output=True)
    finally:
        engine.close()
    got = []
    for a, rows, cur, rep in cells:
        col = enc.by_name[a]
        for r, c, p in zip(rows.tolist(), col.decode(cur), col.decode(rep)):
            got.append((str(r), a, c, p))
    assert sorted(got, key=lambda t: (t[0], t[1])) == want


@pytest.mark.parametrize("mode", ["pmf", "prob", "score", "mlr"])
def test_adult_pmf_modes_parity(mode):
    # model.py:1350-1390; schemas as asserted by tests/test_model.py:1008-1119
    got, want = PU.run_both_pmf(adult(), "tid", [{"type": "null"}], mode, opts=dict(FAST))
    assert len(got) == (7 if mode != "mlr" else len(want)) and len(want) > 0
    assert got == want


def test_pmf_with_cost_weighting_and_topk():
    df = adult()
    df.loc[df.tid == 4, "Sex"] = "Femal"      # a dirty, non-NULL cell so that costs apply
    specs = [{"type": "null"}, {"type": "domain", "attr": "Sex", "values": ["Male", "Female"]}]
    # the regex (Male|Female) finds "Female" inside "Femal"? no: "Femal" does not contain either
    opts = dict(FAST)
    opts.update({"_cost": 1, "repair.pmf.prob_top_k": 3, "repair.pmf.prob_threshold": 0.05})
    got, want = PU.run_both_pmf(df, "tid", specs, "pmf", opts=opts)
    assert got == want
    assert any(g[2] == "Femal" for g in got)


def test_pmf_mixed_types_schema():
    # tests/test_model.py:1060-1082 shape: continuous targets get a single certain candidate
    rows = [(1, 0, 1.0, 1.0, "a"), (2, 1, 1.5, 1.5, "b"), (3, 0, 1.4, None, "b"), (4, 1, 1.3, 1.3, "b"),
            (5, 1, 1.2, 1.1, "b"), (6, 1, 1.1, 1.2, "b"), (7, 0, None, 1.4, "b"), (8, 1, 1.4, 1.0, "b"),
            (9, 0, 1.2, 1.1, "b"), (10, None, 1.3, 1.2, "b"), (11, 0, 1.0, 1.9, "b"), (12, 0, 1.9, 1.2, "b"),
            (13, 0, 1.2, 1.3, "b"), (14, 0, 1.8, 1.2, None), (15, 0, 1.3, 1.1, "b"), (16, 1, 1.3, 1.0, "b"),
            (17, 0, 1.3, 1.0, "b")]
    df = pd.DataFrame(rows, columns=["tid", "v1", "v2", "v3", "v4"])
    got, want = PU.run_both_pmf(df, "tid", [{"type": "null"}], "pmf", opts=dict(FAST))
    assert got == want
    assert sorted((g[0], g[1]) for g in got) == [("10", "v1"), ("14", "v4"), ("3", "v3"), ("7", "v2")]


# ---- rule-based repairs (setRepairByRules) --------------------------------------------------------
def test_repair_by_functional_deps_reference_kat(tmp_path):
    # tests/test_model.py:892-927: y follows x by the constraint; no statistical model is involved
    df = pd.DataFrame({"tid": [1, 2, 3, 4, 5, 6], "x": ["1", "2", "1", "2", "2", "3"],
                       "y": ["test-1", "test-2", None, "test-2", None, None]})
    path = str(tmp_path / "c.txt")
    with open(path, "w") as f:
        f.write("t1&t2&EQ(t1.x,t2.x)&IQ(t1.y,t2.y)")
    specs = [{"type": "null"}, {"type": "constraint", "path": path}]
    given = pd.DataFrame({"tid": [3, 5, 6], "attribute": ["y", "y", "y"]})
    got, want, info = PU.run_both_frame(df, "tid", specs, given=given, rules={"fd": True})
    assert got == want == [("3", "y", None, "test-1"), ("5", "y", None, "test-2"), ("6", "y", None, None)]
    assert info["gpu_launches"] > 0


def test_repair_by_nearest_values_reference_kat():
    # tests/test_model.py:929-973
    df = pd.DataFrame({"tid": [1, 3, 4, 5, 6], "v0": ["100%", "32%", "1xx%", "100x", "12x"],
                       "v1": [100, 101, 1, 2, 300], "v2": ["a", "b", "a", "b", "a"], "v3": [1.0, 1.1, 1.3, 0.6, 0.8]})
    given = pd.DataFrame({"tid": [4, 5, 6, 3, 5, 6, 5], "attribute": ["v0", "v0", "v0", "v1", "v1", "v1", "v2"]})
    rules = {"nearest": True, "cf_targets": ["v0", "v1"], "threshold": 2.0}
    kat = [("3", "v1", "101", "100"), ("4", "v0", "1xx%", "100%"), ("5", "v0", "100x", "100%"),
           ("5", "v1", "2", "1"), ("6", "v0", "12x", "32%"), ("6", "v1", "300", "100")]
    got, want, _ = PU.run_both_frame(df, "tid", [], targets=["v0", "v1"], given=given, rules=rules)
    assert got == want == kat
    # (5, v2) is left to the statistical model (the reference's LightGBM says 'a'; a model that says 'b'
    # reproduces the current value, which the output filter drops)
    got, want, _ = PU.run_both_frame(df, "tid", [], given=given, rules=rules, opts=FAST)
    assert got == want
    assert [g for g in got if g[1] != "v2"] == kat
    assert [g for g in got if g[1] == "v2"] in ([], [("5", "v2", "b", "a")])
    # repair_data: rule repairs and model repairs both land in the table
    rm, out = PU.run_product(df, "tid", [], given=given, rules=rules, opts=FAST, mode="repair_data")
    assert out["v0"].tolist() == ["100%", "32%", "100%", "100%", "32%"] and out["v1"].tolist() == [100, 100, 1, 1, 100]
    # nearest values cannot be combined with the pmf modes (model.py:1501-1507)
    with pytest.raises(ValueError, match="Cannot repair data by nearest values"):
        PU.run_product(df, "tid", [], given=given, rules=rules, mode="pmf")


def test_hospital_rule_based_repair_parity():
    # the reference's best hospital configuration turns the rules on (tests/test_model_perf.py:296-309)
    specs = [{"type": "null"}, {"type": "constraint", "path": os.path.join(GOLDEN, "hospital_constraints.txt")}]
    opts = {"error.pairwise_freq_ratio_threshold": 1.0, "model.lgb.n_estimators": 5}
    targets = ["City", "State", "ZipCode", "EmergencyService", "HospitalType", "CountyName"]
    rules = {"nearest": True, "threshold": 2.0, "fd": True}
    got, want, info = PU.run_both_frame(hospital(), "tid", specs, opts=opts, targets=targets, rules=rules)
    assert got == want
    kinds = {m[0] for _, m in info["rm"].last_run["models"]}
    assert "fd" in kinds and len(got) > 100


def test_rule_based_pmf_with_fd_model():
    # FunctionalDepModel.predict_proba (model.py:89-100): one-hot for a known determinant, nothing otherwise
    df = pd.DataFrame({"tid": list(range(1, 13)), "x": ["1", "2", "1", "2", "2", "3"] * 2,
                       "y": ["p", "q", None, "q", None, None] * 2, "z": ["a", "b", "a", "b", "b", "a"] * 2})
    specs = [{"type": "null"}, {"type": "constraint", "constraints": "x->y"}]
    given = pd.DataFrame({"tid": [3, 5, 6, 9, 11, 12], "attribute": ["y"] * 6})
    rm, out = PU.run_product(df, "tid", specs, given=given, rules={"fd": True}, mode="prob")
    none = lambda v: None if v is None or v != v else v  # noqa: E731
    got = sorted((int(r["tid"]), r["attribute"], none(r["repaired"]), none(r["prob"])) for r in out.to_dict("records"))
    assert got == [(3, "y", "p", 1.0), (5, "y", "q", 1.0), (6, "y", None, None), (9, "y", "p", 1.0),
                   (11, "y", "q", 1.0), (12, "y", None, None)]


def test_repair_by_regex_structure_parity():
    # model.rule.repair_by_regex (model.py:633-651, RepairSuite.scala:514-547): values of the wrong shape
    # are rebuilt from the detector's regex -- some into strings the column never held
    rng = np.random.default_rng(3)
    n = 400
    good = ["%d patients" % v for v in rng.integers(1, 40, size=n)]
    score = ["%d%%" % v for v in rng.integers(1, 9, size=n)]
    kind = [["a", "b", "c"][v] for v in rng.integers(0, 3, size=n)]
    df = pd.DataFrame({"tid": np.arange(n), "sample": good, "score": score, "kind": kind})
    for i in range(0, n, 9):
        df.loc[i, "sample"] = df.loc[i, "sample"].replace("ients", "ixxts")        # right shape, wrong letters
    for i in range(4, n, 23):
        df.loc[i, "sample"] = "x" + df.loc[i, "sample"][1:]                         # not repairable by the regex
    for i in range(2, n, 31):
        df.loc[i, "score"] = df.loc[i, "score"].replace("%", "x")
    df.loc[7, "sample"] = None
    specs = [{"type": "null"}, {"type": "regex", "attr": "sample", "regex": "^[0-9]{1,3} patients$"},
             {"type": "regex", "attr": "score", "regex": "^[0-9]{1,3}%$"},
             {"type": "regex", "attr": "kind", "regex": "^[a-c]{1,"}]                # broken regex: repairs nothing
    got, want, info = PU.run_both_frame(df, "tid", specs[:3], rules={"regex": True}, opts=FAST)
    assert got == want
    fixed = [g for g in got if g[2] is not None and "xx" in g[2] and not g[2].startswith("x")]
    assert len(fixed) >= 40 and all(g[3] == g[2].replace("ixxts", "ients") for g in fixed)
    assert any(g[2] is not None and g[2].startswith("x") for g in got)               # left to the statistical model
    assert info["rm"].last_run["detect"].n_cells["sample"] < len([g for g in got if g[1] == "sample"])


def test_arrow_input_gives_the_same_repairs_as_pandas_input():
    import pyarrow as pa
    from repair import NullErrorDetector, RepairModel
    df = adult()
    want = RepairModel().setInput(df).setRowId("tid").setErrorDetectors([NullErrorDetector()]) \
        .option("model.lgb.n_estimators", "30").option("model.hp.max_evals", "1").run()
    got = RepairModel().setArrowInput(pa.Table.from_pandas(df, preserve_index=False)).setRowId("tid") \
        .setErrorDetectors([NullErrorDetector()]).option("model.lgb.n_estimators", "30") \
        .option("model.hp.max_evals", "1").run()
    assert isinstance(got, pa.Table)   # Arrow in, Arrow out
    got = got.to_pandas()
    assert PU.frame_tuples(got, "tid") == PU.frame_tuples(want, "tid") and len(got) == 7


def test_inequality_denial_constraints_parity():
    """EQ(keys) & LT / GT(b): "another row of the key group has a larger / smaller b" (the oracle
    evaluates these by brute force over all row pairs)."""
    rng = np.random.default_rng(11)
    n = 600
    df = pd.DataFrame({"tid": np.arange(n), "a": rng.integers(0, 12, n).astype(str), "g": rng.integers(0, 3, n).astype(str),
                       "b": rng.integers(0, 30, n), "s": rng.choice(list("pqrstu"), n), "z": rng.integers(0, 5, n).astype(str)})
    df["a"] = df["a"].where(rng.random(n) > 0.05, None)
    df["s"] = df["s"].where(rng.random(n) > 0.05, None)
    df["b"] = df["b"].astype(float).where(rng.random(n) > 0.05, np.nan)
    for cons in ("t1&t2&EQ(t1.a,t2.a)&LT(t1.b,t2.b)", "t1&t2&EQ(t1.a,t2.a)&EQ(t1.g,t2.g)&GT(t1.s,t2.s)",
                 "t1&t2&GT(t1.b,t2.b)&EQ(t1.g,t2.g)"):
        specs = [{"type": "constraint", "constraints": cons}]
        got, want, _ = PU.run_both_frame(df, "tid", specs, mode="detect")
        assert got == want and len(got) > 100, cons


GENERAL_DCS = [
    # several IQs / several inequalities / cross-attribute predicates: every shape the reference's parser
    # accepts (DenialConstraintsSuite.scala:27-52) is evaluated, none raises
    "t1&t2&EQ(t1.a,t2.a)&IQ(t1.b,t2.b)&IQ(t1.s,t2.s)",
    "t1&t2&LT(t1.b,t2.b)&GT(t1.w,t2.w)",
    "t1&t2&LT(t1.b,t2.b)&GT(t1.w,t2.w)&EQ(t1.b,t2.b)",          # DenialConstraintsSuite.scala:41-44: unsatisfiable
    "t1&t2&EQ(t1.a,t2.a)&LT(t1.b,t2.b)&GT(t1.w,t2.w)",
    "t1&t2&EQ(t1.a,t2.z)&IQ(t1.s,t2.s)",                        # cross-attribute equality (strings)
    "t1&t2&LT(t1.b,t2.w)&EQ(t1.g,t2.g)",                        # cross-attribute inequality (numbers)
    "t1&t2&IQ(t1.a,t2.a)&IQ(t1.g,t2.g)",
    "t1&t2&GT(t1.s,t2.s)&LT(t1.a,t2.a)&EQ(t1.g,t2.g)&IQ(t1.z,t2.z)",
]


@pytest.mark.parametrize("cons", GENERAL_DCS)
def test_general_two_tuple_denial_constraints_parity(cons):
    """ErrorDetectorApi.scala:211-229 evaluates ANY parsed predicate list as an EXISTS self semi-join; the
    GPU path decides distinct projections (dr_key_presence / dr_dc_exists / dr_key_flag) and must agree
    with the oracle's brute force over all row pairs."""
    rng = np.random.default_rng(21)
    n = 500
    df = pd.DataFrame({"tid": np.arange(n), "a": rng.integers(0, 9, n).astype(str), "g": rng.integers(0, 3, n).astype(str),
                       "b": rng.integers(0, 12, n), "w": rng.integers(3, 15, n),
                       "s": rng.choice(list("pqrstu"), n), "z": rng.integers(0, 9, n).astype(str)})
    df["a"] = df["a"].where(rng.random(n) > 0.05, None)
    df["s"] = df["s"].where(rng.random(n) > 0.05, None)
    df["b"] = df["b"].astype(float).where(rng.random(n) > 0.05, np.nan)
    df["w"] = df["w"].astype(float).where(rng.random(n) > 0.03, np.nan)
    specs = [{"type": "constraint", "constraints": cons}]
    got, want, _ = PU.run_both_frame(df, "tid", specs, mode="detect")
    assert got == want, cons
    if "EQ(t1.b,t2.b)" not in cons:
        assert len(got) > 20, cons


def test_general_constraint_beyond_the_projection_bitmap(monkeypatch):
    """Projection spaces larger than the presence bitmap fall back to sorted distinct keys."""
    from repair import engine as E
    monkeypatch.setattr(E, "MAX_PROJECTION_BITS", 64)
    test_general_two_tuple_denial_constraints_parity(GENERAL_DCS[0])
    test_general_two_tuple_denial_constraints_parity(GENERAL_DCS[5])


def test_fd_constraint_with_a_huge_key_space_uses_the_hash_table(monkeypatch):
    """Key spaces beyond the direct min / max tables go through the hash table (the limit is lowered here
    so that a three-attribute key of a small table crosses it)."""
    from repair import engine as E
    from repair._native import Context
    rng = np.random.default_rng(12)
    n, groups = 4000, 500
    trip = rng.integers(0, 600, size=(groups, 3))
    pick = rng.integers(0, groups, size=n)
    df = pd.DataFrame({"tid": np.arange(n), "k1": ["a%03d" % v for v in trip[pick, 0]],
                       "k2": ["b%03d" % v for v in trip[pick, 1]], "k3": ["c%03d" % v for v in trip[pick, 2]],
                       "y": (pick % 7).astype(str), "w": rng.integers(0, 40, n)})
    df.loc[rng.random(n) < 0.03, "y"] = "x"
    df.loc[rng.random(n) < 0.02, "k2"] = None
    df.loc[rng.random(n) < 0.02, "y"] = None
    monkeypatch.setattr(E, "MAX_FD_KEY_SPACE", 1 << 12)
    calls = []
    real = Context.dc_hash_build
    monkeypatch.setattr(Context, "dc_hash_build", lambda self, *a: (calls.append(1), real(self, *a))[1])
    for cons in ("t1&t2&EQ(t1.k1,t2.k1)&EQ(t1.k2,t2.k2)&EQ(t1.k3,t2.k3)&IQ(t1.y,t2.y)",
                 "t1&t2&EQ(t1.k1,t2.k1)&EQ(t1.k2,t2.k2)&EQ(t1.k3,t2.k3)&LT(t1.w,t2.w)",
                 "t1&t2&EQ(t1.k1,t2.k1)&EQ(t1.k3,t2.k3)&GT(t1.w,t2.w)"):
        got, want, _ = PU.run_both_frame(df, "tid", [{"type": "constraint", "constraints": cons}], mode="detect")
        assert got == want and len(got) > 50, cons
    assert len(calls) == 3


def test_domain_pruning_in_one_call_equals_the_per_cell_kernel():
    """dr_domain_prune (top-1 per combination of correlated values, bitmap-driven, all targets at once) clears
    exactly the bits the per-cell kernel dr_domain_score + dr_bitmap_clear_rows clears."""
    from repair.engine import Engine
    from repair.errors import ErrorModelOptions
    from repair.table import EncodedTable
    spec, names, codes, enc, specs = PU.synth_inputs(60000, 16, seed=3, c4=True)
    opts = ErrorModelOptions.resolve({"error.pairwise_freq_ratio_threshold": "1.0",
                                      "error.domain_threshold_beta": "0.3"})
    got = {}
    for per_cell in (False, True):
        eng = Engine(enc, 0)
        eng.domain_per_cell = per_cell
        try:
            res = eng.detect(specs, [], 80, opts)
            got[per_cell] = ({a: b.cpu().numpy().copy() for a, b in res.bitmaps.items()}, dict(res.n_cells),
                             int(res.weak_removed_dev.sum().item()), dict(res.n_cells_detected))
        finally:
            eng.close()
    assert got[False][1] == got[True][1] and got[False][2] == got[True][2]
    assert got[False][2] > 0 and got[False][1] != got[False][3]          # something was pruned
    for a in got[True][0]:
        assert np.array_equal(got[False][0][a], got[True][0][a]), a


def test_training_data_rebalancing_reference_kat():
    # tests/test_model.py:1197-1215: mixed_input with setTrainingDataRebalancingEnabled(True) repairs its four cells
    from repair import RepairModel
    rows = [(1, 0, 1.0, 1.0, "a"), (2, 1, 1.5, 1.5, "b"), (3, 0, 1.4, None, "b"), (4, 1, 1.3, 1.3, "b"),
            (5, 1, 1.2, 1.1, "b"), (6, 1, 1.1, 1.2, "b"), (7, 0, None, 1.4, "b"), (8, 1, 1.4, 1.0, "b"),
            (9, 0, 1.2, 1.1, "b"), (10, None, 1.3, 1.2, "b"), (11, 0, 1.0, 1.9, "b"), (12, 0, 1.9, 1.2, "b"),
            (13, 0, 1.2, 1.3, "b"), (14, 0, 1.8, 1.2, None), (15, 0, 1.3, 1.1, "b"), (16, 1, 1.3, 1.0, "b"),
            (17, 0, 1.3, 1.0, "b")]
    df = pd.DataFrame(rows, columns=["tid", "v1", "v2", "v3", "v4"])
    from repair import NullErrorDetector
    out = RepairModel().setInput(df).setRowId("tid").setErrorDetectors([NullErrorDetector()]) \
        .setTrainingDataRebalancingEnabled(True) \
        .option("model.hp.max_evals", "1").option("model.lgb.n_estimators", "30").run()
    none = lambda v: None if v is None or v != v else v  # noqa: E731
    got = sorted((int(r["tid"]), r["attribute"], none(r["current_value"])) for r in out.to_dict("records"))
    assert got == [(3, "v3", None), (7, "v2", None), (10, "v1", None), (14, "v4", None)]
    assert all(none(r) is not None for r in out["repaired"].tolist())
