#!/usr/bin/env python
"""Markdown summary of the committed bench lines (profiles/*.json) for DESIGN.md / README.md.

  python tools/results_table.py profiles/bench_r2_final_100Mx32_1gpu.json [2-GPU line] [12.5M line] [reference line]
"""
import json
import sys


def load(path):
    return json.loads(open(path).read().strip().splitlines()[-1])


def main():
    one = load(sys.argv[1])
    two = load(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2] != "-" else None
    small = load(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3] != "-" else None
    ref = load(sys.argv[4]) if len(sys.argv) > 4 and sys.argv[4] != "-" else None
    k = one["kernels"]
    e = one.get("e2e", {})
    rows = []
    add = lambda a, b: rows.append("| {} | {} |".format(a, b))  # noqa: E731
    add("full pass, table resident in HBM (100M × 32, trained forests)",
        "**{:.0f} ms** = {:.0f} M rows/s (round 1: 464 ms)".format(one["ms_per_step"], one["value"] / 1e6))
    add("detect + statistics phase", "{:.1f} ms = {:.1f} G rows/s (round 1: 25.6 ms)".format(
        one["ms_detect_phase"], one["rows_scanned_per_sec"] / 1e9))
    add("repair phase", "{:.0f} ms; {:.1f} M error cells → {:.0f} M cells/s".format(
        one["ms_repair_phase"], one["error_cells"] / 1e6, one["cells_repaired_per_sec"] / 1e6))
    if e:
        st = e.get("stamps_s", {})
        add("**end to end through the public API** (`setArrowInput(host table).run()` → host Arrow frame; "
            "{:.1f} GB in, {:.0f} MB out per pass)".format(e["h2d_bytes_per_step"] / 1e9, e["d2h_bytes_per_step"] / 1e6),
            "**{:.0f} ms** = {:.0f} M rows/s: ingest {:.0f} ms (copy {:.0f} ms = {:.0f} GB/s, encode {:.0f} ms), "
            "detect + repair {:.0f} ms, egress {:.0f} ms; frame identical to the resident pass: {}".format(
                e["ms_per_step"], e["value"] / 1e6, e["ingest_s"] * 1e3, e["ingest_copy_s"] * 1e3,
                e["h2d_bytes_per_step"] / max(e["ingest_copy_s"], 1e-9) / 1e9, e["ingest_encode_s"] * 1e3,
                (st.get("repair_done", 0) - st.get("engine_ready", 0) - e["egress_s"]) * 1e3, e["egress_s"] * 1e3,
                e.get("frame_equals_resident_pass")))
    if "scan_hist" in k:
        s = k["scan_hist"]
        add("`k_scan_hist` ({:.1f} GB algorithmic)".format(s["algorithmic_gb"]),
            "{:.2f} ms = **{:.2f} TB/s = {:.0f} % of the measured {:.2f} TB/s copy peak**".format(
                s["ms"], s["achieved_gbs"] / 1e3, 100 * s["frac_of_hbm_peak"], one["roofline_scan"]["peak"] / 1e3))
    f = k.get("forest_predict_ranked")
    if f:
        rf = one.get("roofline_forest", {})
        add("`k_forest_predict_ranked` (32 launches)",
            "{:.0f} ms (round 1: 421 ms), {:.2f}·10¹² tree levels/s; ncu: {} shared-memory wavefronts / cycle / SM".format(
                f["ms"], one["forest_tree_levels_per_sec"] / 1e12, rf.get("measured_pipe_utilisation")))
    for name in ("gather_rows_masked", "cooc_skip", "pair_presence", "dc_fd_build", "dc_fd_flag", "domain_prune",
                 "bitmaps_to_rows_many"):
        if name in k:
            add("`{}` ({} call{})".format(name, k[name]["calls"], "" if k[name]["calls"] == 1 else "s"),
                "{:.2f} ms".format(k[name]["ms"]))
    v = one.get("verify")
    if v:
        add("`verify` of the last timed pass", "{} histograms, {} per-attribute cell counts, {} sampled repairs re-evaluated "
            "by the oracle's C forest: {} mismatches".format(v["hist_columns_checked"], v["cell_count_attrs_checked"],
                                                            v["cells_checked"],
                                                            v["hist_mismatches"] + v["cell_count_mismatches"] + v["mismatches"]))
    add("model training (32 models, 232 800 trees)", "{:.0f} s (`dr_gbdt_train`)".format(one["model_training_s"]))
    oc = one.get("other_configs", {})
    if "C2_hospital" in oc:
        c = oc["C2_hospital"]
        add("C2 hospital (1000 × 19, NULL + 15 constraints)", "detect only {:.0f} ms; full run incl. training {:.1f} s "
            "(training {:.1f} s); with frozen models {:.0f} ms".format(c["detect_only_s"] * 1e3, c["full_run_with_training_s"],
                                                                    c["training_s"], (c["full_run_frozen_models_s"] or 0) * 1e3))
    if "C3_10Mx16" in oc:
        c = oc["C3_10Mx16"]
        add("C3 10M × 16, NullErrorDetector, through the API", "{:.0f} ms per pass = {:.0f} M rows/s (ingest {:.0f} ms, egress "
            "{:.0f} ms); first run incl. training {:.1f} s".format(c["ms_per_pass_frozen_models"], c["rows_per_sec"] / 1e6,
                                                                c["ingest_s"] * 1e3, c["egress_s"] * 1e3,
                                                                c["full_run_with_training_s"]))
    if "C5_boston" in oc:
        c = oc["C5_boston"]
        add("C5 boston (506 × 13 numeric, NULL + IQR outlier detectors, regressors)", "detect only {:.0f} ms; full run incl. "
            "training {:.1f} s".format(c["detect_only_s"] * 1e3, c["full_run_with_training_s"]))
    if small:
        add("one GPU's share of the 8-GPU strong-scaling run (12.5M rows)", "{:.1f} ms (detect {:.1f} ms, forest {:.1f} ms) "
            "→ {:.0f} % of 1/8 of the 100M-row pass".format(small["ms_per_step"], small["ms_detect_phase"],
                                                          small["kernels"]["forest_predict_ranked"]["ms"],
                                                          100 * one["ms_per_step"] / 8 / small["ms_per_step"]))
    if two:
        e2 = two.get("e2e", {})
        add("2 GPUs, strong scaling (50M rows each, {} collectives per pass)".format(int(two.get("exchanges_per_step", 0))),
            "{:.1f} ms = {:.0f} M rows/s ({:.0f} % of 2 × 1 GPU); through the API {:.0f} ms = {:.0f} M rows/s".format(
                two["ms_per_step"], two["value"] / 1e6, 100 * one["ms_per_step"] / 2 / two["ms_per_step"],
                e2.get("ms_per_step", 0), e2.get("value", 0) / 1e6))
    cb = one.get("cpu_baseline")
    if cb:
        add("CPU oracle port inside the default run ({} threads)".format(cb["cores"]), "{:.2f} k rows/s ({})".format(
            cb["value"] / 1e3, cb["sample"]))
    if ref:
        ft = ref.get("fit", {})
        add("`--impl reference` ({} threads, same table / detectors / forests)".format(ref["cpu_baseline"]["cores"]),
            "{} rows in {:.1f} s = {:.2f} k rows/s; fit: {:.1f} s fixed + {:.2f} k rows/s marginal".format(
                ref["config"]["rows_in_sample"], ref["ms_per_step"] / 1e3, ref["value"] / 1e3, ft.get("fixed_s", 0),
                (ft.get("marginal_rows_per_sec") or 0) / 1e3))
    print("| quantity | value |\n|---|---|")
    print("\n".join(rows))


if __name__ == "__main__":
    main()
