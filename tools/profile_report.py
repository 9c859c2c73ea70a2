#!/usr/bin/env python
"""Turns the outputs of tools/gpu_profile.sh (gpurun_out/*_<tag>.*) into the committed summaries under
profiles/: launch list, per-kernel ncu figures, the forest kernel's shared-memory accounting, and
profiles/ncu_per_launch.json (read by bench.py for roofline.traffic).

  python tools/profile_report.py <tag> <round-label>
"""
import collections
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
PROF = os.path.join(ROOT, "profiles")
CMD = "python bench.py --forests random --steps 1 --warmup 0 --no-e2e --no-cpu-baseline"
PICK = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum"]


def raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    return idx, units, data


def num(s):
    return float(s.replace(",", "")) if s not in ("", "n/a") else 0.0


def to_bytes(v, unit):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)


def to_ms(v, unit):
    return v * {"ns": 1e-6, "us": 1e-3, "ms": 1, "s": 1e3}.get(unit, 1)


def kernel_section(name, path, lines, per_launch):
    idx, units, data = raw(path)
    for d in data:
        kname = d[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        lines.append("### `{}`".format(kname))
        for k in PICK:
            if k in idx:
                lines.append("- {} = {} {}".format(k, d[idx[k]], units[idx[k]]))
        rd = to_bytes(num(d[idx["dram__bytes_read.sum"]]), units[idx["dram__bytes_read.sum"]])
        wr = to_bytes(num(d[idx["dram__bytes_write.sum"]]), units[idx["dram__bytes_write.sum"]])
        ms = to_ms(num(d[idx["gpu__time_duration.sum"]]), units[idx["gpu__time_duration.sum"]])
        lines.append("- DRAM traffic {:.3f} GB in {:.3f} ms = {:.0f} GB/s (under ncu: serialised, cold caches)".format(
            (rd + wr) / 1e9, ms, (rd + wr) / 1e9 / (ms / 1e3)))
        lines.append("")
        per_launch.setdefault(name, []).append({"kernel": kname, "dram_bytes": rd + wr, "ms": ms})


def main():
    tag, label = sys.argv[1], sys.argv[2]
    per_launch = {}
    # ---- launch list ---------------------------------------------------------------------------
    ll = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), "launches",
                         os.path.join(OUT, "launches_%s.csv" % tag)], capture_output=True, text=True).stdout
    with open(os.path.join(PROF, "%s_launches_100Mx32.md" % label), "w") as f:
        f.write("# ncu launch list, {} (100M x 32 C4, random-init forests)\n\nCommand: `ncu --metrics "
                "gpu__time_duration.sum --clock-control none -k regex:^k_ -c 3000 --csv {}` (set-up detect + model "
                "bookkeeping + 2 full passes; times are cold-cache and serialised: compare shares)\n\n{}".format(
                    label, CMD, ll))
    # ---- streaming kernels --------------------------------------------------------------------------
    lines = ["# ncu --set full, streaming kernels of one pass over the 100M x 32 C4 table ({})\n".format(label),
             "Each kernel captured with `ncu --set full --clock-control none --import-source on -k regex:<kernel> "
             "-s <set-up launches> -c <n> {}`.\n".format(CMD)]
    for name in ("scan", "pairs", "gather", "domain"):
        p = os.path.join(OUT, "prof_%s_%s.csv" % (name, tag))
        if os.path.exists(p) and os.path.getsize(p) > 100:
            kernel_section(name, p, lines, per_launch)
    open(os.path.join(PROF, "%s_ncu_streaming_kernels.md" % label), "w").write("\n".join(lines))
    # ---- forest kernel: all launches of one pass -------------------------------------------------------
    idx, units, data = raw(os.path.join(OUT, "prof_forest_%s.csv" % tag))

    def tot(k):
        return sum(num(d[idx[k]]) for d in data)
    ms = sum(to_ms(num(d[idx["gpu__time_duration.sum"]]), units[idx["gpu__time_duration.sum"]]) for d in data)
    cyc, wf = tot("sm__cycles_elapsed.avg"), tot("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum")
    bc, inst = tot("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"), tot("smsp__inst_executed.sum")
    rd = sum(to_bytes(num(d[idx["dram__bytes_read.sum"]]), units[idx["dram__bytes_read.sum"]]) for d in data)
    wr = sum(to_bytes(num(d[idx["dram__bytes_write.sum"]]), units[idx["dram__bytes_write.sum"]]) for d in data)
    sms = 148
    kname = data[0][idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
    out = ["# ncu --set full, `{}` -- the {} launches of one pass ({})\n".format(kname, len(data), label),
           "Command: `ncu --set full --clock-control none -k regex:k_forest_predict_ranked -s 32 -c 32 {}`\n".format(CMD),
           "| quantity (sum over the launches) | value |\n|---|---:|",
           "| gpu__time_duration | {:.1f} ms |".format(ms),
           "| sm__cycles_elapsed.avg | {:.4g} cycles ({:.3f} GHz) |".format(cyc, cyc / ms / 1e6),
           "| shared-memory wavefronts | {:.4g} = **{:.3f} per cycle per SM** (pipe peak: 1) |".format(wf, wf / sms / cyc),
           "| of which bank-conflict replays | {:.4g} ({:.1f} %) |".format(bc, 100 * bc / wf),
           "| warp instructions executed | {:.4g} = {:.2f} per cycle per SM (issue peak: 4) |".format(inst, inst / sms / cyc),
           "| DRAM read + write | {:.2f} GB + {:.2f} GB |".format(rd / 1e9, wr / 1e9),
           "| registers / thread, CTA threads, grid | {}, {}, {} |".format(
               data[0][idx["launch__registers_per_thread"]], data[0][idx["launch__block_size"]],
               data[0][idx["launch__grid_size"]]), ""]
    src = os.path.join(OUT, "prof_forest1_%s_src.csv" % tag)
    if os.path.exists(src) and os.path.getsize(src) > 100:
        rows = list(csv.reader(open(src)))
        start = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
        hdr = rows[start]
        ix = {h: i for i, h in enumerate(hdr)}
        mix, lds, total = collections.Counter(), collections.defaultdict(lambda: [0.0, 0.0, 0.0]), 0.0
        for r in rows[start + 1:]:
            if len(r) < len(hdr) or r[0] == "Kernel Name":
                break
            toks = r[ix["Source"]].split()
            if not toks:
                continue
            op = toks[1] if toks[0].startswith("@") and len(toks) > 1 else toks[0]
            ex = num(r[ix["Instructions Executed"]])
            total += ex
            mix[op.split(".")[0]] += ex
            if op.startswith("LDS") or op.startswith("STS"):
                e = lds[op]
                e[0] += ex
                e[1] += num(r[ix["L1 Wavefronts Shared"]])
                e[2] += num(r[ix["L1 Wavefronts Shared Ideal"]])
        out += ["Source page of one launch (`--import-source on`), shared-memory instructions:\n",
                "| SASS | executed | wavefronts | ideal | ratio |\n|---|---:|---:|---:|---:|"]
        for op, e in sorted(lds.items(), key=lambda kv: -kv[1][1]):
            out.append("| {} | {:.3g} | {:.3g} | {:.3g} | {:.2f} |".format(op, e[0], e[1], e[2], e[1] / max(e[2], 1)))
        out += ["", "Instruction mix: " + ", ".join("{} {:.1f} %".format(k, 100 * v / total)
                                                      for k, v in mix.most_common(12)), ""]
    open(os.path.join(PROF, "%s_ncu_forest_ranked.md" % label), "w").write("\n".join(out))
    # ---- per-launch figures for bench.py ------------------------------------------------------------------
    js = {"_comment": "per-launch figures read from ncu --set full captures of `bench.py` at the named workload "
                      "(see the .md files next to this one); bench.py copies them into roofline.traffic when its "
                      "workload matches"}
    js["k_forest_predict_ranked"] = {
        "rows_per_gpu": 100000000, "cols": 32, "dram_bytes_per_launch": int((rd + wr) / len(data)),
        "smem_wavefronts_per_cycle_per_sm": round(wf / sms / cyc, 3), "launches_averaged": len(data),
        "source": "profiles/%s_ncu_forest_ranked.md" % label}
    for name, kernel in (("scan", "k_scan_hist"),):
        if name in per_launch:
            e = per_launch[name]
            js[kernel] = {"rows_per_gpu": 100000000, "cols": 32,
                          "dram_bytes_per_launch": int(sum(x["dram_bytes"] for x in e) / len(e)),
                          "launches_averaged": len(e), "source": "profiles/%s_ncu_streaming_kernels.md" % label}
    json.dump(js, open(os.path.join(PROF, "ncu_per_launch.json"), "w"), indent=1)
    print("wrote", [f for f in sorted(os.listdir(PROF)) if f.startswith(label) or f == "ncu_per_launch.json"])


if __name__ == "__main__":
    main()
