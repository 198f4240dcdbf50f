#!/usr/bin/env python3
"""Turn ncu outputs brought back under gpurun_out/ into the small Markdown summaries committed under profiles/.

  python tools/profile_summaries.py launches gpurun_out/bench_launches.csv profiles/r1_bench_launches.md "<command>"
  python tools/profile_summaries.py rep gpurun_out/x.ncu-rep profiles/x.md "<title>" "<free text>"
"""
import collections
import csv
import io
import re
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_registers", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "l1tex__t_sector_pipe_lsu_mem_local_op_ld_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum",
]


def launches(csv_path, out_path, command):
    rows = [r for r in csv.reader(l for l in open(csv_path) if l.startswith('"'))]
    hdr = rows[0]
    iK, iM, iV, iU = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    iID = hdr.index("ID")
    per = collections.defaultdict(dict)
    name = {}
    for r in rows[1:]:
        v = float(r[iV].replace(",", ""))
        u = r[iU]
        if r[iM] == "gpu__time_duration.sum":
            v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
        else:
            v *= {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
        per[r[iID]][r[iM]] = v
        name[r[iID]] = re.sub(r"\(.*", "", r[iK])[:70]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0, 0.0])
    for i, m in per.items():
        a = agg[name[i]]
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0)
        a[2] += m.get("dram__bytes_read.sum", 0.0)
        a[3] += m.get("dram__bytes_write.sum", 0.0)
    total = sum(a[1] for a in agg.values())
    out = ["# ncu launch list of `%s` (aggregated by kernel)" % command, "",
           "Command: `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
           "--csv` (raw CSV: `%s`)." % out_path.replace(".md", ".csv").split("/")[-1],
           "Per-launch times are cold-cache and serialised; compare shares, not absolutes. The timed region of the headline "
           "number contains only `vm_pairing_kernel<BLS381,2,4>` launches (one per step); the other kernels belong to input "
           "generation (`g1_mul_kernel`) and to the secondary configs reported under `extra`.", "",
           "| kernel | launches | total ms | share | dram read MB/launch | dram write MB/launch |", "|---|---|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.3f | %.1f%% | %.1f | %.1f |" % (k, a[0], a[1], 100 * a[1] / total, a[2] / a[0], a[3] / a[0]))
    open(out_path, "w").write("\n".join(out) + "\n")


def _opcode_mix(sass_csv_text, warps):
    """opcode mix + per-warp instruction counts from an `ncu --page source --csv --print-source sass` dump"""
    rows = list(csv.reader(io.StringIO(sass_csv_text)))
    h = None
    for i, r in enumerate(rows):
        if "Instructions Executed" in r:
            h, rows = r, rows[i + 1:]
            break
    if h is None:
        return []
    iS, iE, iT = h.index("Source"), h.index("Instructions Executed"), h.index("Thread Instructions Executed")
    ops, thr, tot = collections.Counter(), collections.Counter(), 0
    for r in rows:
        if len(r) <= iT:
            continue
        t = re.sub(r"^@!?U?P\d+\s+", "", r[iS].strip())
        op = ".".join(t.split()[0].split(".")[:2]) if t else "?"
        e = int(r[iE])
        ops[op] += e
        thr[op] += int(r[iT])
        tot += e
    out = ["", "## Executed warp instructions by opcode (SASS page)", "",
           "%.3f M warp instructions in total%s." % (tot / 1e6, (" = %.2f M per warp (%d warps)" % (tot / warps / 1e6, warps)) if warps else ""),
           "", "| opcode | share | per warp | active lanes |", "|---|---|---|---|"]
    for op, c in ops.most_common(14):
        out.append("| `%s` | %.1f%% | %s | %.1f |" % (op, 100.0 * c / tot, ("%.0f k" % (c / warps / 1e3)) if warps else "-", thr[op] / max(c, 1)))
    return out


def rep(rep_path, out_path, title, text, row=0, warps=0):
    """rep_path: an .ncu-rep, or the CSV of its raw page (`ncu -i x.ncu-rep --page raw --csv`); row = which captured kernel"""
    if rep_path.endswith(".csv"):
        raw = open(rep_path).read()
    else:
        raw = subprocess.run(["ncu", "-i", rep_path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2 + row]
    out = ["# " + title, "", text, "", "Kernel: `%s`" % vals[hdr.index("Kernel Name")][:160] if "Kernel Name" in hdr else "", "",
           "| metric | value | unit |", "|---|---|---|"]
    for m in METRICS:
        if m in hdr:
            i = hdr.index(m)
            out.append("| %s | %s | %s |" % (m, vals[i], units[i]))
    stalls = [(h, vals[hdr.index(h)]) for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h]
    out += ["", "## Warp stall reasons (cycles per issued instruction)", "", "| reason | value |", "|---|---|"]
    for h, v in sorted(stalls, key=lambda kv: -float(kv[1] or 0)):
        if float(v or 0) >= 0.01:
            out.append("| %s | %s |" % (h.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", ""), v))
    sass = None
    if rep_path.endswith(".ncu-rep"):
        cmd = ["ncu", "-i", rep_path, "--page", "source", "--csv", "--print-source", "sass"]
        sass = subprocess.run(cmd, capture_output=True, text=True).stdout
        # several kernels in one report: sections are separated by a "Kernel Name" line
        parts = re.split(r'(?m)^"Kernel Name",', sass)[1:]
        per = max(1, len(parts) // max(1, len(rows) - 2))      # ncu prints every kernel's section `per` times
        if len(parts) > row * per:
            sass = parts[row * per]
    else:
        gz = rep_path.replace("_raw.csv", "_sass.csv.gz")
        try:
            import gzip
            sass = gzip.open(gz, "rt").read()
        except OSError:
            sass = None
    if sass:
        out += _opcode_mix(sass, warps)
    open(out_path, "w").write("\n".join(out) + "\n")


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], sys.argv[3], sys.argv[4])
    else:
        rep(sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5] if len(sys.argv) > 5 else "",
            int(sys.argv[6]) if len(sys.argv) > 6 else 0, int(sys.argv[7]) if len(sys.argv) > 7 else 0)
