"""Turn the ncu artefacts a gpurun call brought back (gpurun_out/) into the tracked summaries
under profiles/ (development tool).

    python tools/summarize_profiles.py TAG launches.csv [name=report.ncu-rep:kernel_substr:warp_steps ...]
"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__grid_size", "grid"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active, % of 64 / SM"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "active lanes / instruction"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots used, %"),
    ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "FP64 pipe, % of peak"),
    ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "ALU pipe, %"),
    ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "FMA pipe, %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "XU pipe, %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "LSU pipe, %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM written"),
    ("lts__t_sectors.sum", "L2 sectors (32 B)"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate, %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate, %"),
    ("smsp__warps_eligible.avg.per_cycle_active", "eligible warps / cycle / SMSP"),
]
STALLS = ["long_scoreboard", "wait", "not_selected", "math_pipe_throttle", "no_instruction",
          "short_scoreboard", "branch_resolving", "dispatch_stall", "barrier", "mio_throttle", "lg_throttle"]


def raw_metrics(report):
    txt = subprocess.run(["ncu", "-i", report, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    return {h: (v, u) for h, u, v in zip(hdr, units, vals)}


def launches_md(tag, path):
    rows = list(csv.DictReader(l for l in open(path) if l.startswith('"')))
    agg = collections.OrderedDict()
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = r["Kernel Name"].split("(")[0]
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(r["Metric Value"].replace(",", "")) / 1e6
    tot = sum(v[1] for v in agg.values())
    out = ["# %s: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-secondary --no-mesh --no-reference-baselines` (1 B200)" % tag,
           "# per-launch times are cold-cache and serialised under the profiler: compare SHARES", "",
           "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.3f | %.2f%% |" % (k, n, ms, 100 * ms / tot))
    return "\n".join(out) + "\n"


def report_md(tag, name, report, kernel, warp_steps, command):
    m = raw_metrics(report)
    out = ["# %s -- `ncu --set full` of %s" % (tag, name), "", "Command (1 B200, under gpurun): `%s`" % command, "",
           "| metric | value |", "|---|---|"]
    for key, label in KEYS:
        if key in m:
            out.append("| %s | %s %s |" % (label, m[key][0], m[key][1]))
    if warp_steps:
        inst = float(m["smsp__inst_executed.sum"][0].replace(",", ""))
        out.append("| warp instructions per warp-step | %.0f |" % (inst / warp_steps))
    out += ["", "Stall reasons (warps per issue-active cycle):", "", "| reason | ratio |", "|---|---|"]
    for s in STALLS:
        k = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % s
        if k in m:
            out.append("| %s | %s |" % (s, m[k][0]))
    lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), report, kernel, "--level", "2",
                            "--top", "22"] + (["--units", str(warp_steps)] if warp_steps else []),
                           capture_output=True, text=True).stdout
    out += ["", "Samples and executed warp instructions by source position (two inlining levels below the kernel body;",
            "`tools/ncu_lines.py`, ncu source page joined with `nvdisasm -gi` line info; /ws = per warp-step):", "", "```"]
    out += [l[:170] for l in lines.splitlines()]
    out += ["```", ""]
    return "\n".join(out)


def main():
    tag, launches = sys.argv[1], sys.argv[2]
    prof = os.path.join(ROOT, "profiles")
    if launches != "-":
        open(os.path.join(prof, tag + "_launches.md"), "w").write(launches_md(tag, launches))
        open(os.path.join(prof, tag + "_launches.csv"), "w").write(open(launches).read())
    for spec in sys.argv[3:]:
        name, rest = spec.split("=", 1)
        report, kernel, ws, command = rest.split(":", 3)
        md = report_md(tag, name, report, kernel, float(ws), command)
        open(os.path.join(prof, "%s_%s_ncu.md" % (tag, name)), "w").write(md)
        print(md[:1500])


if __name__ == "__main__":
    main()
