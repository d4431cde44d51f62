"""Attribute an ncu report's per-instruction samples to source lines (development tool).

    python tools/ncu_lines.py REPORT.ncu-rep KERNEL_SUBSTR [--lib path.so] [--by inner|outer|func]
                              [--top N] [--launch K]

`ncu --page source --csv` lists SASS instructions with sample / execution counts but without
line numbers; `nvdisasm -gi` of the cubin inside the shared library lists the same instructions
with their (inlined) source positions.  The two are aligned by instruction order and opcode.

--by inner : innermost source line (default)        --by outer : line inside the kernel body
--by func  : innermost function-level position, i.e. file:line of the call site one level up
"""
import argparse
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_with_lines(lib, kernel_substr):
    tmp = tempfile.mkdtemp(prefix="ncul_")
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True,
                   stdout=subprocess.DEVNULL)
    out = []
    for cub in sorted(os.listdir(tmp)):
        txt = subprocess.run(["nvdisasm", "-gi", "-c", os.path.join(tmp, cub)], capture_output=True,
                             text=True).stdout.splitlines()
        inside, chain, pending = False, [], []
        for line in txt:
            m = re.match(r"\s*\.text\.(\S+):", line)
            if m:
                inside = kernel_substr in m.group(1)
                if inside and out:
                    raise SystemExit("kernel substring matches more than one function")
                continue
            if not inside:
                continue
            if line.startswith(".section") or re.match(r"\s*\.section", line):
                inside = False
                continue
            m = re.match(r'\s*//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', line)
            if m:
                pending.append((os.path.basename(m.group(1)), int(m.group(2))))
                continue
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
            if m:
                if pending:
                    chain, pending = pending, []
                out.append((int(m.group(1), 16), m.group(2).strip(), list(chain)))
        if out:
            break
    return out


def ncu_sass(report, kernel_substr, launch):
    txt = subprocess.run(["ncu", "-i", report, "--page", "source", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            kernels.append(cur)
        elif cur is not None and r and r[0] == "Address":
            cur["hdr"] = r
        elif cur is not None and cur["hdr"] and len(r) >= len(cur["hdr"]) - 2:
            cur["rows"].append(r)
    base = kernel_substr
    sel = [k for k in kernels if k["hdr"]]
    if not sel:
        raise SystemExit("no source page in report")
    k = sel[min(launch, len(sel) - 1)]
    h = k["hdr"]
    idx = {n: h.index(n) for n in ("Source", "# Samples", "Instructions Executed",
                                   "Thread Instructions Executed")}
    out = []
    for r in k["rows"]:
        out.append((r[idx["Source"]].strip(), int(r[idx["# Samples"]] or 0),
                    int(r[idx["Instructions Executed"]] or 0),
                    int(r[idx["Thread Instructions Executed"]] or 0)))
    return k["name"], out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("report")
    ap.add_argument("kernel")
    ap.add_argument("--lib", default=os.path.join(ROOT, "disimpy_b200", "libdisimpy_b200.so"))
    ap.add_argument("--by", default="inner", choices=["inner", "outer", "func"])
    ap.add_argument("--top", type=int, default=40)
    ap.add_argument("--launch", type=int, default=0)
    ap.add_argument("--level", type=int, default=-1,
                    help="attribute to the position L inlining levels below the kernel body")
    ap.add_argument("--units", type=float, default=0.0,
                    help="warp-steps in the launch: prints warp-instructions per warp-step")
    a = ap.parse_args()
    sass = sass_with_lines(a.lib, a.kernel)
    name, prof = ncu_sass(a.report, a.kernel, a.launch)
    if len(sass) != len(prof):
        print("WARNING: %d instructions in the library, %d in the report -- different builds?"
              % (len(sass), len(prof)), file=sys.stderr)
    n = min(len(sass), len(prof))
    mism = sum(1 for i in range(n) if sass[i][1].split()[0].split(".")[0] not in prof[i][0])
    if mism:
        print("WARNING: %d opcode mismatches while aligning" % mism, file=sys.stderr)
    agg = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for i in range(n):
        chain = sass[i][2]
        if not chain:
            key = "?"
        elif a.level >= 0:
            key = "%s:%d" % chain[max(0, len(chain) - 1 - a.level)]
        elif a.by == "inner":
            key = "%s:%d" % chain[0]
        elif a.by == "outer":
            key = "%s:%d" % chain[-1]
        else:
            key = "%s:%d" % (chain[1] if len(chain) > 1 else chain[0])
        for j in range(3):
            agg[key][j] += prof[i][1 + j]
            tot[j] += prof[i][1 + j]
    print("kernel: %s\nsamples %d, warp-instructions %d, avg active lanes %.1f"
          % (name, tot[0], tot[1], tot[2] / max(tot[1], 1)))
    if a.units:
        print("warp-instructions per warp-step: %.1f" % (tot[1] / a.units))
    src_cache = {}

    def text(key):
        f, _, l = key.partition(":")
        for d in ("disimpy_b200/csrc",):
            p = os.path.join(ROOT, d, f)
            if os.path.exists(p):
                if p not in src_cache:
                    src_cache[p] = open(p).read().splitlines()
                try:
                    return src_cache[p][int(l) - 1].strip()[:90]
                except Exception:
                    return ""
        return ""
    print("%7s %7s %6s  %-24s %s" % ("samp%", "instr%", "lanes", "where", "source"))
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:a.top]:
        extra = "  %.1f/ws" % (v[1] / a.units) if a.units else ""
        print("%6.2f%% %6.2f%% %6.1f  %-24s %s%s" % (100.0 * v[0] / max(tot[0], 1), 100.0 * v[1] / max(tot[1], 1),
                                                   v[2] / max(v[1], 1), key, text(key), extra))


if __name__ == "__main__":
    main()
