#!/usr/bin/env python
"""Text summary of an `ncu --set full --import-source on` report for profiles/ (tracked):
key counters per captured kernel, the executed SASS opcode mix (warp instructions) and the ten
instructions with the most stall samples.

  python profiles/ncu_text.py gpurun_out/<name>.ncu-rep profiles/<tag>.txt ["header line"]
"""
import collections
import csv
import io
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from summarize import KEYS  # noqa: E402

EXTRA = ["sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.avg.pct_of_peak_sustained_active",
         "sm__issue_active.avg.pct_of_peak_sustained_elapsed"]


def page(rep, name, extra=()):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv", *extra], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep, dst = sys.argv[1], sys.argv[2]
    lines = [sys.argv[3] if len(sys.argv) > 3 else f"# {os.path.basename(rep)}", ""]
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    for r in raw[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        lines.append(f"== {d['Kernel Name']}  grid {d.get('launch__grid_size', '?')}")
        for k in KEYS + EXTRA:
            if k in d and d[k] != "":
                lines.append(f"{k:85s} {u.get(k, ''):16s} {d[k]}")
        lines.append("")
    src = page(rep, "source", ["--print-source", "sass"])
    cur, h, seen = None, None, set()
    per = collections.OrderedDict()
    for r in src:
        if r and r[0] == "Kernel Name":
            cur = r[1]
            cur = None if cur in seen else cur  # first instance of every kernel
            if cur:
                seen.add(cur)
                per[cur] = []
            h = None
            continue
        if r and r[0] == "Address":
            h = r
            continue
        if cur and h and len(r) == len(h):
            per[cur].append(dict(zip(h, r)))
    for k, rows in per.items():
        mix, tot, samples = collections.Counter(), 0, 0
        for d in rows:
            op = d["Source"].split()
            if not op:
                continue
            o = (op[1] if op[0].startswith("@") else op[0]).split(".")[0]
            n = int(d["Instructions Executed"])
            mix[o] += n
            tot += n
            samples += int(d["# Samples"])
        lines.append(f"== executed SASS mix (warp instructions), {k}: total {tot}")
        lines.append("   " + "  ".join(f"{o} {n} ({n / max(tot, 1):.3f})" for o, n in mix.most_common(24)))
        lines.append(f"== most-sampled instructions (of {samples} stall samples)")
        for d in sorted(rows, key=lambda d: -int(d["# Samples"]))[:10]:
            st = {x[6:]: int(d[x]) for x in d if x.startswith("stall_") and "Not Issued" not in x and d[x] not in ("", "0")}
            top = ", ".join(f"{a} {b}" for a, b in sorted(st.items(), key=lambda kv: -kv[1])[:3])
            lines.append(f"   {int(d['# Samples']):6d}  {d['Source'].strip()[:60]:60s} {top}")
        lines.append("")
    open(dst, "w").write("\n".join(lines) + "\n")
    print("wrote", dst, len(lines), "lines")


if __name__ == "__main__":
    main()
