#!/usr/bin/env python3
"""Per-source-line summary of an ncu report captured with --import-source on (kernels built with -lineinfo).

usage: ncu_lines.py report.ncu-rep [top_n] [kernel regex] [inst: sort by instructions instead of stall samples]
Prints, for every file the kernel's SASS maps to, the lines with the most warp-stall samples together with the
warp instructions executed there and the mean number of active threads per instruction (lane efficiency).
"""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    kern = ["-k", "regex:" + sys.argv[3]] if len(sys.argv) > 3 else []
    txt = subprocess.run(["ncu", "-i", rep, *kern, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    cur_file, header = None, None
    agg = defaultdict(lambda: [0, 0, 0, ""])     # (file, line) -> samples, inst, thread_inst, source
    stalls = defaultdict(lambda: defaultdict(int))
    tot_s = tot_i = tot_t = 0
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            cur_file = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            header = r
            continue
        if header is None or r[0] == "":
            continue
        d = dict(zip(header, r))
        try:
            line = int(r[0])
            s = int(d["# Samples"]); i = int(d["Instructions Executed"]); t = int(d["Thread Instructions Executed"])
        except ValueError:
            continue
        a = agg[(cur_file, line)]
        a[0] += s; a[1] += i; a[2] += t; a[3] = r[1].strip()
        tot_s += s; tot_i += i; tot_t += t
        for k, v in d.items():
            if k.startswith("stall_") and "Not Issued" not in k:
                try:
                    stalls[(cur_file, line)][k] += int(v)
                except ValueError:
                    pass
    print(f"total samples {tot_s}  warp instructions {tot_i}  lane efficiency {tot_t / max(tot_i, 1):.2f}/32")
    print(f"{'file:line':28s} {'samp%':>6s} {'inst%':>6s} {'lanes':>5s}  top stalls / source")
    for (f, l), a in sorted(agg.items(), key=lambda kv: -kv[1][1 if len(sys.argv) > 4 and sys.argv[4] == 'inst' else 0])[:top]:
        st = sorted(stalls[(f, l)].items(), key=lambda kv: -kv[1])[:3]
        sts = " ".join(f"{k[6:]}={v}" for k, v in st if v)
        print(f"{f + ':' + str(l):28s} {100 * a[0] / max(tot_s, 1):6.2f} {100 * a[1] / max(tot_i, 1):6.2f} "
              f"{a[2] / max(a[1], 1):5.1f}  [{sts}] {a[3][:90]}")


if __name__ == "__main__":
    main()
