#!/usr/bin/env python3
"""Group the per-line numbers of an ncu report (--import-source on, -lineinfo) by code region of the extend kernel.
usage: ncu_regions.py report.ncu-rep"""
import collections
import csv
import re
import subprocess
import sys

SRC = "/root/repo/another_raytracer_b200/csrc/device/rt_intersect.cuh"


def function_ranges(path):
    """(first line, last line, name) of every function in the header, found from 'ARTB_HD... name(' lines"""
    lines = open(path).read().splitlines()
    starts = []
    for i, l in enumerate(lines, 1):
        m = re.match(r"^(?:template.*|ARTB_HD\w*|inline|static)?\s*(?:ARTB_HD\w*\s+)?[\w:<>]+[\s*&]+(\w+)\(", l)
        if m and not l.startswith(" ") and "(" in l:
            starts.append((i, m.group(1)))
    out = []
    for k, (ln, name) in enumerate(starts):
        end = starts[k + 1][0] - 1 if k + 1 < len(starts) else len(lines)
        out.append((ln, end, name))
    return out


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True, check=True).stdout
    ranges = function_ranges(SRC)
    cur, hdr = None, None
    reg = collections.defaultdict(lambda: [0, 0, 0])
    tot = [0, 0, 0]
    for r in csv.reader(txt.splitlines()):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if r[0] in ("Function Name", ""):
            continue
        d = dict(zip(hdr, r))
        try:
            line = int(r[0]); s = int(d["# Samples"]); i = int(d["Instructions Executed"]); t = int(d["Thread Instructions Executed"])
        except ValueError:
            continue
        name = cur
        if cur == "rt_intersect.cuh":
            name = next((n for a, b, n in ranges if a <= line <= b), "rt_intersect.cuh")
        v = reg[name]
        for k, x in enumerate((s, i, t)):
            v[k] += x; tot[k] += x
    print(f"total: samples {tot[0]}  warp instructions {tot[1]}  lane efficiency {tot[2] / max(tot[1], 1):.2f}/32")
    for k, v in sorted(reg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:32s} samples {100 * v[0] / tot[0]:5.1f}%  warp-inst {100 * v[1] / tot[1]:5.1f}%  lanes {v[2] / max(v[1], 1):5.1f}")


if __name__ == "__main__":
    main()
