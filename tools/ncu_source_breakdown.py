#!/usr/bin/env python3
"""Per-source-line view of an `ncu --set full --import-source on` capture (kernels built with -lineinfo): where the
warp-stall samples and the executed instructions of the kernel fall, by source line and by region of issl_triple.cuh.

    python tools/ncu_source_breakdown.py <report.ncu-rep> <out.md> [note]
"""
import collections
import csv
import re
import subprocess
import sys


def regions(path):
    """(first line, last line, name) of the functions of issl_triple.cuh, from the source itself."""
    out, cur = [], None
    pat = re.compile(r"^(?:template\s*<[^>]*>\s*)?(?:__device__|__global__|static|inline|__forceinline__|\s)*[\w:<>\*&\s]+?\b(\w+)\s*\(")
    lines = open(path).read().splitlines()
    for i, l in enumerate(lines, 1):
        if l.startswith(("__device__", "__global__", "template")) or (cur is None and l.startswith("static")):
            m = re.search(r"\b(k_\w+|triple_\w+|score_guide|record_\w+|hit_\w+|heavy_\w+|blocked_residual|bs_full_add|order_slice|occ_\w+|site_text_key|text_key_site|general_key)\s*\(", l)
            if not m and i < len(lines):
                m = re.search(r"\b(k_\w+|triple_\w+|score_guide|record_\w+|hit_\w+|heavy_\w+|blocked_residual|bs_full_add|order_slice|occ_\w+)\s*\(", lines[i])
            if m:
                if cur:
                    out.append((cur[0], i - 1, cur[1]))
                cur = (i, m.group(1))
    if cur:
        out.append((cur[0], len(lines), cur[1]))
    return out


def main():
    rep, out = sys.argv[1], sys.argv[2]
    note = sys.argv[3] if len(sys.argv) > 3 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
    cur, kernel, rows = None, "?", []
    for r in csv.reader(raw.splitlines()):
        if len(r) >= 2 and r[0] == "File Path":
            cur = r[1]
        elif len(r) >= 2 and r[0] == "Function Name":
            kernel = r[1]
        elif len(r) >= 8 and r[0].isdigit():
            try:
                rows.append((cur, int(r[0]), r[1], int(r[4]), int(r[7])))
            except ValueError:
                pass
    ts, ti = sum(r[3] for r in rows) or 1, sum(r[4] for r in rows) or 1
    reg = {}
    for f in {r[0] for r in rows}:
        if f and f.endswith("issl_triple.cuh"):
            try:
                reg[f] = regions(f)
            except OSError:
                reg[f] = []

    def region(f, l):
        for a, b, name in reg.get(f, []):
            if a <= l <= b:
                return name
        return (f or "?").split("/")[-1]
    agg = collections.defaultdict(lambda: [0, 0])
    for f, l, _, s, i in rows:
        a = agg[region(f, l)]
        a[0] += s
        a[1] += i
    with open(out, "w") as o:
        o.write(f"# ncu source page, {kernel}\n\n{note}\n\n{ts} warp-stall samples, {ti} warp-level instructions executed.\n\n"
                "| function | stall samples | % | instructions executed | % |\n|---|---:|---:|---:|---:|\n")
        for k, (s, i) in sorted(agg.items(), key=lambda x: -x[1][1]):
            if 100 * i / ti >= 0.1 or 100 * s / ts >= 0.1:
                o.write(f"| `{k}` | {s} | {100 * s / ts:.1f} | {i} | {100 * i / ti:.1f} |\n")
        o.write("\nTop source lines by stall samples:\n\n| file:line | samples % | instructions % | source |\n|---|---:|---:|---|\n")
        for f, l, src, s, i in sorted(rows, key=lambda x: -x[3])[:30]:
            o.write(f"| {(f or '?').split('/')[-1]}:{l} | {100 * s / ts:.1f} | {100 * i / ti:.1f} | `{src.strip()[:110].replace('|', '¦')}` |\n")
        o.write("\nTop source lines by instructions executed:\n\n| file:line | instructions % | samples % | source |\n|---|---:|---:|---|\n")
        for f, l, src, s, i in sorted(rows, key=lambda x: -x[4])[:30]:
            o.write(f"| {(f or '?').split('/')[-1]}:{l} | {100 * i / ti:.1f} | {100 * s / ts:.1f} | `{src.strip()[:110].replace('|', '¦')}` |\n")


if __name__ == "__main__":
    main()
