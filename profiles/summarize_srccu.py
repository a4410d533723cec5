#!/usr/bin/env python
"""Per-CUDA-source-line totals from `ncu --page source --print-source cuda,sass --csv`:
samples, executed warp instructions and shared-memory wavefronts per source line.
usage: python profiles/summarize_srccu.py <srccu.csv> [top]"""
import csv
import sys
from collections import defaultdict


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return 0.0


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    fpath = func = None
    ix = None
    agg = defaultdict(lambda: [0.0, 0.0, 0.0, 0.0, ""])   # samples, inst, wavefronts, ideal
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fpath = r[1].split("/")[-1]
        elif r[0] == "Function Name":
            func = r[1][:60]
        elif r[0] == "Line No":
            ix = {}
            for k, h in enumerate(r):
                ix.setdefault(h, k)
        elif ix and r[0] not in ("", "-") and r[0].isdigit():
            key = (func, fpath, int(r[0]))
            a = agg[key]
            a[0] += num(r[ix["# Samples"]]); a[1] += num(r[ix["Instructions Executed"]])
            a[2] += num(r[ix["L1 Wavefronts Shared"]]); a[3] += num(r[ix["L1 Wavefronts Shared Ideal"]])
            a[4] = r[1].strip()[:90]
    by_func = defaultdict(list)
    for (fn, fp, ln), a in agg.items():
        by_func[fn].append((fp, ln, a))
    for fn, items in by_func.items():
        ts = sum(a[0] for _, _, a in items) or 1
        ti = sum(a[1] for _, _, a in items) or 1
        print("function:", fn, " samples %d  warp-inst %.3g" % (ts, ti))
        for fp, ln, a in sorted(items, key=lambda x: -x[2][0])[:top]:
            print("  %5.1f%% smp %5.1f%% inst  wf %.3g/%.3g  %s:%d  %s" % (100 * a[0] / ts, 100 * a[1] / ti, a[2], a[3], fp, ln, a[4]))
        print()


if __name__ == "__main__":
    main()
