#!/usr/bin/env python
"""Per-instruction stall samples from `ncu --page source --csv` (SASS view): totals by opcode
and the hottest instructions.   usage: python profiles/summarize_src.py <src.csv> [top]"""
import csv
import sys
from collections import defaultdict


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    # several kernels may be concatenated: split at "Kernel Name" rows
    i = 0
    while i < len(rows):
        if rows[i] and rows[i][0] == "Kernel Name":
            name = rows[i][1][:100]
            hdr = rows[i + 1]
            ix = {h: k for k, h in enumerate(hdr)}
            j = i + 2
            body = []
            while j < len(rows) and not (rows[j] and rows[j][0] == "Kernel Name"):
                if len(rows[j]) == len(hdr):
                    body.append(rows[j])
                j += 1
            report(name, ix, body, top)
            i = j
        else:
            i += 1


def num(s):
    try:
        return float(s.replace(",", ""))
    except ValueError:
        return 0.0


def report(name, ix, body, top):
    print("kernel:", name)
    S = ix["# Samples"]
    tot = sum(num(r[S]) for r in body) or 1
    inst = sum(num(r[ix["Instructions Executed"]]) for r in body)
    print("  SASS instructions: %d, warp-instructions executed: %.3g, samples: %d" % (len(body), inst, tot))
    by_op = defaultdict(lambda: [0.0, 0.0])
    for r in body:
        op = r[ix["Source"]].split()
        op = [t for t in op if not t.startswith("@")]
        o = op[0].split(".")[0] if op else "?"
        by_op[o][0] += num(r[S]); by_op[o][1] += num(r[ix["Instructions Executed"]])
    print("  by opcode (samples %, executed %):")
    for o, (s, e) in sorted(by_op.items(), key=lambda kv: -kv[1][0])[:14]:
        print("    %-10s %5.1f%%  %5.1f%%" % (o, 100 * s / tot, 100 * e / max(inst, 1)))
    stalls = [h for h in ix if h.startswith("stall_") and "Not Issued" not in h]
    tots = {h: sum(num(r[ix[h]]) for r in body) for h in stalls}
    st = sum(tots.values()) or 1
    print("  stall mix: " + ", ".join("%s %.0f%%" % (h[6:], 100 * v / st) for h, v in sorted(tots.items(), key=lambda kv: -kv[1])[:7]))
    print("  hottest instructions:")
    order = sorted(range(len(body)), key=lambda k: -num(body[k][S]))[:top]
    for k in sorted(order):
        r = body[k]
        why = sorted(((num(r[ix[h]]), h[6:]) for h in stalls), reverse=True)[:2]
        print("    #%-5d %5.2f%%  %-70s %s" % (k, 100 * num(r[S]) / tot, r[ix["Source"]][:70], " ".join("%s:%d" % (n, v) for v, n in why if v)))
    print()


if __name__ == "__main__":
    main()
