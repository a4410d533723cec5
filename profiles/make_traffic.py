#!/usr/bin/env python
"""profiles/traffic.json from the `ncu --set full` raw exports: measured DRAM bytes (read + write)
per frame for every kernel of every bench workload.  bench.py multiplies by the frames of
its launch to fill roofline.traffic.   usage: python profiles/make_traffic.py gpurun_out"""
import csv
import json
import os
import re
import sys

FRAMES = {"mfcc_exten": 1996000, "mfcc_d_a": 1996000, "plp": 1996000, "trapdct": 1996000, "exten": 1248000, "fwss_burg": 499000, "tdiir": 1996000}
SRC = {0: "pcm", 1: "spec", 2: "fb"}
DST = {0: "spec", 1: "fb", 2: "fea"}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def label(kernel):
    if "k_frames2<" in kernel:
        return "k_frames<pcm,spec>"
    m = re.match(r"void k_frames<(\d+), (\d+)", kernel)
    if m:
        return "k_frames<%s,%s>" % (SRC[int(m.group(1))], DST[int(m.group(2))])
    m = re.match(r"void k_bank<(\d+), (\d+), (\d+)>", kernel)          # <KIND, DST, NR>: the names ctu_bank.cuh gives its launches
    if m:
        return "k_bank<%s%s>" % ("nr," if int(m.group(3)) else "", DST[int(m.group(2))])
    if "k_nr_scan4<" in kernel:
        return "k_nr_scan"
    m = re.match(r"void (k_[a-z_]+?)(22|64)?<|void (k_[a-z_]+)\(|(k_[a-z_]+)\(", kernel)
    name = next(g for g in (m.group(1), m.group(3), m.group(4)) if g) if m else kernel
    return name


def main():
    d = sys.argv[1]
    tp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
    out = json.load(open(tp)) if os.path.exists(tp) else {}      # workloads without a new capture keep their entries
    for w, frames in FRAMES.items():
        p = os.path.join(d, "raw_p_%s.csv" % w)
        if not os.path.exists(p):
            continue
        rows = list(csv.reader(open(p)))
        hdr, units = rows[0], rows[1]
        ix = {h: i for i, h in enumerate(hdr)}
        for r in rows[2:]:
            rd = float(r[ix["dram__bytes_read.sum"]].replace(",", "")) * UNIT[units[ix["dram__bytes_read.sum"]]]
            wr = float(r[ix["dram__bytes_write.sum"]].replace(",", "")) * UNIT[units[ix["dram__bytes_write.sum"]]]
            out["%s:%s" % (w, label(r[ix["Kernel Name"]]))] = {
                "dram_bytes_per_frame": (rd + wr) / frames, "read": rd / frames, "write": wr / frames,
                "ncu_duration_ms": float(r[ix["gpu__time_duration.sum"]].replace(",", "")), "frames_in_capture": frames}
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json"), "w"), indent=1, sort_keys=True)
    for k, v in sorted(out.items()):
        print("%-40s %8.1f B/frame (r %.1f w %.1f)" % (k, v["dram_bytes_per_frame"], v["read"], v["write"]))


if __name__ == "__main__":
    main()
