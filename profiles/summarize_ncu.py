#!/usr/bin/env python
"""Turns an .ncu-rep, or the `--page raw --csv` export of one (tools/gpu_jobs/ncu_cap.sh writes
those on the GPU box), into the short per-kernel text summary committed under profiles/.
usage: python profiles/summarize_ncu.py <rep | raw.csv> [frames_per_launch]"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("dram__bytes_read.sum.per_second", "dram read rate"),
    ("dram__bytes_write.sum.per_second", "dram write rate"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram % of ncu peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__shared_mem_per_block_dynamic", "dyn smem/block"),
    ("launch__occupancy_limit_registers", "blocks/SM (regs)"),
    ("launch__occupancy_limit_shared_mem", "blocks/SM (smem)"),
    ("launch__grid_size", "grid"),
    ("sm__inst_executed.sum", "warp instructions"),
    ("l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed", "LSU data-pipe wavefronts % of peak"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem wavefronts"),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smem bank conflicts"),
    ("sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "FMA pipe %"),
    ("sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active", "FP64 pipe %"),
    ("sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active", "LSU pipe %"),
]
STALLS = "smsp__pcsamp_warps_issue_stalled_"


def main():
    rep = sys.argv[1]
    frames = float(sys.argv[2]) if len(sys.argv) > 2 else None
    if rep.endswith(".csv"):
        out = open(rep).read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("kernel:", r[ix["Kernel Name"]][:110])
        vals = {}
        for key, label in WANT:
            if key in ix:
                print("  %-28s %s %s" % (label, r[ix[key]], units[ix[key]]))
                vals[key] = r[ix[key]]
        st = []
        for h, i in ix.items():
            if h.startswith(STALLS) and "not_issued" not in h:
                try:
                    st.append((float(r[i].replace(",", "")), h[len(STALLS):]))
                except ValueError:
                    pass
        tot = sum(v for v, _ in st) or 1
        print("  stall reasons: " + ", ".join("%s %.0f%%" % (n, 100 * v / tot) for v, n in sorted(st, reverse=True)[:6]))
        if frames:
            try:
                inst = float(vals["sm__inst_executed.sum"].replace(",", ""))
                wf = float(vals["l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"].replace(",", ""))
                print("  per frame: %.0f warp instructions, %.0f smem wavefronts" % (inst / frames, wf / frames))
            except Exception:
                pass
        print()


if __name__ == "__main__":
    main()
