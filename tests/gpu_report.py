#!/usr/bin/env python
"""Prints, for every golden case and utterance, how far the CUDA path is from the
reference binary's output (run on a GPU box: python tests/gpu_report.py [case ...]).
The summary this prints is what profiles/parity_rNN.txt records."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import ctu_oracle as co  # noqa: E402
import golden_util as gu  # noqa: E402
import ctucopy_b200 as cb  # noqa: E402


def main():
    names = sys.argv[1:] or gu.case_names()
    ins = gu.inputs()
    print("%-26s %3s %11s %9s %9s %9s %8s %s" % ("case", "utt", "shape", "max|err|", "maxrel", "p99.9rel", "nonfin", "notes"))
    for name in names:
        c = gu.Case(name)
        args = c.oracle_args()
        o = co.parse_args(args)
        for i, u in enumerate(ins):
            try:
                res = cb.extract(args, [u], [c.extvad[i]] if c.extvad[i] is not None else None)
            except cb.CtuError as e:
                print("%-26s %3d %s" % (name, i, e.message[:90]))
                continue
            want = c.payload(i)
            if c.kind in ("raw", "wave"):
                got = res.utt_waveform(0)
                if got.shape != want.shape:
                    print("%-26s %3d SHAPE %s %s" % (name, i, got.shape, want.shape)); continue
                d = np.abs(got.astype(np.int32) - want.astype(np.int32))
                print("%-26s %3d %11s %9d %9s %9s %8s frac!=0 %.5f" % (name, i, got.shape, d.max(), "-", "-", "-", (d > 0).mean()))
                continue
            got = res.utt_features(0)
            note = ""
            if c.aux[i] is not None and c.kind != "ark":
                v = np.frombuffer(c.aux[i], dtype=np.uint8) - 48
                gv = res.vad_out[: len(v)]
                note = "vad mismatches %d/%d" % (int((gv != v).sum()) if len(gv) == len(v) else -1, len(v))
            if got.shape != want.shape:
                print("%-26s %3d SHAPE %s %s %s" % (name, i, got.shape, want.shape, note)); continue
            fin = np.isfinite(want) & np.isfinite(got)
            nf = int((np.isfinite(want) != np.isfinite(got)).sum())
            if fin.any():
                err = np.abs(got - want)[fin]
                rel = err / np.maximum(np.abs(want[fin]), 1e-30)
                print("%-26s %3d %11s %9.3g %9.3g %9.3g %8d %s" % (name, i, got.shape, err.max(), rel.max(), np.quantile(rel, 0.999), nf, note))
            else:
                print("%-26s %3d %11s all non-finite %s" % (name, i, got.shape, note))


if __name__ == "__main__":
    main()
