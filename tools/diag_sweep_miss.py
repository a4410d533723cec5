#!/usr/bin/env python
"""Where the CUDA path and the oracle differ most for given option sets (the open classes of tools/parity_sweep.py).
usage: python tools/diag_sweep_miss.py   (GPU box)"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import ctu_oracle as co  # noqa: E402
import golden_util as gu  # noqa: E402
import ctucopy_b200 as cb  # noqa: E402
import importlib.util  # noqa: E402

spec = importlib.util.spec_from_file_location("parity_sweep", os.path.join(ROOT, "tools", "parity_sweep.py"))
ps = importlib.util.module_from_spec(spec); spec.loader.exec_module(ps)

H = ["-format_in", "raw", "-dither", "0", "-format_out", "htk"]
SETS = [
    "-fs 8000 -preem 0 -w 32 -s 10 -remove_dc off -fb_scale bark -fb_shape rect -fb_definition 30filters -fb_norm off -fb_eqld on -fb_inld on -fb_power off -fea_kind spec -fea_ncepcoefs 8 -fea_delta d -d_win 2 -a_win 1 -t_win 2 -nr_mode fwss -nr_p 0.9 -nr_a 1 -nr_b 1 -nr_initsegs 5 -vad burg",
    "-fs 8000 -preem 0.97 -w 25 -s 8 -remove_dc on -fb_scale expolog -fb_shape triang -fb_definition 40filters -fb_norm on -fb_eqld off -fb_inld on -fb_power off -fea_kind trapdct,11,3 -nr_mode exten -nr_p 0.9 -nr_a 2 -nr_b 1 -nr_initsegs 10 -nr_when afterFB",
]
for sset in SETS:
    tok = sset.split()
    args = tok[:2] + H + tok[2:]
    o = co.parse_args(args)
    ins = [gu.inputs()[i] for i in (0, 5)]
    res = cb.extract(args, ins)
    print("==", sset)
    for i, u in enumerate(ins):
        ref = co.run_pipeline(u, o)
        got, want = res.utt_features(i), ref.features
        pre = ps.sensitivity(args, o, u, ref)
        ok, why = ps.tol_ok(got, want, o.fea_kind, pre)
        print(" input", i, "ok" if ok else "MISS", why, "shape", want.shape)
        err = np.abs(got.astype(np.float64) - want)
        for k in np.argsort(err.ravel())[::-1][:6]:
            t, c = divmod(int(k), want.shape[1])
            lo, hi = max(0, t - 3), min(want.shape[0], t + 4)
            print("   t=%d col=%d got %.6g want %.6g err %.3g | want[t-3..t+3, col] = %s" % (t, c, got[t, c], want[t, c], err[t, c], np.array2string(want[lo:hi, c], precision=4)))
            if ref.internal is not None and o.fea_kind == "spec":
                cc = c % (o.fea_ncepcoefs + 1)
                print("      static col %d: want %s got %s" % (cc, np.array2string(want[lo:hi, cc], precision=6), np.array2string(got[lo:hi, cc], precision=6)))
        if o.fea_kind == "trapdct" and ref.fb_out is not None:
            Y = ref.fb_out
            print("   band values after NR: min %.3g, max %.3g; log range %.3g .. %.3g" % (Y.min(), Y.max(), np.log(max(Y.min(), 1e-300)), np.log(Y.max())))
