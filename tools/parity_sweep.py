#!/usr/bin/env python
"""Randomised option sweep.  Draws valid command lines from the hot-path option space and compares
   cpu : the numpy oracle against the REFERENCE BINARY (oracle/_ref, needs /root/reference built here)
   gpu : the CUDA path against the oracle
on two short inputs.  usage: python tools/parity_sweep.py cpu|gpu [n] [seed] [share of waveform-output draws] [share of draws at other sampling rates]"""
import os
import random
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import ctu_oracle as co  # noqa: E402
import golden_util as gu  # noqa: E402

B = ["-fs", "16000", "-format_in", "raw", "-dither", "0", "-format_out", "htk"]
SIGNAL_SHARE = 0.0      # share of enhanced-waveform draws (set from the command line: 4th argument)


def draw_signal(rng):
    """enhanced-waveform command lines (src/io/out.cc:346-451): raw output, NR mode x window/shift"""
    a = ["-fs", "16000", "-format_in", "raw", "-dither", "0", "-format_out", "raw"]
    w, s_ = rng.choice([("32", "16"), ("32", "16"), ("25", "10"), ("32", "8"), ("20", "10"), ("30", "10"), ("25", "12.5")])
    a += ["-w", w, "-s", s_, "-preem", rng.choice(["0", "0.97"]), "-remove_dc", rng.choice(["on", "off"])]
    nr = rng.choice(["exten", "exten", "fwss", "2fwss", "none"])
    a += ["-nr_mode", nr, "-nr_p", rng.choice(["0.95", "0.9"]), "-nr_a", rng.choice(["1", "2"]), "-nr_b", rng.choice(["1", "1.5"])]
    if nr in ("fwss", "2fwss"):
        a += ["-vad", "burg"]
    return a


def draw(rng):
    a = draw16(rng)
    # other sampling rates: the same samples read at another rate -> 256- to 2048-point frames (general kernels)
    if rng.random() < OTHER_FS_SHARE:
        a[1] = rng.choice(["8000", "8000", "11025", "22050", "44100"])
    return a


OTHER_FS_SHARE = 0.0    # 5th command-line argument


def draw16(rng):
    if rng.random() < SIGNAL_SHARE:
        return draw_signal(rng)
    a = list(B)
    kind = rng.choice(["dctc", "dctc", "lpc", "lpc", "spec", "logspec", "lpa", "trapdct"])
    scale = rng.choice(["mel", "bark", "lin", "expolog"])
    shape = rng.choice(["triang", "triang", "rect", "trapez"])
    a += ["-preem", rng.choice(["0", "0.95", "0.97"]), "-w", rng.choice(["25", "25", "20", "30", "32"]), "-s", rng.choice(["10", "10", "8", "16"])]
    a += ["-remove_dc", rng.choice(["on", "on", "off"])]
    if shape == "trapez":
        a += ["-fb_shape", "trapez"]
    else:
        nb = rng.choice([15, 20, 23, 26, 30, 40])
        a += ["-fb_scale", scale, "-fb_shape", shape, "-fb_definition", rng.choice(["%dfilters" % nb, "1-%d/%dfilters" % (nb - 2, nb)])]
        a += ["-fb_norm", rng.choice(["on", "off"]), "-fb_eqld", rng.choice(["on", "off"]), "-fb_inld", rng.choice(["on", "off"])]
    a += ["-fb_power", rng.choice(["on", "on", "off"])]
    ncep = rng.choice([8, 12, 12, 13, 16])
    if kind == "trapdct":
        a += ["-fea_kind", "trapdct,%d,%d" % (rng.choice([11, 21, 31]), rng.choice([3, 6]))]
    else:
        a += ["-fea_kind", kind]
    if kind in ("dctc", "lpc", "lpa"):
        order = ncep if kind == "lpa" else rng.choice([ncep, ncep, 10, 14])
        a += ["-fea_ncepcoefs", str(ncep), "-fea_lporder", str(order), "-fea_lifter", rng.choice(["22", "0", "1", "30"]), "-fea_c0", "on"]
    if kind in ("spec", "logspec") and rng.random() < 0.5:
        # the delta chain / context stacking take the first ncep+1 bands
        a += ["-fea_ncepcoefs", str(rng.choice([8, 12, 14]))]
        kind_chain = True
    else:
        kind_chain = kind in ("dctc", "lpc")
    if kind_chain and rng.random() < 0.6:
        if rng.random() < 0.35:
            a += ["-fea_trap", rng.choice(["3", "5", "7", "11"])]
        else:
            a += ["-fea_delta", rng.choice(["d", "d_a", "d_a_t"]), "-d_win", rng.choice(["2", "3"]), "-a_win", rng.choice(["2", "1"]), "-t_win", "2"]
    nr = rng.choice(["none", "none", "exten", "exten", "fwss", "hwss", "2fwss"])
    if nr != "none":
        a += ["-nr_mode", nr, "-nr_p", rng.choice(["0.95", "0.9"]), "-nr_a", rng.choice(["1", "2"]), "-nr_b", rng.choice(["1", "1.5"]),
              "-nr_initsegs", rng.choice(["10", "5"])]
        if nr != "exten":
            a += ["-vad", "burg"]
        elif rng.random() < 0.4:
            a += ["-nr_when", "afterFB"]
    if kind != "trapdct" and rng.random() < 0.25:
        a += ["-fea_E", "on"] + (["-fea_rawenergy", "on"] if rng.random() < 0.3 else [])
    if (kind in ("dctc", "lpc") or "-fea_trap" in a or "-fea_delta" in a) and rng.random() < 0.2 and "-fea_E" not in a:
        a += ["-fea_Z_exp", rng.choice(["300", "1000"])]
    # VAD module (src/vad/vad.cc): criterion x threshold x majority filter x apply mode
    if rng.random() < 0.35 and "-fea_Z_exp" not in a and kind != "trapdct":
        cri = rng.choice(["energy", "energy", "cepdist"])
        a += ["-vad_out_mode", "vad", "-vad_cri_mode", cri, "-vad_thr_mode", rng.choice(["absolute", "perc", "adapt", "dyn"]),
              "-vad_filter_order", rng.choice(["1", "3", "5"]), "-vad_apply_mode", rng.choice(["none", "none", "drop", "silence"])]
        if cri == "cepdist":
            mode = rng.choice(["lpc", "fea"]) if kind in ("dctc", "lpc") else "lpc"
            a += ["-vad_cepdist_mode", mode]
            if mode == "lpc" and "-vad" not in a:
                a += ["-vad", "burg"]
        else:
            a += ["-vad_energy_db", rng.choice(["on", "off"])]
        a += ["-vad_absolute_thr", rng.choice(["60", "100"]), "-vad_perc_thr", rng.choice(["50", "30"])]
    return a


def sensitivity(args, o, pcm, ref=None, ext_vad=None, navg0=None):
    """How far the oracle's own output for this configuration and input moves when the spectrum that leaves its front end
    moves by one rounding error of the CUDA front end (co.run_pipeline(perturb=...)): |out(perturbed) - out|, the larger of
    two sign patterns.  None when the configuration has no spectral subtraction (every other chain is well conditioned and
    is held to north_star's bar alone) or when the VAD module changes rows (drop / silence).
    The rounding error: the device computes the spectrum in fp32 (two ulps of the power spectrum = 2.4e-7) except on the
    fp64 band path taken by noise reduction after the filter bank (4.4e-16).  A spectral subtraction |X| - b N that cancels
    k digits turns that into 1e-7 x 10^k of its result, and a logarithm or a cube root behind it amplifies it again; the
    reference itself, run against another FFT library, moves by as much there."""
    if o.nr_mode == "none" or o.format_out in ("raw", "wave") or o.vad_apply_mode != "none":
        return None
    if ref is None:
        ref = co.run_pipeline(pcm, o, ext_vad, navg0=navg0)
    eps = 4.4e-16 if o.nr_when == "afterFB" else 2.4e-7
    sens = None
    for seed in (1, 2):
        with np.errstate(all="ignore"):
            r = co.run_pipeline(pcm, o, ext_vad, perturb=(eps, seed), force_vad_nr=ref.vad_nr, navg0=navg0)
        if r.features.shape != ref.features.shape:
            return None
        d = np.abs(r.features.astype(np.float64) - ref.features.astype(np.float64))
        d = np.where(np.isfinite(d), d, 0.0)
        sens = d if sens is None else np.maximum(sens, d)
    return sens


def tol_ok(got, want, kind, sens=None):
    """north_star's bar: 1e-4 relative, or 1e-3 absolute in the log domain.  Linear kinds get a floor of 1e-5 of the row
    maximum (= 100 dB below the frame's peak).  sens (see sensitivity()): behind a spectral subtraction the tolerance of an
    entry is widened by four times the oracle's own response to one front-end rounding error of its input spectrum."""
    if got.shape != want.shape:
        return False, "shape %s vs %s" % (got.shape, want.shape)
    if not gu.same_nonfinite(got, want):
        return False, "non-finite positions differ"
    fin = np.isfinite(want)
    err = np.abs(got - want)[fin]
    if kind in ("dctc", "lpc", "logspec", "trapdct", "td-iir-mfcc"):
        tol = 1e-4 * np.abs(want) + 1e-3
    else:
        rowmax = np.max(np.where(fin, np.abs(want), 0), axis=1, keepdims=True) * np.ones_like(want)
        tol = 1e-4 * np.abs(want) + 1e-5 * rowmax
    if sens is not None:
        tol = tol + 4.0 * sens
    tol = tol[fin]
    bad = err > tol
    return (not bad.any()), ("max err %.3g at %d entries (max |want| %.3g)" % (err.max() if err.size else 0, int(bad.sum()), np.abs(want[fin]).max() if fin.any() else 0))


def main():
    mode = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rng = random.Random(int(sys.argv[3]) if len(sys.argv) > 3 else 1)
    global SIGNAL_SHARE
    SIGNAL_SHARE = float(sys.argv[4]) if len(sys.argv) > 4 else 0.0
    global OTHER_FS_SHARE
    OTHER_FS_SHARE = float(sys.argv[5]) if len(sys.argv) > 5 else 0.0
    ins = [gu.inputs()[i] for i in (0, 5)]
    nbad = nrun = nskip = 0
    if mode == "cpu":
        import ref_runner as rr
    else:
        import ctucopy_b200 as cb
    for it in range(n):
        args = draw(rng)
        try:
            o = co.parse_args(args)
            refs = [co.run_pipeline(u, o) for u in ins]
        except Exception as e:  # the option set is invalid for the reference too
            nskip += 1
            continue
        if mode == "cpu":
            do_vad = "-vad_out_mode" in args
            r = rr.run_reference(args, ins, opt="O0", one_per_process=True, vad_out=do_vad)
            if r["returncode"] != 0 or any(x is None for x in r["outputs"]):
                print("REFERENCE FAILED rc=%s: %s\n   %s" % (r["returncode"], ("-fs %s " % args[1]) + " ".join(args[7:]), r["stderr"].strip()[-120:]))
                nskip += 1
                continue
            for i in range(len(ins)):
                if o.format_out == "raw":
                    want = np.frombuffer(r["outputs"][i], dtype="<i2")
                    got = refs[i].waveform
                    nrun += 1
                    if got.shape != want.shape or not np.array_equal(got, want):
                        nbad += 1
                        print("ORACLE != REFERENCE waveform (input %d, %s vs %s, differing %s): %s" % (i, got.shape, want.shape, int((got != want).sum()) if got.shape == want.shape else -1, ("-fs %s " % args[1]) + " ".join(args[7:])))
                    continue
                want = rr.parse_htk(r["outputs"][i])[1]
                got = refs[i].features
                ok = got.shape == want.shape and gu.same_nonfinite(got, want) and np.allclose(got[np.isfinite(want)], want[np.isfinite(want)], rtol=3e-6, atol=3e-6)
                if do_vad and ok:
                    v = np.frombuffer(r["vad"][i]["vad"], dtype=np.uint8) - 48
                    ok = np.array_equal(v, refs[i].vad.vad.astype(np.uint8))
                nrun += 1
                if not ok:
                    nbad += 1
                    d = np.abs(got - want)[np.isfinite(want)].max() if got.shape == want.shape else -1
                    print("ORACLE != REFERENCE (input %d, shape %s vs %s, max diff %.3g): %s" % (i, got.shape, want.shape, d, ("-fs %s " % args[1]) + " ".join(args[7:])))
        else:
            try:
                res = cb.extract(args, ins)
            except cb.CtuError as e:
                if e.status == 3:
                    nskip += 1
                    continue
                print("CUDA PATH ERROR %s: %s" % (e.message[:80], ("-fs %s " % args[1]) + " ".join(args[7:])))
                nbad += 1
                continue
            for i in range(len(ins)):
                if o.format_out == "raw":
                    g, w_ = res.utt_waveform(i).astype(np.int32), refs[i].waveform.astype(np.int32)
                    ok = g.shape == w_.shape and np.abs(g - w_).max() <= 1 and (g != w_).mean() < 0.02
                    why = "waveform: shape %s vs %s, max |diff| %s, share differing %.4f" % (g.shape, w_.shape, np.abs(g - w_).max() if g.shape == w_.shape else -1, (g != w_).mean() if g.shape == w_.shape else 1)
                else:
                    ok, why = tol_ok(res.utt_features(i), refs[i].features, o.fea_kind, sensitivity(args, o, ins[i], refs[i]))
                nrun += 1
                r0 = int(res.row_offsets[i])
                if refs[i].vad_nr is not None:
                    gv = res.vad_nr[r0: r0 + refs[i].nframes].astype(bool)
                    if not np.array_equal(gv, refs[i].vad_nr):
                        ok, why = False, why + "; detector decisions differ at %d frames" % int((gv != refs[i].vad_nr).sum())
                if refs[i].vad is not None:
                    gv = res.vad_out[r0: r0 + refs[i].nframes].astype(bool)
                    if not np.array_equal(gv, refs[i].vad.vad):
                        ok, why = False, why + "; VAD-module decisions differ at %d frames" % int((gv != refs[i].vad.vad).sum())
                if not ok:
                    nbad += 1
                    print("CUDA != ORACLE (input %d): %s\n   %s" % (i, why, ("-fs %s " % args[1]) + " ".join(args[7:])))
    print("sweep %s: %d comparisons, %d mismatches, %d option sets skipped" % (mode, nrun, nbad, nskip))
    return 1 if nbad else 0


if __name__ == "__main__":
    sys.exit(main())
