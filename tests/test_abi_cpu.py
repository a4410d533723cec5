"""CPU-only checks of the C-ABI library: it loads, exports every symbol the header
declares, parses options like the reference, designs filter banks bit-exactly, and REFUSES
to run without a GPU (no CPU fallback)."""
import json
import os
import re

import numpy as np
import pytest

import ctu_oracle as co
import golden_util as gu
import ref_runner as rr
import ctucopy_b200 as cb

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "ctucopy_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ctu_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = cb.lib()
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libctucopy_b200.so does not export %s" % n


FIELDS = ["fs", "preem", "dither", "remove_dc", "remove_dc1", "window_ms", "wshift_ms", "fb_scale", "fb_shape", "fb_norm", "fb_power",
          "fb_eqld", "fb_inld", "fb_definition", "vadmode", "nr_mode", "nr_p", "nr_q", "nr_a", "nr_b", "nr_initsegs", "fea_kind",
          "fea_lporder", "fea_ncepcoefs", "fea_c0", "fea_E", "fea_rawenergy", "fea_lifter", "d_win", "a_win", "t_win", "n_order",
          "vad_apply_mode", "vad_out_mode", "vad_cri_mode", "vad_thr_mode", "vad_energy_db", "vad_cepdist_mode", "vad_cepdist_p",
          "vad_cepdist_init", "vad_lpc_coefs", "vad_absolute_thr", "vad_perc_init", "vad_perc_thr", "vad_adapt_init", "vad_adapt_q",
          "vad_adapt_za", "vad_dyn_init", "vad_dyn_perc", "vad_dyn_min", "vad_dyn_qmaxinc", "vad_dyn_qmaxdec", "vad_dyn_qmindec",
          "vad_dyn_qmininc", "vad_filter_order", "window", "wshift", "wfft", "wfftby2", "phase_needed", "format_out",
          "fea_delta", "fea_trap", "trap_win", "nfeacoefs", "weight_of_td_iir_mfcc_bank"]


@pytest.mark.parametrize("name", gu.case_names() + gu.feain_case_names())
def test_config_parser_agrees_with_oracle_parser(name):
    c = gu.Case(name)
    args = c.oracle_args()
    o = co.parse_args(args)
    g = cb.parse_config(args)
    for f in FIELDS:
        a, b = getattr(o, f), getattr(g, f)
        if isinstance(b, bytes):
            b = b.decode()
        if isinstance(a, bool):
            a = int(a)
        assert a == b, (name, f, a, b)
    assert (o.nr_when == "afterFB") == bool(g.nr_when)
    assert (o.format_in == "htk") == bool(g.fea_in)
    assert o.ffilters == g.filters.decode()
    if o.fea_kind == "trapdct":
        assert (o.fea_trapdct_traplen, o.fea_trapdct_ndct) == (g.fea_trapdct_traplen, g.fea_trapdct_ndct)


def test_config_file_and_errors(tmp_path):
    p = tmp_path / "a.ctuconf"
    p.write_text("# CtuCopy config file\n-fs 16000       # sampling freq\n-preset mfcc\n-preem 0.97 \n-fb_definition 30filters \n")
    g = cb.parse_config(["-C", str(p), "-fea_delta", "d_a", "-w", "20"])
    assert g.fs == 16000 and g.fea_kind == b"dctc" and g.fb_definition == b"30filters" and g.n_order == 2 and g.window == 320
    with pytest.raises(cb.CtuError) as e:
        cb.parse_config(["-fs", "16000", "-bogus", "1"])
    assert "Syntax error" in e.value.message
    with pytest.raises(cb.CtuError) as e:
        cb.parse_config(["-preset", "mfcc"])
    assert "sampling rate" in e.value.message
    with pytest.raises(cb.CtuError):
        cb.parse_config(["-fs", "16000", "-preem", "1.5"])
    # a value starting with '-' is never consumed (src/io/opts.cc:185-192): "-3" is then
    # parsed as an (unknown) option, exactly as the reference does
    with pytest.raises(cb.CtuError) as e:
        cb.parse_config(["-fs", "16000", "-nr_b", "-3", "-preset", "mfcc"])
    assert '"-3"' in e.value.message


def test_filter_bank_design_bit_exact_vs_oracle_and_printself():
    z = np.load(os.path.join(gu.GOLDEN, "fb_design.npz"))
    for nm in [f for f in z.files if not f.endswith("_args")]:
        args = json.loads(str(z[nm + "_args"]))
        o = co.parse_args(args)
        want = co.fb_design(o)
        mat, lo, hi = cb.design_filter_bank(cb.parse_config(args))
        assert mat.shape == want.mat.shape, nm
        assert np.array_equal(mat, want.mat), nm           # same fp64 expressions -> same bits
        assert np.array_equal(lo, want.lo) and np.array_equal(hi, want.hi), nm
        np.testing.assert_allclose(mat, z[nm], rtol=2e-5, atol=0)   # reference's own printout (6 digits)


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(cb.CtuError) as e:
        cb.Handle(["-fs", "16000", "-preset", "mfcc", "-format_out", "htk"])
    assert e.value.status == 4 and "no CPU fallback" in e.value.message


def test_fft_index_algebra_on_the_cpu(tmp_path):
    """ctu_fft.cuh is __host__ __device__: the four-step 256-point FFT, the on-the-fly twiddles, the real split
    and the inverse pre-split (shared-memory and register-exchange variants), and the 16 x 8 decomposition of the
    256-point front end (8 threads per frame), run here thread by thread against a naive DFT (tests/emu/emu_fft.cpp)."""
    import subprocess
    exe = str(tmp_path / "emu_fft")
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", exe, os.path.join(ROOT, "tests", "emu", "emu_fft.cpp")])
    pr = subprocess.run([exe], capture_output=True, text=True)
    assert pr.returncode == 0 and "OK" in pr.stdout, pr.stdout


def test_td_iir_mfcc_refusals_and_filter_file_errors(tmp_path):
    """-fea_kind td-iir-mfcc: what the reference does not survive on this branch is refused before any device is
    touched (status 3, with the reference line), and the coefficient file is read like rawIN::loadf_filters reads it, minus its
    undefined cases (status 2).  A valid configuration then fails only for the missing GPU (status 4)."""
    base = ["-fs", "16000", "-format_in", "raw", "-format_out", "htk", "-w", "30", "-s", "10", "-fea_kind", "td-iir-mfcc", "-fea_ncepcoefs", "12"]
    good = base + ["-filters", rr.TDIIR_FILTERS]
    for extra in (["-fea_E", "on"], ["-fea_delta", "d_a"], ["-fea_trap", "5"], ["-fea_c0", "off"], ["-fea_ncepcoefs", "13"],
                  ["-stat_cmvn", "x.stat"], ["-vad_out_mode", "vad"], ["-format_out", "raw"]):
        with pytest.raises(cb.CtuError) as e:
            cb.Handle(good + extra)
        assert e.value.status == 3 and "td-iir-mfcc" in e.value.message, (extra, e.value.message)
    with pytest.raises(cb.CtuError) as e:
        cb.Handle(base + ["-filters", str(tmp_path / "missing.asc")])
    assert e.value.status == 2 and "filter coefficient file" in e.value.message
    lines = open(rr.TDIIR_FILTERS).read().splitlines()
    short = tmp_path / "short.asc"
    short.write_text("\n".join(lines[:23]) + "\n")
    with pytest.raises(cb.CtuError) as e:
        cb.Handle(base + ["-filters", str(short)])
    assert e.value.status == 2 and "fewer than 24" in e.value.message
    ragged = tmp_path / "ragged.asc"
    ragged.write_text("\n".join(lines[:5] + ["\t".join(lines[5].split("\t")[:9])] + lines[6:]) + "\n")
    with pytest.raises(cb.CtuError) as e:
        cb.Handle(base + ["-filters", str(ragged)])
    assert e.value.status == 2 and "fewer than 10" in e.value.message
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(cb.CtuError) as e:
            cb.Handle(good)
        assert e.value.status == 4 and "no CPU fallback" in e.value.message
