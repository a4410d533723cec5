"""The reference's own example suite (egs/run.sh, egs/conf/NN_*.ctuconf) through the command-line host, stage by stage,
against the reference binary run on the same lists.  The option lists below are the configuration files of the examples
(comments stripped); list formats follow egs/conf/NN_list.scp.  Left out: 04, 05 and 12 apply CMVN from an EXISTING
statistics file, which makes the reference itself write +-inf for every value; 20 needs a filter
file the reference does not ship."""
import os
import subprocess

import numpy as np
import pytest

import golden_util as gu
import ref_runner as rr

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "host", "ctucopy_b200")

RAW = ("-fs 16000 -format_in raw -format_out {OUT} -endian_in little -endian_out little -w 25 -s 10 -preem 0.97 -fb_scale mel -fb_shape triang "
       "-fb_power on -fb_definition 30filters -nr_mode none -fb_eqld off -fb_inld off -fea_kind dctc -fea_ncepcoefs 12 -fea_c0 on -fea_E off "
       "-fea_lifter 22 -fea_rawenergy off")
HTK = "-fs 16000 -format_in htk -format_out {OUT} -endian_in little -endian_out little -nfeacoefs 13"
# name: (options, input kind, list columns, what is produced)
EXAMPLES = [
    ("01", RAW.format(OUT="htk"), "raw", "in out", "htk"),
    ("02", RAW.format(OUT="htk") + " -stat_cmvn {D}/02.stat", "raw", "in spk", "stat"),
    ("03", HTK.format(OUT="htk") + " -stat_cmvn {D}/03.stat", "fea", "in spk", "stat"),
    ("06", RAW.format(OUT="htk") + " -apply_cmvn {D}/06.stat", "raw", "in out spk", "stat+htk"),
    ("07", HTK.format(OUT="htk") + " -apply_cmvn {D}/07.stat", "fea", "in out spk", "stat+htk"),
    ("08", RAW.format(OUT="htk") + " -stat_cmvn {D}/08.stat -apply_cmvn {D}/08.stat", "raw", "in out spk", "stat+htk"),
    ("09", HTK.format(OUT="htk") + " -stat_cmvn {D}/09.stat -apply_cmvn {D}/09.stat", "fea", "in out spk", "stat+htk"),
    ("10", HTK.format(OUT="ark={D}/10.ark") + " -stat_cmvn {D}/10.stat -apply_cmvn {D}/10.stat", "fea", "in out spk", "stat+ark"),
    ("11", RAW.format(OUT="ark={D}/11.ark"), "raw", "in out", "ark"),
    ("15", RAW.format(OUT="htk") + " -fea_delta d_a -d_win 2 -a_win 2 -t_win 2", "raw", "in out", "htk"),
    ("16", RAW.format(OUT="ark={D}/16.ark") + " -fea_delta d_a -fea_Z_exp 1000", "raw", "in out", "ark"),
    ("17", RAW.format(OUT="ark={D}/17.ark") + " -fea_trap 5", "raw", "in out", "ark"),
    ("18", HTK.format(OUT="ark={D}/18.ark") + " -fea_trap 5", "fea", "in out", "ark"),
    ("19", HTK.format(OUT="ark={D}/19.ark"), "fea", "in out", "ark"),
    ("21", "-fs 16000 -format_in raw -format_out raw -preset exten", "raw", "in out", "raw"),
]
SPK = ["SA000", "SA001", "SA000"]


def _numbers(text):
    names, vals = [], []
    for ln in text.splitlines():
        if ln.startswith("mean\t") or ln.startswith("var\t"):
            vals.append([float(x) for x in ln.split("\t", 1)[1].split()])
        elif ln.strip():
            names.append(ln.strip())
    return names, np.array(vals)


def _close(a, b):
    return a.shape == b.shape and bool(np.all(np.abs(a - b) <= 1e-4 * np.abs(b) + 1e-3))


@pytest.mark.skipif(rr.ref_binary("O0") is None, reason="oracle/_ref not built")
def test_reference_example_suite_through_the_cli(tmp_path):
    ins = gu.inputs()
    utts = [ins[5], ins[4], ins[0]]                    # real speech (egs/sig/SA000CB1.CS0, 1.2 s), pink noise + tone, tone/chirp
    dirs = {}
    for who, exe in (("ref", rr.ref_binary("O0")), ("cli", EXE)):
        d = str(tmp_path / who)
        os.makedirs(d)
        dirs[who] = d
        for i, u in enumerate(utts):
            np.asarray(u).astype("<i2").tofile(os.path.join(d, "u%d.raw" % i))
        for name, opts, kind, cols, prod in EXAMPLES:
            # feature input = this implementation's own output of example 01 (like egs/conf/03_list.scp points into data/01)
            src = ["u%d.raw" % i if kind == "raw" else "ex01_%d.htk" % i for i in range(3)]
            lst = os.path.join(d, "list%s.scp" % name)
            with open(lst, "w") as fh:
                for i in range(3):
                    c = [os.path.join(d, src[i])]
                    if "out" in cols:
                        c.append(os.path.join(d, "ex%s_%d.%s" % (name, i, "htk" if prod != "raw" else "raw")) if "ark" not in prod else "utt%d" % i)
                    if "spk" in cols:
                        c.append(SPK[i])
                    fh.write(" ".join(c) + "\n")
            pr = subprocess.run([exe] + opts.replace("{D}", d).split() + ["-S", lst], capture_output=True, cwd=d)
            assert pr.returncode == 0, (who, name, pr.stderr.decode()[-300:])
    r, c = dirs["ref"], dirs["cli"]
    for name, opts, kind, cols, prod in EXAMPLES:
        if "stat" in prod:
            rn, rv = _numbers(open(os.path.join(r, name + ".stat")).read())
            cn, cv = _numbers(open(os.path.join(c, name + ".stat")).read())
            assert rn == cn and rv.shape == cv.shape, name
            np.testing.assert_allclose(cv, rv, rtol=2e-4, atol=2e-4, err_msg="example " + name)
        if "htk" in prod:
            for i in range(3):
                a, b = open(os.path.join(c, "ex%s_%d.htk" % (name, i)), "rb").read(), open(os.path.join(r, "ex%s_%d.htk" % (name, i)), "rb").read()
                assert len(a) == len(b) and a[:12] == b[:12], (name, i)
                assert _close(rr.parse_htk(a)[1], rr.parse_htk(b)[1]), (name, i)
        if "ark" in prod:
            ma, mb = rr.parse_ark(open(os.path.join(c, name + ".ark"), "rb").read()), rr.parse_ark(open(os.path.join(r, name + ".ark"), "rb").read())
            assert list(ma) == list(mb) == ["utt0", "utt1", "utt2"], name
            for k in ma:
                assert _close(ma[k], mb[k]), (name, k)
            sa, sb = open(os.path.join(c, name + ".scp")).read(), open(os.path.join(r, name + ".scp")).read()
            assert sa.replace(c, "") == sb.replace(r, ""), name          # same keys, same byte offsets
        if prod == "raw":
            for i in range(3):
                a = np.fromfile(os.path.join(c, "ex21_%d.raw" % i), "<i2").astype(np.int32)
                b = np.fromfile(os.path.join(r, "ex21_%d.raw" % i), "<i2").astype(np.int32)
                assert a.shape == b.shape and np.abs(a - b).max() <= 1, (name, i)
