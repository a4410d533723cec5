"""
ref_runner.py -- TEST INFRASTRUCTURE ONLY.

Drives the reference binary built by oracle/build_ref.sh (oracle/_ref/ctucopy4_O0|_O2)
and parses what it writes.  Always list mode: single-file and online modes segfault in
CtuCopy 4.0.2 (src/io/batch.cc:43-44, 310; SURVEY.md finding 7).
"""
from __future__ import annotations

import os
import struct
import subprocess
import tempfile
from typing import Dict, List, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# coefficient file of the td-iir-mfcc goldens (tests/golden/make_tdiir_filters.py)
TDIIR_FILTERS = os.path.join(os.path.dirname(HERE), "tests", "golden", "tdiir_filters.asc")


def ref_binary(opt: str = "O0") -> Optional[str]:
    p = os.path.join(HERE, "_ref", "ctucopy4_" + opt)
    return p if os.path.exists(p) else None


def parse_htk(b: bytes, endian: str = "<"):
    n, period, size, kind = struct.unpack(endian + "IIHH", b[:12])
    dim = size // 4
    a = np.frombuffer(b[12:12 + n * size], dtype=endian + "f4").reshape(n, dim)
    return dict(nsamples=n, period=period, size=size, kind=kind), a


def parse_ark(b: bytes) -> Dict[str, np.ndarray]:
    out = {}
    i = 0
    while i < len(b):
        j = b.index(b" ", i)
        key = b[i:j].decode()
        assert b[j + 1:j + 6] == b"\x00BFM ", b[j:j + 8]
        p = j + 6
        assert b[p] == 4
        rows = struct.unpack("<i", b[p + 1:p + 5])[0]
        assert b[p + 5] == 4
        cols = struct.unpack("<i", b[p + 6:p + 10])[0]
        p += 10
        out[key] = np.frombuffer(b[p:p + rows * cols * 4], dtype="<f4").reshape(rows, cols)
        i = p + rows * cols * 4
    return out


def parse_pfile(b: bytes):
    hdr = b[:32768].split(b"\x00", 1)[0].decode()
    f = {}
    for line in hdr.splitlines():
        tok = line.split()
        f[tok[0]] = tok[1:]
    nfr = int(f["-num_frames"][0]); nfea = int(f["-num_features"][0]); ns = int(f["-num_sentences"][0])
    ncol = nfea + 2
    rows = np.frombuffer(b[32768:32768 + nfr * ncol * 4], dtype=">u4").reshape(nfr, ncol)
    feat = rows[:, 2:].copy().view(">f4").astype("<f4")
    table = np.frombuffer(b[32768 + nfr * ncol * 4: 32768 + nfr * ncol * 4 + (ns + 1) * 4], dtype=">u4")
    return dict(header=hdr, sent_id=rows[:, 0].astype(np.int64), frame_id=rows[:, 1].astype(np.int64),
                table=table.astype(np.int64)), feat


def run_reference(args: Sequence[str], pcms: List[np.ndarray], *, opt: str = "O0", out_ext: str = "out",
                  vad_out: bool = False, ext_vad_bytes: Optional[bytes] = None, one_per_process: bool = False,
                  endian_in: str = "<"):
    """Runs `ctucopy4 <args> -S list` on the given utterances.  Returns a dict with, per
    utterance, the raw bytes of every file written, plus container files when `args`
    names ark=/pfile= targets via the placeholders {ARK} / {PFILE} / {VADIN}."""
    exe = ref_binary(opt)
    if exe is None:
        raise FileNotFoundError("oracle/_ref not built (run oracle/build_ref.sh where /root/reference exists)")
    res = dict(outputs=[], vad=[], stderr="", returncode=0, files={})
    with tempfile.TemporaryDirectory() as d:
        groups = [[i] for i in range(len(pcms))] if one_per_process else [list(range(len(pcms)))]
        for i, p in enumerate(pcms):
            p = np.asarray(p)
            if p.dtype.kind == "f":
                # a feature matrix: written as an HTK parameter file (input of `-format_in htk`, src/io/in.cc:630-680;
                # only the vector size of the header is used by the reader)
                with open(os.path.join(d, "u%d.raw" % i), "wb") as fh:
                    fh.write(struct.pack(endian_in + "IIHH", p.shape[0], 100000, 4 * p.shape[1], 6))
                    fh.write(p.astype(endian_in + "f4").tobytes())
                continue
            if p.dtype == np.uint8:                 # 8-bit G.711 codes (-format_in alaw | mulaw): the file is the byte stream
                p.tofile(os.path.join(d, "u%d.raw" % i))
                continue
            p.astype(endian_in + "i2").tofile(os.path.join(d, "u%d.raw" % i))
        if ext_vad_bytes is not None:
            open(os.path.join(d, "vadin.bin"), "wb").write(ext_vad_bytes)
        for gi, g in enumerate(groups):
            with open(os.path.join(d, "list%d.scp" % gi), "w") as fh:
                for i in g:
                    line = "%s/u%d.raw %s/u%d.%s" % (d, i, d, i, out_ext)
                    if vad_out:
                        line += " spk %s/u%d.vad" % (d, i)
                    fh.write(line + "\n")
            a = [s.replace("{ARK}", os.path.join(d, "out%d.ark" % gi)).replace("{PFILE}", os.path.join(d, "out%d.pfile" % gi))
                 .replace("{VADIN}", os.path.join(d, "vadin.bin")).replace("{FILTERS}", TDIIR_FILTERS) for s in args]
            pr = subprocess.run([exe] + a + ["-S", os.path.join(d, "list%d.scp" % gi)], capture_output=True, cwd=d)
            res["stderr"] += pr.stderr.decode(errors="replace")
            res["returncode"] = pr.returncode or res["returncode"]
        for i in range(len(pcms)):
            p = os.path.join(d, "u%d.%s" % (i, out_ext))
            res["outputs"].append(open(p, "rb").read() if os.path.exists(p) else None)
            v = {}
            for fn in os.listdir(d):
                if fn.startswith("u%d.vad" % i):
                    v[fn[len("u%d." % i):]] = open(os.path.join(d, fn), "rb").read()
            res["vad"].append(v)
        for fn in os.listdir(d):
            if fn.startswith("out") and (fn.endswith(".ark") or fn.endswith(".pfile") or fn.endswith(".scp")):
                res["files"][fn] = open(os.path.join(d, fn), "rb").read()
        res["tmpdir"] = d
    return res
