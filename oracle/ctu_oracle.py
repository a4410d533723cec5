"""
ctu_oracle.py -- TEST INFRASTRUCTURE ONLY.  Not part of the product.

CPU (numpy, float64) restatement of CtuCopy 4.0.2's framewise analysis / enhancement /
feature path, written function-by-function after the reference sources (file:line cited
on every function; paths are relative to /root/reference).  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module; the
product (ctucopy_b200/, libctucopy_b200.so, the ctucopy_b200 CLI) never does.

Pinning: the reference ships NO tests, golden vectors or known-answer files for this path
(SURVEY.md section 4), so this restatement is pinned against outputs of the reference
ITSELF, built here from its unmodified sources by oracle/build_ref.sh (FFTW replaced by
oracle/shim/fftw3.h, the published FFTW conventions).  tests/golden/make_golden.py runs
that binary on egs/sig/* and on the synthetic parity set and commits the results;
tests/test_oracle_vs_golden.py checks this module against them (bit-identical float32 /
int16 for every pipeline except where noted there).

Third-party arithmetic: FFTW 3.x (version unpinned by the reference: README:24,
src/objects.mk:7).  Here numpy.fft (pocketfft, float64) plays that role.
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass, field, replace
from typing import List, Optional, Tuple

import numpy as np

# --------------------------------------------------------------------------------------
# options  (src/io/opts.h:36-166, defaults src/io/opts.cc:34-146)
# --------------------------------------------------------------------------------------


@dataclass
class Opts:
    format_in: str = ""
    format_out: str = ""
    endian_in: str = "little"
    endian_out: str = "little"
    preem: float = 0.0            # stored as C float (src/io/opts.h:47)
    fs: int = 0
    dither: float = 0.0
    remove_dc: bool = True
    remove_dc1: bool = False
    pfilename: str = ""
    arkfilename: str = ""
    window_ms: float = 25.0
    wshift_ms: float = 10.0
    fb_scale: str = "mel"
    fb_shape: str = "triang"
    fb_norm: bool = True
    fb_power: bool = True
    fb_eqld: bool = True
    fb_inld: bool = True
    fb_definition: str = "26filters"
    vadmode: str = "none"
    filevad: str = ""
    nr_mode: str = "none"
    nr_p: float = 0.95
    nr_q: float = 0.99
    nr_a: float = 1.0
    nr_b: float = 1.0
    nr_initsegs: int = 10
    nr_when: str = "beforeFB"
    fea_kind: str = "lpc"
    fea_lporder: int = 12
    fea_ncepcoefs: int = 12
    fea_c0: bool = True
    fea_E: bool = False
    fea_Z_exp: float = -1.0            # stored as float (src/io/opts.h); check_config turns the time constant into the coefficient
    fea_Z_block: float = -1.0
    length_b: int = 0
    cms_exp_coef: float = -1.0
    stat_cmvn: bool = False
    apply_cmvn: bool = False
    fea_rawenergy: bool = False
    fea_lifter: int = 22
    fea_trapdct_traplen: int = 0
    fea_trapdct_ndct: int = 0
    d_win: int = 2
    a_win: int = 2
    t_win: int = 2
    fea_delta: bool = False
    n_order: int = 0
    fea_trap: bool = False
    nfeacoefs: int = 13
    trap_win: int = 5
    ffilters: str = ""                          # -filters (src/io/opts.cc:100, 705-706)
    weight_of_td_iir_mfcc_bank: float = float(np.float32(2.026))    # C float (src/io/opts.h:164, src/io/opts.cc:104)
    vad_apply_mode: str = "none"
    vad_out_mode: str = "none"
    vad_cri_mode: str = "energy"
    vad_thr_mode: str = "perc"
    vad_energy_db: bool = True
    vad_cepdist_mode: str = "lpc"
    vad_cepdist_p: float = 0.8
    vad_cepdist_init: int = 4
    vad_lpc_coefs: int = 14
    vad_absolute_thr: float = 1.0
    vad_perc_init: int = 10
    vad_perc_thr: float = 50.0
    vad_adapt_init: int = 20
    vad_adapt_q: float = 0.9
    vad_adapt_za: float = 2.0
    vad_dyn_init: int = 5
    vad_dyn_perc: float = 50.0
    vad_dyn_min: float = 1.0
    vad_dyn_qmaxinc: float = 0.8
    vad_dyn_qmaxdec: float = 0.995
    vad_dyn_qmindec: float = 0.8
    vad_dyn_qmininc: float = 0.9999
    vad_filter_order: int = 3
    # derived (src/io/opts.cc:255-325)
    window: int = 0
    wshift: int = 0
    wfft: int = 0
    wfftby2: int = 0
    phase_needed: bool = False


def _onoff(v: str, cur: bool) -> bool:
    if v == "on":
        return True
    if v == "off":
        return False
    return cur


def _set_preset(o: Opts, name: str) -> None:
    """src/io/opts.cc:196-253"""
    if name == "mfcc":
        o.fb_scale, o.fb_shape, o.fb_power = "mel", "triang", True
        o.fb_definition = "1-26/26filters"
        o.nr_mode = "none"
        o.fb_eqld = o.fb_inld = False
        o.fea_kind, o.fea_ncepcoefs, o.fea_c0, o.fea_E = "dctc", 12, True, False
        o.fea_lifter, o.fea_rawenergy = 22, False
    elif name == "plpc":
        o.fb_scale, o.fb_shape, o.fb_power = "bark", "trapez", True
        o.fb_definition = "1-15/15filters"
        o.nr_mode = "none"
        o.fb_eqld = o.fb_inld = True
        o.fea_kind, o.fea_lporder, o.fea_ncepcoefs = "lpc", 12, 12
        o.fea_c0, o.fea_E, o.fea_lifter, o.fea_rawenergy = True, False, 22, False
    elif name == "exten":
        o.window_ms, o.wshift_ms = 32.0, 16.0
        o.fb_definition = o.fb_scale = o.fb_shape = "none"
        o.nr_a = 2.0
        o.fb_eqld = o.fb_inld = o.fb_power = o.fb_norm = False
        o.nr_mode, o.fea_kind = "exten", "none"
        o.fea_c0 = o.fea_E = False
        o.fea_lifter, o.fea_rawenergy = 0, False
    else:
        raise ValueError("OPTS: Unknown preset!")


def _parse_one(o: Opts, l: str, r: Optional[str]) -> None:
    """src/io/opts.cc:644-846 (hot-path subset; list/-i/-o handled by the caller)"""
    f = float
    if l in ("-S", "-i", "-o", "-C", "-v", "-verbose", "-quiet", "-info", "-online_in", "-online_out",
             "-fb_printself"):
        return
    if r is None:
        # every remaining option needs a value; the reference silently ignores a missing one
        if l == "-fea_kind":
            raise ValueError("OPTS: Missing argument to '-fea_kind' option!")
        return
    if l == "-format_in": o.format_in = r
    elif l == "-format_out":
        if "pfile=" in r: o.pfilename, o.format_out = r.split("=", 1)[1], "pfile"
        elif "ark=" in r: o.arkfilename, o.format_out = r.split("=", 1)[1], "ark"
        else: o.format_out = r
    elif l == "-endian_in":
        if r in ("big", "little"): o.endian_in = r
    elif l == "-endian_out":
        if r in ("big", "little"): o.endian_out = r
    elif l == "-preem": o.preem = float(np.float32(f(r)))
    elif l == "-fea_delta":
        o.fea_delta, o.fea_trap = True, False
        if r == "d": o.n_order = 1
        elif r == "d_a": o.n_order = 2
        elif r == "d_a_t": o.n_order = 3
        else: o.fea_delta = False
    elif l == "-fea_trap":
        if not o.fea_delta:
            o.fea_trap, o.trap_win, o.fea_delta, o.n_order = True, int(r), True, 1
            o.d_win = (o.trap_win - 1) // 2
    elif l == "-nfeacoefs": o.nfeacoefs = int(r)
    elif l == "-filters":
        if r is not None: o.ffilters = r
    elif l == "-weight_of_td_iir_mfcc_bank":
        if r is not None: o.weight_of_td_iir_mfcc_bank = float(np.float32(float(r)))
    elif l == "-fs": o.fs = int(r)
    elif l == "-dither": o.dither = f(r)
    elif l == "-remove_dc": o.remove_dc = _onoff(r, o.remove_dc)
    elif l == "-remove_dc1": o.remove_dc1 = _onoff(r, o.remove_dc1)
    elif l == "-w": o.window_ms = f(r)
    elif l == "-s": o.wshift_ms = f(r)
    elif l == "-fb_scale": o.fb_scale = r
    elif l == "-fb_shape": o.fb_shape = r
    elif l == "-fb_norm": o.fb_norm = _onoff(r, o.fb_norm)
    elif l == "-fb_power": o.fb_power = _onoff(r, o.fb_power)
    elif l == "-fb_eqld": o.fb_eqld = _onoff(r, o.fb_eqld)
    elif l == "-fb_inld": o.fb_inld = _onoff(r, o.fb_inld)
    elif l == "-fb_definition": o.fb_definition = r
    elif l == "-vad":
        if r == "burg": o.vadmode = "burg"
        elif "file=" in r: o.filevad, o.vadmode = r.split("=", 1)[1], "file"
        else: raise ValueError("OPTS: Syntax error in option -vad !")
    elif l == "-nr_mode": o.nr_mode = r
    elif l == "-nr_p": o.nr_p = f(r)
    elif l == "-nr_q": o.nr_q = f(r)
    elif l == "-nr_a": o.nr_a = f(r)
    elif l == "-nr_b": o.nr_b = f(r)
    elif l == "-nr_initsegs": o.nr_initsegs = int(r)
    elif l == "-nr_when":
        if r in ("beforeFB", "afterFB"): o.nr_when = r
    elif l == "-fea_kind":
        if "trapdct" in r:
            parts = r.split(",")
            if len(parts) < 3:
                raise ValueError("OPTS: Syntax error in option -fea_kind!")
            o.fea_kind, o.fea_trapdct_traplen, o.fea_trapdct_ndct = parts[0], int(parts[1]), int(parts[2])
        else:
            o.fea_kind = r
    elif l == "-d_win": o.d_win = int(r)
    elif l == "-a_win": o.a_win = int(r)
    elif l == "-t_win": o.t_win = int(r)
    elif l == "-fea_lporder": o.fea_lporder = int(r)
    elif l == "-fea_ncepcoefs": o.fea_ncepcoefs = int(r)
    elif l == "-fea_c0": o.fea_c0 = _onoff(r, o.fea_c0)
    elif l == "-fea_E": o.fea_E = _onoff(r, o.fea_E)
    elif l == "-stat_cmvn":
        if r is not None: o.stat_cmvn = True
    elif l == "-apply_cmvn":
        if r is not None: o.apply_cmvn = True
    elif l == "-fea_Z_exp":
        if r is not None: o.fea_Z_exp = float(np.float32(float(r)))
    elif l == "-fea_Z_block":
        if r is not None: o.fea_Z_block = float(np.float32(float(r)))
    elif l == "-fea_rawenergy": o.fea_rawenergy = _onoff(r, o.fea_rawenergy)
    elif l == "-fea_lifter": o.fea_lifter = int(r)
    elif l == "-vad_apply_mode": o.vad_apply_mode = r
    elif l == "-vad_out_mode": o.vad_out_mode = r
    elif l == "-vad_cri_mode": o.vad_cri_mode = r
    elif l == "-vad_thr_mode": o.vad_thr_mode = r
    elif l == "-vad_energy_db": o.vad_energy_db = _onoff(r, o.vad_energy_db)
    elif l == "-vad_cepdist_mode": o.vad_cepdist_mode = r
    elif l == "-vad_cepdist_p": o.vad_cepdist_p = f(r)
    elif l == "-vad_cepdist_init": o.vad_cepdist_init = int(r)
    elif l == "-vad_lpc_coefs": o.vad_lpc_coefs = int(r)
    elif l == "-vad_absolute_thr": o.vad_absolute_thr = f(r)
    elif l == "-vad_perc_init": o.vad_perc_init = int(r)
    elif l == "-vad_perc_thr": o.vad_perc_thr = f(r)
    elif l == "-vad_adapt_init": o.vad_adapt_init = int(r)
    elif l == "-vad_adapt_q": o.vad_adapt_q = f(r)
    elif l == "-vad_adapt_za": o.vad_adapt_za = f(r)
    elif l == "-vad_dyn_init": o.vad_dyn_init = int(r)
    elif l == "-vad_dyn_perc": o.vad_dyn_perc = f(r)
    elif l == "-vad_dyn_min": o.vad_dyn_min = f(r)
    elif l == "-vad_dyn_qmaxinc": o.vad_dyn_qmaxinc = f(r)
    elif l == "-vad_dyn_qmaxdec": o.vad_dyn_qmaxdec = f(r)
    elif l == "-vad_dyn_qmindec": o.vad_dyn_qmindec = f(r)
    elif l == "-vad_dyn_qmininc": o.vad_dyn_qmininc = f(r)
    elif l == "-vad_filter_order": o.vad_filter_order = int(r)
    elif l == "-preset": _set_preset(o, r)
    else:
        raise ValueError('OPTS: Syntax error in option "%s".' % l)


def check_config(o: Opts) -> Opts:
    """src/io/opts.cc:255-325 -- derived sizes and forced settings."""
    if o.fs == 0:
        raise ValueError("OPTS: Please specify sampling rate!")
    o.window = int(math.floor(0.5 + o.window_ms / 1000.0 * float(o.fs)))
    o.wshift = int(math.floor(0.5 + o.wshift_ms / 1000.0 * float(o.fs)))
    i = 1048576
    while i > 4:  # src/io/opts.cc:274-278
        if o.window // i == 1:
            o.wfft = i * (1 + (1 if (o.window % i) != 0 else 0))
        i //= 2
    o.wfftby2 = o.wfft // 2 + 1
    # CMS (src/io/opts.cc:270-274): block length in frames; the exp time constant becomes the (float) coefficient
    o.cms_exp_coef = -1.0
    if o.fea_Z_block != -1:
        o.length_b = int(math.floor((o.fea_Z_block - o.window_ms) / o.wshift_ms)) + 1
    if o.fea_Z_exp != -1:
        o.cms_exp_coef = float(np.float32(1.0 - (2 * o.wshift_ms) / o.fea_Z_exp))
    o.phase_needed = o.format_out in ("raw", "wave")
    if o.vadmode == "burg":
        o.phase_needed = True
    if o.preem >= 1.0 or o.preem < 0.0:
        raise ValueError("OPTS: Preemphasis not in range <0,1)!")
    if o.format_out in ("raw", "wave") and o.fb_power:
        o.fb_power = False
    return o


def parse_args(argv: List[str]) -> Opts:
    """Mimics opts::opts (src/io/opts.cc:27-194): defaults, then -C file, then argv.
    An argument that begins with '-' is never taken as a value (src/io/opts.cc:185-192)."""
    o = Opts()
    for j, a in enumerate(argv):
        if a == "-C":
            with open(argv[j + 1]) as fh:
                for line in fh:
                    line = line.split("#", 1)[0]
                    tok = line.split()
                    if tok:
                        _parse_one(o, tok[0], tok[1] if len(tok) > 1 else None)
    for j, a in enumerate(argv):
        if a.startswith("-"):
            r = None
            if j + 1 < len(argv) and not argv[j + 1].startswith("-"):
                r = argv[j + 1]
            _parse_one(o, a, r)
    return check_config(o)


# --------------------------------------------------------------------------------------
# front end: framing, pre-emphasis, Hamming, DC removal, FFT  (src/io/in.cc)
# --------------------------------------------------------------------------------------


def hamming(w: int, alpha: float = 0.54) -> np.ndarray:
    """_IN::hamming, src/io/in.cc:139-144 (pi = 2*asin(1))."""
    pi = 2.0 * math.asin(1.0)
    j = np.arange(w, dtype=np.float64)
    return alpha - (1 - alpha) * np.cos(2 * pi * j / (w - 1.0))


def num_frames(nsamples: int, o: Opts) -> int:
    """rawIN::new_file / get_frame, src/io/in.cc:264-279, 314: the first (w-s) samples are
    primed, then every get_frame needs s more.  Fewer than (w-s) samples throws."""
    if nsamples < o.window - o.wshift:
        raise ValueError("IO: Signal shorter than one frame!")
    return (nsamples - (o.window - o.wshift)) // o.wshift


def c_ph(re: np.ndarray, im: np.ndarray) -> np.ndarray:
    """_IN::c_ph, src/io/in.cc:187-200 (15-digit constants)."""
    hpi, pi = 1.57079632679490, 3.14159265358979
    with np.errstate(divide="ignore", invalid="ignore"):
        y = np.arctan(im / re)
    y = np.where((re < 0.0) & (im >= 0.0), y + pi, y)
    y = np.where((re < 0.0) & (im < 0.0), y - pi, y)
    y = np.where(re == 0.0, np.where(im > 0.0, hpi, -hpi), y)
    return y


@dataclass
class FrontEnd:
    Xabs: np.ndarray                  # [T, wfftby2] power (fb_power) or magnitude
    Xph: Optional[np.ndarray]         # [T, wfftby2] or None
    E: Optional[np.ndarray]           # [T] or None (src/io/in.cc:353-361, 403-413)
    spec: np.ndarray                  # complex [T, wfftby2] raw DFT (for tests)


def glibc_rand(n: int, skip: int = 0, seed: int = 1) -> np.ndarray:
    """The first skip+n values of glibc's rand() after srand(seed), values [skip:] (TYPE_3 additive feedback
    generator of random_r: r[i] = r[i-3] + r[i-31] mod 2^32, output r[i] >> 1, 310 values discarded after
    seeding).  The reference seeds it once per process (src/io/in.cc:205) and draws one value per loaded
    sample (src/io/in.cc:452-455); tests/test_oracle_vs_golden.py checks this against libc itself."""
    tot = skip + n
    r = np.zeros(tot + 344, dtype=np.uint64)
    r[0] = seed
    for i in range(1, 31):
        r[i] = (16807 * int(r[i - 1])) % 2147483647
    for i in range(31, 34):
        r[i] = r[i - 31]
    rl = [int(v) for v in r[:34]] + [0] * (tot + 310)
    for i in range(34, tot + 344):
        rl[i] = (rl[i - 31] + rl[i - 3]) & 0xFFFFFFFF
    return np.array([v >> 1 for v in rl[344 + skip: 344 + tot]], dtype=np.int64)


def dither_noise(nloaded: int, o: Opts, rand_offset: int = 0) -> np.ndarray:
    """(2*rand()/RAND_MAX - 1) * dither for the nloaded samples the reference reads of one file
    (src/io/in.cc:452-455); rand_offset = values the process has drawn for earlier files."""
    rv = glibc_rand(nloaded, rand_offset).astype(np.float64)
    return (2.0 * rv / 2147483647.0 - 1.0) * o.dither


def g711_expand(codes: np.ndarray, alaw: bool) -> np.ndarray:
    """alaw2lin, src/io/amulaw.h:19-56 (mode 1 = A-law, 0 = mu-law), as rawIN::loadframe applies it to every byte of an
    `-format_in alaw | mulaw` file (src/io/in.cc:470-500): bit-level restatement on signed chars, 16-bit result."""
    a = np.asarray(codes, dtype=np.uint8).astype(np.int8).astype(np.int64)       # `char` is signed: >> is arithmetic
    sgn = (~(a >> 7)) & 1
    if not alaw:
        chord = (~(a >> 4)) & 7
        step = (~a) & 0xF
        mag = (((2 * step) + 33) << chord) - 33
    else:
        x = a ^ 0x55
        chord = (x >> 4) & 7
        step = x & 0xF
        mag = (step << 1) + 1
        mag = np.where(chord > 0, mag + 32, mag)
        chord = np.where(chord > 0, chord, 1)
        mag = mag << chord
    out = ((1 - 2 * sgn) * mag) & 0xFFFF
    out = (out << 2) & 0xFFFF
    out = np.where(out & 0x8000, out - 65536, out)
    return out.astype(np.int16)


def front_end(pcm: np.ndarray, o: Opts, rand_offset: int = 0) -> FrontEnd:
    """rawIN::get_frame, src/io/in.cc:305-419.  Dither draws from glibc rand() in list order
    (src/io/in.cc:205,454): rand_offset places this file in the process-wide stream."""
    w, s, nfft = o.window, o.wshift, o.wfft
    x = np.asarray(pcm, dtype=np.float64)
    T = num_frames(len(x), o)
    if o.dither != 0.0:
        nloaded = (w - s) + T * s                      # samples the reference actually loads (one rand() each)
        x = x.copy()
        x[:nloaded] += dither_noise(nloaded, o, rand_offset)
    W = hamming(w)
    alpha = float(np.float32(o.preem))
    Xabs = np.empty((T, o.wfftby2))
    spec = np.empty((T, o.wfftby2), dtype=np.complex128)
    Xph = np.empty((T, o.wfftby2)) if o.phase_needed else None
    E = np.empty(T) if o.fea_E else None
    ring = None
    preemtmp = 0.0
    if o.remove_dc1:
        ring = np.zeros(w)
        ring[: w - s] = x[: w - s]
    for t in range(T):
        if o.remove_dc1:
            # src/io/in.cc:343-350: the ring itself is de-meaned every frame (cumulative)
            pos = (t * s + w - s) % w
            idx = (pos + np.arange(s)) % w
            ring[idx] = x[t * s + w - s: t * s + w]
            ring -= _seq_sum(ring) / w
            fr = ring[((t * s) % w + np.arange(w)) % w].copy()
        else:
            fr = x[t * s: t * s + w]
        if o.fea_E and o.fea_rawenergy:
            E[t] = math.log(_seq_sum(fr[1:] * fr[1:]))  # src/io/in.cc:353-361 (skips sample 0)
        if alpha > 0.0:
            y = np.empty(w)
            y[0] = W[0] * (fr[0] - alpha * preemtmp)
            y[1:] = W[1:] * (fr[1:] - alpha * fr[:-1])
        else:
            y = W * fr
        if o.remove_dc:
            y = y - _seq_sum(y) / w
        preemtmp = fr[s - 1]
        buf = np.zeros(nfft)
        buf[:w] = y
        F = np.fft.rfft(buf)
        spec[t] = F
        P = F.real * F.real + F.imag * F.imag
        if o.remove_dc:
            P[0] = 1e-10
        if o.phase_needed:
            ph = c_ph(F.real, F.imag)
            ph[0] = 0.0
            ph[-1] = 0.0 if F.real[-1] >= 0 else 3.14159265358979
            Xph[t] = ph
        if o.fea_E and not o.fea_rawenergy:
            E[t] = math.log(_seq_sum(P[1:-1], P[0] / 2.0 + P[-1] / 2.0) * 2.0)  # src/io/in.cc:403-413
        if not o.fb_power:
            P = np.sqrt(P)
        Xabs[t] = P
    return FrontEnd(Xabs, Xph, E, spec)


def _seq_sum(v: np.ndarray, init: float = 0.0) -> float:
    """Left-to-right double accumulation exactly as the reference's scalar loops do
    (np.sum is pairwise and may differ in the last bit)."""
    acc = float(init)
    for a in v.tolist():
        acc += a
    return acc


# --------------------------------------------------------------------------------------
# filter bank  (src/fea/fb.cc)
# --------------------------------------------------------------------------------------


@dataclass
class FilterBank:
    mat: np.ndarray      # [nb, wfftby2] float64 weights (eq-loudness folded in)
    lo: np.ndarray       # first tap of each band   (mat[i][wfftby2+1], src/fea/fb.cc:439-447)
    hi: np.ndarray       # last tap of the first contiguous non-zero run
    inld: bool           # apply ^0.33 after projection (src/fea/fb.cc:81-83)
    size: int = field(init=False)

    def __post_init__(self):
        self.size = self.mat.shape[0]


def _warp(scale: str, hz):
    """FB::init_scale / get_filter scale formulas, src/fea/fb.cc:100-132, 311-338.
    Uses the C library's log/log10/pow through `math` (numpy's vectorised versions can
    differ from glibc in the last bit, which the triangle-edge differences amplify)."""
    scalar = np.ndim(hz) == 0
    hzv = np.atleast_1d(np.asarray(hz, dtype=np.float64))
    out = np.empty_like(hzv)
    for i, f in enumerate(hzv.tolist()):
        if scale == "lin":
            out[i] = f
        elif scale == "bark":
            out[i] = 6.0 * math.log(f / 600.0 + math.sqrt((f / 600.0) * (f / 600.0) + 1.0))
        elif scale == "expolog":
            out[i] = 700.0 * (math.pow(10.0, f / 3988.0) - 1.0) if f <= 2000 else 2595.0 * math.log10(1.0 + f / 700.0)
        elif scale == "mel":
            out[i] = 2595 * math.log10(1.0 + f / 700.0)
        else:
            raise ValueError("FB: Unknown frequency scale!")
    return float(out[0]) if scalar else out


def _eqloud(om: float, fs: int) -> float:
    """src/fea/fb.cc:157-164, 375-383."""
    eqnum = om * om * om * om * (om * om + 5.68e7)
    if fs <= 10000:
        eqden = (om * om + 6.3e6) * (om * om + 6.3e6) * (om * om + 3.8e8)
    else:
        eqden = (om * om + 6.3e6) * (om * om + 6.3e6) * (om * om + 3.8e8) * (om * om * om * om * om * om + 9.58e26)
    return eqnum / eqden


def parse_fb_definition(defn: str, fs: int):
    """FB::parse, src/fea/fb.cc:186-253: tokens [[X-YHz:]K-L/]Nfilters separated by ','."""
    import re
    banks = []
    for tok in defn.split(","):
        if tok == "":
            continue
        m = re.fullmatch(r"([0-9.]+)-([0-9.]+)Hz:([0-9]+)-([0-9]+)/([0-9]+)filters.*", tok)
        if m:
            banks.append((float(m.group(1)), float(m.group(2)), int(m.group(5)), int(m.group(3)), int(m.group(4))))
            continue
        m = re.fullmatch(r"([0-9.]+)-([0-9.]+)/([0-9]+)filters.*", tok)
        if m:
            banks.append((0.0, fs / 2.0, int(m.group(3)), int(float(m.group(1))), int(float(m.group(2)))))
            continue
        m = re.fullmatch(r"([0-9.]+)filters.*", tok)
        if m:
            n = int(float(m.group(1)))
            banks.append((0.0, fs / 2.0, n, 1, n))
            continue
        raise ValueError("FB: Filter bank specification parse error!")
    return banks  # (f_start, f_stop, bands, band_first, band_last)


def fb_design(o: Opts) -> FilterBank:
    """FB::FB + init_scale/plp_design/parse/check_and_design/get_filter/optimize,
    src/fea/fb.cc:20-66, 100-457."""
    nb2 = o.wfftby2
    scale, eqld, inld = o.fb_scale, o.fb_eqld, o.fb_inld
    plp = o.fb_shape == "trapez"
    if plp:  # src/fea/fb.cc:44-54
        scale, eqld, inld = "bark", True, True
    hz = np.arange(nb2, dtype=np.float64) * o.fs / float(o.wfft)
    warp = _warp(scale, hz)
    rows = []
    if plp:  # plp_design, src/fea/fb.cc:134-184
        maxBark = 6 * math.log(o.fs / 1200.0 + math.sqrt((o.fs / 1200.0) * (o.fs / 1200.0) + 1.0))
        nBark = int(math.floor(maxBark + 0.5))
        step = maxBark / float(nBark)
        for i in range(nBark - 1):
            Om = (i + 1) * step
            om = 3.1415926535898 * 1200 * math.sinh(Om / 6)
            eq = _eqloud(om, o.fs)
            v = np.zeros(nb2)
            for k in range(nb2):
                d = warp[k] - Om
                if -1.3 <= d <= -0.5:
                    v[k] = math.pow(10.0, 2.5 * (0.5 + d))
                elif abs(d) < 0.5:
                    v[k] = 1
                elif 0.5 <= d <= 2.5:
                    v[k] = math.pow(10.0, 0.5 - d)
                else:
                    v[k] = 0
                if eqld:
                    v[k] *= eq
            rows.append(v)
    else:
        banks = [list(b) for b in parse_fb_definition(o.fb_definition, o.fs)]
        if o.fb_shape == "rect":  # src/fea/fb.cc:257-281
            df = o.fs / float(o.wfft)
            joins = [any(bi[1] == bj[0] for bj in banks) for bi in banks]
            for bi, j in zip(banks, joins):
                if not j:
                    bi[1] += df
        for (f_start, f_stop, bands, first, last) in banks:
            w_high = float(_warp(scale, f_stop))
            w_low = float(_warp(scale, f_start))
            for b in range(first, last + 1):
                if o.fb_shape == "rect":
                    w_start = w_low + (b - 1.0) * (w_high - w_low) / float(bands)
                    w_end = w_low + (b + 0.0) * (w_high - w_low) / float(bands)
                elif o.fb_shape == "triang":
                    w_start = w_low + (b - 1.0) * (w_high - w_low) / float(bands + 1)
                    w_end = w_low + (b + 1.0) * (w_high - w_low) / float(bands + 1)
                else:
                    raise ValueError("FB: Unknown filter shape!")
                eq = 1.0
                if eqld:  # src/fea/fb.cc:352-384
                    w_mid = w_start + (w_end - w_start) / 2.0
                    f_mid = 0.0
                    if scale == "lin": f_mid = w_mid
                    if scale == "bark": f_mid = 600 * math.sinh(w_mid / 6.0)
                    if scale == "expolog":
                        f_mid = 3988.0 * math.log10(1.0 + (w_mid / 700.0)) if w_mid <= 1521.4 \
                            else 700.0 * (math.pow(10.0, w_mid / 2595.0) - 1)
                    if scale == "mel": f_mid = 700.0 * (math.pow(10.0, w_mid / 2595.0) - 1.0)
                    eq = _eqloud(2 * 3.141592653589793 * f_mid, o.fs)
                v = np.zeros(nb2)
                area = 0.0
                if o.fb_shape == "rect":
                    for i in range(nb2):
                        if w_start <= warp[i] < w_end:
                            v[i] = 1
                            area += 1
                else:
                    w_mid = w_start + (w_end - w_start) / 2.0
                    for i in range(nb2):
                        if not (warp[i] < w_start or warp[i] > w_end):
                            v[i] = 1.0 - 2.0 * abs(w_mid - warp[i]) / (w_end - w_start)
                            area += v[i]
                with np.errstate(divide="ignore", invalid="ignore"):
                    v = v * (eq / area) if o.fb_norm else v * eq
                rows.append(v)
    mat = np.array(rows).reshape(len(rows), nb2)
    lo = np.zeros(len(rows), dtype=np.int32)
    hi = np.zeros(len(rows), dtype=np.int32)
    for b in range(len(rows)):  # optimize(), src/fea/fb.cc:432-447
        ext = np.concatenate([mat[b], [0.0]])
        k = 0
        while ext[k] == 0:
            k += 1
        lo[b] = k
        while ext[k] != 0:
            k += 1
        hi[b] = k - 1
    return FilterBank(mat, lo, hi, inld)


def fb_project(X: np.ndarray, fb: FilterBank) -> np.ndarray:
    """FB::project_frame, src/fea/fb.cc:72-86 (sequential accumulation order kept)."""
    T = X.shape[0]
    Y = np.zeros((T, fb.size))
    for b in range(fb.size):
        acc = np.zeros(T)
        for k in range(int(fb.lo[b]), int(fb.hi[b]) + 1):
            acc = acc + X[:, k] * fb.mat[b, k]
        Y[:, b] = acc
    if fb.inld:
        with np.errstate(invalid="ignore"):
            Y = np.power(Y, 0.33)
    return Y


# --------------------------------------------------------------------------------------
# features  (src/fea/fea_impl.cc, fea_trap.cc, fea_delta.cc)
# --------------------------------------------------------------------------------------


def _lifter(o: Opts, n_out: int) -> np.ndarray:
    n = np.arange(n_out - 1, dtype=np.float64)
    return 1 + (float(o.fea_lifter)) / 2 * np.sin(3.141592653589793 * (n + 1.0) / float(o.fea_lifter)) \
        if o.fea_lifter != 0 else np.ones(n_out - 1)


def fea_dctc(Y: np.ndarray, o: Opts) -> np.ndarray:
    """dctcFEA, src/fea/fea_impl.cc:81-131. Returns [T, ncep+1] with c0 FIRST."""
    T, Nin = Y.shape
    Nout = o.fea_ncepcoefs + 1
    wdct = np.cos(3.1415926535898 * np.arange(4 * Nin, dtype=np.float64) / (2 * Nin))
    with np.errstate(divide="ignore", invalid="ignore"):
        L = np.log(Y)
    c = np.zeros((T, Nout))
    norm = math.sqrt(2.0 / Nin)
    with np.errstate(invalid="ignore"):
        for i in range(Nout):
            acc = np.zeros(T)
            for k in range(1, Nin + 1):
                acc = acc + L[:, k - 1] * wdct[((2 * k - 1) * i) % (4 * Nin)]
            c[:, i] = acc * norm
        if o.fea_lifter > 1:
            c[:, 1:] *= _lifter(o, Nout)
    return c


def fea_lpa(Y: np.ndarray, o: Opts, inld: bool):
    """lpaFEA::process_frame/idft/LevDurb, src/fea/fea_impl.cc:163-222.
    Returns (a [T, order+1], P [T, order+1], R0 [T])."""
    X = Y if inld else Y * Y
    T, Nin = X.shape
    p = o.fea_lporder
    Nf = (Nin - 1) * 2
    WRe = np.cos(2 * 3.141592653589793 * np.arange(Nf, dtype=np.float64) / Nf)
    R = np.zeros((T, p + 1))
    for k in range(p + 1):
        acc = X[:, 0] / 2.0
        for n in range(1, Nin - 1):
            acc = acc + X[:, n] * WRe[(n * k) % Nf]
        acc = acc + (1 - 2 * (k % 2)) * X[:, Nin - 1] / 2.0
        R[:, k] = acc / (float(Nf) / 2)
    a = np.zeros((T, p + 1)); aa = np.zeros((T, p + 1)); P = np.zeros((T, p + 1))
    with np.errstate(divide="ignore", invalid="ignore"):
        P[:, 0] = R[:, 0]
        rc = -R[:, 1] / R[:, 0]
        P[:, 1] = P[:, 0] * (1 - rc * rc)
        aa[:, 1] = rc
        a[:, 1] = rc  # (a[1] is only assigned inside the loop for order>=2; equal for order 1)
        a[:, 0] = aa[:, 0] = 1
        for ik in range(2, p + 1):
            dm = R[:, ik].copy()
            for n in range(1, ik):
                dm = dm + aa[:, n] * R[:, ik - n]
            rc = -dm / P[:, ik - 1]
            a[:, ik] = rc
            for n in range(1, ik):
                a[:, n] = aa[:, n] + rc * aa[:, ik - n]
            aa[:, 1:ik + 1] = a[:, 1:ik + 1]
            P[:, ik] = P[:, ik - 1] * (1 - rc * rc)
    return a, P, R[:, 0]


def fea_lpc(Y: np.ndarray, o: Opts, inld: bool) -> np.ndarray:
    """lpcFEA::process_frame/a2c, src/fea/fea_impl.cc:251-284. [T, ncep+1], c0 first."""
    a, P, _ = fea_lpa(Y, o, inld)
    T = Y.shape[0]
    p = o.fea_lporder
    Nout = o.fea_ncepcoefs + 1
    c = np.zeros((T, Nout))
    with np.errstate(divide="ignore", invalid="ignore"):
        c[:, 0] = np.log(P[:, p])
        for n in range(1, Nout):
            s = np.zeros(T)
            if n <= p:
                for k in range(1, n):
                    s = s + (n - k) * c[:, n - k] * a[:, k]
                c[:, n] = -a[:, n] - s / n
            else:
                for k in range(1, p + 1):
                    s = s + (n - k) * c[:, n - k] * a[:, k]
                c[:, n] = -s / n
        if o.fea_lifter > 1:
            c[:, 1:] *= _lifter(o, Nout)
    return c


def fea_trapdct(Y: np.ndarray, o: Opts) -> np.ndarray:
    """trapdctFEA, src/fea/fea_trap.cc:20-127 (+ new_file in fea_trap.h:43).
    log, ring of traplen frames, first frame replicated left, last replicated right by
    flush; per band: mean removal, Hamming (pi=3.14159265359), REDFT10, keep 1..ndct.
    Output [T, nb*ndct] band-major.  NB the reference never resets `avail` between files
    and returns nothing for utterances shorter than (traplen+1)/2 frames; the oracle
    covers T >= htraplen (standalone file)."""
    L, n = o.fea_trapdct_traplen, o.fea_trapdct_ndct
    if L % 2 == 0:
        raise ValueError("FEA: TRAP length must be odd!")
    if n >= L:
        raise ValueError("FEA: Number of DCT coeffs must be less than TRAP length (c0 is not output)!")
    T, nb = Y.shape
    h = (L + 1) // 2
    with np.errstate(divide="ignore", invalid="ignore"):
        G = np.log(Y)
    if T == 0:
        return np.zeros((0, nb * n))
    hamm = 0.54 - (1 - 0.54) * np.cos(2 * 3.14159265359 * np.arange(L, dtype=np.float64) / (L - 1.0))
    j = np.arange(L, dtype=np.float64)
    out = np.zeros((T, nb * n))
    cosm = np.array([np.cos(math.pi * (j + 0.5) * k / L) for k in range(1, n + 1)])  # [n, L]
    with np.errstate(invalid="ignore"):
        for t in range(T):
            if T >= h - 1:
                win = G[np.clip(t + np.arange(L) - (h - 1), 0, T - 1)]      # [L, nb]
            else:
                # fewer than h-1 frames: nothing is emitted before flush, and flush r sees
                # Z = h-T-(r+1) never-written ring rows (zeros in a fresh process) ahead of
                # the replicated context (src/fea/fea_trap.cc:53-82, 111-127)
                Z = h - T - (t + 1)
                win = G[np.clip(np.arange(L) - Z - (h - 1), 0, T - 1)].copy()
                win[:Z] = 0.0
            for b in range(nb):
                v = win[:, b]
                m = _seq_sum(v) / L
                u = (v - m) * hamm
                out[t, b * n:(b + 1) * n] = 2.0 * (cosm @ u)
    return out


def fea_delta_block(C: np.ndarray, win: int) -> np.ndarray:
    """Closed form of one deltaFEA stage in its regular regime (rows >= win + 2):
    HTK regression d_t = sum_i i (c_{t+i} - c_{t-i}) / (2 sum i^2) with the first/last row
    replicated (src/fea/fea_delta.cc:146-164 + edge logic :70-130, 178-206).  One quirk
    survives into the regular regime: with win == 1 the flush writes the replica over the
    slot the last real row sits in (`index=end`, src/fea/fea_delta.cc:182-185), so the LAST
    row's delta is exactly 0.  tests/test_oracle_vs_golden.py checks this closed form
    against the state machine below; the CUDA kernel implements the closed form."""
    T = C.shape[0]
    den = 2.0 * sum(i * i for i in range(1, win + 1))
    idx = np.arange(T)
    acc = np.zeros_like(C)
    for i in range(1, win + 1):
        acc = acc + i * (C[np.minimum(idx + i, T - 1)] - C[np.maximum(idx - i, 0)])
    acc = acc / den
    if win == 1 and T > 0:
        acc[-1] = 0.0
    return acc


class DeltaFEA:
    """Faithful port of the deltaFEA state machine, src/fea/fea_delta.cc:20-206 (ring of
    2*win+1 input rows, `avail`/`index`/`start`/`end` bookkeeping, first row written `win`
    times, second row twice, the b(dwlen-1)=b(dwlen-2) patch, flush replicas).  Needed for
    utterances shorter than win+2 rows, where the output depends on never-written ring
    rows (zeros in a fresh process: Mat zero-fills all but column 0, src/base/types.h:72)."""

    def __init__(self, fea_c: int, n_order: int, delta_w: int, trap: bool = False):
        if delta_w < 1:
            raise ValueError("FEA: Trap window size must be >= 3!" if trap else "FEA: Delta window size must be > 1!")
        self.fea_c, self.delta_w, self.dwlen, self.n_order = fea_c, delta_w, 2 * delta_w + 1, n_order
        self.nfea = fea_c * (n_order - 1)
        self.trap = trap
        if trap:                                   # src/fea/fea_delta.cc:40 (after nfea was taken from n_order = 2)
            self.n_order = self.dwlen
        self.buf = np.zeros((self.dwlen, self.nfea))
        self.fvec = np.zeros(fea_c * self.n_order)
        self.den = 2.0 * sum(i * i for i in range(1, delta_w + 1))
        self.new_file()

    def new_file(self):
        self.start = self.end = 0
        self.endframe = True
        self.frame = 1
        self.avail = self.delta_w + 1
        self.index = self.avail - 1
        self.num_c = 1

    def _init_cbuffer(self, f):
        for _ in range(self.num_c):
            self.buf[self.index] = f[: self.nfea]
            self.index = (self.index + 1) % self.dwlen

    def _delta(self):
        d = self.buf[(self.start + np.arange(self.dwlen)) % self.dwlen]
        if self.trap:                              # deltaFEA::trap, src/fea/fea_delta.cc:166-176: fvec[i*dwlen+j] = ring row j, column i
            self.fvec[:] = d[:, : self.fea_c].T.reshape(-1)
            return
        lo, hi = self.fea_c * (self.n_order - 2), self.fea_c * (self.n_order - 1)
        x = np.zeros(hi - lo)
        for i in range(1, self.delta_w + 1):
            x = x + i * (d[self.delta_w + i, lo:hi] - d[self.delta_w - i, lo:hi])
        self.fvec[lo + self.fea_c: hi + self.fea_c] = x / self.den

    def process_frame(self, f) -> bool:
        if self.avail != 0:
            self.num_c = self.delta_w if self.frame == 1 else (2 if self.frame == 2 else 1)
            self._init_cbuffer(f)
            self.avail -= 1
            ready = not self.avail
            if ready:
                self._delta()
                self.buf[self.dwlen - 1] = self.buf[self.dwlen - 2]
                self.fvec[: self.nfea] = self.buf[self.delta_w]
            self.start = (self.start + 1) % self.dwlen
            self.frame += 1
            return ready
        fea_x = self.dwlen - 1 if self.start == 0 else self.start - 1
        self.buf[fea_x] = f[: self.nfea]
        self._delta()
        if not self.trap:                          # src/fea/fea_delta.cc:120-124; NOT guarded in the fill / flush branches
            self.fvec[: self.nfea] = self.buf[(fea_x - self.delta_w) % self.dwlen]
        self.end = self.start
        self.start = (self.start + 1) % self.dwlen
        self.frame += 1
        return True

    def flush_frame(self, f) -> bool:
        if self.endframe:
            self.index = self.end
            self.endframe = False
        if self.avail < self.delta_w:
            self._init_cbuffer(f)
            self._delta()
            self.fvec[: self.nfea] = self.buf[(self.start - self.delta_w - 1) % self.dwlen]
            self.start = (self.start + 1) % self.dwlen
            self.avail += 1
            return True
        return False


def add_deltas(C: np.ndarray, o: Opts) -> np.ndarray:
    """BATCH::init_delta/fea_delta/flush_fea chain, src/io/batch.cc:122-130, 172-192,
    251-296, driving DeltaFEA stages exactly as the reference does: blocks [c, d, dd, ddd];
    each higher order is the same operator applied to the previous stage's rows."""
    wins = [o.d_win, o.a_win, o.t_win][: o.n_order]
    n = len(wins)
    fea_c = C.shape[1]
    st = [DeltaFEA(fea_c, k + 2, wins[k], trap=o.fea_trap) for k in range(n)]
    out = []

    def push(level, vec):
        if level == n:
            out.append(np.array(vec, copy=True))
        elif st[level].process_frame(vec):
            push(level + 1, st[level].fvec)

    for t in range(C.shape[0]):
        push(0, C[t])
    for level in range(n):
        src = (C[-1] if C.shape[0] else np.zeros(fea_c)) if level == 0 else st[level - 1].fvec
        while st[level].flush_frame(src):
            push(level + 1, st[level].fvec)
    return np.array(out).reshape(len(out), fea_c * (st[-1].n_order if o.fea_trap else n + 1))


def trap_stack_closed_form(C: np.ndarray, win: int) -> np.ndarray:
    """What the CUDA path computes for `-fea_trap 2*win+1` (deltaFEA in trap mode, T >= win + 2 rows):
    out[t, i*L + j] = C[clip(t + j - win, 0, T-1), i], L = 2 win + 1, plus the reference's edge behaviour:
      * the ring is primed with  row0 x win, row1 x 2, row2, ...  (src/fea/fea_delta.cc:73-77), so the FIRST output row
        stacks rows [0]*win + [1, 1, 2, .., win] instead of [0]*(win+1) + [1, .., win];
      * with win == 1 the flush replica lands on the slot of the last real row (`index=end`, :182-185): the LAST output
        row stacks row T-1 three times;
      * in the first row and in the `win` flush rows the first fea_c elements are then overwritten with that row's own
        columns C[t, :] (the copies `ofvec[i] = b(...)` at :88-90 and :197-199 are not guarded by !fea_trap).
    tests/test_oracle_vs_golden.py checks this against the state machine and the reference binary."""
    T, fea_c = C.shape
    L = 2 * win + 1
    out = np.empty((T, fea_c * L))
    for t in range(T):
        idx = np.clip(t + np.arange(L) - win, 0, T - 1)
        if t == 0:
            idx = np.array(([0] * win + [1, 1] + list(range(2, win + 1)))[:L])
        if win == 1 and t == T - 1:
            idx[:] = T - 1
        out[t] = C[idx].T.reshape(-1)
        if t == 0 or t >= T - win:
            out[t, :fea_c] = C[t]
    return out


def add_deltas_closed_form(C: np.ndarray, o: Opts) -> np.ndarray:
    """What the CUDA path computes: valid when every stage sees >= win+2 rows."""
    blocks = [C]
    for k in range(o.n_order):
        blocks.append(fea_delta_block(blocks[-1], [o.d_win, o.a_win, o.t_win][k]))
    return np.concatenate(blocks, axis=1)


# --------------------------------------------------------------------------------------
# Burg cepstral detector  (src/vdet/Burg.h, src/vdet/CepstralDet.h)
# --------------------------------------------------------------------------------------


def burg_cepstrum(x: np.ndarray, ncoefs: int) -> np.ndarray:
    """DSP::Burg::Process + Burg2Cepstrum::Process, src/vdet/Burg.h:49-95, 141-152.
    x: [npoints] -> c[ncoefs] (c[0] = log alpha)."""
    npts = len(x)
    ef = x.astype(np.float64).copy()
    eb = ef.copy()
    alpha = _seq_sum(np.power(ef, 2.0)) / npts
    a = np.zeros(ncoefs); aa = np.zeros(ncoefs)
    a[0] = 1.0
    for ik in range(1, ncoefs):
        e1 = ef[ik:]
        e2 = eb[ik - 1: npts - 1]
        den = 0.0
        num = 0.0
        for u, v in zip(e1.tolist(), e2.tolist()):
            den += u * u + v * v
            num += u * v
        num *= 2.0
        with np.errstate(divide="ignore", invalid="ignore"):
            rc = -np.float64(num) / np.float64(den)
        a[ik] = rc
        alpha *= 1 - rc * rc
        efo, ebo = ef.copy(), eb.copy()
        ef[1:] = efo[1:] + rc * ebo[:-1]
        eb[1:] = ebo[:-1] + rc * efo[1:]
        for i in range(1, ik):
            a[i] = aa[i] + rc * aa[ik - i]
        aa[1:ik + 1] = a[1:ik + 1]
    c = np.zeros(ncoefs)
    for n in range(1, ncoefs):
        s = 0.0
        for k in range(1, n):
            s += (n - k) * c[n - k] * a[k]
        c[n] = -a[n] - s / n
    with np.errstate(divide="ignore", invalid="ignore"):
        c[0] = np.log(np.float64(alpha))
    return c


def burg_time_signal(Xa: np.ndarray, Xph: np.ndarray, o: Opts) -> np.ndarray:
    """The "brutal hack" of hwssNR::vad_get_frame (src/nr/nr.cc:281-292) and
    VADcri_cepdist::process_frame (src/vad/vad.cc:222-233): half-complex from (|X|, phase),
    UNNORMALISED inverse real DFT, first `window` samples."""
    n = o.wfft
    re = Xa * np.cos(Xph)
    im = Xa * np.sin(Xph)
    spec = re + 1j * im
    spec[0] = re[0]
    spec[-1] = re[-1]
    return np.fft.irfft(spec, n)[: o.window] * n


class CepstralDetector:
    """Voice::CepstralDetector<BurgCepstrumEstimator>, src/vdet/CepstralDet.h:85-216."""

    def __init__(self, npoints: int, ninit: int, ncoefs: int, p: float, q: float):
        self.n, self.ninit, self.nc, self.P, self.Q = npoints, ninit, ncoefs, p, q
        m = 2 * 3.141592653 / npoints
        self.han = 0.5 * (1 - np.cos(m * np.arange(npoints, dtype=np.float64)))
        self.c0 = np.zeros(ncoefs)
        self.dMean = self.dMean2 = self.dVar = self.thr = 0.0
        self.nseg = 0
        self.last_ci = None
        self.last_dist = 0.0

    @staticmethod
    def dist(ci, c0) -> float:
        """CepstralDistance::Compute, src/vdet/CepstralDet.h:64-82."""
        d = ci[1:] - c0[1:]
        return 4.3429 * math.sqrt(2 * _seq_sum(d * d))

    def process(self, x: np.ndarray) -> bool:
        ci = burg_cepstrum(self.han * x, self.nc)
        return self.process_cepstrum(ci)

    def process_cepstrum(self, ci: np.ndarray) -> bool:
        self.last_ci = ci
        res = False
        if self.nseg == 0:
            self.c0 = ci.copy()
        elif self.nseg == 1:
            self.c0 = (self.c0 + ci) / 2.0
            d = self.dist(ci, self.c0)
            self.dMean, self.dMean2, self.thr = d, d * d, d
        else:
            d = self.dist(ci, self.c0)
            self.last_dist = d
            res = (self.nseg > self.ninit) and (d >= self.thr)
            if not res:
                self.c0 = self.P * self.c0 + (1 - self.P) * ci
                self.dMean = self.Q * self.dMean + (1 - self.Q) * d
                self.dMean2 = self.Q * self.dMean2 + (1 - self.Q) * d * d
                self.dVar = self.dMean2 - self.dMean * self.dMean
                with np.errstate(invalid="ignore"):
                    self.thr = self.dMean + 2.0 * float(np.sqrt(np.float64(self.dVar)))
        self.nseg += 1
        return res


# --------------------------------------------------------------------------------------
# noise reduction  (src/nr/nr.cc)
# --------------------------------------------------------------------------------------


def compute_E(X: np.ndarray) -> float:
    """_NR::compute_E, src/nr/nr.cc:36-45 (squares X even if it already is power)."""
    return math.log(_seq_sum(X[1:-1] * X[1:-1], X[0] * X[0] / 2.0 + X[-1] * X[-1] / 2.0) * 2.0)


def nr_exten(X: np.ndarray, o: Opts) -> np.ndarray:
    """extenNR::new_file/process_frame, src/nr/nr.cc:86-140.  X: [T, size] -> enhanced."""
    a, p = o.nr_a, o.nr_p
    T, size = X.shape
    Navg = np.full(size, 0.95)
    Yavg = np.full(size, 0.05)
    out = np.empty_like(X)
    with np.errstate(divide="ignore", invalid="ignore"):
        for t in range(T):
            x = X[t]
            if a == 1.0:
                H = Navg / (Navg + Yavg)
            elif a == 2.0:
                H = Navg / np.sqrt(Navg * Navg + Yavg * Yavg)
            else:
                H = Navg / np.power(np.power(Navg, a) + np.power(Yavg, a), 1.0 / a)
            N = H * x
            Navg = p * Navg + (1 - p) * N
            Yavg = np.where(x > Navg, x - Navg, Navg - x)
            out[t] = x - N
    return out


def nr_ss(X: np.ndarray, o: Opts, vad_flags=None, Xph: Optional[np.ndarray] = None,
          navg0: Optional[np.ndarray] = None):
    """hwssNR / fwssNR / dfwssNR, src/nr/nr.cc:212-261, 331-369, 397-442.
    X: [T,size] in, returns (enhanced [T,size], vad [T] bool).
    Per-utterance definition: Navg starts from the buffer contents at
    new_file, i.e. ZEROS for a standalone file (src/nr/nr.cc:212-222; the carry-over from
    a previous file of the same list is a cross-utterance quirk that no sharded run can
    reproduce -- SURVEY.md finding 4).  navg0 lets a test model that carry-over.
    vad_flags: external per-frame bytes (vadmode file) or None for the internal Burg
    detector (needs Xph)."""
    mode, a, b, p = o.nr_mode, o.nr_a, o.nr_b, o.nr_p
    T, size = X.shape
    ninit = o.nr_initsegs
    Navg = np.zeros(size) if navg0 is None else np.power(navg0, a) if mode != "2fwss" else navg0.copy()
    Nravg = np.zeros(size)
    det = None
    if vad_flags is None:
        if o.vadmode != "burg":
            raise ValueError("NR: Please specify Voice Activity Detector!")
        if Xph is None or Xph.shape[1] != size:
            raise ValueError("NR: Cannot use Burg detector after filter bank!")
        det = CepstralDetector(o.window, o.nr_initsegs, o.fea_ncepcoefs, p, o.nr_q)
    out = np.empty_like(X)
    vout = np.zeros(T, dtype=bool)
    with np.errstate(divide="ignore", invalid="ignore"):
        for t in range(T):
            x = X[t].copy()
            if mode == "hwss":
                ninit -= 1                      # src/nr/nr.cc:226 (decrement BEFORE use)
            if mode in ("hwss", "fwss"):
                if a == 2.0: x = x * x
                elif a != 1.0: x = np.power(x, a)
            if det is not None:
                vad = det.process(burg_time_signal(x, Xph[t], o))
            else:
                vad = bool(vad_flags[t])
            vout[t] = vad
            upd = (not vad) or ninit > 0
            if mode == "2fwss":
                if upd: Navg = p * Navg + (1 - p) * x
                x = x - Navg
                x = np.where(x < 0.0, -x, x)
                if upd: Nravg = p * Nravg + (1 - p) * x
                x = x - Nravg
                x = np.where(x < 0.0, -x, x)
            else:
                if upd: Navg = p * Navg + (1 - p) * x
                x = x - b * Navg
                if mode == "hwss": x = np.where(x < 0.0, 0.0, x)
                else: x = np.where(x < 0.0, -x, x)
                if a == 2.0: x = np.sqrt(x)
                elif a != 1.0: x = np.power(x, 1.0 / a)
            if mode != "hwss":
                ninit -= 1                      # src/nr/nr.cc:367, 440 (decrement AFTER)
            out[t] = x
    return out, vout


def apply_nr(X, o: Opts, Xph=None, vad_flags=None, force_flags=None, navg0=None):
    """force_flags: detector decisions to use instead of running the Burg detector (the sensitivity probe of
    run_pipeline(perturb=...) keeps the decisions of the unperturbed run)."""
    if o.nr_mode == "none":
        return X, None
    if force_flags is not None and o.nr_mode in ("hwss", "fwss", "2fwss"):
        return nr_ss(X, o, vad_flags=force_flags, navg0=navg0)
    if o.nr_mode == "exten":
        return nr_exten(X, o), None
    if o.nr_mode in ("hwss", "fwss", "2fwss"):
        if o.vadmode == "file":
            if vad_flags is None:
                raise ValueError("NR: Unable to open VAD file!\n")
            return nr_ss(X, o, vad_flags=vad_flags, navg0=navg0)
        return nr_ss(X, o, Xph=Xph, navg0=navg0)
    raise ValueError("NR: Unknown noise reduction mode!")


# --------------------------------------------------------------------------------------
# VAD module  (src/vad/vad.cc, src/vad/vad.h)
# --------------------------------------------------------------------------------------


def _DB(x: float) -> float:
    """src/vad/vad.cc:54-57."""
    v = np.finfo(np.float64).tiny + x
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(10.0 * np.log10(np.float64(v)))


@dataclass
class VadResult:
    vad0: np.ndarray        # unfiltered decisions per input frame           [T]
    vad: np.ndarray         # median-filtered decisions, one per OUTPUT frame [T]
    cri: np.ndarray         # criterion per input frame                      [T]
    thr: np.ndarray         # threshold per input frame                      [T]
    keep: np.ndarray        # rows written to the feature file (drop mode)   [T] bool
    debug: Optional[dict] = None   # -vad_out_mode debug: side file suffix -> array (float64 values / uint8 '0','1' chars)


def vad_module(o: Opts, Xs: np.ndarray, Xph: Optional[np.ndarray], feat: np.ndarray) -> VadResult:
    """VAD::process_frame chain (src/vad/vad.cc:692-745) + medianFilter
    (src/vad/vad.h:79-176) driven as BATCH::save_frame / flush_vad do
    (src/io/batch.cc:230-249).
    Xs: in->_Xsabs per INPUT frame as it stands when BATCH::save_frame runs (after NR when
    NR is before the FB).  feat: the final feature rows [T, dim] (the same vector is passed
    as InFvec and FeaFvec, src/io/batch.cc:75).  When the feature chain has latency
    (deltas, trapdct) the reference pairs output row r with spectrum frame r+latency; the
    caller passes Xs already aligned that way (see run_pipeline)."""
    T = feat.shape[0]
    cri = np.zeros(T); thr = np.zeros(T); vad0 = np.zeros(T, dtype=bool)
    st = {}
    c0 = None
    rec = {k: np.zeros(T) for k in ("a", "b", "c", "d")}     # threshold internals per step (debug side files)
    for t in range(T):
        # ---- criterion
        if o.vad_cri_mode == "energy":                      # src/vad/vad.cc:96-107
            e = _seq_sum(Xs[t] * Xs[t])
            c = _DB(e) if o.vad_energy_db else e
            ci = None
        elif o.vad_cri_mode == "cepdist":                   # src/vad/vad.cc:220-276
            if o.vad_cepdist_mode == "lpc":
                if Xph is None:
                    raise ValueError("VADcri_cepdist: cannot perform iFFT!")
                ci = burg_cepstrum(burg_time_signal(Xs[t], Xph[t], o), o.vad_lpc_coefs)
            elif o.vad_cepdist_mode in ("fea", "in"):
                ci = feat[t].copy()
            else:
                raise ValueError("VADcri_cepdist: unknown vad_cepdist_mode!")
            if t == 0:
                c0 = ci.copy(); c = 0.0
            else:
                if t == 1:
                    c0 = (c0 + ci) / 2.0
                d = ci[1:] - c0[1:]
                with np.errstate(invalid="ignore"):
                    c = 4.3429 * float(np.sqrt(np.float64(2 * _seq_sum(d * d))))
        else:
            raise ValueError("VAD: unknown vad_cri_mode!")
        cri[t] = c
        # ---- threshold
        m = o.vad_thr_mode
        if m == "absolute":                                  # src/vad/vad.cc:329-331
            th = o.vad_absolute_thr; v = c >= th
        elif m == "perc":                                    # src/vad/vad.cc:384-398
            if t == 0 or t < float(o.vad_perc_init):
                st["min"] = st["max"] = c
            else:
                st["min"] = c if c < st["min"] else st["min"]
                st["max"] = c if c > st["max"] else st["max"]
            th = st["min"] + (o.vad_perc_thr / 100.0) * (st["max"] - st["min"])
            v = c >= th
        elif m == "adapt":                                   # src/vad/vad.cc:469-495
            if t == 0:
                th = c; st.update(mean=c, mean2=c * c, var=0.0); v = False
            else:
                with np.errstate(invalid="ignore"):
                    th = st["mean"] + o.vad_adapt_za * float(np.sqrt(np.float64(st["var"])))
                if (c < th) or (t <= o.vad_adapt_init):
                    q = o.vad_adapt_q
                    st["mean"] = q * st["mean"] + (1.0 - q) * c
                    st["mean2"] = q * st["mean2"] + (1.0 - q) * c * c
                    st["var"] = st["mean2"] - st["mean"] * st["mean"]
                    v = False
                else:
                    v = True
        elif m == "dyn":                                     # src/vad/vad.cc:578-625
            i0 = max(1, o.vad_dyn_init)
            if t < i0:
                st.update(dmax=c, dmin=c, dyn=0.0); th = c; v = False
            elif t == i0:
                st["dmax"] = max(st["dmax"], c) + o.vad_dyn_min / 10.0
                st["dmin"] = min(st["dmin"], c) - o.vad_dyn_min / 10.0
                st["dyn"] = st["dmax"] - st["dmin"]; th = c; v = False
            else:
                if st["dmax"] < c: st["dmax"] = o.vad_dyn_qmaxinc * st["dmax"] + (1.0 - o.vad_dyn_qmaxinc) * c
                else: st["dmax"] = o.vad_dyn_qmaxdec * st["dmax"] + (1.0 - o.vad_dyn_qmaxdec) * c
                if st["dmin"] > c: st["dmin"] = o.vad_dyn_qmindec * st["dmin"] + (1.0 - o.vad_dyn_qmindec) * c
                else: st["dmin"] = o.vad_dyn_qmininc * st["dmin"] + (1.0 - o.vad_dyn_qmininc) * c
                st["dyn"] = st["dmax"] - st["dmin"]
                th = st["dmin"] + (o.vad_dyn_perc / 100.0) * st["dyn"]
                v = (c > th) and (st["dyn"] > o.vad_dyn_min)
        else:
            raise ValueError("VAD: unknown vad_thr_mode!")
        thr[t] = th; vad0[t] = v
        if m == "perc":
            rec["a"][t], rec["b"][t] = st["min"], st["max"]
        elif m == "adapt":
            rec["a"][t], rec["b"][t], rec["c"][t] = st["mean"], st["mean2"], st["var"]
        elif m == "dyn":
            rec["a"][t], rec["b"][t], rec["c"][t] = st["dmin"], st["dmax"], st["dyn"]
        # ---- consume_vad (cepdist background update, src/vad/vad.cc:289-294)
        if o.vad_cri_mode == "cepdist" and not (v and t > o.vad_cepdist_init):
            c0 = o.vad_cepdist_p * c0 + (1.0 - o.vad_cepdist_p) * ci
    # ---- median filter with flush (zeros pushed at the end), src/vad/vad.h:126-175
    order = o.vad_filter_order
    if order < 1 or order % 2 == 0:
        raise ValueError("medianFilter: filter order must be positive, odd number!")
    h = (order - 1) // 2
    ext = np.concatenate([vad0.astype(np.float64), np.zeros(h)])
    vad = np.zeros(T, dtype=bool)
    for r in range(T):
        # output r is emitted when input r+h is pushed; history = the last `order` pushes
        # (zeros before the start)
        lo = r + h - order + 1
        seg = ext[max(lo, 0): r + h + 1]
        vad[r] = (seg.sum() / float(order)) >= 0.5
    keep = vad | (o.vad_apply_mode != "drop")
    debug = None
    if o.vad_out_mode == "debug" and T > h:
        # VAD::save_frame (src/vad/vad.cc:710-725) runs once per written row, after process_frame has advanced the state:
        # row i carries the filtered decision of frame i but the criterion / threshold state of step i+h, and the h rows
        # written by flush_vad (src/io/batch.cc:243-249) repeat the state of the last step.  The *init flags compare the
        # already incremented frame index (src/vad/vad.cc:703, 281-284, 498-501).
        step = np.minimum(np.arange(T) + h, T - 1)
        ch = lambda b: np.where(b, ord("1"), ord("0")).astype(np.uint8)
        debug = {"vad0": ch(vad0[step])}
        if o.vad_cri_mode == "energy":
            debug["energy"] = cri[step].copy()
        else:
            debug["cepdist"] = cri[step].copy()
            debug["c0init"] = ch((step + 1) <= o.vad_cepdist_init)
        m = o.vad_thr_mode
        if m == "absolute":
            debug["thr"] = np.full(T, float(o.vad_absolute_thr))
        elif m == "perc":
            debug.update(crimin=rec["a"][step].copy(), crimax=rec["b"][step].copy(), thr=thr[step].copy())
        elif m == "adapt":
            debug.update(init=ch((step + 1) <= o.vad_adapt_init), crimean=rec["a"][step].copy(), crimean2=rec["b"][step].copy(),
                         crivar=rec["c"][step].copy(), thr=thr[step].copy())
        else:
            debug.update(dmin=rec["a"][step].copy(), dmax=rec["b"][step].copy(), dyn=rec["c"][step].copy(),
                         dynmin=np.full(T, float(o.vad_dyn_min)), thr=thr[step].copy())
    return VadResult(vad0, vad, cri, thr, keep, debug)


# --------------------------------------------------------------------------------------
# signal synthesis  (src/io/out.cc:346-451)
# --------------------------------------------------------------------------------------


def ola_correction(o: Opts) -> float:
    """sigOUT::sigOUT, src/io/out.cc:346-372."""
    s, w = o.wshift, o.window
    pi = 2.0 * math.asin(1.0)
    corr = 0.0
    for i in range(s):
        x, y = i, 0.0
        while x < w:
            y += 0.54 - (1 - 0.54) * math.cos(2 * pi * float(x) / (w - 1.0))
            x += s
        if y > corr:
            corr = y
    return corr


def synth(Xa: np.ndarray, Xph: np.ndarray, o: Opts) -> np.ndarray:
    """sigOUT::save_frame + fill_cache + close, src/io/out.cc:405-451, 483-487:
    (|X|, phase) -> half-complex / nfft -> HC2R -> overlap-add (no synthesis window) ->
    floor(x / correction) -> clip +-32767 -> int16; T*s + (w-s) samples."""
    T = Xa.shape[0]
    w, s, n = o.window, o.wshift, o.wfft
    corr = ola_correction(o)
    acc = np.zeros(T * s + w)
    if o.fb_power:
        Xa = np.sqrt(Xa)
    for t in range(T):
        amp = Xa[t] / float(n)
        spec = amp * np.cos(Xph[t]) + 1j * (amp * np.sin(Xph[t]))
        spec[0] = amp[0]
        spec[-1] = amp[-1]          # Nyquist always non-negative (src/io/out.cc:419)
        y = np.fft.irfft(spec, n) * n
        acc[t * s: t * s + w] += y[:w]
    nout = T * s + (w - s) if T > 0 else (w - s)
    v = np.floor(acc[:nout] / corr)
    v = np.where(np.abs(v.astype(np.float32)) > 32767, np.where(v < 0, -32767, 32767), v)
    return v.astype(np.int16)


# --------------------------------------------------------------------------------------
# whole pipeline  (BATCH ctor / process_frame / flush_fea, src/io/batch.cc)
# --------------------------------------------------------------------------------------


@dataclass
class Result:
    nframes: int
    features: Optional[np.ndarray] = None   # float32 [rows, dim], writer column order
    waveform: Optional[np.ndarray] = None   # int16
    vad_nr: Optional[np.ndarray] = None     # NR-internal detector decisions
    vad: Optional[VadResult] = None         # VAD-module results
    fb_out: Optional[np.ndarray] = None     # [T, nb] float64 (post NR if afterFB)
    spectrum: Optional[np.ndarray] = None   # [T, bins] float64 after NR (beforeFB)
    internal: Optional[np.ndarray] = None   # feature matrix before column reorder
    navg0: Optional[np.ndarray] = None      # run_list_carry: the buffer this file's noise estimate started from


def writer_order(F: np.ndarray, o: Opts) -> np.ndarray:
    """htkOUT::save_frame column order, src/io/out.cc:183-202: spec/logspec/trapdct as is;
    lpc/dctc blocks are written c1..cN then c0.  (fea_E / fea_c0 off are handled by the
    caller: E appended last, see out_dim.)"""
    if o.fea_kind in ("spec", "logspec", "trapdct") or o.fea_trap:   # -fea_trap: kind forced to "spec", src/io/out.cc:182
        return F
    n = o.fea_ncepcoefs + 1
    cols = []
    for j in range(o.n_order + 1):
        cols += list(range(n * j + 1, n * (j + 1)))
        if o.fea_c0:
            cols.append(n * j)
    return F[:, cols]


def cms(F: np.ndarray, o: Opts) -> np.ndarray:
    """cms_POST::process_frame, exponential version (src/fea/post_impl.cc:203-209), on the finished
    feature vector (after the deltas, src/io/batch.cc:159-163, 198-199): only the first ncep+1 elements
    (internal order: c0..cN) are normalised.  `sumM` is float and so is the coefficient
    (src/fea/post_impl.h, src/io/opts.h): sumM = sumM*Z + F*(1-Z); F -= sumM.
    The block version (-fea_Z_block) is not restated: the reference binary dies with SIGSEGV on it (its
    ring of row pointers is allocated with sizeof(float) per pointer, src/fea/post_impl.cc:179)."""
    if o.fea_Z_block > 0:
        raise ValueError("CTU: -fea_Z_block: the reference crashes in this mode (src/fea/post_impl.cc:179); nothing to match")
    F = F.copy()
    n = o.fea_ncepcoefs + 1
    f32 = np.float32
    Z = f32(o.cms_exp_coef)
    omZ = f32(1) - Z
    sumM = np.zeros(n, dtype=f32)
    for t in range(F.shape[0]):
        sumM = (np.float64(sumM * Z) + F[t, :n] * np.float64(omZ)).astype(f32)
        F[t, :n] -= sumM.astype(np.float64)
    return F


def cmvn_stat_order(F: np.ndarray) -> np.ndarray:
    """Statistics are kept as F[1], ..., F[size-1], F[0] of the INTERNAL vector -- c0 of the static block
    moves to the very end, everything else (deltas included) keeps its place (src/fea/post_impl.cc:60-64)."""
    return np.concatenate([F[:, 1:], F[:, :1]], axis=1)


def cmvn_stats(feats: List[np.ndarray], spk: List[str]):
    """cmvn_POST::sum_fea / stat_cm / sum_cv / stat_cv (src/fea/post_impl.cc:52-104) driven as BATCH::process
    does for -stat_cmvn (src/io/batch.cc:339-420): pass 0 sums every frame of a speaker's files (list order),
    mean = sum / count; pass 1 sums squared deviations, var = sum / (count - 1).  feats: internal-order
    feature matrices (after deltas), one per list line.  Returns (speaker names in order of first
    appearance, mean [n_spk, size], var [n_spk, size]) in the statistics order of cmvn_stat_order."""
    names: List[str] = []
    for s_ in spk:
        if s_ not in names:
            names.append(s_)
    size = feats[0].shape[1]
    mean = np.zeros((len(names), size)); var = np.zeros((len(names), size)); cnt = np.zeros(len(names))
    for F, s_ in zip(feats, spk):
        j = names.index(s_)
        G = cmvn_stat_order(F)
        for t in range(G.shape[0]):
            mean[j] += G[t]
        cnt[j] += G.shape[0]
    mean /= cnt[:, None]
    for F, s_ in zip(feats, spk):
        j = names.index(s_)
        G = cmvn_stat_order(F)
        for t in range(G.shape[0]):
            d = G[t] - mean[j]
            var[j] += d * d
    var /= (cnt[:, None] - 1)
    return names, mean, var


def cmvn_stat_text(names: List[str], mean: np.ndarray, var: np.ndarray) -> str:
    """cmvnOUT::save_frame, src/io/out.cc:591-615: "<id>\nmean\t%f %f ... %f\nvar\t%f ... %f\n" per speaker."""
    out = ""
    for j, n in enumerate(names):
        out += "%s\nmean\t%s\nvar\t%s\n" % (n, " ".join("%f" % v for v in mean[j]), " ".join("%f" % v for v in var[j]))
    return out


def cmvn_apply(F: np.ndarray, mean: np.ndarray, var: np.ndarray) -> np.ndarray:
    """cmvn_POST::process_frame, src/fea/post_impl.cc:106-118: (F - mean) / var -- the VARIANCE, not its root."""
    G = (cmvn_stat_order(F) - mean) / var
    return np.concatenate([G[:, -1:], G[:, :-1]], axis=1)


def run_list_cmvn(pcms: List[np.ndarray], spk: List[str], o: Opts):
    """A whole list through -stat_cmvn (statistics only) or -apply_cmvn with a statistics file that does not
    exist yet (three passes: statistics, then normalised features; src/io/batch.cc:136-152, 339-420).
    Returns (statistics text, per-utterance float32 feature matrices in writer order or None)."""
    feats = []
    for pcm in pcms:
        fe = front_end(pcm, o)
        fb = fb_design(o)
        if o.nr_when == "afterFB":
            Y, _ = apply_nr(fb_project(fe.Xabs, fb), o, None, None)
        else:
            Xs, _ = apply_nr(fe.Xabs, o, fe.Xph, None)
            Y = fb_project(Xs, fb)
        k = o.fea_kind
        if k == "spec": F = Y.copy()
        elif k == "logspec": F = np.log(Y)
        elif k == "dctc": F = fea_dctc(Y, o)
        elif k == "lpc": F = fea_lpc(Y, o, fb.inld)
        else: raise ValueError("CTU: CMVN oracle covers spec, logspec, dctc and lpc")
        if o.fea_delta and o.n_order > 0:
            if F.shape[1] < o.fea_ncepcoefs + 1:
                raise ValueError("FEA: deltas / stacking read beyond the feature vector for this kind")
            F = add_deltas(F[:, : o.fea_ncepcoefs + 1], o)
        feats.append(F)
    names, mean, var = cmvn_stats(feats, spk)
    text = cmvn_stat_text(names, mean, var)
    if not o.apply_cmvn:
        return text, None
    outs = []
    for F, s_ in zip(feats, spk):
        j = names.index(s_)
        G = cmvn_apply(F, mean[j], var[j])
        outs.append((writer_order(G, o)).astype(np.float32))
    return text, outs


def run_list_cmvn_features(mats: List[np.ndarray], spk: List[str], o: Opts):
    """run_list_cmvn for `-format_in htk`: the statistics cover the whole vector the chain produced, in the order it
    comes -- with feature-file input POST neither skips nor rotates element 0 (the `format_in == "htk"` branches of
    cmvn_POST::sum_fea / sum_cv / process_frame, src/fea/post_impl.cc:56-58, 83-85, 111-113)."""
    feats = []
    for M in mats:
        F = np.asarray(M, dtype=np.float64)
        if o.fea_delta and o.n_order > 0:
            F = add_deltas(F[:, : o.fea_ncepcoefs + 1], o)
        feats.append(np.concatenate([F[:, -1:], F[:, :-1]], axis=1))      # undo cmvn_stat_order's rotation: identity order
    names, mean, var = cmvn_stats(feats, spk)
    text = cmvn_stat_text(names, mean, var)
    if not o.apply_cmvn:
        return text, None
    outs = []
    for F, s_ in zip(feats, spk):
        j = names.index(s_)
        G = cmvn_apply(F, mean[j], var[j])
        outs.append(np.concatenate([G[:, 1:], G[:, :1]], axis=1).astype(np.float32))
    return text, outs


def energy_column(o: Opts, fe: "FrontEnd", Xs: np.ndarray, Y: np.ndarray, kind: str, inld: bool, lat: int) -> np.ndarray:
    """The optional _E column.  Which stage's E the writer points at is decided in
    BATCH::init_out (src/io/batch.cc:98-118): raw energy -> IN (src/io/in.cc:353-361); dctc ->
    NR::compute_E on the spectrum handed to the filter bank (src/nr/nr.cc:36-45; it squares X
    again even when X is already power) or, with NR after the FB, IN's power-domain energy
    (src/io/in.cc:403-413); lpc/lpa -> log R0 (src/fea/fea_impl.cc:177); spec/logspec -> the
    same half-spectrum formula applied to the BAND values (src/fea/fea_impl.cc:45-50, 69-74).
    htkOUT::save_frame reads *E when a (possibly delayed) row is written (src/io/out.cc:183-202),
    so row r carries the energy of input frame min(r + latency, T-1)."""
    T = Xs.shape[0]

    def half_spectrum_energy(X):
        with np.errstate(divide="ignore", invalid="ignore"):
            return np.array([math.log(_seq_sum(X[t, 1:-1] * X[t, 1:-1], X[t, 0] * X[t, 0] / 2.0 + X[t, -1] * X[t, -1] / 2.0) * 2.0)
                             if True else 0.0 for t in range(X.shape[0])])

    do_vad = o.vad_apply_mode != "none" or o.vad_out_mode != "none"
    if do_vad and o.fea_rawenergy and not (kind == "dctc" and o.nr_when == "afterFB") and kind not in ("lpa", "lpc"):
        # with the VAD module BATCH::init_out never looks at fea_rawenergy (src/io/batch.cc:74-96): it points at
        # nr->E / fea->E, which those stages leave unset when raw energy is asked for
        raise ValueError("CTU: -fea_rawenergy with the VAD module: the reference writes an energy that was never computed")
    if o.fea_rawenergy and not (do_vad and kind in ("lpa", "lpc")):
        E = fe.E
    elif kind == "dctc":
        E = fe.E if o.nr_when == "afterFB" else half_spectrum_energy(Xs)
    elif kind in ("lpa", "lpc"):
        X = Y if inld else Y * Y
        Nin = X.shape[1]; Nf = (Nin - 1) * 2
        R0 = np.array([(_seq_sum(X[t, 1:Nin - 1], X[t, 0] / 2.0) + X[t, Nin - 1] / 2.0) / (float(Nf) / 2) for t in range(T)])
        with np.errstate(divide="ignore", invalid="ignore"):
            E = np.log(R0)
    elif kind in ("spec", "logspec"):
        E = half_spectrum_energy(Y)
    else:
        raise ValueError("CTU: -fea_E with -fea_kind %s: the reference never sets that energy" % kind)
    idx = np.minimum(np.arange(T) + lat, T - 1)
    return np.asarray(E, dtype=np.float64)[idx]


def load_iir_filters(path: str) -> np.ndarray:
    """rawIN::loadf_filters, src/io/in.cc:242-262: one filter per line, ten TAB-separated numbers
    (b0 b1 b2 b3 b4 | input gain | a1 a2 a3 a4); the first 24 lines are the bank."""
    rows = []
    with open(path) as fh:
        for line in fh:
            tok = line.rstrip("\n").split("\t")
            if len(tok) < 10:
                raise ValueError("IN: filter line with fewer than 10 coefficients (the reference hands strtok's NULL to atof)")
            rows.append([_atof(t) for t in tok[:10]])
            if len(rows) == 24:
                break
    if len(rows) < 24:
        raise ValueError("IN: fewer than 24 filters (the reference would use coefficient rows it never set)")
    return np.array(rows, dtype=np.float64)


def _atof(t: str) -> float:
    """C atof: the longest numeric prefix, 0.0 when there is none"""
    import re
    m = re.match(r"\s*[-+]?(\d+\.?\d*([eE][-+]?\d+)?|\.\d+([eE][-+]?\d+)?|inf(inity)?|nan)", t, re.I)
    return float(m.group(0)) if m else 0.0


def td_iir_mfcc(pcm: np.ndarray, o: Opts, coefs: np.ndarray) -> np.ndarray:
    """-fea_kind td-iir-mfcc: 24 fourth-order IIR band filters in the time domain -> windowed band
    energies per frame -> log -> 13-point DCT.  rawIN::compute_td_iir_mfcc src/io/in.cc:281-303, the td-iir branch of
    rawIN::get_frame :317-340, tables :234-239; no pre-emphasis, no dither, no DC removal on this branch.
    The filter state (second canonical form, Mat(24,5)) is zero at the first sample: the reference leaves column 0 of
    its Mat unset (src/base/types.h:72) and reads it once, at the very first sample of a PROCESS; a fresh heap gives 0
    there (the default run is byte-identical to one under MALLOC_PERTURB_=255, which zero-fills).  The state is
    also carried from file to file of a list; like the *ss modes this path is defined per utterance (= the first file of
    a list)."""
    w, s = o.window, o.wshift
    x = np.asarray(pcm, dtype=np.float64)
    n_first = w - s
    if len(x) < n_first:
        raise ValueError("IO: Signal shorter than one frame!")
    T = (len(x) - n_first) // s
    N = n_first + T * s
    W = hamming(w)
    weight = float(o.weight_of_td_iir_mfcc_bank)
    wdct = np.cos(3.14159265358979 * np.arange(4 * 24, dtype=np.float64) / (2 * 24))     # src/io/in.cc:236-237
    normcoef = math.sqrt(2.0 / 24)
    # filtered, windowed samples: the window weight goes by the sample's position in the circular buffer, n % window
    # (src/io/in.cc:291-299) -- not by its position inside a frame
    widx = np.arange(N) % w
    # the 24 filters side by side (same operations in the same order per filter as the reference's scalar loops)
    c = np.ascontiguousarray(coefs.T)                 # c[j] = coefficient j of every filter
    s0 = np.zeros(24); s1 = np.zeros(24); s2 = np.zeros(24); s3 = np.zeros(24)
    y = np.empty((N, 24), dtype=np.float64)
    for n, xn in enumerate(x[:N].tolist()):
        v = c[5] * xn
        v = v - c[6] * s3
        v = v - c[7] * s2
        v = v - c[8] * s1
        v = v - c[9] * s0
        acc = c[0] * v
        acc = acc + c[4] * s0
        acc = acc + c[3] * s1
        acc = acc + c[2] * s2
        acc = acc + c[1] * s3
        s0, s1, s2, s3 = s1, s2, s3, v
        y[n] = acc
    z = (y * W[widx][:, None]).T
    out = np.empty((T, 13), dtype=np.float64)
    for t in range(T):
        # frame t holds samples t*s .. t*s + w - 1; the reference sums them in buffer order i = n % w (src/io/in.cc:322-324)
        seg = z[:, t * s: t * s + w]
        order = np.argsort((np.arange(t * s, t * s + w)) % w, kind="stable")
        E = np.zeros(24)
        for ff in range(24):
            e = _seq_sum(seg[ff, order] * seg[ff, order])
            with np.errstate(divide="ignore"):
                E[23 - ff] = np.log((w * w * e / w) / weight)
        for i in range(13):
            acc = 0.0
            for kk in range(1, 25):
                acc += E[kk - 1] * wdct[((2 * kk - 1) * i) % (4 * 24)]
            out[t, i] = acc * normcoef
    return out


def run_pipeline(pcm: np.ndarray, o: Opts, ext_vad: Optional[np.ndarray] = None, rand_offset: int = 0,
                 perturb: Optional[Tuple[float, int]] = None, force_vad_nr: Optional[np.ndarray] = None,
                 navg0: Optional[np.ndarray] = None) -> Result:
    """One utterance through the chain BATCH builds (src/io/batch.cc:24-69, 205-296).
    perturb = (eps, seed): CONDITIONING PROBE, not part of the reference's algorithm -- the spectrum that leaves the front
    end is moved by eps (relative per bin, plus eps of the frame's mean level, random signs) before anything else sees it,
    with the noise-reduction detector's decisions held at force_vad_nr.  The difference to the unperturbed output says how
    far the result of THIS configuration on THIS input moves when its input spectrum moves by one rounding error of the
    front end: where a spectral subtraction cancels many digits that is far more than the rounding error itself, and no
    implementation that computes the spectrum in the stated precision can agree better (tools/parity_sweep.py).  With
    noise reduction after the filter bank the band vector -- the subtraction's input there -- is what is moved."""
    if o.fea_kind == "td-iir-mfcc" and o.format_out in ("htk", "pfile", "ark"):
        # BATCH::BATCH src/io/batch.cc:61-62, process_frame :221-222: IN's vector goes straight to the writer
        if o.fea_ncepcoefs != 12:
            raise ValueError("CTU: td-iir-mfcc writes 13 coefficients into a vector of fea_ncepcoefs+1 (src/io/in.cc:168-170, 328)")
        if not o.fea_c0 or o.fea_E or (o.fea_delta and o.n_order > 0) or o.stat_cmvn or o.apply_cmvn or \
                o.vad_apply_mode != "none" or o.vad_out_mode != "none":
            raise ValueError("CTU: td-iir-mfcc with -fea_c0 off / -fea_E / deltas / CMVN / the VAD module: the reference reads unset memory or null objects")
        F = td_iir_mfcc(pcm, o, load_iir_filters(o.ffilters))
        out = np.concatenate([F[:, 1:], F[:, :1]], axis=1)           # writer: c1..c12, c0 (src/io/out.cc:189-197)
        return Result(F.shape[0], features=out.astype(np.float32))
    fe = front_end(pcm, o, rand_offset)

    def _perturbed(A):
        eps, seed = perturb
        rng = np.random.default_rng(seed)
        s1 = rng.choice([-1.0, 1.0], size=A.shape)
        s2 = rng.choice([-1.0, 1.0], size=A.shape)
        return np.abs(A * (1.0 + eps * s1) + eps * np.mean(A, axis=1, keepdims=True) * s2)

    if perturb is not None and not (o.nr_when == "afterFB" and o.format_out not in ("raw", "wave")):
        fe.Xabs = _perturbed(fe.Xabs)
    T = fe.Xabs.shape[0]
    signal_out = o.format_out in ("raw", "wave")
    if signal_out:
        X, vnr = apply_nr(fe.Xabs, o, fe.Xph, ext_vad, force_vad_nr, navg0)
        return Result(T, waveform=synth(X, fe.Xph, o), vad_nr=vnr, spectrum=X)
    fb = fb_design(o)
    if o.nr_when == "afterFB":
        Y = fb_project(fe.Xabs, fb)
        if perturb is not None:
            Y = _perturbed(Y)                       # the subtraction's input is the band vector here
        if o.nr_mode in ("hwss", "fwss", "2fwss") and o.vadmode == "burg":
            raise ValueError("NR: Cannot use Burg detector after filter bank!")
        Y, vnr = apply_nr(Y, o, None, ext_vad, force_vad_nr, navg0)
        Xs = fe.Xabs
    else:
        Xs, vnr = apply_nr(fe.Xabs, o, fe.Xph, ext_vad, force_vad_nr, navg0)
        Y = fb_project(Xs, fb)
    k = o.fea_kind
    latency = 0
    if k == "spec": F = Y.copy()
    elif k == "logspec":
        with np.errstate(divide="ignore", invalid="ignore"):
            F = np.log(Y)
    elif k == "dctc": F = fea_dctc(Y, o)
    elif k == "lpa": F = fea_lpa(Y, o, fb.inld)[0]
    elif k == "lpc": F = fea_lpc(Y, o, fb.inld)
    elif k == "trapdct":
        F = fea_trapdct(Y, o)
    else:
        raise ValueError("FEA: Unknown feature kind!")
    internal = F
    if o.fea_delta and o.n_order > 0:
        # deltaFEA works on the first fea_ncepcoefs+1 elements of whatever vector FEA produced (src/fea/fea_delta.cc:23,
        # 31) and the writer then sees a vector of (ncep+1)*(n_order+1) [or *(2 win + 1) when stacking] elements
        fea_c = o.fea_ncepcoefs + 1
        if k == "lpa" or F.shape[1] < fea_c:
            raise ValueError("FEA: deltas / stacking read beyond the feature vector for this kind")
        if k == "trapdct":
            raise ValueError("CTU: trapdct with deltas: the reference never flushes the TRAP ring (src/io/batch.cc:253), the last rows are lost")
        if o.fea_trap and k in ("dctc", "lpc") and not o.fea_c0:
            raise ValueError("CTU: -fea_trap with -fea_c0 off: the reference writes past its output buffer (src/io/out.cc:184)")
        F = add_deltas(F[:, :fea_c], o)
    if o.cms_exp_coef > 0 or o.fea_Z_block > 0:
        F = cms(F, o)
    out = writer_order(F, o) if k != "lpa" else F[:, 1:]
    lat = sum([o.d_win, o.a_win, o.t_win][: o.n_order]) if o.fea_delta else 0
    if k == "trapdct":
        lat += (o.fea_trapdct_traplen + 1) // 2 - 1
    do_vad = o.vad_apply_mode != "none" or o.vad_out_mode != "none"
    if o.fea_E:
        # rows also wait in the VAD module's majority filter, (order-1)/2 frames (src/vad/vad.h:126-175)
        elat = lat + ((o.vad_filter_order - 1) // 2 if do_vad else 0)
        out = np.concatenate([out, energy_column(o, fe, Xs, Y, k, fb.inld, elat)[:, None]], axis=1)
    res = Result(T, features=out.astype(np.float32), vad_nr=vnr, fb_out=Y, spectrum=Xs, internal=internal)
    if o.vad_apply_mode != "none" or o.vad_out_mode != "none":
        # BATCH::save_frame (src/io/batch.cc:230-241) runs when a feature row leaves the
        # delta / TRAP-DCT delay lines, so the VAD criterion sees in->_Xsabs of the frame
        # that is `latency` frames AHEAD of the row (and the last frame during flush).
        idx = np.minimum(np.arange(F.shape[0]) + lat, T - 1)
        res.vad = vad_module(o, Xs[idx], fe.Xph[idx] if fe.Xph is not None else None, F)
        res.features = res.features[res.vad.keep]
    return res


def run_list_carry(pcms: List[np.ndarray], o: Opts, ext_vads: Optional[List[np.ndarray]] = None) -> List[Result]:
    """The files of ONE reference process in list order, for the VAD-driven subtraction modes: hwssNR / dfwssNR::new_file
    (src/nr/nr.cc:212-222, 397-408) start a file's noise estimate from whatever the shared spectrum buffer holds -- the
    ENHANCED last frame of the file before (zeros for the first file: Vec's constructor, src/base/types.h:35-38).  With
    noise reduction after the filter bank the buffer is the band vector.  run_pipeline alone is the per-utterance definition
    (every file = the first file of a process)."""
    out, last = [], None
    for k, u in enumerate(pcms):
        r = run_pipeline(u, o, ext_vads[k] if ext_vads is not None else None, navg0=last)
        r.navg0 = last
        out.append(r)
        if o.nr_mode in ("hwss", "fwss", "2fwss"):
            buf = r.fb_out if (o.nr_when == "afterFB" and o.format_out not in ("raw", "wave")) else r.spectrum
            if buf is not None and len(buf):
                last = np.array(buf[-1], dtype=np.float64)
                if o.nr_when == "afterFB" and o.format_out not in ("raw", "wave"):
                    # FEA works IN PLACE on the band vector: dctcFEA takes its logarithm (src/fea/fea_impl.cc:108), lpaFEA
                    # squares it when the cube-root law is off (:166-169)
                    with np.errstate(divide="ignore", invalid="ignore"):
                        if o.fea_kind == "dctc": last = np.log(last)
                        elif o.fea_kind in ("lpa", "lpc") and not fb_design(o).inld: last = last * last
                if o.format_out in ("raw", "wave") and front_end(u, o).Xph[-1][-1] != 0:
                    last[-1] = -last[-1]           # sigOUT::save_frame negates the Nyquist bin IN the shared buffer (src/io/out.cc:414)
    return out


def run_features(Fin: np.ndarray, o: Opts) -> np.ndarray:
    """`-format_in htk`: an existing feature file through deltas / stacking / CMS (BATCH::BATCH src/io/batch.cc:55-60,
    process_frame :217-218, htkIN src/io/in.cc:623-690).  The reader's vector holds -nfeacoefs elements; deltaFEA
    takes its first fea_ncepcoefs+1; the writer copies the result in the order it comes (src/io/out.cc:177-179) and
    cuts it to get_fea_size() elements (src/io/out.cc:95-112)."""
    if o.fea_E or o.vad_apply_mode != "none" or o.vad_out_mode != "none":
        raise ValueError("CTU: no spectrum / energy exists for feature-file input")
    if o.fea_kind == "dctc" and not o.fea_rawenergy:
        raise ValueError("CTU: the reference dereferences a null NR here (src/io/batch.cc:108)")
    F = np.asarray(Fin, dtype=np.float64)
    if F.shape[1] > o.nfeacoefs:
        raise ValueError("IN: feature file wider than -nfeacoefs")
    if o.fea_delta and o.n_order > 0:
        fea_c = o.fea_ncepcoefs + 1
        if F.shape[1] < fea_c:
            raise ValueError("IN: feature file narrower than fea_ncepcoefs+1")
        F = add_deltas(F[:, :fea_c], o)
    elif F.shape[1] != o.nfeacoefs:
        raise ValueError("IN: feature file width differs from -nfeacoefs")
    if (o.cms_exp_coef > 0 or o.fea_Z_block > 0) and o.fea_delta and o.n_order > 0:
        F = cms(F, o)                                # the plain-copy branch never calls POST (src/io/batch.cc:223-227)
    size = F.shape[1]
    if o.fea_kind == "lpa" or (o.fea_kind in ("lpc", "dctc") and not o.fea_c0):
        size -= 1
    return F[:, :size].astype(np.float32)


# --------------------------------------------------------------------------------------
# file formats  (src/io/out.cc, src/io/pfile.cc) -- byte-exact writers used by the tests
# --------------------------------------------------------------------------------------


def htk_parmkind(o: Opts, file_index: int = 0) -> int:
    """htkOUT::new_file, src/io/out.cc:145-158.  With -fea_trap (raw input) the first written row renames the kind to
    "spec" (src/io/out.cc:182), so every file after the first of a process carries base kind 8."""
    kind = {"lpc": 11, "dctc": 6, "trapdct": 9, "spec": 8, "logspec": 7}.get(o.fea_kind, 9)
    if o.fea_trap and file_index > 0 and o.format_in != "htk":
        kind = 8
    c0 = o.fea_c0 and o.fea_kind not in ("lpa", "spec", "logspec")
    if c0: kind |= 0o20000
    if o.fea_E: kind |= 0o100
    if o.fea_delta and o.n_order >= 1: kind |= 0o400
    if o.fea_delta and o.n_order >= 2: kind |= 0o1000
    if o.fea_delta and o.n_order == 3: kind = (kind | 100000) & 0xFFFF   # DECIMAL 100000 in the reference (src/io/out.cc:158), cut to 16 bits
    return kind


def write_htk(feat: np.ndarray, o: Opts) -> bytes:
    """htkOUT, src/io/out.cc:128-213."""
    e = "<" if o.endian_out == "little" else ">"
    period = int(math.floor(0.5 + 10000000.0 * o.wshift / float(o.fs)))
    hdr = struct.pack(e + "IIHH", feat.shape[0], period, 4 * feat.shape[1], htk_parmkind(o))
    return hdr + feat.astype(e + "f4").tobytes()


def write_ark(items: List[Tuple[str, np.ndarray]], arkname: str) -> Tuple[bytes, str]:
    """arkOUT, src/io/out.cc:680-781: returns (ark bytes, scp text)."""
    ark = b""
    scp = ""
    for key, feat in items:
        head = key.encode() + b" \x00BFM \x04"
        off = len(ark) + len(head) + 4 + 1 + 4 - 15
        scp += "%s %s:%d\n" % (key, arkname, off)
        ark += head + struct.pack("<i", feat.shape[0]) + b"\x04" + struct.pack("<i", feat.shape[1])
        ark += feat.astype("<f4").tobytes()
    return ark, scp


def write_pfile(sents: List[np.ndarray]) -> bytes:
    """pfileOUT + PFile, src/io/out.cc:222-313, src/io/pfile.cc:435-468, 470-592."""
    nfea = sents[0].shape[1]
    nframes = sum(s.shape[0] for s in sents)
    ncol = nfea + 2
    data = b""
    table = [0]
    for sid, s in enumerate(sents):
        T = s.shape[0]
        rows = np.zeros((T, ncol), dtype=">u4")
        rows[:, 0] = sid
        rows[:, 1] = np.arange(T)
        rows[:, 2:] = s.astype(">f4").view(">u4")
        data += rows.tobytes()
        table.append(table[-1] + T)
    hdr = "-pfile_header version 0 size 32768\n"
    hdr += "-num_sentences %d\n-num_frames %d\n" % (len(sents), nframes)
    hdr += "-first_feature_column 2\n-num_features %d\n" % nfea
    hdr += "-first_label_column %d\n-num_labels 0\n" % (2 + nfea)
    hdr += "-format dd" + "f" * nfea + "\n"
    hdr += "-data size %d offset 0 ndim 2 nrow %d ncol %d\n" % (ncol * nframes, nframes, ncol)
    hdr += "-sent_table_data size %d offset %d ndim 1\n" % (len(sents) + 1, ncol * nframes)
    hdr += "-end\n"
    hb = hdr.encode()
    hb += b"\x00" * (32768 - len(hb))
    return hb + data + np.array(table, dtype=">u4").tobytes()


def write_wave(pcm: np.ndarray, fs: int) -> bytes:
    """waveOUT, src/io/out.cc:519-557."""
    data = pcm.astype("<i2").tobytes()
    return (b"RIFF" + struct.pack("<I", len(data) + 36) + b"WAVEfmt " +
            struct.pack("<ihhiihh", 16, 1, 1, fs, fs * 2, 2, 16) + b"data" + struct.pack("<I", len(data)) + data)
