#!/usr/bin/env python
"""Writes tests/golden/tdiir_filters.asc: the coefficient file `-fea_kind td-iir-mfcc` needs (`-filters <file>`,
rawIN::loadf_filters, src/io/in.cc:242-262).  The reference repository names one (egs/conf/20_td-iir-mfcc.ctuconf:
conf/filterbank_coefs.asc) but does not ship it, so the goldens use this one: 24 fourth-order Butterworth band-pass
filters (scipy.signal.butter, order 2) whose edges are 26 points equally spaced on the mel scale between 64 Hz and
0.4875 fs, normalised frequencies (the same file serves 16 kHz and 8 kHz runs).  One filter per line, TAB separated:
    b0 b1 b2 b3 b4  g  a1 a2 a3 a4        (g = input gain 1/a0; second canonical form, src/io/in.cc:286-296)
"""
import os

import numpy as np
import scipy.signal as ss


def mel(f):
    return 2595.0 * np.log10(1.0 + f / 700.0)


def imel(m):
    return 700.0 * (10.0 ** (m / 2595.0) - 1.0)


def main():
    fs = 16000.0
    edges = imel(np.linspace(mel(64.0), mel(0.4875 * fs), 26))
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "tdiir_filters.asc"), "w") as fh:
        for k in range(24):
            b, a = ss.butter(2, [edges[k] / (fs / 2), edges[k + 2] / (fs / 2)], btype="band")
            row = list(b) + [1.0 / a[0]] + list(a[1:])
            fh.write("\t".join("%.17g" % v for v in row) + "\n")


if __name__ == "__main__":
    main()
