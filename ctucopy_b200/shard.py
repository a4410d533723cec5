"""Utterance sharding across GPUs and the ordered merge of per-rank containers.

The path has no exchange step: every rank processes a contiguous range of
list lines with its own handle and the only post-step is putting per-rank outputs back in
input order.  HTK / raw / wave outputs are one file per utterance (nothing to merge);
Kaldi ark+scp and ICSI pfile are single containers and are merged here byte-exactly, i.e.
the merged file equals what one process would have written for the whole list
(src/io/out.cc:680-700, 720-754; src/io/pfile.cc:435-468, 470-539, 573-592).
host/ctucopy_main.cc implements the same three functions in C++ (-shard / -gpus / -merge).
"""
from __future__ import annotations

import struct
from typing import List, Sequence, Tuple

import numpy as np

PFILE_HEADER = 32768


def partition(samples: Sequence[int], n: int) -> List[int]:
    """Cut points (n+1 of them) of contiguous list ranges balanced by cumulative sample count:
    rank r takes lines [cut[r], cut[r+1])."""
    cut = [len(samples)] * (n + 1)
    cut[0] = 0
    total = float(sum(samples))
    acc = 0.0
    r = 1
    for i, v in enumerate(samples):
        if r >= n:
            break
        acc += float(v)
        while r < n and acc >= total * r / n:
            cut[r] = i + 1
            r += 1
    return cut


def rank_range(samples: Sequence[int], rank: int, world: int) -> Tuple[int, int]:
    cut = partition(samples, world)
    return cut[rank], cut[rank + 1]


def split_ext_vad(ext_vad: bytes, frames_per_utt: Sequence[int], cut: Sequence[int]) -> List[bytes]:
    """One global VAD byte file (one byte per frame for the whole list, src/nr/nr.cc:297-302)
    split at the same utterance boundaries as the list."""
    off = np.concatenate([[0], np.cumsum(frames_per_utt)]).astype(np.int64)
    return [ext_vad[int(off[cut[r]]): int(off[cut[r + 1]])] for r in range(len(cut) - 1)]


def merge_ark(arks: Sequence[bytes], scps: Sequence[str]) -> Tuple[bytes, str]:
    """Byte-concatenate per-rank arks; every scp offset moves by the bytes before its shard."""
    out, lines, base = [], [], 0
    for a, s in zip(arks, scps):
        for ln in s.splitlines():
            if ":" not in ln:
                continue
            head, off = ln.rsplit(":", 1)
            lines.append("%s:%d" % (head, int(off) + base))
        out.append(a)
        base += len(a)
    return b"".join(out), "".join(l + "\n" for l in lines)


def pfile_header(nsent: int, frames: int, dim: int) -> bytes:
    ncol = dim + 2
    h = ("-pfile_header version 0 size 32768\n-num_sentences %d\n-num_frames %d\n-first_feature_column 2\n-num_features %d\n"
         "-first_label_column %d\n-num_labels 0\n-format dd%s\n-data size %d offset 0 ndim 2 nrow %d ncol %d\n"
         "-sent_table_data size %d offset %d ndim 1\n-end\n") % (nsent, frames, dim, dim + 2, "f" * dim, ncol * frames, frames, ncol,
                                                               nsent + 1, ncol * frames)
    b = h.encode()
    return b + b"\x00" * (PFILE_HEADER - len(b))


def _pfile_fields(b: bytes):
    hdr = b[:PFILE_HEADER].split(b"\x00", 1)[0].decode()
    f = {ln.split()[0]: ln.split()[1:] for ln in hdr.splitlines() if ln.strip()}
    return int(f["-num_sentences"][0]), int(f["-num_frames"][0]), int(f["-num_features"][0])


def merge_pfile(shards: Sequence[bytes]) -> bytes:
    """Rows concatenated in rank order with the sentence id (big-endian u32, first column)
    shifted by the sentences before the shard; cumulative sentence table and ASCII header
    rebuilt."""
    rows_out, table, frames, dim = [], [0], 0, None
    for b in shards:
        ns, nf, d = _pfile_fields(b)
        if dim is not None and d != dim:
            raise ValueError("pfile shards differ in feature count")
        dim = d
        ncol = d + 2
        rows = np.frombuffer(b[PFILE_HEADER: PFILE_HEADER + 4 * ncol * nf], dtype=">u4").reshape(nf, ncol).copy()
        st = np.frombuffer(b[PFILE_HEADER + 4 * ncol * nf: PFILE_HEADER + 4 * (ncol * nf + ns + 1)], dtype=">u4")
        rows[:, 0] += len(table) - 1
        rows_out.append(rows.tobytes())
        table += [frames + int(v) for v in st[1:]]
        frames += nf
    tab = struct.pack(">%dI" % len(table), *table)
    return pfile_header(len(table) - 1, frames, dim or 0) + b"".join(rows_out) + tab
