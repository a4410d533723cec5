#!/usr/bin/env python
"""File-to-file throughput of the command-line host against the reference binary on the same
list (both reading and writing /dev/shm).   usage: python tools/cli_e2e.py [n_utts]"""
import os
import shutil
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from ctucopy_b200 import synthetic  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
d = "/dev/shm/ctu_cli"
shutil.rmtree(d, ignore_errors=True)
os.makedirs(d + "/in"); os.makedirs(d + "/out"); os.makedirs(d + "/ref")
uniq = [synthetic.utterance(k, 10.0).astype("<i2").tobytes() for k in range(16)]
for i in range(n):
    open("%s/in/u%05d.raw" % (d, i), "wb").write(uniq[i % 16])
args = ["-fs", "16000", "-format_in", "raw", "-dither", "0", "-preset", "mfcc", "-preem", "0.97", "-nr_mode", "exten", "-format_out", "htk", "-fea_delta", "d_a"]
with open(d + "/list.scp", "w") as fh:
    for i in range(n):
        fh.write("%s/in/u%05d.raw %s/out/u%05d.htk\n" % (d, i, d, i))
frames = n * 998
exe = os.path.join(ROOT, "host", "ctucopy_b200")
for rep in range(2):
    t0 = time.perf_counter()
    subprocess.check_call([exe] + args + ["-S", d + "/list.scp"], env=dict(os.environ, CTU_TIMING="1"))
    dt = time.perf_counter() - t0
    print("ctucopy_b200 CLI: %d files, %.2f s, %.3g frames/s (run %d)" % (n, dt, frames / dt, rep))
ref = os.path.join(ROOT, "oracle", "_ref", "ctucopy4_O2")
cores = len(os.sched_getaffinity(0))
m = min(n, 64 * cores)
procs = []
t0 = time.perf_counter()
for p in range(cores):
    lst = "%s/l%d.scp" % (d, p)
    with open(lst, "w") as fh:
        for i in range(p, m, cores):
            fh.write("%s/in/u%05d.raw %s/ref/u%05d.htk\n" % (d, i, d, i))
    procs.append(subprocess.Popen([ref] + args + ["-S", lst]))
for p in procs:
    p.wait()
dt = time.perf_counter() - t0
print("reference binary x%d processes: %d files, %.2f s, %.3g frames/s" % (cores, m, dt, m * 998 / dt))
a = open("%s/out/u00003.htk" % d, "rb").read(); b = open("%s/ref/u00003.htk" % d, "rb").read()
print("same size:", len(a) == len(b), "same header:", a[:12] == b[:12])
shutil.rmtree(d, ignore_errors=True)
