"""steady-state cost of one ctu_run call (plan + H2D + kernels + D2H + destroy) on a batch of 800 utterances"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, ctypes as C
import ctucopy_b200 as cb
from ctucopy_b200 import synthetic
args = ["-fs", "16000", "-format_in", "raw", "-dither", "0", "-preset", "mfcc", "-preem", "0.97", "-nr_mode", "exten", "-format_out", "htk", "-fea_delta", "d_a"]
hd = cb.Handle(args)
pcm, lens = synthetic.batch(800, 10.0, unique=16)
hp = torch.empty(len(pcm), dtype=torch.int16).pin_memory(); hp.numpy()[:] = pcm
off = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
T = 800 * 998
fea = torch.empty((T, hd.feature_dim), dtype=torch.float32).pin_memory()
fr = np.zeros(800, np.int64); rows = np.zeros(800, np.int64)
L = hd.L
def run():
    st = L.ctu_run(hd.h, hp.data_ptr(), off.ctypes.data, 800, None, fea.data_ptr(), T, None, 0, None, None, fr.ctypes.data, rows.ctypes.data)
    assert st == 0, L.ctu_last_error(hd.h)
for i in range(6):
    t0 = time.perf_counter(); run(); print("ctu_run %.1f ms" % (1e3 * (time.perf_counter() - t0)))
p = hd.plan(lens)
for i in range(4):
    t0 = time.perf_counter(); p.run_host(hp.numpy(), features=fea.numpy(), want_vad=False); print("plan.run_host %.1f ms" % (1e3 * (time.perf_counter() - t0)))
t0 = time.perf_counter(); p2 = hd.plan(lens); print("plan create %.1f ms" % (1e3 * (time.perf_counter() - t0)))
t0 = time.perf_counter(); p2.close(); print("plan destroy %.1f ms" % (1e3 * (time.perf_counter() - t0)))
