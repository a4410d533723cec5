"""End-to-end (host buffers) rate of the headline pipeline with 16-bit samples against 8-bit G.711 codes expanded on the device."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden"))
import numpy as np, torch
import ctucopy_b200 as cb
from ctucopy_b200 import synthetic

N = 10000
pcm, lens = synthetic.batch(N, 10.0, unique=16)
table = cb.g711_table(True).astype(np.int64)
order = np.argsort(table, kind="stable"); tv = table[order]
j = np.clip(np.searchsorted(tv, pcm.astype(np.int64)), 1, 255)
codes = order[np.where(np.abs(tv[j - 1] - pcm) <= np.abs(tv[j] - pcm), j - 1, j)].astype(np.uint8)
for fmt, buf in (("raw", pcm), ("alaw", codes)):
    args = ["-fs", "16000", "-format_in", fmt, "-dither", "0", "-preset", "mfcc", "-preem", "0.97", "-nr_mode", "exten", "-format_out", "htk", "-fea_delta", "d_a"]
    hd = cb.Handle(args); plan = hd.plan(lens)
    hin = torch.empty(len(buf), dtype=torch.uint8 if fmt == "alaw" else torch.int16).pin_memory(); hin.numpy()[:] = buf
    hout = torch.empty((plan.total_frames, hd.feature_dim), dtype=torch.float32).pin_memory()
    kw = {"g711": "alaw"} if fmt == "alaw" else {}
    plan.run_host(hin.numpy(), features=hout.numpy(), want_vad=False, **kw)
    t0 = time.perf_counter()
    for _ in range(3): plan.run_host(hin.numpy(), features=hout.numpy(), want_vad=False, **kw)
    dt = (time.perf_counter() - t0) / 3
    print("%-5s h2d %.2f GB  %.1f ms per %d frames -> %.3e frames/s" % (fmt, hin.numel() * hin.element_size() / 1e9, dt * 1e3, plan.total_frames, plan.total_frames / dt))
    plan.close(); hd.close()
