"""Per-kernel device times (CUDA events inside the library) and host-path wall time for any command line:
   python tools/time_args.py [n_utts] -- <ctucopy options>"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ctucopy_b200 as cb
from ctucopy_b200 import synthetic

i = sys.argv.index("--")
n_utts = int(sys.argv[1]) if i > 1 else 4000
args = sys.argv[i + 1:]
hd = cb.Handle(args)
pcm, lens = synthetic.batch(n_utts, 10.0, unique=16)
plan = hd.plan(lens)
hp = torch.empty(len(pcm), dtype=torch.int16).pin_memory(); hp.numpy()[:] = pcm
dp = hp.cuda()
sig = hd.signal_output
if sig:
    d_out = torch.empty(plan.total_output_samples, dtype=torch.int16, device="cuda")
    h_out = torch.empty(plan.total_output_samples, dtype=torch.int16).pin_memory()
else:
    d_out = torch.empty((plan.total_frames, hd.feature_dim), dtype=torch.float32, device="cuda")
    h_out = torch.empty((plan.total_frames, hd.feature_dim), dtype=torch.float32).pin_memory()
d_v = torch.zeros(plan.total_frames, dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
def step():
    if sig: plan.run_device(dp.data_ptr(), d_waveform=d_out.data_ptr(), d_vad_nr=d_v.data_ptr(), d_vad_out=d_v.data_ptr(), stream=s)
    else: plan.run_device(dp.data_ptr(), d_features=d_out.data_ptr(), d_vad_nr=d_v.data_ptr(), d_vad_out=d_v.data_ptr(), stream=s)
for _ in range(2): step()
torch.cuda.synchronize(); hd.profile(True)
R = 3
for _ in range(R): step()
torch.cuda.synchronize()
acc = {}
for n, ms in hd.profile_records(): acc[n] = acc.get(n, 0) + ms / R
hd.profile(False)
tot = sum(acc.values())
print("frames %d  device total %.2f ms -> %.3e frames/s" % (plan.total_frames, tot, plan.total_frames / tot * 1e3))
for k, v in acc.items(): print("   %-28s %.3f ms" % (k, v))
kw = {"waveform": h_out.numpy()} if sig else {"features": h_out.numpy()}
plan.run_host(hp.numpy(), want_vad=True, **kw)
t0 = time.perf_counter()
for _ in range(R): plan.run_host(hp.numpy(), want_vad=True, **kw)
dt = (time.perf_counter() - t0) / R
print("host path (pinned buffers, chunked pipeline): %.2f ms -> %.3e frames/s" % (dt * 1e3, plan.total_frames / dt))
