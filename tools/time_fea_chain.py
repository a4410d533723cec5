"""Device-resident timing of the 8f.3 row: deltas / stacking on existing feature rows (ctu_plan_run_device_fea) and
stacking behind the sample front end.  Prints per-kernel times (CUDA events inside the library) and the HBM rate of
the gather / delta kernels against their algorithmic bytes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import ctucopy_b200 as cb
from ctucopy_b200 import synthetic

N_UTTS, ROWS = 10000, 998
BH = ["-fs", "16000", "-format_in", "htk", "-fea_ncepcoefs", "12", "-fea_kind", "lpc", "-format_out", "htk"]
x = torch.randn((N_UTTS * ROWS, 13), dtype=torch.float32, device="cuda")


def kernel_times(hd, fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    hd.profile(True)
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    acc = {}
    for name, ms in hd.profile_records():
        acc[name] = acc.get(name, 0.0) + ms / reps
    hd.profile(False)
    return acc


for tag, extra in (("copy", []), ("d_a", ["-fea_delta", "d_a"]), ("d_a_t", ["-fea_delta", "d_a_t"]), ("trap5", ["-fea_trap", "5"]), ("trap11", ["-fea_trap", "11"]),
                   ("d_a+cms", ["-fea_delta", "d_a", "-fea_Z_exp", "500"])):
    hd = cb.Handle(BH + extra)
    plan = hd.plan([ROWS] * N_UTTS)
    y = torch.empty((plan.total_frames, hd.feature_dim), dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    t = kernel_times(hd, lambda: plan.run_device_fea(x.data_ptr(), y.data_ptr(), s))
    rows = plan.total_frames
    # algorithmic bytes: every input row read once, every output row written once (the delta kernel works in place on the
    # output rows: reads the static block, writes the delta blocks)
    alg = rows * 4 * (13 + hd.feature_dim)
    tot = sum(t.values())
    print("%-8s out_dim %3d  total %.3f ms  %s  -> %.0f GB/s of algorithmic bytes, %.2e rows/s" % (
        tag, hd.feature_dim, tot, " ".join("%s=%.3f" % kv for kv in t.items()), alg / tot / 1e6, rows / tot * 1e3))
    plan.close(); hd.close()

# stacking behind the sample front end (BASELINE config 1 + -fea_trap 5)
pcm, lens = synthetic.batch(N_UTTS, 10.0, unique=16)
dp = torch.from_numpy(pcm).cuda()
for tag, extra in (("mfcc", []), ("mfcc+trap5", ["-fea_trap", "5"]), ("logspec+d_a", ["-fea_kind", "logspec", "-fea_delta", "d_a"])):
    args = ["-fs", "16000", "-format_in", "raw", "-dither", "0", "-preset", "mfcc", "-preem", "0.97", "-format_out", "htk"] + extra
    hd = cb.Handle(args)
    plan = hd.plan(lens)
    y = torch.empty((plan.total_frames, hd.feature_dim), dtype=torch.float32, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    t = kernel_times(hd, lambda: plan.run_device(dp.data_ptr(), y.data_ptr(), stream=s), reps=5)
    print("%-12s out_dim %3d  total %.3f ms  %s" % (tag, hd.feature_dim, sum(t.values()), " ".join("%s=%.3f" % kv for kv in t.items())))
    plan.close(); hd.close()
