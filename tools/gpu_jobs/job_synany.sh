timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "8k or 44k or 22k or exten" > gpurun_out/synany_pytest.log 2>&1; tail -15 gpurun_out/synany_pytest.log
timeout 300 python - > gpurun_out/synany_time.log 2>&1 <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch, ctucopy_b200 as cb
from ctucopy_b200 import synthetic
pcm, lens = synthetic.batch(4000, 10.0, unique=16)
dp = torch.from_numpy(pcm).cuda()
for fs in ("8000", "16000", "44100"):
    args = ["-fs", fs, "-format_in", "raw", "-dither", "0", "-preset", "exten", "-format_out", "raw"]
    hd = cb.Handle(args); plan = hd.plan(lens)
    out = torch.empty(plan.total_output_samples, dtype=torch.int16, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(2): plan.run_device(dp.data_ptr(), d_waveform=out.data_ptr(), stream=s)
    torch.cuda.synchronize(); hd.profile(True)
    for _ in range(3): plan.run_device(dp.data_ptr(), d_waveform=out.data_ptr(), stream=s)
    torch.cuda.synchronize()
    acc = {}
    for n, ms in hd.profile_records(): acc[n] = acc.get(n, 0) + ms / 3
    print(fs, "frames", plan.total_frames, "total %.2f ms" % sum(acc.values()), " ".join("%s=%.2f" % kv for kv in acc.items()), "-> %.2e frames/s" % (plan.total_frames / sum(acc.values()) * 1e3))
    plan.close(); hd.close()
PY
cat gpurun_out/synany_time.log | tail -5
