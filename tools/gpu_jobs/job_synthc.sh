timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_cli.py -x -q -k "raw or wave or exten or waveform" > gpurun_out/synthc_pytest.log 2>&1; tail -3 gpurun_out/synthc_pytest.log
for v in pcm cspec; do
  if [ $v = pcm ]; then export CTU_SYNTH_FROM_PCM=1; else unset CTU_SYNTH_FROM_PCM; fi
  python bench.py --workload exten --no-cpu-baseline --steps 10 --e2e-steps 2 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.readline()); print('$v %.2f ms  %.3e frames/s  e2e %.3e' % (d['ms_per_step'], d['value'], d['e2e']['value']), {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()})"
done
python - <<'PY'
# bit-identity of the two synthesis paths on the parity set
import os, sys, subprocess
sys.path.insert(0, os.getcwd())
code = """
import sys, os, numpy as np
sys.path.insert(0, os.getcwd())
import ctucopy_b200 as cb
from ctucopy_b200 import synthetic
utts = [synthetic.utterance(k, 2.0) for k in range(12)]
for args in (['-fs','16000','-format_in','raw','-dither','0','-preset','exten','-format_out','raw'],
             ['-fs','16000','-format_in','raw','-dither','0','-w','25','-s','10','-nr_mode','fwss','-vad','burg','-format_out','raw']):
    r = cb.extract(args, utts)
    np.save(sys.argv[1] + '_%d.npy' % len(args), r.waveform)
"""
for tag, env in (("a", {"CTU_SYNTH_FROM_PCM": "1"}), ("b", {})):
    e = dict(os.environ); e.pop("CTU_SYNTH_FROM_PCM", None); e.update(env)
    subprocess.run([sys.executable, "-c", code, "/tmp/w" + tag], check=True, env=e)
import numpy as np
for n in (10, 16):
    a, b = np.load("/tmp/wa_%d.npy" % n), np.load("/tmp/wb_%d.npy" % n)
    print("args", n, "identical:", np.array_equal(a, b), "differing samples:", int((a != b).sum()), "of", a.size)
PY
