mkdir -p /tmp/su && cd /tmp/su
python - <<'PY'
import sys, numpy as np
sys.path.insert(0, "/root/repo")
from ctucopy_b200 import synthetic
for i in range(3): synthetic.utterance(i, 2.0).astype("<i2").tofile("/tmp/su/u%d.raw" % i)
open("/tmp/su/l.scp", "w").write("".join("/tmp/su/u%d.raw /tmp/su/o%d.htk\n" % (i, i) for i in range(3)))
PY
for k in 1 2 3; do
  ( time env CTU_TIMING=1 /root/repo/host/ctucopy_b200 -fs 16000 -format_in raw -preset mfcc -fea_delta d_a -format_out htk -S /tmp/su/l.scp ) 2>&1 | tail -16
  echo ----
done
