python - <<'PY'
import os, sys, subprocess, shutil, filecmp
sys.path.insert(0, ".")
import numpy as np
from ctucopy_b200 import synthetic
d = "/dev/shm/g2"; shutil.rmtree(d, ignore_errors=True); os.makedirs(d + "/in")
n = 300
for i in range(n):
    synthetic.utterance(i % 16, 2.0 + (i % 5)).astype("<i2").tofile("%s/in/u%03d.raw" % (d, i))
exe = "host/ctucopy_b200"
for fmt, extra in (("htk", []), ("ark={D}/o.ark", []), ("pfile={D}/o.pfile", []), ("htk", ["-dither", "1.0"])):
    res = {}
    for tag, gp in (("one", []), ("two", ["-gpus", "2"])):
        od = "%s/%s" % (d, tag); shutil.rmtree(od, ignore_errors=True); os.makedirs(od)
        with open(od + "/list.scp", "w") as fh:
            for i in range(n):
                fh.write("%s/in/u%03d.raw %s/u%03d.out\n" % (d, i, od, i))
        a = ["-fs", "16000", "-format_in", "raw", "-preset", "plpc", "-format_out", fmt.replace("{D}", od)] + extra + ["-S", od + "/list.scp"] + gp
        pr = subprocess.run([exe] + a, capture_output=True)
        assert pr.returncode == 0, pr.stderr.decode()
        res[tag] = od
    names = sorted(f for f in os.listdir(res["one"]) if f != "list.scp")
    same = all(open(res["one"] + "/" + f, "rb").read().replace(res["one"].encode(), b"@") == open(res["two"] + "/" + f, "rb").read().replace(res["two"].encode(), b"@") for f in names)
    print(fmt, extra, "files:", len(names), "identical across -gpus 1 / 2:", same, "leftover shards:", [f for f in os.listdir(res["two"]) if f.startswith("shard")])
shutil.rmtree(d, ignore_errors=True)
PY
