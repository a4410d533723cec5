#!/bin/bash
# round 2, job 48: td-iir filter with warp-private staging (no CTA barrier per chunk, lib_w.so) against the CTA-wide staging
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 --others none --no-selfcheck --cli-utts 0"
L=ctucopy_b200/libctucopy_b200.so
show() { python - "$1" <<'P'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d["ms_per_step"], 3), {k: round(v["ms"], 3) for k, v in d["roofline"]["kernels"].items()})
P
}
cp $L /tmp/cur.so
$B --workload tdiir > gpurun_out/ab48_cur.json 2> gpurun_out/ab48_cur.err; show gpurun_out/ab48_cur.json
cp ctucopy_b200/lib_w.so $L
$B --workload tdiir > gpurun_out/ab48_w.json 2> gpurun_out/ab48_w.err; show gpurun_out/ab48_w.json
timeout 600 python -m pytest tests -x -q -m gpu -k "tdiir" > gpurun_out/r2_pytest48.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest48.log
cp /tmp/cur.so $L
