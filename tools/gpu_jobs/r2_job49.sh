#!/bin/bash
# round 2, job 49: td-iir filter (warp-private staging) at 9 CTAs per SM with 320-sample chunks (lib_w9.so) against 8 with 480
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
$B --workload tdiir > gpurun_out/ab49_cur.json 2> gpurun_out/ab49_cur.err; show gpurun_out/ab49_cur.json
$B --workload tdiir --utts 9472 > gpurun_out/ab49_cur_b.json 2> gpurun_out/ab49_cur_b.err; show gpurun_out/ab49_cur_b.json
cp ctucopy_b200/lib_w9.so $L
$B --workload tdiir > gpurun_out/ab49_w9.json 2> gpurun_out/ab49_w9.err; show gpurun_out/ab49_w9.json
cp /tmp/cur.so $L
