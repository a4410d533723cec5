#!/bin/bash
# round 2, job 46: k_frames2 with its window read from shared memory and five CTAs per SM (92 registers, lib_vd.so) against the
# window in registers and four CTAs (121 registers)
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 --others none --no-selfcheck --cli-utts 0"
L=ctucopy_b200/libctucopy_b200.so
show() { python - "$1" <<'P'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d["ms_per_step"], 3), {k: round(v["ms"], 3) for k, v in d["roofline"]["kernels"].items()})
P
}
run() { for w in mfcc_exten exten trapdct; do $B --workload $w > gpurun_out/ab46_$1_$w.json 2> gpurun_out/ab46_$1_$w.err; show gpurun_out/ab46_$1_$w.json; done; }
cp $L /tmp/cur.so
echo "== current"; run cur
cp ctucopy_b200/lib_vd.so $L
echo "== vd"; run vd
cp /tmp/cur.so $L
