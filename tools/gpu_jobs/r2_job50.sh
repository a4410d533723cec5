#!/bin/bash
# round 2, job 50: k_bank without the barrier at the end of a tile (feature path): quick A/B against the previous build, then the
# driver's round-end sequence (full GPU suite, smoke, both bench arms) on the new build
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 --others none --no-selfcheck --cli-utts 0"
L=ctucopy_b200/libctucopy_b200.so
show() { python - "$1" <<'P'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d["ms_per_step"], 3), {k: round(v["ms"], 3) for k, v in d["roofline"]["kernels"].items()})
P
}
cp $L /tmp/new.so
cp ctucopy_b200/lib_prev.so $L
for w in mfcc_exten mfcc_d_a; do $B --workload $w > gpurun_out/ab50_prev_$w.json 2>/dev/null; show gpurun_out/ab50_prev_$w.json; done
cp /tmp/new.so $L
for w in mfcc_exten mfcc_d_a; do $B --workload $w > gpurun_out/ab50_new_$w.json 2>/dev/null; show gpurun_out/ab50_new_$w.json; done
bash tools/gpu_jobs/r2_job26.sh
