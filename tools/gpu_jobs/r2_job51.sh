#!/bin/bash
# round 2, job 51: cepstra behind a delta chain as compact rows (CTU_COMPACT_STATIC, default on) against the in-place layout, then the
# driver's round-end sequence on the new build
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 --others none --no-selfcheck --cli-utts 0"
show() { python - "$1" <<'P'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d["ms_per_step"], 3), {k: round(v["ms"], 3) for k, v in d["roofline"]["kernels"].items()})
P
}
for c in 0 1; do for w in mfcc_exten mfcc_d_a; do CTU_COMPACT_STATIC=$c $B --workload $w > gpurun_out/ab51_${c}_$w.json 2>/dev/null; show gpurun_out/ab51_${c}_$w.json; done; done
bash tools/gpu_jobs/r2_job26.sh
