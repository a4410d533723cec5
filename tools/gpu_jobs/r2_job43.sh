#!/bin/bash
# round 2, job 43: (1) synthesis kernel with packed FP32: scalar translation unit (v1) against packed with two CTAs per SM (v2);
# (2) td-iir filter with and without the chunk prefetch; (3) parity tests of what changed; (4) ncu of the packed front end
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 --others none --no-selfcheck --cli-utts 0"
L=ctucopy_b200/libctucopy_b200.so
show() { python - "$1" <<'P'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d["ms_per_step"], 3), {k: round(v["ms"], 3) for k, v in d["roofline"]["kernels"].items()})
P
}
for v in v2 v1; do
  cp ctucopy_b200/lib_$v.so $L
  $B --workload exten > gpurun_out/ab43_${v}_exten.json 2> gpurun_out/ab43_$v.err; echo "== $v"; show gpurun_out/ab43_${v}_exten.json
done
for pf in 0 1; do
  CTU_TDIIR_PF=$pf $B --workload tdiir > gpurun_out/ab43_tdiir$pf.json 2> gpurun_out/ab43_tdiir$pf.err; echo "== tdiir pf=$pf"; show gpurun_out/ab43_tdiir$pf.json
done
timeout 900 python -m pytest tests -x -q -m gpu -k "tdiir or exten or alternative or golden or trap or cli" > gpurun_out/r2_pytest43.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest43.log
B2="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --utts 2000 --others none --no-selfcheck --cli-utts 0"
tools/gpu_jobs/ncu_cap.sh p_mfcc_exten "k_frames|k_bank|k_delta" 12 3 $B2 --workload mfcc_exten
rm -f gpurun_out/srccu_p_mfcc_exten.csv
