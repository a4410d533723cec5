#!/bin/bash
# round 2, job 45: TRAP-DCT with four rows per thread; packed windowing of the two front ends against the scalar windowing
# (ctucopy_b200/lib_vc.so: packed FFT, windowing as before); parity tests of the TRAP-DCT and golden cases
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 --others none --no-selfcheck --cli-utts 0"
L=ctucopy_b200/libctucopy_b200.so
show() { python - "$1" <<'P'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(d["ms_per_step"], 3), {k: round(v["ms"], 3) for k, v in d["roofline"]["kernels"].items()})
P
}
run() {
  for w in $2; do $B --workload $w > gpurun_out/ab45_$1_$w.json 2> gpurun_out/ab45_$1_$w.err; show gpurun_out/ab45_$1_$w.json; done
  python tools/time_args.py 10000 -- -fs 8000 -format_in raw -preset mfcc -preem 0.97 -fea_delta d_a -format_out htk 2>&1 | grep -E "k_frames<pcm"
}
cp $L /tmp/cur.so
echo "== current"; run cur "mfcc_exten trapdct"
cp ctucopy_b200/lib_vc.so $L
echo "== vc (scalar windowing)"; run vc "mfcc_exten"
cp /tmp/cur.so $L
echo "== current again"; run cur2 "mfcc_exten"
timeout 900 python -m pytest tests -x -q -m gpu -k "trap or golden" > gpurun_out/r2_pytest45.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest45.log
