#!/bin/bash
# round 2, job 42: A/B of the packed FP32 complex arithmetic (FADD2 / FMUL2 / FFMA2, ctu_fft.cuh) against the scalar build
# (-DCTU_NO_F32X2, ctucopy_b200/libctucopy_b200_scalar.so), device-resident step times; then the parity tests of the packed build
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 --others none --no-selfcheck --cli-utts 0"
L=ctucopy_b200/libctucopy_b200.so
cp $L /tmp/packed.so
run() {  # label
  for w in mfcc_exten exten trapdct; do
    $B --workload $w > gpurun_out/ab42_$1_$w.json 2> gpurun_out/ab42_$1_$w.err
    python - "$1" $w <<'P'
import json, sys
d = json.loads(open("gpurun_out/ab42_%s_%s.json" % (sys.argv[1], sys.argv[2])).read().strip().splitlines()[-1])
print(sys.argv[1], sys.argv[2], round(d["ms_per_step"], 3), {k: round(v["ms"], 3) for k, v in d["roofline"]["kernels"].items()})
P
  done
  python tools/time_args.py 10000 -- -fs 8000 -format_in raw -preset mfcc -preem 0.97 -fea_delta d_a -format_out htk 2>&1 | tail -3
}
echo "== packed"; run packed
cp ctucopy_b200/libctucopy_b200_scalar.so $L
echo "== scalar"; run scalar
cp /tmp/packed.so $L
echo "== packed again"; run packed2
timeout 1500 python -m pytest tests -x -q -m gpu -k "not sweeps" > gpurun_out/r2_pytest42.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest42.log
