#!/bin/bash
# round 2, job 44: (1) td-iir filter at 8 / 9 / 10 resident CTAs per SM; (2) the other kernels of the packed translation unit
# (k_frames256 at 8 kHz, the general kernel at 32 kHz, the single fused kernel k_frames<pcm,fea>) against the scalar build
mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-cpu-baseline --e2e-steps 0 --others none --no-selfcheck --cli-utts 0"
L=ctucopy_b200/libctucopy_b200.so
for m in 8 9 10; do
  CTU_TDIIR_MINB=$m $B --workload tdiir > gpurun_out/ab44_tdiir$m.json 2> gpurun_out/ab44_tdiir$m.err
  python - $m <<'P'
import json, sys
d = json.loads(open("gpurun_out/ab44_tdiir%s.json" % sys.argv[1]).read().strip().splitlines()[-1])
print("tdiir minb", sys.argv[1], round(d["ms_per_step"], 3), {k: round(v["ms"], 3) for k, v in d["roofline"]["kernels"].items()})
P
done
cp $L /tmp/packed.so
run() {
  echo "-- 8 kHz mfcc_d_a"; python tools/time_args.py 10000 -- -fs 8000 -format_in raw -preset mfcc -preem 0.97 -fea_delta d_a -format_out htk 2>&1 | grep -E "k_|device total"
  echo "-- 32 kHz logspec"; python tools/time_args.py 4000 -- -fs 32000 -format_in raw -preset mfcc -fea_kind logspec -preem 0.97 -format_out htk 2>&1 | grep -E "k_|device total"
  echo "-- 16 kHz mfcc_d_a, single fused kernel"; CTU_SPLIT_FRONT=0 python tools/time_args.py 10000 -- -fs 16000 -format_in raw -preset mfcc -preem 0.97 -fea_delta d_a -format_out htk 2>&1 | grep -E "k_|device total"
  echo "-- 8 kHz exten waveform"; python tools/time_args.py 4000 -- -fs 8000 -format_in raw -preset exten -format_out raw 2>&1 | grep -E "k_|device total"
}
echo "== packed"; run
cp ctucopy_b200/libctucopy_b200_scalar.so $L
echo "== scalar"; run
cp /tmp/packed.so $L
