set -x
timeout 600 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -8 > gpurun_out/t5.log
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0"
$B --utts 1000 --workload exten > gpurun_out/plain_exten.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_synth -s 3 -c 1 -o gpurun_out/prof_synth $B --utts 1000 --workload exten > gpurun_out/ncu_synth.log 2>&1
$B --utts 500 --workload fwss_burg > gpurun_out/plain_burg.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_burg -s 3 -c 1 -o gpurun_out/prof_burg $B --utts 500 --workload fwss_burg > gpurun_out/ncu_burg.log 2>&1
$B --utts 2000 --workload mfcc_d_a > gpurun_out/plain_mfcc.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_frames|k_delta" -s 6 -c 2 -o gpurun_out/prof_frames $B --utts 2000 --workload mfcc_d_a > gpurun_out/ncu_frames.log 2>&1
$B --utts 2000 --workload trapdct > gpurun_out/plain_trap.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_trapdct" -s 3 -c 1 -o gpurun_out/prof_trap $B --utts 2000 --workload trapdct > gpurun_out/ncu_trap.log 2>&1
tail -3 gpurun_out/t5.log
