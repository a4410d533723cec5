# round-1 evidence: launch lists (device time of every launch of one step) and one ncu --set full capture per workload
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --utts 2000"
for w in mfcc_exten mfcc_d_a plp trapdct exten fwss_burg; do
  $B --workload $w > gpurun_out/plain_$w.log 2>&1 || { echo "plain failed $w"; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" --csv --log-file gpurun_out/launches_$w.csv $B --workload $w > gpurun_out/nl_$w.log 2>&1
done
tools/gpu_jobs/ncu_cap.sh p_mfcc_exten "k_frames|k_nr_scan|k_delta" 12 4 $B --workload mfcc_exten
tools/gpu_jobs/ncu_cap.sh p_mfcc_d_a "k_frames|k_delta" 6 2 $B --workload mfcc_d_a
tools/gpu_jobs/ncu_cap.sh p_plp "k_frames|k_lpc" 6 2 $B --workload plp
tools/gpu_jobs/ncu_cap.sh p_trapdct "k_trapdct" 3 1 $B --workload trapdct
tools/gpu_jobs/ncu_cap.sh p_exten "k_synth" 3 1 $B --workload exten
tools/gpu_jobs/ncu_cap.sh p_fwss_burg "k_burg|k_cepdet" 6 2 $B --utts 500 --workload fwss_burg
rm -f gpurun_out/srccu_p_*.csv gpurun_out/src_p_*.csv
du -sh gpurun_out
