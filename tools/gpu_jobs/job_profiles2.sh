# refresh of the round-1 evidence after the split front end and the stored-spectrum synthesis became the defaults
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 0 --utts 2000"
for w in mfcc_d_a plp trapdct exten; do
  $B --workload $w > gpurun_out/plain_$w.log 2>&1 || { echo "plain failed $w"; continue; }
  ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_" --csv --log-file gpurun_out/launches_$w.csv $B --workload $w > gpurun_out/nl_$w.log 2>&1
done
tools/gpu_jobs/ncu_cap.sh p_exten "k_synth_c|k_frames2" 6 2 $B --workload exten
tools/gpu_jobs/ncu_cap.sh p_mfcc_d_a "k_frames|k_delta" 9 3 $B --workload mfcc_d_a
rm -f gpurun_out/srccu_p_*.csv gpurun_out/src_p_*.csv
du -sh gpurun_out
