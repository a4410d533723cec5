timeout 900 python -m pytest tests -m gpu -q --maxfail=10 -p no:cacheprovider 2>&1 | tail -25 > gpurun_out/t9.log
tail -3 gpurun_out/t9.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 0"
$B --workload exten > gpurun_out/b_exten.json 2> gpurun_out/b_exten.err
$B --utts 4000 --workload fwss_burg > gpurun_out/b_burg.json 2> gpurun_out/b_burg.err
CTU_BURG_MINB=3 $B --utts 4000 --workload fwss_burg > gpurun_out/b_burg3.json 2> gpurun_out/b_burg3.err
