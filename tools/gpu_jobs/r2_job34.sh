#!/bin/bash
# round 2, job 34: the CLI's I/O thread count on the 16-core box (20 000 files on /dev/shm); the bench line's cli block at its new default
mkdir -p gpurun_out
nproc
for t in 8 12 16; do
  echo "== CTU_IO_THREADS=$t"
  CTU_IO_THREADS=$t python tools/cli_e2e.py 20000 2>&1 | grep -v "reference binary\|same size" | grep -v "batch" | tail -4
  CTU_IO_THREADS=$t python tools/cli_e2e.py 8000 2>&1 | grep "batch" | tail -9 | awk '{print $3, $4, $5}' | sort | uniq -c | sort -rn | head -6
done
python bench.py --others none --no-cpu-baseline --e2e-steps 1 --steps 5 > gpurun_out/r2_bench34.json 2> gpurun_out/r2_bench34.err; echo "bench rc=$?"
python -c "
import json; d=json.load(open('gpurun_out/r2_bench34.json')); print(d.get('cli'))"
