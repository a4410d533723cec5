#!/bin/bash
# round 2, job 10: where the wall-clock time of a CLI run goes (stage times on stderr), 4000 and 20000 files on /dev/shm
mkdir -p gpurun_out
nproc; free -g | head -2
python tools/cli_e2e.py 4000 > gpurun_out/r2_cli_e2e_4000.txt 2>&1; echo "rc=$?"
python tools/cli_e2e.py 20000 > gpurun_out/r2_cli_e2e_20000.txt 2>&1; echo "rc=$?"
grep -v "batch" gpurun_out/r2_cli_e2e_20000.txt | tail -20
