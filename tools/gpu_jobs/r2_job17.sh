#!/bin/bash
# round 2, job 17: ss_carry (list semantics of hwss / fwss / 2fwss) against the reference's list-mode files, library + CLI
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "carry or fwss or burg" > gpurun_out/r2_pytest17.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2_pytest17.log
