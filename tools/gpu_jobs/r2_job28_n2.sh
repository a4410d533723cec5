#!/bin/bash
# round 2, job 28 (2 GPUs): the bench line under torchrun as the driver launches it; CLI -gpus 2 containers against one GPU
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2_bench28_n2.json 2> gpurun_out/r2_bench28_n2.err; echo "bench n=2 rc=$?"; tail -2 gpurun_out/r2_bench28_n2.err
python -c "
import json; d=json.load(open('gpurun_out/r2_bench28_n2.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['selfcheck'], 'e2e', d['e2e']['value'], d['e2e']['frac_of_copy_only'], list(d.get('workloads',{}).keys()))"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2_ref28_n2.json 2> gpurun_out/r2_ref28_n2.err; echo "ref n=2 rc=$?"; cut -c1-200 gpurun_out/r2_ref28_n2.json
python -m pytest tests -m gpu -q -k "gpus or shard" 2>&1 | tail -2
