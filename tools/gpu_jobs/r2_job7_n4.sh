#!/bin/bash
# round 2, job 7 (4 GPUs): the box's host <-> device copy ceiling with 1 / 2 / 4 GPUs copying at once, and the bench line at N = 2 and 4
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r2_topo_n4.txt 2>&1
lscpu | head -25 >> gpurun_out/r2_topo_n4.txt
tools/copy_probe.sh 4 gpurun_out/r2_copy_probe_n4.jsonl 3 > /dev/null 2>&1; echo "probe4 rc=$?"
tools/copy_probe.sh 2 gpurun_out/r2_copy_probe_n2.jsonl 3 > /dev/null 2>&1; echo "probe2 rc=$?"
for n in 2 4; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 10 --warmup 3 --others none --no-selfcheck > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err; echo "bench n=$n rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n$n.json')); e=d['e2e']; print('N=$n value', d['value'], 'e2e', e['value'], 'ms', e['ms_per_step'], 'copy_only_ms', e['copy_only_ms'], 'frac', e['frac_of_copy_only'], 'GB/s/gpu', e['copy_only_gbs_per_gpu'])"
done
grep -h '"concurrent": [24]' gpurun_out/r2_copy_probe_n4.jsonl gpurun_out/r2_copy_probe_n2.jsonl | python -c "
import sys, json, collections
agg=collections.defaultdict(list)
for l in sys.stdin:
    d=json.loads(l); agg[(d['concurrent'], d['alloc'], d['dirs'], d['chunk_mb'])].append(d)
for k,v in sorted(agg.items()): print(k, 'n=%d' % len(v), 'sum h2d %.1f d2h %.1f total %.1f GB/s' % (sum(x['h2d_gbs'] for x in v), sum(x['d2h_gbs'] for x in v), sum(x['total_gbs'] for x in v)))"
