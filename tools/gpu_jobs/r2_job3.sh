#!/bin/bash
# round 2, job 3: k_burg v2 + single-buffer k_bank: Burg / NR parity tests first, then the whole suite, then bench with both scan placements
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -k "burg or fwss or hwss or vad or exten" > gpurun_out/r2_pytest3a.log 2>&1; echo "pytest(burg,nr) rc=$?"; tail -3 gpurun_out/r2_pytest3a.log
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest3.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest3.log
python bench.py --no-cpu-baseline --e2e-steps 1 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc=$?"
CTU_FUSE_NR=0 python bench.py --no-cpu-baseline --others none --e2e-steps 0 > gpurun_out/r2_bench3_nofuse.json 2> gpurun_out/r2_bench3_nofuse.err; echo "bench nofuse rc=$?"
python - <<'PY'
import json
for f in ('gpurun_out/r2_bench3.json','gpurun_out/r2_bench3_nofuse.json'):
    try:
        d=json.load(open(f))
    except Exception as e:
        print(f, 'unreadable', e); continue
    print(f, 'value',d['value'],'ms',d['ms_per_step'],'self',d['selfcheck'], d.get('selfcheck_detail'))
    print(d['kernel_ms_per_step'])
    for k,v in d.get('workloads',{}).items(): print(k, v.get('value'), v.get('ms_per_step'), v.get('selfcheck'), v.get('kernel_ms_per_step'), v.get('error'))
PY
tail -3 gpurun_out/r2_bench3.err
