for mb in 16 32 64 128 256; do
CTU_CHUNK_MB=$mb python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 5 --workload mfcc_d_a 2>/dev/null | python -c "import json,sys; j=json.loads(sys.stdin.read()); print('$mb', j['e2e'])"
done
