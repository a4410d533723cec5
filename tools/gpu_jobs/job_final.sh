timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/final_pytest.log 2>&1; tail -4 gpurun_out/final_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
( time python bench.py --others ) > gpurun_out/b_all.json 2> gpurun_out/b_all.err; tail -3 gpurun_out/b_all.err
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; tail -3 gpurun_out/bench_ref.err
