# the driver's round-end sequence on one GPU: pytest -m gpu, smoke(), default bench (both arms)
cd /root/repo
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r02b_full_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02b_smoke.log 2>&1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_bench_ref.json 2> gpurun_out/r02b_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
tail -3 gpurun_out/r02b_full_pytest.log; cat gpurun_out/r02b_smoke.log | tail -2; tail -c 600 gpurun_out/r02b_bench.json
