cd /root/repo
# the driver's round-end sequence on one GPU: pytest -m gpu, smoke(), default bench (both arms)

mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests/ -x -q -m gpu ) > gpurun_out/r02b_full_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02b_smoke.log 2>&1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_bench_ref.json 2> gpurun_out/r02b_bench_ref.err
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err
tail -3 gpurun_out/r02b_full_pytest.log; cat gpurun_out/r02b_smoke.log | tail -2; tail -c 600 gpurun_out/r02b_bench.json
# final-state captures of the C4 step: ncu --set full (traffic + summary), launch list, default bench (both arms)

mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --no-extras"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hvb_ -s 15 -c 5 -o gpurun_out/r02b_c4 -f python bench.py --steps 3 --warmup 3 $B > gpurun_out/r02b_c4_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 36 --csv --log-file gpurun_out/r02b_c4_launches.csv python bench.py --steps 8 --warmup 3 $B > /dev/null 2>&1
timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02b_bench_reference.json 2>/dev/null
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r02b_bench.json 2>/dev/null
timeout 300 python bench.py --workload C1 --steps 300 --warmup 10 $B > gpurun_out/r02b_bench_c1.json 2>/dev/null
timeout 300 python bench.py --workload C5 --frames 16 --steps 300 --warmup 10 $B > gpurun_out/r02b_bench_c5.json 2>/dev/null
ls -la gpurun_out/r02b_c4.ncu-rep; tail -c 300 gpurun_out/r02b_bench.json
