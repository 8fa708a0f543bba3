cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp10d.log
: > $L
timeout 1200 python -m pytest tests/test_gpu_voxel.py tests/test_gpu_packed.py tests/test_gpu_vfe.py tests/test_gpu_random.py -m gpu -x -q -k "not exhaustive" 2>&1 | tail -3 >> $L
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items() if k.startswith("hvb")})'
for fr in 16 16 128; do
  echo -n "C5 $fr frames: " >> $L
  timeout 300 python bench.py --workload C5 --frames $fr --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-extras 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
done
cat $L
