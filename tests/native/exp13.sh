cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp13.log
: > $L
timeout 1500 python -m pytest tests/test_gpu_voxel.py tests/test_gpu_packed.py tests/test_gpu_vfe.py tests/test_gpu_random.py -m gpu -x -q -k "not exhaustive" 2>&1 | tail -3 >> $L
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items() if k in ("hvb_expand","hvb_bucket")})'
for fr in 16 128; do
for ct in 16 12 8 4; do
  echo -n "C5 $fr frames words_ctas=$ct: " >> $L
  timeout 300 python bench.py --workload C5 --frames $fr --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-extras --debug hv_words_ctas=$ct 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
done
done
for fr in 64 8; do
  echo -n "C4 $fr frames (map 2 default): " >> $L
  timeout 300 python bench.py --frames $fr --steps 300 --warmup 10 --no-cpu-baseline --no-e2e --no-extras 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
done
cat $L
