cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp7b.log
: > $L
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], {k:v["ms_per_step"] for k,v in d["kernels"].items() if k in ("hvb_expand","hvb_bin","hvb_bucket")})'
for rep in 1 2; do
for pf in 1 2 3 0; do
  echo -n "frames 64 pol2 prefetch $pf: " >> $L
  timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-e2e --no-extras --debug hv_expand_prefetch=$pf 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
done
done
for fr in 8 16 32; do
  echo -n "frames $fr pol2: " >> $L
  timeout 300 python bench.py --frames $fr --steps 300 --warmup 10 --no-cpu-baseline --no-e2e --no-extras 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
done
timeout 600 python -m pytest tests/test_gpu_voxel.py tests/test_gpu_packed.py tests/test_gpu_vfe.py -m gpu -x -q -k "(launches or record) and not exhaustive" 2>&1 | tail -2 >> $L
cat $L
