cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp20.log
: > $L
timeout 900 python -m pytest tests/test_gpu_voxel.py -m gpu -x -q -k "(long_voxels or full_c5 or caps or shapes or random_small or (full_size and C5) or heavy or unaligned or repeatable) and (launches or bucket_general or fallback)" 2>&1 | tail -3 >> $L
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items() if k in ("hvb_bucket","hvb_expand")})'
for rep in 1 2; do
for v in libpcfe_e64 libpcfe libpcfe_e8 libpcfe_e32 libpcfe_e16w8; do
  echo -n "C5 16 frames $v: " >> $L
  PCFE_LIB=$PWD/detmatch_b200/lib/$v.so timeout 300 python bench.py --workload C5 --frames 16 --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-extras 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
done
done
echo -n "C5 128 frames libpcfe: " >> $L
timeout 300 python bench.py --workload C5 --frames 128 --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-extras 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
cat $L
