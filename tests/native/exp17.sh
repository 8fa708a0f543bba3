cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp17.log
: > $L
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], {k:v["ms_per_step"] for k,v in d["kernels"].items() if k in ("hvb_expand",)})'
for v in libpcfe libpcfe_s2m10 libpcfe_s2m12 libpcfe_s3m12; do
for t in 3 4; do
  echo -n "frames 64 $v tiles=$t: " >> $L
  PCFE_LIB=$PWD/detmatch_b200/lib/$v.so timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-e2e --no-extras --debug hv_expand_tiles=$t 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
done
done
cat $L
