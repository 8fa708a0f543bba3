cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp8.log
: > $L
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], {k:v["ms_per_step"] for k,v in d["kernels"].items() if k in ("hvb_expand","hvb_bin","hvb_bucket")})'
for rep in 1 2; do
for v in libpcfe libpcfe_rec1 libpcfe_rec1pol6 libpcfe_st1 libpcfe_st2 libpcfe_pol10; do
  echo -n "frames 64 $v: " >> $L
  PCFE_LIB=$PWD/detmatch_b200/lib/$v.so timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-e2e --no-extras 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
done
done
cat $L
