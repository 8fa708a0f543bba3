cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp21.log
: > $L
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], {k:v["ms_per_step"] for k,v in d["kernels"].items() if k in ("hvb_expand",)})'
for rep in 1 2; do
for v in libpcfe libpcfe_pol0 libpcfe_pol1 libpcfe_pol3 libpcfe_pol6; do
  echo -n "frames 64 (map 2) $v: " >> $L
  PCFE_LIB=$PWD/detmatch_b200/lib/$v.so timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-e2e --no-extras 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
done
done
cat $L
