cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp4.log
: > $L
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], {k:v["ms_per_step"] for k,v in d["kernels"].items() if k.startswith("hvb")})'
for fr in 64 8; do
for v in libpcfe libpcfe_b128x16 libpcfe_b128x32 libpcfe_b256x8 libpcfe_b256x16m5 libpcfe_bt512 libpcfe_bt384 libpcfe_bt128 libpcfe; do
  echo -n "frames $fr $v: " >> $L
  PCFE_LIB=$PWD/detmatch_b200/lib/$v.so timeout 300 python bench.py --frames $fr --steps 300 --warmup 10 --no-cpu-baseline --no-e2e --no-extras 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
done
done
cat $L
