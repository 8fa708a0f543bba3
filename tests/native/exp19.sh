cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp19b.log
: > $L
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items() if k in ("hvb_bin","hvb_bucket","hvb_expand")})'
for rep in 1 2 3; do
for v in libpcfe_head libpcfe; do
  echo -n "C4 64 frames $v: " >> $L
  PCFE_LIB=$PWD/detmatch_b200/lib/$v.so timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-e2e --no-extras 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
done
done
echo -n "C4 64 frames libpcfe exact 800: " >> $L
timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-e2e --no-extras --debug hv_nb_exact=1 --hv-bucket-avg 800 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
echo -n "C4 64 frames libpcfe exact 860: " >> $L
timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-e2e --no-extras --debug hv_nb_exact=1 --hv-bucket-avg 860 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
echo -n "C4 64 frames libpcfe exact 760: " >> $L
timeout 300 python bench.py --steps 300 --warmup 10 --no-cpu-baseline --no-e2e --no-extras --debug hv_nb_exact=1 --hv-bucket-avg 760 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
cat $L
