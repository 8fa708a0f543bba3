cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp12.log
: > $L
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["fused"]["packed_ms_per_step"], d["fused"]["packed_mean_ms_per_step"])'
for rep in 1 2; do
for v in libpcfe_pol0 libpcfe; do
for m in "0 3" "2 4" "2 3"; do
  set -- $m
  echo -n "$v map=$1 tiles=$2 (step, packed, packed+mean): " >> $L
  PCFE_LIB=$PWD/detmatch_b200/lib/$v.so timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --debug hv_expand_map=$1 --debug hv_expand_tiles=$2 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
done
done
done
cat $L
