cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp14.log
: > $L
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], {k:v["ms_per_step"] for k,v in d["kernels"].items() if k.startswith("hvb")})'
run() { fr=$1; shift
  echo -n "C1 $fr frames $*: " >> $L
  timeout 300 python bench.py --workload C1 --frames $fr --steps 500 --warmup 20 --no-cpu-baseline --no-e2e --no-extras "$@" 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
}
for fr in 16 4; do
run $fr
run $fr --debug hv_expand_tiles=1
run $fr --debug hv_expand_tiles=2
run $fr --debug hv_expand_tiles=4
run $fr --debug hv_expand_map=0
run $fr --hv-bucket-avg 512
run $fr --hv-bucket-avg 2048
run $fr --debug hv_bin_small=0
run $fr --debug hv_bin_small=1
run $fr --debug hv_expand_prefetch=0
done
cat $L
