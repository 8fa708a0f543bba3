cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp11b.log
: > $L
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], {k:v["ms_per_step"] for k,v in d["kernels"].items() if k in ("hvb_expand",)})'
run() { fr=$1; shift
  echo -n "frames $fr $*: " >> $L
  timeout 300 python bench.py --frames $fr --steps 300 --warmup 10 --no-cpu-baseline --no-e2e --no-extras "$@" 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
}
for fr in 64 32 16 4; do
run $fr
for t in 3 4 5; do run $fr --debug hv_expand_map=2 --debug hv_expand_tiles=$t; done
done
run 64 --debug hv_expand_map=2 --debug hv_expand_tiles=4 --debug hv_expand_prefetch=2
run 64 --debug hv_expand_map=2 --debug hv_expand_tiles=4 --debug hv_expand_prefetch=0
run 64 --debug hv_expand_map=2 --debug hv_expand_tiles=8
echo "packed / mean:" >> $L
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["fused"])' >> $L 2>&1
timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --debug hv_expand_map=2 --debug hv_expand_tiles=4 2>&1 | tail -1 | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["fused"])' >> $L 2>&1
cat $L
