cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp6.log
: > $L
timeout 900 python -m pytest tests/test_gpu_voxel.py tests/test_gpu_packed.py -m gpu -x -q -k "async and not exhaustive" 2>&1 | tail -4 >> $L
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], {k:v["ms_per_step"] for k,v in d["kernels"].items() if k in ("hvb_expand","hvb_bucket")})'
run() { fr=$1; shift
  echo -n "frames $fr $*: " >> $L
  timeout 300 python bench.py --frames $fr --steps 300 --warmup 10 --no-cpu-baseline --no-e2e --no-extras "$@" 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
}
run 64
for t in 3 4 6 8 12; do run 64 --debug hv_expand_async=1 --debug hv_expand_tiles=$t; done
run 64 --debug hv_expand_async=1 --debug hv_expand_tiles=6 --debug hv_expand_prefetch=0
run 64 --debug hv_expand_async=1 --debug hv_expand_tiles=6 --debug hv_expand_prefetch=2
run 8
for t in 2 3 4 6; do run 8 --debug hv_expand_async=1 --debug hv_expand_tiles=$t; done
run 64
cat $L
