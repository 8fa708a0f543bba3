# round-2 (second session) experiment batch 1: tail variants of the record path.  gpurun -- bash tests/native/exp1.sh
cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp1.log
: > $L
echo "== parity (tail mode, quick tests)" >> $L
timeout 600 python -m pytest tests/test_gpu_voxel.py tests/test_gpu_packed.py tests/test_gpu_vfe.py -m gpu -x -q \
  -k "tail and not full_c4 and not full_c5 and not exhaustive" 2>&1 | tail -5 >> $L
B="--no-cpu-baseline --no-e2e --no-extras"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})'
run() { # frames, debug flags...
  fr=$1; shift
  echo -n "frames $fr $* : " >> $L
  timeout 300 python bench.py --frames $fr --steps 300 --warmup 10 $B "$@" 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
}
for fr in 64 8; do
  run $fr
  run $fr --debug hv_walk2=1
  run $fr --debug hv_scan_fold=1
  run $fr --debug hv_walk2=1 --debug hv_scan_fold=1
  run $fr --debug hv_expand_rev=1
  run $fr --debug hv_ent_evict=1
  run $fr --debug hv_expand_rev=1 --debug hv_ent_evict=1
  run $fr --debug hv_walk2=1 --debug hv_scan_fold=1 --debug hv_expand_rev=1 --debug hv_ent_evict=1
done
run 64 --debug hv_carveout=1
run 64 --debug hv_carveout=1 --hv-wave 32
run 64 --hv-wave 32
run 64 --debug hv_carveout=1 --hv-wave 16
run 64 --hv-wave 16
echo "== C1 / C5" >> $L
for wl in C1; do
  echo -n "$wl base: " >> $L; timeout 300 python bench.py --workload $wl --steps 300 --warmup 10 $B 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
  echo -n "$wl tail: " >> $L; timeout 300 python bench.py --workload $wl --steps 300 --warmup 10 $B --debug hv_walk2=1 --debug hv_scan_fold=1 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
done
cat $L
