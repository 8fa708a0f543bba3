# experiment batch 2: persistent TMA-ring partition kernel, two-words-per-thread numbering kernel
cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp2.log
: > $L
echo "== parity (ring mode)" >> $L
timeout 900 python -m pytest tests/test_gpu_voxel.py tests/test_gpu_packed.py tests/test_gpu_vfe.py -m gpu -x -q \
  -k "ring and not full_c4 and not full_c5 and not exhaustive" 2>&1 | tail -5 >> $L
timeout 900 python -m pytest tests/test_gpu_voxel.py -m gpu -x -q -k "launches and (full_size or sweep)" 2>&1 | tail -3 >> $L
B="--no-cpu-baseline --no-e2e --no-extras"
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})'
run() { wl=$1; fr=$2; shift; shift
  echo -n "$wl frames $fr $* : " >> $L
  timeout 300 python bench.py --workload $wl --frames $fr --steps 300 --warmup 10 $B "$@" 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
}
for rep in 1 2; do
run C4 64 --debug hv_bin_ring=0 --debug hv_scan_wpt=1
run C4 64 --debug hv_bin_ring=1 --debug hv_scan_wpt=1
run C4 64 --debug hv_bin_ring=0 --debug hv_scan_wpt=2
run C4 64 --debug hv_bin_ring=1 --debug hv_scan_wpt=2
done
run C4 32 --debug hv_bin_ring=0 --debug hv_scan_wpt=1
run C4 32
run C4 16 --debug hv_bin_ring=0 --debug hv_scan_wpt=1
run C4 16 --debug hv_bin_small=0
run C4 16 --debug hv_scan_wpt=2
run C5 16 --debug hv_bin_ring=0
run C5 16
run C1 64 --debug hv_bin_ring=0 --debug hv_scan_wpt=1
run C1 64
cat $L
