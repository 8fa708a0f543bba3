# ncu --set full of the two reverted experiment kernels (variant libraries built from their commits)
cd /root/repo
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --no-extras --steps 3 --warmup 3"
PCFE_LIB=$PWD/detmatch_b200/lib/libpcfe_ring.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:hvb_bin -s 3 -c 1 \
  -o gpurun_out/r02b_bin_ring -f python bench.py $B --debug hv_bin_ring=1 > gpurun_out/r02b_ncu_ring.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hvb_bin -s 3 -c 1 \
  -o gpurun_out/r02b_bin_plain -f python bench.py $B > gpurun_out/r02b_ncu_plain.log 2>&1
PCFE_LIB=$PWD/detmatch_b200/lib/libpcfe_tail.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:hvb_bucket_rec -s 3 -c 1 \
  -o gpurun_out/r02b_bucket_walk2 -f python bench.py $B --debug hv_walk2=1 > gpurun_out/r02b_ncu_walk2.log 2>&1
PCFE_LIB=$PWD/detmatch_b200/lib/libpcfe_tail.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:hvb_bucket_rec -s 3 -c 1 \
  -o gpurun_out/r02b_bucket_fold -f python bench.py $B --debug hv_scan_fold=1 > gpurun_out/r02b_ncu_fold.log 2>&1
ls -la gpurun_out/r02b_*.ncu-rep
tail -2 gpurun_out/r02b_ncu_ring.log
