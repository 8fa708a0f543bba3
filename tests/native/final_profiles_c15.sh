cd /root/repo
mkdir -p gpurun_out
B="--no-cpu-baseline --no-e2e --no-extras"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hv -s 21 -c 7 -o gpurun_out/r02b_c5 -f python bench.py --workload C5 --frames 16 --steps 3 --warmup 3 $B > gpurun_out/r02b_c5_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:hv -s 18 -c 6 -o gpurun_out/r02b_c1 -f python bench.py --workload C1 --steps 3 --warmup 3 $B > gpurun_out/r02b_c1_ncu.log 2>&1
ls -la gpurun_out/r02b_c5.ncu-rep gpurun_out/r02b_c1.ncu-rep
