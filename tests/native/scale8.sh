cd /root/repo
mkdir -p gpurun_out
N=$1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02b_bench_${N}gpu.json 2> gpurun_out/r02b_bench_${N}gpu.err
tail -c 400 gpurun_out/r02b_bench_${N}gpu.json
if [ "$N" = "8" ]; then
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --workload C5 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/r02b_bench_c5_${N}gpu.json 2>> gpurun_out/r02b_bench_${N}gpu.err
fi
