cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp18.log
: > $L
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items() if k.startswith("hvb")})'
for fr in 16; do
for ba in 1024 512 640 768 896 1200; do
  echo -n "C5 $fr frames hv_bucket_avg=$ba: " >> $L
  timeout 300 python bench.py --workload C5 --frames $fr --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-extras --hv-bucket-avg $ba 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
done
done
cat $L
