cd /root/repo
mkdir -p gpurun_out
L=gpurun_out/r02b_exp5c.log
: > $L
P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})'
run() { fr=$1; shift
  echo -n "C5 $fr frames $*: " >> $L
  timeout 300 python bench.py --workload C5 --frames $fr --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-extras "$@" 2>&1 | tail -1 | python -c "$P" >> $L 2>&1
}
run 16 --debug hv_prefill=0
run 16 --debug hv_prefill=16 --debug hv_prefill_after=1
run 16 --debug hv_prefill=8 --debug hv_prefill_after=1
run 16 --debug hv_prefill=0 --hv-wave 8
run 16 --debug hv_prefill=0 --hv-wave 4
run 16 --debug hv_prefill=16 --hv-wave 8
run 128 --debug hv_prefill=0
run 128 --debug hv_prefill=0 --hv-wave 32
run 128 --debug hv_prefill=0 --hv-wave 16
run 128 --debug hv_prefill=16 --debug hv_prefill_after=1
cat $L
