B="python bench.py --no-cpu-baseline --no-e2e --workload C5 --frames 16"
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items() if "expand" in k})'
for v in 1 2 4 8 16 32; do echo -n "vpw $v: "; $B --steps 100 --warmup 3 --debug hv_expand_vpw=$v 2>&1 | tail -1 | python -c "$P"; done
