B="python bench.py --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items() if "expand" in k})'
for v in 0 1 2 4 8; do echo -n "prefetch $v: "; $B --steps 300 --warmup 5 --debug hv_expand_prefetch=$v 2>&1 | tail -1 | python -c "$P"; done
