B="python bench.py --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})'
for a in 400 512 700 1024 1400; do echo -n "bucket avg $a: "; $B --steps 100 --warmup 3 --hv-bucket-avg $a 2>&1 | tail -1 | python -c "$P"; done
