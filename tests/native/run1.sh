B="python bench.py --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})'
for v in libpcfe libpcfe_b512; do for a in 1024 1600; do echo -n "$v avg $a: "; PCFE_LIB=$PWD/detmatch_b200/lib/$v.so $B --steps 200 --warmup 5 --hv-bucket-avg $a 2>&1 | tail -1 | python -c "$P"; done; done
