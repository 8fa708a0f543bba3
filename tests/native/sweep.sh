P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items() if k.startswith("hvb")})'
for v in $(ls detmatch_b200/lib/ | sed 's/.so//'); do echo -n "$v: "; PCFE_LIB=$PWD/detmatch_b200/lib/$v.so python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "$P"; done
