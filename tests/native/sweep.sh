P='import json,sys; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items() if k.startswith("hvb")})'
for v in libpcfe v_bt128 v_bt512 v_escalar v_echunk4 v_etiles4 v_ewarps8; do echo -n "$v: "; PCFE_LIB=$PWD/detmatch_b200/lib/$v.so python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "$P"; done
for a in 512 1400; do echo -n "avg $a: "; python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-e2e --hv-bucket-avg $a 2>&1 | tail -1 | python -c "$P"; done
