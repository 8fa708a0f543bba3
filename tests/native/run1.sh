B="python bench.py --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items() if "expand" in k})'
for w in 64 32 16; do
for cfg in "hv_expand_pad_kb=0" "hv_expand_pad_kb=16" "hv_expand_pad_kb=24" "hv_expand_pad_kb=43"; do
  a=""; for kv in $cfg; do a="$a --debug $kv"; done
  echo -n "== wave $w $cfg: "; $B --steps 100 --warmup 3 --hv-wave $w $a 2>&1 | tail -1 | python -c "$P"
done; done
