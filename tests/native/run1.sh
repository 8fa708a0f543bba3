B="python bench.py --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["roofline"]["frac"])'
for v in libpcfe libpcfe_m5 libpcfe_m6; do
for cfg in "mega_d1=3 mega_d2=6 mega_d3=9 mega_ring=12" "mega_d1=4 mega_d2=8 mega_d3=12 mega_ring=16"; do
  a=""; for kv in $cfg; do a="$a --debug $kv"; done
  echo "== $v $cfg"; PCFE_LIB=$PWD/detmatch_b200/lib/$v.so $B --steps 50 --warmup 3 $a 2>&1 | tail -1 | python -c "$P"
  PCFE_LIB=$PWD/detmatch_b200/lib/$v.so $B --steps 3 --warmup 2 --debug mega_stats=1 $a 2>&1 | grep hvm | tail -5
done; done
