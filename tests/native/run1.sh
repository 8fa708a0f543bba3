B="python bench.py --no-cpu-baseline --no-e2e"
P='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["roofline"]["frac"], {k:v["ms_per_step"] for k,v in d["kernels"].items()})'
timeout 600 python -m pytest tests/test_gpu_voxel.py -m gpu -x -q -k "not exhaustive" 2>&1 | tail -2
$B --steps 500 --warmup 5 2>&1 | tail -1 | python -c "$P"
