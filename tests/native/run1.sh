timeout 900 python -m pytest tests/test_gpu_voxel.py -m gpu -x -q -k "bucket and not general and not exhaustive" 2>&1 | tail -3
python tests/native/bench_ops.py 2>&1 | grep -A1 -E "C5"
