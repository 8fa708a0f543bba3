timeout 900 python -m pytest tests/test_gpu_pib.py -m gpu -x -q 2>&1 | tail -3
python tests/native/bench_ops.py 2>&1 | grep -E "points_in_boxes"
