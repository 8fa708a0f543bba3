# compute-sanitizer pass over a small but representative slice of the GPU parity tests
set -x
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_voxel.py -m gpu -x -q -k "(caps or special or empty or unaligned or contention or shapes or golden or batched) and not exhaustive" 2>&1 | tail -15
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_pib.py -m gpu -x -q -k "golden or shapes or empty or faces or margin" 2>&1 | tail -8
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 python -m pytest tests/test_gpu_voxel.py -m gpu -x -q -k "test_caps and (bucket-3-150 or bucket-5-100000 or bucket-64-50 or pipeline-5-100000)" 2>&1 | tail -25
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_voxel.py -m gpu -x -q -k "test_caps and (bucket-5-100000 or bucket-64-50 or pipeline-5-100000)" 2>&1 | tail -10
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 7 python -m pytest tests/test_gpu_voxel.py -m gpu -x -q -k "test_caps and (bucket-5-100000 or bucket-64-50)" 2>&1 | tail -12
