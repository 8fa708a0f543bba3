# compute-sanitizer pass over the round-2 kernels: cluster grouping (DSMEM), RoI-aware pooling, the
# lane = word pillar expansion, rolled bitonic merges, host-buffer pipeline
set -x
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_roiaware.py -m gpu -x -q -k "not reference_cuda" 2>&1 | tail -6
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_voxel.py -m gpu -x -q -k "cluster and (caps or special or empty or unaligned or contention or shapes or golden or batched or random_small)" 2>&1 | tail -6
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_voxel.py -m gpu -x -q -k "(launches or bucket_general) and (caps or empty_frames or workspace_query or random_small)" 2>&1 | tail -6
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 python -m pytest tests/test_gpu_voxel.py -m gpu -x -q -k "test_caps and (cluster-5-100000 or cluster-3-150 or bucket_general-64-50 or bucket_general-200-3)" 2>&1 | tail -12
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 python -m pytest tests/test_gpu_roiaware.py -m gpu -x -q -k "forward_vs_oracle and (1-16 or 3-7)" 2>&1 | tail -12
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_voxel.py tests/test_gpu_roiaware.py -m gpu -x -q -k "(test_caps and (cluster-5-100000 or bucket_general-64-50)) or (forward_vs_oracle and 1-16)" 2>&1 | tail -8
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 7 python -m pytest tests/test_gpu_voxel.py tests/test_gpu_roiaware.py tests/test_gpu_packed.py -m gpu -x -q -k "(test_caps and (cluster-5-100000 or bucket_general-64-50)) or (forward_vs_oracle and 1-16) or (host_buffer and C1 and record)" 2>&1 | tail -10
