# compute-sanitizer pass over the kernels changed in the second session of round 2: record expansion with the
# round-robin tile mapping (all three maps), the flat pillar expansion, the evict_last prefetch
cd /root/repo
mkdir -p gpurun_out
set -x
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_voxel.py -m gpu -x -q -k "(launches or map0 or map1) and (caps or empty or unaligned or shapes or batched or random_small or many_frames or golden)" 2>&1 | tail -5
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_voxel.py tests/test_gpu_packed.py tests/test_gpu_vfe.py -m gpu -x -q -k "(long_voxels and launches) or (packed_ragged and (record or map0)) or (fused_batch and record)" 2>&1 | tail -5
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 python -m pytest tests/test_gpu_voxel.py -m gpu -x -q -k "(test_caps and (launches-5-100000 or launches-3-150 or map1-5-100000 or bucket_general-64-50)) or (long_voxels and launches and 64-5)" 2>&1 | tail -8
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_voxel.py -m gpu -x -q -k "(test_caps and (launches-5-100000 or bucket_general-64-50)) or (long_voxels and launches and 64-5)" 2>&1 | tail -5
