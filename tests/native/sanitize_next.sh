# compute-sanitizer pass over the packed / mean / OpenPCDet additions (full-size C4 tests excluded)
set -x
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_packed.py tests/test_gpu_vfe.py tests/test_gpu_pcdet_pib.py -m gpu -x -q -k "not full_c4 and not reference_test_shape" 2>&1 | tail -12
timeout 600 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 7 python -m pytest tests/test_gpu_packed.py -m gpu -x -q -k "overflowing and (record or fallback)" 2>&1 | tail -15
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_packed.py tests/test_gpu_vfe.py -m gpu -x -q -k "(overflowing or fused_batch) and (record or fallback)" 2>&1 | tail -8
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 7 python -m pytest tests/test_gpu_packed.py tests/test_gpu_vfe.py -m gpu -x -q -k "(ragged and record and C4) or (fused_batch and record and C1)" 2>&1 | tail -10
