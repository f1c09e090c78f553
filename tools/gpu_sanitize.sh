#!/bin/bash
# compute-sanitizer (memcheck, then racecheck on the shared-memory heavy kernels) over a small, representative subset of the GPU tests
set -u
OUT=gpurun_out/sanitize
mkdir -p $OUT
SEL='golden_files or fused_idct or refinement_three_phase or progressive_ac_encoders or not_whole_rows or odd_restart or parallel_variants or first_scan_extend'
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck.log \
    python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q -k "$SEL and not 3840 and not 1936 and not 2048" > $OUT/memcheck_pytest.log 2>&1
echo "memcheck exit $?" | tee -a $OUT/memcheck_pytest.log
tail -3 $OUT/memcheck_pytest.log
grep -c "Invalid\|out of bounds\|misaligned" $OUT/memcheck.log
tail -5 $OUT/memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck.log \
    python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "refinement_three_phase and 1-0 or all_kinds and progressive and 0" > $OUT/racecheck_pytest.log 2>&1
echo "racecheck exit $?" | tee -a $OUT/racecheck_pytest.log
tail -3 $OUT/racecheck_pytest.log
tail -5 $OUT/racecheck.log
