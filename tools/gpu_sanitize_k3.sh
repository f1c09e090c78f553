#!/bin/bash
# compute-sanitizer over the tests that drive the subsequence-parallel entropy decoder (memcheck, then racecheck)
set -u
OUT=gpurun_out/sanitize_k3
mkdir -p $OUT
SEL='parallel_variants or sequential_scan_geometries or first_scan_extend or scan_decode_errors'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $OUT/memcheck.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > $OUT/memcheck_pytest.log 2>&1
echo "memcheck exit $?"; tail -2 $OUT/memcheck_pytest.log; tail -2 $OUT/memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $OUT/racecheck.log python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "parallel_variants and 4-32 or sequential_scan_geometries and rows1" > $OUT/racecheck_pytest.log 2>&1
echo "racecheck exit $?"; tail -2 $OUT/racecheck_pytest.log; tail -4 $OUT/racecheck.log
