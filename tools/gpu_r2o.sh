#!/bin/bash
set -u
OUT=gpurun_out/r2o
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest exit $?" >> $OUT/pytest.log
tail -8 $OUT/pytest.log
timeout 600 python tools/time_configs.py 4 > $OUT/cfg4.log 2>&1
tail -1 $OUT/cfg4.log
JPEG_SM100_TRACE=1 timeout 600 python tools/time_configs.py 4 > $OUT/cfg4_trace.log 2>&1
grep "encode scan" $OUT/cfg4_trace.log | tail -4
timeout 600 python tools/time_configs.py 3 5 > $OUT/cfg35.log 2>&1
tail -2 $OUT/cfg35.log
