#!/bin/bash
set -u
OUT=gpurun_out/r2k
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest exit $?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
JPEG_SM100_PAR_STATS=1 timeout 600 python tools/trace_layer_a.py > $OUT/layer_a.json 2> $OUT/layer_a.err
grep -B4 "band 0..1 bits -1/1" $OUT/layer_a.err | tail -5
grep "bits 1/0" $OUT/layer_a.err | tail -5
grep -A8 progressive_4k $OUT/layer_a.json | head -5
timeout 600 python tools/time_configs.py 4 > $OUT/cfg4.log 2>&1
tail -3 $OUT/cfg4.log
