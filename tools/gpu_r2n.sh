#!/bin/bash
set -u
OUT=gpurun_out/r2n
mkdir -p $OUT
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest exit $?" >> $OUT/pytest.log
tail -8 $OUT/pytest.log
timeout 600 python tools/trace_layer_a.py > $OUT/layer_a.json 2> $OUT/layer_a.err
grep "bits 1/0" $OUT/layer_a.err | tail -5
grep -A5 progressive_4k $OUT/layer_a.json | head -6
timeout 600 python tools/time_configs.py 4 > $OUT/cfg4.log 2>&1
tail -2 $OUT/cfg4.log
