#!/bin/bash
set -u
OUT=gpurun_out/r2q
mkdir -p $OUT
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "refinement or all_kinds or golden_files or resident or online or custom_format or odd_restart or not_whole" > $OUT/pytest.log 2>&1
echo "pytest exit $?" >> $OUT/pytest.log
tail -4 $OUT/pytest.log
timeout 600 python tools/trace_layer_a.py > $OUT/layer_a.json 2> $OUT/layer_a.err
grep "bits 1/0" $OUT/layer_a.err | tail -5
grep -A5 progressive_4k $OUT/layer_a.json | head -6
