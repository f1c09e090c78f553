#!/bin/bash
set -u
OUT=gpurun_out/r2l
mkdir -p $OUT
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest exit $?" >> $OUT/pytest.log
tail -8 $OUT/pytest.log
timeout 600 python tools/trace_layer_a.py > $OUT/layer_a.json 2> $OUT/layer_a.err
grep "bits 1/0" $OUT/layer_a.err | tail -5
grep -A5 progressive_4k $OUT/layer_a.json | head -6
for st in 0 1; do
  JPEG_BENCH_STAGED=$st timeout 600 python bench.py --quick --steps 10 --warmup 3 > $OUT/quick_staged$st.json 2> $OUT/quick_staged$st.err
  tail -c 600 $OUT/quick_staged$st.json; tail -3 $OUT/quick_staged$st.err
done
for b in 8 12 24 34; do
  JPEG_SM100_FUSE_BAND=$b timeout 600 python bench.py --quick --steps 10 --warmup 3 2>/dev/null | tail -c 330
done
