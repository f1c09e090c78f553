#!/bin/bash
# one gpurun call: all GPU tests, then the bench line (N = 1)
set -u
TAG=${1:-check}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 1800 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest exit $?" >> $OUT/pytest.log
tail -6 $OUT/pytest.log
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
tail -c 4500 $OUT/bench.json
tail -3 $OUT/bench.err
