#!/bin/bash
# one gpurun call: GPU tests, K3p sweep, the bench line, the ncu launch list.  Usage: tools/gpu_check.sh <tag> [sweep]
set -u
TAG=${1:-run}
SWEEP=${2:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1
echo "pytest exit $?" >> $OUT/pytest.log
tail -5 $OUT/pytest.log
if [ -n "$SWEEP" ]; then
  timeout 600 python bench.py --quick --steps 5 --warmup 2 --sweep "$SWEEP" > $OUT/sweep.log 2>&1
  cat $OUT/sweep.log | tail -20
  JPEG_SM100_PAR_STATS=1 timeout 600 python bench.py --quick --steps 1 --warmup 1 --sweep "$SWEEP" 2>&1 | grep k_decode_par | uniq > $OUT/stats.log
  cat $OUT/stats.log | tail -20
fi
timeout 900 python bench.py > $OUT/bench.json 2> $OUT/bench.err
tail -c 3000 $OUT/bench.json
tail -5 $OUT/bench.err
