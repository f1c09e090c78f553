#!/bin/bash
# one gpurun call at the end of a session: GPU tests, smoke, the bench line, then the ncu launch list and --set full capture
set -u
OUT=gpurun_out/final
mkdir -p $OUT
timeout 600 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -3 $OUT/pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/smoke.log 2>&1; tail -1 $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench_n1.json 2> $OUT/bench_n1.err; tail -c 2500 $OUT/bench_n1.json; tail -3 $OUT/bench_n1.err
bash tools/gpu_profile.sh > $OUT/profile.log 2>&1; tail -4 $OUT/profile.log
