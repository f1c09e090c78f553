#!/bin/bash
# A/B of library variants: tools/ab.sh "<variant names>" "<sweep>"  (device-resident stage times only)
for v in $1; do
  if [ "$v" = base ]; then unset JPEG_SM100_LIB; else export JPEG_SM100_LIB=$PWD/jpeg_b200/libjpeg_sm100_$v.so; fi
  echo "== $v"
  python bench.py --quick --steps 5 --warmup 2 --sweep "$2" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); e = d['env']; print(e.get('JPEG_SM100_PAR_T'), e.get('JPEG_SM100_PAR_WARM'), d['ms_per_step'], d['stages_ms']['huffman'])
    elif 'Error' in l or 'error' in l: print(l.strip())
"
done
