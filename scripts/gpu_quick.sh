#!/bin/bash
# quick GPU check: tests + probe + short bench variants.  usage: scripts/gpu_quick.sh <tag>
T=${1:-q}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
[ -x scripts/fp64_probe ] && timeout 60 scripts/fp64_probe > $O/${T}_fp64.txt 2>&1
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sharded > $O/${T}_bench.json 2> $O/${T}_bench.err
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sharded --ba-handles 4 --streams 592 > $O/${T}_bench_h4.json 2> $O/${T}_bench_h4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/${T}_ba_launches.csv python scripts/prof_ba.py 148 1 > $O/${T}_ncu_ba.log 2>&1
tail -15 $O/${T}_pytest.log; cat $O/${T}_fp64.txt; for f in $O/${T}_bench.json $O/${T}_bench_h4.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(sys.argv[1], 'value',round(d['value']),'e2e',round(d['e2e']['value']),'klt_only',round(d['klt_only']['value']),'ba',d['ba_only']['ms_per_batch'],round(d['ba_only']['solves_per_s']), 'cost', d['ba_only']['final_cost_mean'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
