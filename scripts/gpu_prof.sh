#!/bin/bash
# in-situ BA stage profile (ICG_BA_PROFILE) + BA/marg tests + short bench (+ optional ncu of one kernel: NCU_K=regex).  usage: scripts/gpu_prof.sh <tag>
T=${1:-p}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest ${PYTEST_SEL:-tests/test_ba_gpu.py tests/test_marg_gpu.py} -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
ICG_BA_PROFILE=1 timeout 300 python scripts/prof_ba.py 148 4 > $O/${T}_prof1.log 2>&1
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-sharded > $O/${T}_bench.json 2> $O/${T}_bench.err
if [ -n "${NCU_K:-}" ]; then
timeout 500 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"$NCU_K" -s ${NCU_S:-5} -c ${NCU_C:-1} -f -o $O/${T}_ncu python scripts/prof_ba.py 148 1 > $O/${T}_ncu.log 2>&1
fi
if [ -n "${NCU_KLT:-}" ]; then
timeout 500 ncu --set full --clock-control none --import-source on -k regex:klt_track -s 1 -c 1 -f -o $O/${T}_klt python bench.py --steps 1 --warmup 1 --streams 148 --no-cpu-baseline --no-sharded --no-ba --no-detect > $O/${T}_ncu_klt.log 2>&1
fi
tail -25 $O/${T}_pytest.log; cat $O/${T}_prof1.log
python - "$O/${T}_bench.json" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read())
    print(sys.argv[1], 'value',round(d['value']),'e2e',round(d['e2e']['value']),'klt_only',round(d['klt_only']['value']),'ba',d['ba_only']['ms_per_batch'],round(d['ba_only']['solves_per_s']), 'cost', d['ba_only']['final_cost_mean'], 'marg', d.get('marginalization'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
tail -3 $O/${T}_bench.err
