#!/bin/bash
# Round-2 single-GPU session F (short): GPU tests, smoke, cfg-4 stage profile + ba_solve_cam phase clocks, short bench.   usage: scripts/gpu_r2_f.sh <tag>
set -u
T=${1:-r2f}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${T}_pytest.log
timeout 200 python __graft_entry__.py --smoke > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${T}_smoke.log
ICG_BA_PROFILE=1 timeout 200 python scripts/prof_ba.py 32 3 20 2000 > $O/${T}_ba_stages_cfg4.log 2>&1
ICG_LIB_VARIANT=prof ICG_BA_PROFILE=1 timeout 200 python scripts/prof_ba.py 32 3 20 2000 > $O/${T}_ba_phase_clocks_cfg4.log 2>&1
timeout 400 python bench.py --no-cpu-baseline --no-marg --no-detect --no-clahe --no-keyframe > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
tail -4 $O/${T}_pytest.log; tail -2 $O/${T}_smoke.log; tail -30 $O/${T}_ba_stages_cfg4.log; tail -30 $O/${T}_ba_phase_clocks_cfg4.log
python - <<PY
import json
d=json.loads(open('$O/${T}_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms/step',d['ms_per_step'],'e2e',round(d['e2e']['value']),'ba_only',d.get('ba_only',{}).get('solves_per_s'))
print('sharded',{k:v for k,v in d.get('sharded_ba',{}).items() if k!='workload'})
PY
