#!/bin/bash
# Closing check of a ba_solve_cam_dsm variant + the block-detection section of the bench once more.
set -u
T=${1:-r2x}
O=gpurun_out
mkdir -p $O
timeout 400 python -m pytest tests -q -m gpu > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${T}_pytest.log
tail -3 $O/${T}_pytest.log
ICG_BA_PROFILE=1 timeout 100 python scripts/prof_ba.py 32 3 20 2000 > $O/${T}_ba_stages_cfg4.log 2>&1
ICG_LIB_VARIANT=prof ICG_BA_PROFILE=1 timeout 100 python scripts/prof_ba.py 32 3 20 2000 > $O/${T}_ba_phase_clocks_cfg4.log 2>&1
grep -E "  solve|step_lm" $O/${T}_ba_stages_cfg4.log | tail -2; sed -n '/ba_solve_cam_dsm phases/,$p' $O/${T}_ba_phase_clocks_cfg4.log
timeout 200 python bench.py --no-cpu-baseline --no-marg --no-clahe --no-keyframe > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('$O/${T}_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),'ba_only',d.get('ba_only',{}).get('solves_per_s'))
print('sharded',{k:v for k,v in d.get('sharded_ba',{}).items() if k!='workload'})
print('detect',{k:v for k,v in d.get('detection',{}).items() if k!='workload'})
PY
