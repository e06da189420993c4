#!/bin/bash
# Round-2 closing single-GPU session: GPU suite, smoke, the default bench (the record kept as profiles/r2_bench.json), cfg-4 stage table + phase clocks.
set -u
T=${1:-r2z}
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${T}_pytest.log
tail -4 $O/${T}_pytest.log
timeout 200 python __graft_entry__.py --smoke > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${T}_smoke.log
timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
ICG_BA_PROFILE=1 timeout 200 python scripts/prof_ba.py 32 3 20 2000 > $O/${T}_ba_stages_cfg4.log 2>&1
ICG_LIB_VARIANT=prof ICG_BA_PROFILE=1 timeout 200 python scripts/prof_ba.py 32 3 20 2000 > $O/${T}_ba_phase_clocks_cfg4.log 2>&1
tail -16 $O/${T}_ba_stages_cfg4.log; sed -n '/ba_solve_cam_dsm phases/,$p' $O/${T}_ba_phase_clocks_cfg4.log
python - <<PY
import json
d=json.loads(open('$O/${T}_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms/step',d['ms_per_step'],'e2e',round(d['e2e']['value']),'ba_only',d.get('ba_only',{}).get('solves_per_s'))
print('keyframe',d.get('keyframe_path',{}).get('value')); print('marg',{k:v for k,v in d.get('marginalization',{}).items() if k!='workload'})
print('sharded',{k:v for k,v in d.get('sharded_ba',{}).items() if k!='workload'}); print('roofline',d.get('roofline')); print('cpu',d.get('cpu_baseline',{}).get('value'))
PY
