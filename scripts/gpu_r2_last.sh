#!/bin/bash
# Last single-GPU session of round 2: GPU suite, short bench, cfg-4 stage table + phase clocks, ncu of ba_solve_cam_dsm (full set, one launch) and the
# launch list of one cfg-4 two-pass solve.
set -u
T=${1:-r2y}
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests -q -m gpu > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${T}_pytest.log
tail -4 $O/${T}_pytest.log
timeout 100 python __graft_entry__.py --smoke > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${T}_smoke.log
timeout 300 python bench.py --no-cpu-baseline --no-marg --no-detect --no-clahe --no-keyframe > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open('$O/${T}_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms/step',d['ms_per_step'],'e2e',round(d['e2e']['value']),'ba_only',d.get('ba_only',{}).get('solves_per_s'))
print('sharded',{k:v for k,v in d.get('sharded_ba',{}).items() if k!='workload'})
PY
ICG_BA_PROFILE=1 timeout 100 python scripts/prof_ba.py 32 3 20 2000 > $O/${T}_ba_stages_cfg4.log 2>&1
ICG_LIB_VARIANT=prof ICG_BA_PROFILE=1 timeout 100 python scripts/prof_ba.py 32 3 20 2000 > $O/${T}_ba_phase_clocks_cfg4.log 2>&1
grep -E "  solve|step_lm" $O/${T}_ba_stages_cfg4.log | tail -2; sed -n '/ba_solve_cam_dsm phases/,$p' $O/${T}_ba_phase_clocks_cfg4.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:ba_solve_cam_dsm -s 3 -c 1 -f -o $O/${T}_solve_cam_dsm python scripts/prof_ba.py 32 1 20 2000 > $O/${T}_ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_cfg4_launches.csv python scripts/prof_ba.py 32 1 20 2000 > $O/${T}_ncu_list.log 2>&1; echo "ncu list rc=$?"
ls -la $O | tail -12
