#!/bin/bash
# Round-2 single-GPU session E: GPU tests, bench, BA stage profile + handle / batch sweep, marginalization timing.   usage: scripts/gpu_r2_e.sh <tag>
set -u
T=${1:-r2e}
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${T}_pytest.log
timeout 300 python __graft_entry__.py --smoke > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${T}_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/${T}_launches.csv \
   python bench.py --steps 2 --warmup 1 --streams 148 --no-cpu-baseline --no-sharded --no-marg --no-detect --no-clahe --no-keyframe > $O/${T}_ncu_bench.log 2>&1
ICG_BA_PROFILE=1 timeout 300 python scripts/prof_ba.py 148 4 > $O/${T}_ba_stages.log 2>&1
ICG_LIB_VARIANT=prof ICG_BA_PROFILE=1 timeout 300 python scripts/prof_ba.py 148 4 > $O/${T}_ba_phase_clocks.log 2>&1
ICG_BA_PROFILE=1 timeout 300 python scripts/prof_ba.py 32 3 20 2000 > $O/${T}_ba_stages_cfg4.log 2>&1
timeout 400 python scripts/prof_ba_handles.py 296 1 2 3 > $O/${T}_ba_handles.log 2>&1
timeout 400 python scripts/prof_ba_handles.py 592 2 4 >> $O/${T}_ba_handles.log 2>&1
timeout 300 python scripts/prof_marg.py 148 3 > $O/${T}_marg.log 2>&1
tail -5 $O/${T}_pytest.log; tail -2 $O/${T}_smoke.log; cat $O/${T}_ba_handles.log; tail -3 $O/${T}_marg.log; tail -12 $O/${T}_ba_stages.log
python - <<PY
import json
d=json.loads(open('$O/${T}_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'ms/step',d['ms_per_step'],'e2e',round(d['e2e']['value']),d['e2e'].get('host_ms_per_step'),'ba_only',d.get('ba_only',{}).get('solves_per_s'))
print('keyframe',d.get('keyframe_path',{}).get('value')); print('marg',{k:v for k,v in d.get('marginalization',{}).items() if k!='workload'}); print('sharded',{k:v for k,v in d.get('sharded_ba',{}).items() if k!='workload'})
PY
