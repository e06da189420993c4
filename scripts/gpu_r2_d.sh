#!/bin/bash
# Round-2 single-GPU session: GPU tests, smoke, bench (both arms), KLT A/B, BA handle sweep + stage profile, marginalization profile,
# ncu --set full of ba_solve / ba_lin_vis / marg_jacobi_cta.   usage: scripts/gpu_r2_d.sh <tag>
set -u
T=${1:-r2d}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/${T}_smi.txt 2>&1
timeout 1500 python -m pytest tests -q -m gpu > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $O/${T}_pytest.log
timeout 300 python __graft_entry__.py --smoke > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/${T}_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?"
for MB in 4 5; do
  ICG_KLT_MINB=$MB timeout 300 python bench.py --steps 10 --warmup 3 --no-ba --no-cpu-baseline --no-detect --no-clahe --no-sharded --no-marg --no-keyframe > $O/${T}_klt_minb$MB.json 2> $O/${T}_klt_minb$MB.err
done
timeout 400 python scripts/prof_ba_handles.py 296 1 2 3 4 6 8 > $O/${T}_ba_handles.log 2>&1
ICG_BA_PROFILE=1 timeout 300 python scripts/prof_ba.py 148 4 > $O/${T}_ba_stages.log 2>&1
timeout 300 python scripts/prof_marg.py 148 4 > $O/${T}_marg.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${T}_marg_launches.csv python scripts/prof_marg.py 148 1 > $O/${T}_ncu_marg.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^ba_solve$ -s 20 -c 1 -f -o $O/${T}_ba_solve python scripts/prof_ba.py 148 1 > $O/${T}_ncu_ba_solve.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^ba_lin_vis$ -s 20 -c 1 -f -o $O/${T}_ba_lin_vis python scripts/prof_ba.py 148 1 > $O/${T}_ncu_ba_lin_vis.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:marg_jacobi_cta -s 2 -c 1 -f -o $O/${T}_marg_jacobi python scripts/prof_marg.py 148 1 > $O/${T}_ncu_marg_jacobi.log 2>&1
tail -5 $O/${T}_pytest.log; tail -2 $O/${T}_smoke.log; cat $O/${T}_ba_handles.log; cat $O/${T}_marg.log | tail -5
for f in $O/${T}_klt_minb4.json $O/${T}_klt_minb5.json; do python -c "
import json,sys
d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', round(d['value']), d['roofline']['kernel_ms'], round(d['roofline']['frac'],4))"; done
python - <<PY
import json
d=json.loads(open('$O/${T}_bench.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'e2e',round(d['e2e']['value']),d['e2e'].get('host_ms_per_step'),'ba_only',d.get('ba_only',{}).get('solves_per_s'))
print('keyframe',d.get('keyframe_path')); print('marg',d.get('marginalization')); print('sharded',d.get('sharded_ba'))
PY
