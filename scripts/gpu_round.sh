#!/bin/bash
# One GPU session: tests, smoke, bench (both arms), ncu launch list of the bench command, ncu full capture of the KLT kernel.
# usage: scripts/gpu_round.sh <tag>     (outputs under gpurun_out/<tag>_*)
set -u
T=${1:-r1}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $O/${T}_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
timeout 300 python __graft_entry__.py --smoke > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${T}_smoke.log
timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?" >> $O/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
if [ "${NCU:-1}" = "1" ]; then
ICG_BA_PROFILE=1 timeout 300 python scripts/prof_ba.py 148 4 > $O/${T}_ba_stages.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/${T}_launches.csv \
   python bench.py --steps 2 --warmup 1 --streams 148 --no-cpu-baseline --no-sharded --no-marg --no-detect --no-clahe > $O/${T}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:klt_track -s 1 -c 1 -f -o $O/${T}_klt_track \
   python bench.py --steps 1 --warmup 1 --streams 148 --no-cpu-baseline --no-sharded --no-ba --no-detect --no-clahe > $O/${T}_ncu_klt.log 2>&1
fi
tail -8 $O/${T}_pytest.log; tail -2 $O/${T}_smoke.log; cat $O/${T}_bench.json | cut -c1-400; tail -3 $O/${T}_bench.err
