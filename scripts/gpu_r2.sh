#!/bin/bash
# Round-2 GPU session: tests, smoke, bench (both arms).  usage: scripts/gpu_r2.sh <tag> [pytest-args]
set -u
T=${1:-r2}
shift || true
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv > $O/${T}_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q --durations=15 "$@" > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
timeout 300 python __graft_entry__.py --smoke > $O/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> $O/${T}_smoke.log
if [ "${BENCH:-1}" = "1" ]; then
timeout 900 python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err; echo "bench rc=$?" >> $O/${T}_bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
fi
tail -15 $O/${T}_pytest.log; tail -2 $O/${T}_smoke.log; cut -c1-600 $O/${T}_bench.json; tail -3 $O/${T}_bench.err; cut -c1-300 $O/${T}_bench_ref.json
