#!/bin/bash
# 8-GPU session (charged 8x: kept short): the all-GPU sharded parity test and the bench at N = 8.   usage: scripts/gpu_r2_multi8.sh <tag> [N]
set -u
T=${1:-r2m8}
N=${2:-8}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/${T}_smi.txt 2>&1
timeout 600 python -m pytest "tests/test_ba_multigpu.py::test_sharded_cfg4_matches_oracle[0]" -m gpu -q > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 6 --warmup 3 \
    --no-detect --no-clahe --no-marg --no-keyframe > $O/${T}_bench${N}.json 2> $O/${T}_bench${N}.err; echo "bench rc=$?" >> $O/${T}_bench${N}.err
tail -5 $O/${T}_pytest.log; python - <<PY
import json
try:
    d = json.loads(open("$O/${T}_bench${N}.json").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "sharded", {k: v for k, v in d["sharded_ba"].items() if k != "workload"})
except Exception as e:
    print("bench parse failed", e)
PY
tail -4 $O/${T}_bench${N}.err
