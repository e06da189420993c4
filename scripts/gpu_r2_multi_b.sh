#!/bin/bash
# Short multi-GPU session (charged N x): the bench at N ranks first (cfg-4 landmark-sharded section on the peer-memory transport), then the sharded parity test.
set -u
T=${1:-r2mb}
N=${2:-2}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m > $O/${T}_smi.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 6 --warmup 3 \
    --no-detect --no-clahe --no-marg --no-keyframe > $O/${T}_bench${N}.json 2> $O/${T}_bench${N}.err; echo "bench rc=$?" >> $O/${T}_bench${N}.err
python - <<PY
import json
try:
    d = json.loads(open("$O/${T}_bench${N}.json").read().strip().splitlines()[-1])
    print("N", d["n_gpus"], "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "sharded", {k: v for k, v in d["sharded_ba"].items() if k != "workload"})
except Exception as e:
    print("bench parse failed", e)
PY
tail -3 $O/${T}_bench${N}.err
timeout 200 python -m pytest "tests/test_ba_multigpu.py" -m gpu -q -x -k "cfg4" > $O/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> $O/${T}_pytest.log
tail -4 $O/${T}_pytest.log
