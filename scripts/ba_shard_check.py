"""torchrun --nproc-per-node N scripts/ba_shard_check.py : landmark-sharded window solve vs the CPU oracle (run on N GPUs)."""
import copy, ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from ic_gvins_b200.ba import WindowSolver, merge_shard, nccl_unique_id, shard_window
from datagen import synth_ba
from tests import oracle_api as oa
import oracle
olib = C.CDLL(oracle.build()); oa.declare(olib); oa.declare_ba(olib)
K, L = (20, 2000) if "--cfg4" in sys.argv else (10, 300)
prob, _ = synth_ba.make_window(lambda *a: oa.preintegrate(olib, *a), K=K, L=L, seed=2024)
prob["ext_const"], prob["td_const"] = 1, 1
ids = [nccl_unique_id() if rank == 0 else None]
dist.broadcast_object_list(ids, src=0)
sh = shard_window(prob, rank, world)
s = WindowSolver(max_windows=1, max_K=K, max_L=sh["L"], max_F=max(1, sh["F"]), max_gnss=16, max_marg_r=1, device=local)
s.set_shard(rank, world, ids[0])
summ = s.solve(sh, 20)[0]
# gather the landmark results
parts = [None] * world
dist.all_gather_object(parts, (sh["lm_lo"], sh["lm_hi"], sh["invdepth"], sh["pose"]))
if rank == 0:
    full = copy.deepcopy(prob)
    for lo, hi, rho, pose in parts:
        full["invdepth"][lo:hi] = rho
        assert np.array_equal(pose, parts[0][3]), "camera blocks differ between shards"
    full["pose"] = parts[0][3]
    po = copy.deepcopy(prob)
    so = oa.ba_solve(olib, po, 20)
    rel = lambda a, b: float(np.abs(a - b).max() / np.abs(b).max())
    print(f"world={world} K={K} L={L} F={prob['F']}: gpu {summ['iterations']} its cost {summ['final_cost']:.9f} | oracle {so['iterations']} its cost {so['final_cost']:.9f}")
    print("  rel err pose", rel(full["pose"], po["pose"]), "invdepth", rel(full["invdepth"], po["invdepth"]))
    assert summ["iterations"] == so["iterations"] and rel(full["pose"], po["pose"]) < 1e-6 and rel(full["invdepth"], po["invdepth"]) < 1e-6
    print("SHARD PARITY OK")
dist.destroy_process_group()
