"""N > 1 host logic on CPU (gloo, world_size 2): the landmark partition used by the sharded window solve covers every landmark
and factor exactly once, the camera-side blocks are replicated, and merging the shards reproduces the problem."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from datagen import synth_ba


def _fake_preintegrate(st, iewn, g, nz, imu):
    return np.zeros(480), np.zeros((imu.shape[0] - 1, 4)), np.zeros(10)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import torch
    from ic_gvins_b200.ba import merge_shard, shard_window
    prob, _ = synth_ba.make_window(_fake_preintegrate, K=10, L=301, seed=5)
    sh = shard_window(prob, rank, world)
    counts = torch.tensor([sh["L"], sh["F"]], dtype=torch.int64)
    dist.all_reduce(counts)
    assert counts.tolist() == [prob["L"], prob["F"]]
    # every factor of a local landmark is local, indices remapped into [0, L_local)
    assert sh["f_lm"].min() >= 0 and sh["f_lm"].max() < sh["L"]
    assert np.array_equal(np.asarray(prob["f_lm"])[sh["f_index"]] - sh["lm_lo"], sh["f_lm"])
    assert np.array_equal(sh["pose"], prob["pose"]) and np.array_equal(sh["mix"], prob["mix"])
    # "solve": perturb the local landmarks, gather, merge
    sh["invdepth"] = sh["invdepth"] * (1.0 + 0.01 * (rank + 1))
    parts = [None] * world
    dist.all_gather_object(parts, {k: sh[k] for k in ("lm_lo", "lm_hi", "invdepth", "f_index", "f_active", "pose", "mix", "ext", "gnss_std")})
    full = {k: np.array(v, copy=True) if isinstance(v, np.ndarray) else v for k, v in prob.items()}
    for p in parts:
        merge_shard(full, p)
    lo = [(prob["L"] * r) // world for r in range(world + 1)]
    for r in range(world):
        assert np.allclose(full["invdepth"][lo[r]:lo[r + 1]], prob["invdepth"][lo[r]:lo[r + 1]] * (1.0 + 0.01 * (r + 1)))
    q.put(rank)
    dist.destroy_process_group()


def test_landmark_partition_world2():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get() for _ in range(2)) == [0, 1]
