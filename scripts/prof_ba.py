"""BA-only profiling driver: B windows (cfg 3), one warm-up gvinsOptimization, then one more (profile with ncu -s <launches of the first>)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ic_gvins_b200.ba import WindowSolver, imu_preintegrate
from datagen import synth_ba

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10      # 20 2000: cfg 4 (split pipeline)
L = int(sys.argv[4]) if len(sys.argv) > 4 else 300
NDIST = min(B, 16)                                     # distinct windows, repeated to fill the batch (generation is host-bound)


def pre(st, iewn, g, nz, imu):
    blob, end = imu_preintegrate(st, iewn, g, nz, imu)
    return blob, np.zeros((imu.shape[0] - 1, 4)), end


import copy
base = [synth_ba.make_window(pre, K=K, L=L, seed=2024 + b)[0] for b in range(NDIST)]
wins = [copy.deepcopy(base[b % NDIST]) for b in range(B)]
s = WindowSolver(max_windows=B, max_K=K, max_L=L, max_F=max(w["F"] for w in wins), max_gnss=16, max_marg_r=1)
s.upload(wins)
import time
for r in range(reps):
    s.sync()
    t0 = time.perf_counter()
    s.run_gvins(20, restart=True)
    s.sync()
    print("rep", r, "ms", (time.perf_counter() - t0) * 1e3)
print(s.download(write_back=False)[0])
