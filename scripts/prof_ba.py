"""BA-only profiling driver: B windows (cfg 3), one warm-up gvinsOptimization, then one more (profile with ncu -s <launches of the first>)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ic_gvins_b200.ba import WindowSolver, imu_preintegrate
from datagen import synth_ba

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2


def pre(st, iewn, g, nz, imu):
    blob, end = imu_preintegrate(st, iewn, g, nz, imu)
    return blob, np.zeros((imu.shape[0] - 1, 4)), end


wins = [synth_ba.make_window(pre, K=10, L=300, seed=2024 + b)[0] for b in range(B)]
s = WindowSolver(max_windows=B, max_K=10, max_L=300, max_F=max(w["F"] for w in wins), max_gnss=8, max_marg_r=1)
s.upload(wins)
import time
for r in range(reps):
    s.sync()
    t0 = time.perf_counter()
    s.run_gvins(20, restart=True)
    s.sync()
    print("rep", r, "ms", (time.perf_counter() - t0) * 1e3)
print(s.download(write_back=False)[0])
