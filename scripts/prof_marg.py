"""Marginalization profiling driver: B cfg-3 windows through icg_ba_marginalize (host API), timed per call.  Run under
`ncu --metrics gpu__time_duration.sum` for the kernel list."""
import copy
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from datagen import synth_ba
from ic_gvins_b200.ba import WindowSolver, imu_preintegrate

B = int(sys.argv[1]) if len(sys.argv) > 1 else 148
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3


def pre(st, iewn, g, nz, imu):
    blob, end = imu_preintegrate(st, iewn, g, nz, imu)
    return blob, np.zeros((imu.shape[0] - 1, 4)), end


base = [synth_ba.make_window(pre, K=10, L=300, seed=2024 + b)[0] for b in range(min(B, 16))]
wins = [copy.deepcopy(base[b % len(base)]) for b in range(B)]
s = WindowSolver(max_windows=B, max_K=10, max_L=300, max_F=max(w["F"] for w in wins), max_gnss=8, max_marg_r=1)
s.marginalize(wins[:2], 1, want_schur=False)
call = s.marg_prepare(wins, 1, want_schur=False)
for r in range(reps):
    t0 = time.perf_counter()
    s.marg_run(call)
    t1 = time.perf_counter()
    s.marg_run(call, resident=True)
    t2 = time.perf_counter()
    pri = s.marg_collect(call)
    print("rep", r, "icg_ba_marginalize ms", (t1 - t0) * 1e3, "resident ms", (t2 - t1) * 1e3, "m", [p["m"] for p in pri[:8]], "r", [p["r"] for p in pri[:8]])
