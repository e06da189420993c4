"""Two (or more) solver handles on separate streams, B windows each: do latency-bound BA kernels overlap across streams?"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ic_gvins_b200.ba import WindowSolver, imu_preintegrate
from datagen import synth_ba
B = int(sys.argv[1]); NS = int(sys.argv[2])
def pre(st, iewn, g, nz, imu):
    blob, end = imu_preintegrate(st, iewn, g, nz, imu)
    return blob, np.zeros((imu.shape[0] - 1, 4)), end
wins = [synth_ba.make_window(pre, K=10, L=300, seed=2024 + b)[0] for b in range(min(B, 32))]
wins = [wins[i % len(wins)] for i in range(B)]
solvers = []
for k in range(NS):
    s = WindowSolver(max_windows=B, max_K=10, max_L=300, max_F=max(w["F"] for w in wins), max_gnss=8, max_marg_r=1)
    s.upload(wins)
    solvers.append(s)
for r in range(3):
    for s in solvers: s.sync()
    t0 = time.perf_counter()
    for s in solvers: s.run_gvins(20, restart=True)
    for s in solvers: s.sync()
    dt = (time.perf_counter() - t0) * 1e3
    print(f"B={B} x {NS} streams: {dt:.2f} ms -> {B*NS/dt*1e3:.0f} solves/s")
