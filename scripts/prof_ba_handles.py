"""BA-only throughput against the number of solver handles (CUDA streams) the batch is split over: B cfg-3 windows in total, two-pass
gvinsOptimization (5 + 15 iterations) with restart, device-timed.  usage: prof_ba_handles.py [B] [handles ...]"""
import copy
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from datagen import synth_ba
from ic_gvins_b200.ba import WindowSolver, imu_preintegrate

B = int(sys.argv[1]) if len(sys.argv) > 1 else 296
NHS = [int(a) for a in sys.argv[2:]] or [1, 2, 3, 4, 6, 8]


def pre(st, iewn, g, nz, imu):
    blob, end = imu_preintegrate(st, iewn, g, nz, imu)
    return blob, np.zeros((imu.shape[0] - 1, 4)), end


base = [synth_ba.make_window(pre, K=10, L=300, seed=2024 + b)[0] for b in range(16)]
wins = [copy.deepcopy(base[b % 16]) for b in range(B)]
maxF = max(w["F"] for w in wins)
dev = torch.device("cuda", 0)
main = torch.cuda.Stream(device=dev)
for NH in NHS:
    bounds = [(B * k) // NH for k in range(NH + 1)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(NH)]
    svs = []
    for k in range(NH):
        part = wins[bounds[k]:bounds[k + 1]]
        sv = WindowSolver(max_windows=len(part), max_K=10, max_L=300, max_F=maxF, max_gnss=8, max_marg_r=1, stream=streams[k].cuda_stream)
        sv.upload(part)
        sv.sync()
        svs.append(sv)
    ms = []
    for rep in range(6):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record(main)
        for st in streams:
            st.wait_stream(main)
        for sv in svs:
            sv.run_gvins(20, restart=True)
        for st in streams:
            main.wait_stream(st)
        b.record(main)
        torch.cuda.synchronize()
        ms.append(a.elapsed_time(b))
    t = float(np.median(ms[2:]))
    print(f"handles {NH:2d}  windows/handle {B // NH:4d}  ms/batch {t:8.3f}  solves/s {B / t * 1e3:9.0f}", flush=True)
    for sv in svs:
        sv.close()
