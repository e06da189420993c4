import sys, os, copy, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from tests import oracle_api as oa
from datagen import synth_ba
import oracle
olib = C.CDLL(oracle.build()); oa.declare(olib); oa.declare_ba(olib)
from ic_gvins_b200.ba import WindowSolver
prob, truth = synth_ba.make_window(lambda *a: oa.preintegrate(olib, *a), K=10, L=300, seed=2024)
prob["ext_const"], prob["td_const"] = 1, 1
s = WindowSolver(max_windows=2)
for it in (1, 5, 20):
    po, pg = copy.deepcopy(prob), copy.deepcopy(prob)
    so = oa.ba_solve(olib, po, it)
    sg = s.solve(pg, it)[0]
    print(it, "oracle", so)
    print(it, "gpu   ", sg)
    for k in ("pose", "mix", "invdepth"):
        print("   ", k, np.abs(pg[k] - po[k]).max() / np.abs(po[k]).max())
