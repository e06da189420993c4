import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ic_gvins_b200.klt import KltTracker
g = np.load("tests/golden/klt_golden.npz")
name = "small_plain"
f0, f1, p0, init = g[name + "_f0"], g[name + "_f1"], g[name + "_p0"], g[name + "_init"]
t = KltTracker(320, 240)
q, st, err = t.calcOpticalFlowPyrLK(f0, f1, p0, init, flags=4)
print("status", st.sum(), (st == g[name + "_st"]).all(), np.abs(q - g[name + "_fwd"]).max())
