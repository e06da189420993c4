"""The CLAHE oracle (oracle/clahe_ref.c) against the cv2 golden vectors: bit-exact.  CPU only."""
import os

import numpy as np
import pytest

from tests import oracle_api as oa

GOLD = os.path.join(os.path.dirname(__file__), "golden", "clahe_golden.npz")


def cases():
    g = np.load(GOLD)
    return sorted(k[:-3] for k in g.files if k.endswith("_in"))


@pytest.mark.parametrize("name", cases())
def test_clahe_oracle_matches_cv2_golden(oracle, name):
    g = np.load(GOLD)
    img, ref = g[name + "_in"], g[name + "_out"]
    tx, ty, clip = g[name + "_par"]
    out = oa.clahe_apply(oracle, img, clip, int(tx), int(ty))
    assert np.array_equal(out, ref)


def test_clahe_oracle_in_place(oracle):
    g = np.load(GOLD)
    img, ref = g["ref_1280x560_in"].copy(), g["ref_1280x560_out"]
    oa.clahe_apply(oracle, img, 3.0, 21, 21, in_place=True)
    assert np.array_equal(img, ref)
