"""Shared by the CPU (oracle) and GPU (product) tests of the 1000-scene evaluation set (BASELINE.json config #5):
the fixture with the unmodified reference's outputs (tests/golden/eval1k.npz, made by make_golden_eval1k.py) and the
comparison rules.  Inputs are rebuilt from (seed, scene index) by pose2room_b200.synthetic.make_eval_scene."""
import os.path as osp

import numpy as np

GOLDEN = osp.join(osp.dirname(osp.abspath(__file__)), "golden", "eval1k.npz")
K = 128


def load():
    g = np.load(GOLDEN)
    mask = np.unpackbits(g["pred_mask_bits"], axis=1)[:, :K]
    return g, mask


def check_ap(got_ap, got_map, want_ap, want_map, tol):
    """got_ap: {class: AP} (missing or NaN = class absent); want_ap: (22,) with NaN for absent classes."""
    for c in range(22):
        if np.isnan(want_ap[c]):
            assert c not in got_ap or np.isnan(got_ap[c]), c
        else:
            assert abs(got_ap[c] - want_ap[c]) <= tol, (c, got_ap[c], want_ap[c])
    assert abs(got_map - want_map) <= tol, (got_map, want_map)
