"""CPU: oracle/dataloader_ref.py against goldens made by the unmodified reference dataset class
(tests/golden/make_golden_dataloader.py).  Bit-exact, every output array, every (flip, angle) pair."""
import random

import numpy as np

from oracle import dataloader_ref as R
from tests import dataloader_helpers as H


def test_oracle_get_item_matches_reference_goldens():
    g = H.load()
    n = 0
    for name, tag, draws, nf in H.cases(g):
        item = R.get_item(H.raw_sample(g, name), draws, nf)
        assert item["sample_idx"] == name
        for k in H.KEYS:
            H.assert_same(item[k], g["%s_%s_%s" % (name, tag, k)], "%s/%s/%s" % (name, tag, k))
        n += 1
    assert n == 4 * 9


def test_oracle_draw_order_matches_reference_seeds():
    g = H.load()
    for name in g["names"]:
        for seed, d in zip(g["%s_seeds" % name], g["%s_draws" % name]):
            random.seed(int(seed))
            np.random.seed(int(seed))
            flip, angle, scale = R.draw_augmentation(random, np.random)
            assert (flip, R.ROT_ANGLES.index(angle), scale) == (int(d[0]), int(d[1]), float(d[2]))


def test_oracle_height_channel_and_collate():
    g = H.load()
    s = H.raw_sample(g, "s0")
    H.assert_same(R.get_item(s, None, 16, use_height=True)["input_joints"], g["s0_height_noaug_input_joints"], "height/noaug")
    random.seed(int(g["s0_height_aug_seed"]))
    np.random.seed(int(g["s0_height_aug_seed"]))
    d = R.draw_augmentation(random, np.random)
    H.assert_same(R.get_item(s, d, 16, use_height=True)["input_joints"], g["s0_height_aug_input_joints"], "height/aug")
    batch = R.collate([R.get_item(H.raw_sample(g, n), None, 16) for n in ("s0", "s2")])
    for k in H.KEYS:
        H.assert_same(batch[k], g["collate_%s" % k], "collate/" + k)
    assert batch["sample_idx"] == ["s0", "s2"]


def test_frame_ids_properties():
    for n_raw, nf in [(1, 1), (5, 1), (40, 16), (23, 32), (1000, 768), (1024, 1024), (70000, 8)]:
        ids = R.frame_ids(n_raw, nf)
        assert ids.dtype == np.uint16 and len(ids) == nf and ids[0] == 0
        if n_raw <= 65536:
            assert np.all(np.diff(ids.astype(np.int64)) >= 0) and (nf == 1 or ids[-1] == n_raw - 1)
