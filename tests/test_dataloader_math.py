"""CPU: the host-side half of the sample -> batch path (pose2room_b200/dataloader.py) and the kernel's
per-joint arithmetic (csrc/augment_math.h, compiled for the host by this test) against goldens made by the
unmodified reference dataset class.  Bit-exact.  The CUDA kernel itself is covered by test_dataloader_gpu.py."""
import ctypes
import os.path as osp
import random
import subprocess

import numpy as np
import pytest
import torch

from pose2room_b200 import dataloader as DL
from pose2room_b200 import synthetic
from tests import dataloader_helpers as H

ROOT = osp.dirname(osp.dirname(osp.abspath(__file__)))


@pytest.fixture(scope="module")
def host_lib(tmp_path_factory):
    so = str(tmp_path_factory.mktemp("dl") / "make_batch_host.so")
    subprocess.run(["gcc", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-I", osp.join(ROOT, "pose2room_b200", "csrc"),
                    osp.join(ROOT, "tests", "csrc", "make_batch_host.c"), "-o", so, "-lm"], check=True)
    lib = ctypes.CDLL(so)
    lib.host_frame_id.restype = ctypes.c_int
    return lib


def host_make_batch(lib, store, indices, draws, num_frames, use_height=False):
    ds = H.dataset_for(store, num_frames, use_height, aug=any(d is not None for d in draws))
    params, labels, classes = H.host_side(ds, indices, draws)
    B, J, C = len(indices), store.num_joints, 4 if use_height else 3
    ids = np.asarray(indices, np.int32)
    oj = np.empty((B, num_frames, J, C), np.float32)
    ov = np.empty((B, num_frames, J, 9), np.float32)
    om = np.empty((B, num_frames, J), np.int64)
    ptr = lambda a: a.ctypes.data_as(ctypes.c_void_p)
    joints, votes = np.ascontiguousarray(store.joints), np.ascontiguousarray(store.votes)
    lib.host_make_batch(ptr(joints), ptr(votes), ptr(store.frame_start), ptr(ids), ptr(params), B, num_frames, J, C,
                        ptr(oj), ptr(ov), ptr(om))
    return oj, ov, om, labels, classes


def test_per_joint_arithmetic_and_box_labels_match_reference_goldens(host_lib):
    g = H.load()
    names = [str(n) for n in g["names"]]
    stores = {n: DL.PackedSamples.from_samples([H.raw_sample(g, n)]) for n in names}
    for name, tag, draws, nf in H.cases(g):
        oj, ov, om, labels, classes = host_make_batch(host_lib, stores[name], [0], [draws], nf)
        pre = "%s_%s_" % (name, tag)
        H.assert_same(oj[0], g[pre + "input_joints"], pre + "input_joints")
        H.assert_same(ov[0], g[pre + "vote_label"], pre + "vote_label")
        H.assert_same(om[0], g[pre + "vote_label_mask"], pre + "vote_label_mask")
        H.assert_same(labels[0, :, 0], g[pre + "box_label_mask"], pre + "box_label_mask")
        H.assert_same(labels[0, :, 1:4], g[pre + "center_label"], pre + "center_label")
        H.assert_same(labels[0, :, 4:7], g[pre + "size"], pre + "size")
        H.assert_same(labels[0, :, 7:9], g[pre + "heading"], pre + "heading")
        H.assert_same(classes[0], g[pre + "sem_cls_label"], pre + "sem_cls_label")


def test_packed_batch_of_ragged_samples_and_height_channel(host_lib):
    g = H.load()
    store = DL.PackedSamples.from_samples([H.raw_sample(g, n) for n in ("s0", "s1", "s2")])   # 40 / 23 / 64 raw frames
    assert list(store.frame_start) == [0, 40, 63, 127] and store.names == ["s0", "s1", "s2"]
    oj, ov, om, labels, _ = host_make_batch(host_lib, store, [2, 0], [None, None], 16)
    for b, n in enumerate(("s2", "s0")):
        H.assert_same(oj[b], g["%s_noaug_input_joints" % n], "ragged/" + n)
        H.assert_same(ov[b], g["%s_noaug_vote_label" % n], "ragged/" + n)
    H.assert_same(labels[:, :, 1:4], g["collate_center_label"][::-1], "ragged/center")
    # optional height channel, un-augmented (float32 path) and augmented (float64 path)
    oj, _, _, _, _ = host_make_batch(host_lib, store, [0], [None], 16, use_height=True)
    H.assert_same(oj[0], g["s0_height_noaug_input_joints"], "height/noaug")
    random.seed(int(g["s0_height_aug_seed"]))
    np.random.seed(int(g["s0_height_aug_seed"]))
    oj, _, _, _, _ = host_make_batch(host_lib, store, [0], [DL.draw_augmentation()], 16, use_height=True)
    H.assert_same(oj[0], g["s0_height_aug_input_joints"], "height/aug")


def test_frame_id_matches_numpy_linspace(host_lib):
    rng = np.random.default_rng(5)
    pairs = [(1, 1), (1, 7), (2, 2), (5, 1), (40, 16), (23, 32), (341, 768), (1000, 768), (1024, 1024), (1531, 1024), (65536, 3)]
    pairs += [(int(a), int(b)) for a, b in zip(rng.integers(1, 5000, 300), rng.integers(1, 1200, 300))]
    for n_raw, nf in pairs:
        want = np.linspace(0, n_raw - 1, nf).round().astype(np.uint16)
        got = np.array([host_lib.host_frame_id(n_raw, nf, t) for t in range(nf)])
        assert np.array_equal(got, want), (n_raw, nf)


def test_draws_consume_the_rngs_like_the_reference():
    g = H.load()
    for seed, d in zip(g["s0_seeds"], g["s0_draws"]):
        random.seed(int(seed))
        np.random.seed(int(seed))
        flip, angle, scale = DL.draw_augmentation()
        assert (flip, DL.ROT_ANGLES.index(angle), scale) == (int(d[0]), int(d[1]), float(d[2]))


def test_pack_file_round_trip(tmp_path):
    rng = np.random.default_rng(1)
    samples = [synthetic.make_raw_sample(rng, F, 25, name="r%d" % i) for i, F in enumerate((17, 5, 33))]
    store = DL.PackedSamples.from_samples(samples)
    path = str(tmp_path / "train.p2rpack")
    store.save(path)
    for mmap in (True, False):
        back = DL.PackedSamples.load(path, mmap=mmap)
        assert back.names == store.names
        for name in DL.PackedSamples.ARRAYS:
            H.assert_same(np.asarray(getattr(back, name)), getattr(store, name), name)
    with open(str(tmp_path / "bad"), "wb") as f:
        f.write(b"not a pack")
    with pytest.raises(ValueError):
        DL.PackedSamples.load(str(tmp_path / "bad"))
    with pytest.raises(ValueError):   # the reference cannot represent a sample without boxes either
        DL.PackedSamples.from_samples([dict(samples[0], object_nodes=[])])
    with pytest.raises(ImportError):  # no h5py in this image: the error must say what to do
        DL.PackedSamples.from_hdf5(["/nonexistent.hdf5"])


def test_loader_mirrors_reference_sampler_logic_and_refuses_cpu():
    rng = np.random.default_rng(2)
    store = DL.PackedSamples.from_samples([synthetic.make_raw_sample(rng, 12, 25, name="n%d" % i) for i in range(7)])
    cfg = H.Cfg(num_frames=8, batch_size=3)
    seen = []
    loader = DL.P2RNet_dataloader(cfg, "val", packed=store)
    assert isinstance(loader, DL.Custom_Dataloader) and isinstance(loader.sampler, torch.utils.data.SequentialSampler)
    loader.dataloader.dataset.make_batch = lambda idx: seen.append(list(idx)) or {"n": len(idx)}
    assert len(loader.dataloader) == 3 and [b["n"] for b in loader.dataloader] == [3, 3, 1]
    assert seen == [[0, 1, 2], [3, 4, 5], [6]]
    train = DL.P2RNet_dataloader(cfg, "train", packed=store)
    assert isinstance(train.sampler, torch.utils.data.RandomSampler) and train.dataloader.dataset.aug
    torch.manual_seed(0)
    order_a = [i for b in train.dataloader.batch_sampler for i in b]
    torch.manual_seed(0)
    order_b = list(torch.utils.data.RandomSampler(range(7)))
    assert order_a == order_b and sorted(order_a) == list(range(7))
    ds = DL.P2RNet_VirtualHome(cfg, "test", packed=store, device="cpu")
    with pytest.raises(RuntimeError, match="CUDA device only"):
        ds.make_batch([0])
    with pytest.raises(IndexError):
        ds.make_batch([7])
