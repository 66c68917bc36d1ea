"""Sample -> batch path with the raw dataset resident in HBM.

Mirrors /root/reference/models/p2rnet/dataloader.py -- `P2RNet_VirtualHome` (cfg, mode; `__len__`,
`__getitem__`), `collate_fn`, `Custom_Dataloader`, `P2RNet_dataloader(cfg, mode)` -- so that
train_epoch.py / test_epoch.py iterate `loader.dataloader` and call `loader.sampler.set_epoch` unchanged.
What differs is where the work happens:

  reference (per sample, in 12 worker processes)      here (per batch, one kernel launch)
  ---------------------------------------------      -----------------------------------------------
  h5py open + read of (F,J,3) and (F,J,10)            `PackedSamples`: every raw frame of the split packed
                                                      once into two arrays that live in device memory
                                                      (2.7 MB a sample at F=1000, J=53: 60 k samples in 180 GB)
  numpy flip / rotate / translate, float64            `p2r_make_batch` (csrc/dataloader_ops.cu): frame picking,
  frame picking, .astype casts, default_collate,      augmentation with the reference's exact rounding points,
  pinned H2D copy of 1.4 MB per sample                casts and batching straight into the batch tensors;
                                                      host -> device traffic is 132 B of parameters per sample
  rot2head / log size of <= 10 boxes                  same numpy calls on the host (`_box_labels`), <= 10 boxes

The augmentation parameters are drawn on the host with the reference's own RNG calls in the reference's
order (dataloader.py:33-35: random.randint, np.random.choice, random.uniform -- the global `random` and
`np.random` states), one sample after the other in batch order, i.e. the same stream of draws as the
reference with `num_workers: 0`.

There is no CPU path: batches are built by the CUDA kernel or not at all (`_lib.call` raises when the
library is missing; a CPU device raises here).
"""
import json
import os
import random

import numpy as np
import torch
import torch.utils.data
import torch.utils.data.distributed

from . import _lib

FLIP_MATRIX = np.array([[0, 0, 1], [0, 1, 0], [1, 0, 0]])          # dataloader.py:25 (int64 on purpose)
ROT_ANGLES = [-np.pi, -0.5 * np.pi, 0, 0.5 * np.pi]                 # dataloader.py:34
_MAGIC = b"P2RPACK1"
_ALIGN = 4096
PARAM_STRIDE = 16  # doubles per batch item, layout in csrc/augment_math.h


def rot_func(theta):
    """dataloader.py:26-28."""
    return np.array([[np.cos(theta), 0., -np.sin(theta)],
                     [0., 1., 0.],
                     [np.sin(theta), 0, np.cos(theta)]])


def offset_func(scale):
    """dataloader.py:29."""
    return np.array([1., 0., 1.]) * scale


class PackedSamples:
    """Every raw sample of a split, packed: frames back to back, boxes padded to `max_obj` per sample.

    joints f32 [F_total, J, 3] and votes f32 [F_total, J, 10] are the two arrays that go to the device;
    frame_start i64 [N+1] delimits the samples.  Box arrays stay on the host."""

    ARRAYS = ["joints", "votes", "frame_start", "n_obj", "class_id", "centroid", "R_mat", "size"]

    def __init__(self, joints, votes, frame_start, n_obj, class_id, centroid, R_mat, size, names):
        self.joints, self.votes, self.frame_start = joints, votes, frame_start
        self.n_obj, self.class_id, self.centroid, self.R_mat, self.size = n_obj, class_id, centroid, R_mat, size
        self.names = list(names)
        assert joints.dtype == np.float32 and votes.dtype == np.float32 and frame_start.dtype == np.int64
        assert joints.shape[0] == votes.shape[0] == frame_start[-1] and votes.shape[2] == 10
        assert len(self.names) == len(n_obj) == len(frame_start) - 1
        self._floor = {}
        self._device = {}

    def __len__(self):
        return len(self.names)

    @property
    def num_joints(self):
        return self.joints.shape[1]

    # ---- construction ------------------------------------------------------------------------
    @classmethod
    def from_samples(cls, samples, max_obj=10):
        """samples: iterable of dict(skeleton_joints (F,J,3), skeleton_joint_votes (F,J,10), object_nodes =
        list of dict(class_id, centroid, R_mat, size), name) -- the HDF5 schema of
        utils/virtualhome/3_generate_samples.py:188-193, values stored as float32 like utils/tools.py:139."""
        samples = list(samples)
        n = len(samples)
        frames = [int(s["skeleton_joints"].shape[0]) for s in samples]
        frame_start = np.zeros(n + 1, np.int64)
        frame_start[1:] = np.cumsum(frames)
        J = samples[0]["skeleton_joints"].shape[1]
        joints = np.empty((int(frame_start[-1]), J, 3), np.float32)
        votes = np.empty((int(frame_start[-1]), J, 10), np.float32)
        n_obj = np.zeros(n, np.int32)
        class_id = np.zeros((n, max_obj), np.int32)
        centroid = np.zeros((n, max_obj, 3), np.float32)
        R_mat = np.zeros((n, max_obj, 3, 3), np.float32)
        size = np.ones((n, max_obj, 3), np.float32)
        for i, s in enumerate(samples):
            if s["skeleton_joints"].shape[1:] != (J, 3) or s["skeleton_joint_votes"].shape != (frames[i], J, 10):
                raise ValueError("sample %d: inconsistent joint / vote shapes" % i)
            joints[frame_start[i]:frame_start[i + 1]] = s["skeleton_joints"]
            votes[frame_start[i]:frame_start[i + 1]] = s["skeleton_joint_votes"]
            nodes = s["object_nodes"]
            if not 1 <= len(nodes) <= max_obj:
                # the reference would fail too: zero boxes breaks boxes3D[:, 0:3] (dataloader.py:110,124),
                # more than max_gt_boxes breaks the slice assignment (dataloader.py:123)
                raise ValueError("sample %d has %d boxes; need 1..%d" % (i, len(nodes), max_obj))
            n_obj[i] = len(nodes)
            for k, node in enumerate(nodes):
                class_id[i, k] = node["class_id"]
                centroid[i, k] = node["centroid"]
                R_mat[i, k] = node["R_mat"]
                size[i, k] = node["size"]
        return cls(joints, votes, frame_start, n_obj, class_id, centroid, R_mat, size,
                   [str(s.get("name", i)) for i, s in enumerate(samples)])

    @classmethod
    def from_hdf5(cls, paths, max_obj=10):
        """Pack the reference's per-sample HDF5 files (read like dataloader.py:88-101, nodes in h5py's key
        order).  Needs h5py, which this image does not have: pack once where it is available, `save`, and
        ship the pack."""
        try:
            import h5py
            h5py.File
        except (ImportError, AttributeError) as e:      # absent, or a stub module standing in for it
            raise ImportError("h5py is needed to read the reference's .hdf5 samples; pack them on a machine "
                              "that has it (PackedSamples.from_hdf5(...).save(path)) and load the pack here") from e

        def read(path):
            with h5py.File(path, "r") as f:
                nodes = [dict(class_id=f["object_nodes"][k]["class_id"][0], centroid=f["object_nodes"][k]["centroid"][:],
                              R_mat=f["object_nodes"][k]["R_mat"][:], size=f["object_nodes"][k]["size"][:])
                         for k in f["object_nodes"].keys()]
                return dict(skeleton_joints=f["skeleton_joints"][:], skeleton_joint_votes=f["skeleton_joint_votes"][:],
                            object_nodes=nodes, name=".".join(os.path.basename(path).split(".")[:-1]))
        return cls.from_samples((read(p) for p in paths), max_obj=max_obj)

    # ---- on-disk pack: one file, JSON header + page-aligned raw arrays (np.memmap-able) -------
    def save(self, path):
        header = {"names": self.names, "arrays": {}}
        offset = 0
        for name in self.ARRAYS:
            a = np.ascontiguousarray(getattr(self, name))
            header["arrays"][name] = {"dtype": a.dtype.str, "shape": list(a.shape), "offset": offset}
            offset += -(-a.nbytes // _ALIGN) * _ALIGN
        blob = json.dumps(header).encode()
        data_start = -(-(len(_MAGIC) + 8 + len(blob)) // _ALIGN) * _ALIGN
        with open(path, "wb") as f:
            f.write(_MAGIC)
            f.write(np.array([len(blob)], np.int64).tobytes())
            f.write(blob)
            for name in self.ARRAYS:
                f.seek(data_start + header["arrays"][name]["offset"])
                f.write(np.ascontiguousarray(getattr(self, name)).tobytes())
            f.truncate(data_start + offset)

    @classmethod
    def load(cls, path, mmap=True):
        with open(path, "rb") as f:
            if f.read(len(_MAGIC)) != _MAGIC:
                raise ValueError("%s is not a P2RPACK1 file" % path)
            n = int(np.frombuffer(f.read(8), np.int64)[0])
            header = json.loads(f.read(n).decode())
        data_start = -(-(len(_MAGIC) + 8 + n) // _ALIGN) * _ALIGN
        arrays = {}
        for name in cls.ARRAYS:
            d = header["arrays"][name]
            shape = tuple(d["shape"])
            if int(np.prod(shape)) == 0:
                arrays[name] = np.zeros(shape, np.dtype(d["dtype"]))
            elif mmap:
                arrays[name] = np.memmap(path, np.dtype(d["dtype"]), "r", data_start + d["offset"], shape)
            else:
                arrays[name] = np.fromfile(path, np.dtype(d["dtype"]), int(np.prod(shape)),
                                           offset=data_start + d["offset"]).reshape(shape)
        return cls(names=header["names"], **arrays)

    # ---- residency ---------------------------------------------------------------------------
    def device_arrays(self, device):
        """(joints, votes, frame_start) as device tensors; uploaded once per device, then resident."""
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("pose2room_b200.dataloader builds batches on a CUDA device only (got %s)" % device)
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        if key not in self._device:
            self._device[key] = tuple(torch.from_numpy(np.ascontiguousarray(a)).to(device)
                                      for a in (self.joints, self.votes, self.frame_start))
        return self._device[key]

    def floor_height(self, idx, augmented):
        """np.percentile(joints[..., 1], 0.99) of dataloader.py:113.  Height is untouched by the augmentation
        (flip swaps x/z, the rotation is about y, the shift has no y part), so it is a property of the raw
        sample; only its dtype differs: float64 once augment_data has run, float32 otherwise."""
        key = (int(idx), bool(augmented))
        if key not in self._floor:
            y = np.asarray(self.joints[self.frame_start[idx]:self.frame_start[idx + 1], :, 1])
            self._floor[key] = float(np.percentile(y.astype(np.float64) if augmented else y, 0.99))
        return self._floor[key]


def draw_augmentation():
    """The three draws of augment_data, same calls and order (dataloader.py:33-35)."""
    if_flip = random.randint(0, 1)
    rot_angle = np.random.choice(ROT_ANGLES)
    offset_scale = random.uniform(-1., 1.)
    return if_flip, rot_angle, offset_scale


def _box_labels(store, idx, draws, max_num_obj):
    """center / log-size / heading / class / mask rows of one sample (dataloader.py:46-51,74-76,81-82,103-126).
    Only the first row of R_mat reaches the output (rot2head reads R[0], utils/pc_utils.py:44), so the
    `R_mat[2] = cross(...)` repair of the flip (dataloader.py:50) has no effect on the labels and is skipped."""
    n = int(store.n_obj[idx])
    centroid = np.asarray(store.centroid[idx, :n])
    heading_vec = np.asarray(store.R_mat[idx, :n, 0])
    size = np.asarray(store.size[idx, :n])
    if draws is not None:
        if_flip, rot_angle, offset_scale = draws
        rot_mat = rot_func(rot_angle)
        if if_flip:
            centroid = np.dot(centroid, FLIP_MATRIX)
            heading_vec = np.dot(heading_vec, FLIP_MATRIX)
        centroid = np.dot(centroid, rot_mat) + offset_func(offset_scale)
        heading_vec = np.dot(heading_vec, rot_mat)
    heading = np.arctan2(-heading_vec[:, 2], heading_vec[:, 0])
    boxes = np.hstack([centroid, np.log(size), np.sin(heading)[:, None], np.cos(heading)[:, None]])
    labels = np.zeros((max_num_obj, 9), np.float32)      # mask, centre, log size, sin, cos
    labels[:n, 0] = 1
    labels[:n, 1:] = boxes
    classes = np.zeros(max_num_obj, np.int64)
    classes[:n] = store.class_id[idx, :n]
    return labels, classes


class P2RNet_VirtualHome:
    """Dataset with the reference's constructor and item schema (dataloader.py:16-146); items and batches are
    CUDA tensors.  `cfg.config['data']` keys read: `split` (directory holding <mode>.json, models/datasets.py:18-19),
    `num_frames`, `no_height`, `max_gt_boxes`, plus `packed` (optional: a PackedSamples file covering the split;
    without it the .hdf5 files of the split are packed at start-up, which needs h5py)."""

    def __init__(self, cfg, mode, packed=None, device=None):
        self.config = cfg.config
        self.dataset_config = getattr(cfg, "dataset_config", None)
        self.mode = mode
        data = cfg.config["data"]
        self.aug = mode == "train"
        self.num_frames = data["num_frames"]
        self.use_height = not data["no_height"]
        self.max_num_obj = data["max_gt_boxes"]
        if packed is None:
            if data.get("packed"):
                packed = PackedSamples.load(os.path.join(data["packed"], mode + ".p2rpack")
                                            if os.path.isdir(data["packed"]) else data["packed"])
            else:
                with open(os.path.join(data["split"], mode + ".json")) as f:
                    packed = PackedSamples.from_hdf5(json.load(f), max_obj=self.max_num_obj)
        self.packed = packed
        self.split = list(packed.names)
        self.device = None if device is None else torch.device(device)   # None = the current CUDA device

    def __len__(self):
        return len(self.packed)

    def host_side(self, indices, draws):
        """Everything the host contributes to a batch: the per-item parameter blocks of the kernel
        (csrc/augment_math.h) and the box labels.  Returns (params f64 [B,16], labels f32 [B,max_obj,9] =
        (mask, centre, log size, sin, cos), classes i64 [B,max_obj])."""
        B = len(indices)
        params = np.zeros((B, PARAM_STRIDE), np.float64)
        labels = np.zeros((B, self.max_num_obj, 9), np.float32)
        classes = np.zeros((B, self.max_num_obj), np.int64)
        for b, (i, d) in enumerate(zip(indices, draws)):
            if d is not None:
                params[b, 0] = 1.0
                params[b, 1] = float(d[0])
                params[b, 2:11] = rot_func(d[1]).reshape(9)
                params[b, 11:14] = offset_func(d[2])
            if self.use_height:
                params[b, 14] = self.packed.floor_height(i, d is not None)
            labels[b], classes[b] = _box_labels(self.packed, i, d, self.max_num_obj)
        return params, labels, classes

    def make_batch(self, indices, draws=None):
        """Collated batch (dataloader.py:148-160) of the samples `indices`, as CUDA tensors.
        draws: None = draw here when mode == 'train'; or a list of (if_flip, rot_angle, offset_scale) / None."""
        indices = [int(i) for i in indices]
        B, T, J = len(indices), self.num_frames, self.packed.num_joints
        for i in indices:
            if not 0 <= i < len(self.packed):
                raise IndexError("sample index %d out of range" % i)
        if draws is None:
            draws = [draw_augmentation() if self.aug else None for _ in indices]
        dev = self.device if self.device is not None else torch.device("cuda", torch.cuda.current_device())
        joints, votes, frame_start = self.packed.device_arrays(dev)
        C = 4 if self.use_height else 3
        params, labels, classes = self.host_side(indices, draws)
        with torch.cuda.device(dev):
            params_d = torch.from_numpy(params).to(dev, non_blocking=True)
            ids_d = torch.tensor(indices, dtype=torch.int32).to(dev, non_blocking=True)
            labels_d = torch.from_numpy(labels).to(dev, non_blocking=True)
            classes_d = torch.from_numpy(classes).to(dev, non_blocking=True)
            input_joints = torch.empty(B, T, J, C, dtype=torch.float32, device=dev)
            vote_label = torch.empty(B, T, J, 9, dtype=torch.float32, device=dev)
            vote_label_mask = torch.empty(B, T, J, dtype=torch.int64, device=dev)
            _lib.call("p2r_make_batch", joints.data_ptr(), votes.data_ptr(), frame_start.data_ptr(),
                      ids_d.data_ptr(), params_d.data_ptr(), B, T, J, C, input_joints.data_ptr(),
                      vote_label.data_ptr(), vote_label_mask.data_ptr(), torch.cuda.current_stream().cuda_stream)
        return {
            "input_joints": input_joints,
            "box_label_mask": labels_d[:, :, 0].contiguous(),
            "sem_cls_label": classes_d,
            "center_label": labels_d[:, :, 1:4].contiguous(),
            "size": labels_d[:, :, 4:7].contiguous(),
            "heading": labels_d[:, :, 7:9].contiguous(),
            "vote_label": vote_label,
            "vote_label_mask": vote_label_mask,
            "sample_idx": [self.split[i] for i in indices],
        }

    def __getitem__(self, idx):
        return {k: v[0] for k, v in self.make_batch([idx]).items()}


def collate_fn(batch):
    """dataloader.py:148-160 for items that are already CUDA tensors."""
    out = {}
    for key in batch[0]:
        if key == "sample_idx":
            out[key] = [elem[key] for elem in batch]
        else:
            out[key] = torch.stack([elem[key] for elem in batch])
    return out


class DeviceBatchLoader:
    """What `Custom_Dataloader.dataloader` is here: iterating it yields one collated CUDA batch per entry of
    the batch sampler, `len()` is the number of batches (train_epoch.py:33,46)."""

    def __init__(self, dataset, batch_sampler):
        self.dataset = dataset
        self.batch_sampler = batch_sampler

    def __len__(self):
        return len(self.batch_sampler)

    def __iter__(self):
        for indices in self.batch_sampler:
            yield self.dataset.make_batch(indices)


class Custom_Dataloader(object):
    """dataloader.py:162-165."""

    def __init__(self, dataloader, sampler):
        self.dataloader = dataloader
        self.sampler = sampler


def P2RNet_dataloader(cfg, mode="train", packed=None, device=None):
    """dataloader.py:172-199: same samplers (so the same sample order under the same torch seed), same return
    type; `num_workers` is ignored -- there are no worker processes to feed."""
    if cfg.config["data"]["dataset"] != "virtualhome":
        raise NotImplementedError
    dataset = P2RNet_VirtualHome(cfg, mode, packed=packed, device=device)
    if cfg.config["device"]["distributed"]:
        sampler = torch.utils.data.distributed.DistributedSampler(dataset, shuffle=(mode == "train"))
    elif mode == "train":
        sampler = torch.utils.data.RandomSampler(dataset)
    else:
        sampler = torch.utils.data.SequentialSampler(dataset)
    batch_sampler = torch.utils.data.BatchSampler(sampler, batch_size=cfg.config[mode]["batch_size"], drop_last=False)
    return Custom_Dataloader(DeviceBatchLoader(dataset, batch_sampler), sampler)
