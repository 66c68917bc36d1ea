"""CPU, gloo, world size 2: the data-parallel host logic (parameter broadcast, flat gradient all-reduce, AP-state
gather).  The kernels themselves never cross ranks -- the path shards by sample."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from pose2room_b200 import parallel
    from pose2room_b200.ap_helper import APCalculator
    torch.manual_seed(100 + rank)                      # different replicas on purpose
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.BatchNorm1d(5), torch.nn.Linear(5, 3)).double()
    extra = torch.nn.Parameter(torch.randn(4, dtype=torch.float32))   # second dtype bucket
    parallel.broadcast_parameters(net)
    w0 = net[0].weight.detach().clone()
    x = torch.randn(8, 6, dtype=torch.float64, generator=torch.Generator().manual_seed(rank))
    (net(x).sum() + (extra * (rank + 1)).sum()).backward()
    local = [p.grad.clone() for p in net.parameters()] + [extra.grad.clone()]
    parallel.allreduce_gradients(list(net.parameters()) + [extra])
    reduced = [p.grad.clone() for p in net.parameters()] + [extra.grad.clone()]
    calc = APCalculator(0.25)
    calc.pred_map_cls = {0: [(rank, np.full((8, 3), float(rank)), 0.5)]}
    calc.gt_map_cls = {0: [(rank, np.full((8, 3), float(rank)))]}
    calc.scan_cnt = 1
    parallel.gather_ap_state(calc)
    # sample -> batch path: the loader shards the split with the reference's DistributedSampler (dataloader.py:179-180)
    from pose2room_b200 import dataloader as DL, synthetic
    from tests.dataloader_helpers import Cfg
    rng = np.random.default_rng(0)                                        # the same split on every rank
    store = DL.PackedSamples.from_samples([synthetic.make_raw_sample(rng, 6, 25, name="n%d" % i) for i in range(9)])
    loader = DL.P2RNet_dataloader(Cfg(num_frames=4, batch_size=2, distributed=True), "train", packed=store)
    loader.sampler.set_epoch(3)                                           # train_epoch.py:24-25
    shard = [i for b in loader.dataloader.batch_sampler for i in b]
    torch.save(dict(w0=w0, local=local, reduced=reduced, scans=calc.scan_cnt, shard=shard, batches=len(loader.dataloader),
                    classes=sorted(v[0][0] for v in calc.gt_map_cls.values())), os.path.join(out, "r%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_broadcast_allreduce_and_ap_gather_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / ("r%d.pt" % i), weights_only=False) for i in range(world)]
    assert torch.equal(r[0]["w0"], r[1]["w0"])                                   # identical replicas after broadcast
    for i, (a, b) in enumerate(zip(r[0]["reduced"], r[1]["reduced"])):
        assert torch.equal(a, b)                                                 # every rank holds the same gradient
        want = (r[0]["local"][i] + r[1]["local"][i]) / 2
        assert torch.allclose(a, want, rtol=1e-12, atol=1e-12), i               # ... and it is the mean
    assert r[0]["reduced"][-1].dtype == torch.float32 and torch.allclose(r[0]["reduced"][-1], torch.full((4,), 1.5))
    assert r[0]["scans"] == 2 and r[0]["classes"] == [0, 1] and r[1]["classes"] == [0, 1]
    # 9 samples over 2 ranks: 5 each (one padded duplicate, like the reference), disjoint otherwise, 3 batches of <= 2
    assert len(r[0]["shard"]) == len(r[1]["shard"]) == 5 and r[0]["batches"] == r[1]["batches"] == 3
    assert set(r[0]["shard"]) | set(r[1]["shard"]) == set(range(9))
    assert len(set(r[0]["shard"]) & set(r[1]["shard"])) <= 1


def test_single_process_is_a_no_op():
    from pose2room_b200 import parallel
    net = torch.nn.Linear(3, 2)
    net(torch.randn(4, 3)).sum().backward()
    g = net.weight.grad.clone()
    parallel.broadcast_parameters(net)
    parallel.allreduce_gradients(list(net.parameters()))
    assert torch.equal(net.weight.grad, g)
