"""GPU, 2 ranks over NCCL (skipped on a one-GPU box; run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`):
the data-parallel step of the product (the reference's DDP axis, net_utils/utils.py:251).

  * identical batch on both ranks  => loss and every gradient after the flat all-reduce are what a single GPU computes
    (AVG of two identical values is that value), in fp32 and in bf16 mode;
  * different batches per rank     => the reduced gradient is the mean of the two local gradients, identical on both ranks,
    and after AdamW the replicas still hold identical parameters.
"""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from pose2room_b200 import _lib, gemm_sm100, parallel, synthetic
    from pose2room_b200.config import P2RConfig
    from pose2room_b200.p2rnet import P2RNet
    _lib.load()
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    T, J, S, P, B = 256, 25, 128, 32, 4
    res = {}
    for precision in ("fp32", "bf16"):
        if precision == "bf16":
            gemm_sm100.install()
        for case in ("same_batch", "different_batches"):
            torch.manual_seed(0)
            np.random.seed(0)
            net = P2RNet(P2RConfig(mode="train", joint_num=J, num_frames=T, precision=precision, num_seeds=S, num_target=P))
            sd = synthetic.deterministic_state_dict(net.state_dict(), seed=7 + rank)      # different replicas on purpose
            net.load_state_dict(sd)
            net = net.to(dev).train()
            parallel.broadcast_parameters(net)                                            # -> rank 0's weights everywhere
            params = [p for p in net.parameters() if p.requires_grad]
            opt = torch.optim.AdamW(params, lr=1e-3)
            seed = 50 if case == "same_batch" else 50 + rank
            data = {k: (v.to(dev) if isinstance(v, torch.Tensor) else v) for k, v in synthetic.make_batch(B, T, J, seed=seed).items()}
            torch.manual_seed(1)                                                          # same mixture-head noise on both ranks
            loss = net.loss(net(data), data)["total"]
            loss.backward()
            local = [p.grad.detach().clone() for p in params]
            parallel.allreduce_gradients(params)
            reduced = [p.grad.detach().clone() for p in params]
            opt.step()
            gathered = [[torch.empty_like(g) for _ in range(world)] for g in local]
            for g, bucket in zip(local, gathered):
                dist.all_gather(bucket, g)
            res[precision, case] = dict(loss=float(loss), local=[g.cpu() for g in local], reduced=[g.cpu() for g in reduced],
                                        mean=[(sum(b.double() for b in bucket) / world).cpu() for bucket in gathered],
                                        after=[p.detach().cpu().clone() for p in params])
        if precision == "bf16":
            gemm_sm100.uninstall()
    torch.save(res, os.path.join(out, "r%d.pt" % rank))
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_rank_step_matches_single_gpu(cuda, tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    import torch.multiprocessing as mp
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    r = [torch.load(tmp_path / ("r%d.pt" % i), weights_only=False) for i in range(world)]
    for precision in ("fp32", "bf16"):
        same = [x[precision, "same_batch"] for x in r]
        assert abs(same[0]["loss"] - same[1]["loss"]) <= 1e-6 * abs(same[0]["loss"]), (precision, same[0]["loss"], same[1]["loss"])
        # the single-GPU gradient (rank 0's local one) == the all-reduced one, on both ranks.  Bit-exact whenever the two
        # GPUs computed bit-identical local gradients (the BatchNorm statistics are double-precision atomics: their
        # summation order, hence the last bit, may differ between two devices), else to rounding
        tol = 1e-5 if precision == "fp32" else 5e-3
        exact = 0
        for i, (loc, red0, red1) in enumerate(zip(same[0]["local"], same[0]["reduced"], same[1]["reduced"])):
            assert torch.equal(red0, red1), (precision, i)
            exact += int(torch.equal(loc, red0))
            scale = float(loc.abs().max()) + 1e-12
            assert float((loc.double() - red0.double()).abs().max()) <= tol * scale, (precision, i)
        print("%s: %d of %d gradient tensors bit-identical to the single-GPU ones" % (precision, exact, len(same[0]["local"])))
        diff = [x[precision, "different_batches"] for x in r]
        assert diff[0]["loss"] != diff[1]["loss"]
        for i, (red0, red1, mean) in enumerate(zip(diff[0]["reduced"], diff[1]["reduced"], diff[0]["mean"])):
            assert torch.equal(red0, red1), (precision, i)
            scale = float(mean.abs().max()) + 1e-12
            assert float((red0.double() - mean).abs().max()) <= 2e-6 * scale, (precision, i)
        for i, (a, b) in enumerate(zip(diff[0]["after"], diff[1]["after"])):
            assert torch.equal(a, b), (precision, i)          # replicas stay identical after the optimiser step
