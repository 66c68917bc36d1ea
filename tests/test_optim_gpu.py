"""pose2room_b200.optim.AdamW (csrc/optim_ops.cu) against torch.optim.AdamW -- the optimiser the reference's factory builds
(models/optimizers.py:90) -- on the same float32 tensors and gradients: same parameters after several updates, with weight
decay, odd sizes (scalar tails, unaligned views), more tensors than one launch takes, a state-dict round trip, and a
captured CUDA graph."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture
def cuda():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def _params(dev, seed, n_tensors):
    g = torch.Generator().manual_seed(seed)
    shapes = [(1,), (3,), (7, 5), (64, 64), (1000,), (4097,), (259, 256), (704, 64, 1, 1), (11, 25, 25)]
    out = []
    for i in range(n_tensors):
        t = torch.randn(shapes[i % len(shapes)], generator=g).to(dev)
        out.append(t.double() if i % 17 == 5 else t)       # (P2RNet has one float64 parameter: the torch-kernel route)
    return out


@pytest.mark.parametrize("n_tensors,wd", [(9, 1e-2), (131, 0.0), (60, 0.1)])
def test_adamw_matches_torch(cuda, n_tensors, wd):
    from pose2room_b200.optim import AdamW
    init = _params(cuda, 1, n_tensors)
    mine = [torch.nn.Parameter(t.clone()) for t in init]
    ref = [torch.nn.Parameter(t.clone()) for t in init]
    o_mine = AdamW(mine, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    o_ref = torch.optim.AdamW(ref, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    g = torch.Generator().manual_seed(2)
    for step in range(6):
        for a, b in zip(mine, ref):
            gr = (torch.randn(a.shape, generator=g) * (10.0 ** (step % 3 - 1))).to(cuda).to(a.dtype)
            a.grad = gr.clone()
            b.grad = gr.clone()
        o_mine.step()
        o_ref.step()
    for i, (a, b) in enumerate(zip(mine, ref)):
        assert torch.allclose(a, b, rtol=2e-6, atol=2e-7), (i, tuple(a.shape), (a - b).abs().max().item())
        sm, sr = o_mine.state[a], o_ref.state[b]
        assert torch.allclose(sm["exp_avg"], sr["exp_avg"], rtol=2e-6, atol=1e-9)
        assert torch.allclose(sm["exp_avg_sq"], sr["exp_avg_sq"], rtol=2e-6, atol=1e-12)
        assert float(sm["step"]) == float(sr["step"]) == 6.0


def test_adamw_state_dict_round_trip_and_cuda_graph(cuda):
    from pose2room_b200.optim import AdamW
    init = _params(cuda, 3, 20)
    ps = [torch.nn.Parameter(t.clone()) for t in init]
    ref = [torch.nn.Parameter(t.clone()) for t in init]
    grads = [torch.randn_like(p) for p in ps]
    opt = AdamW(ps, lr=2e-3, weight_decay=0.05)
    o_ref = torch.optim.AdamW(ref, lr=2e-3, weight_decay=0.05)
    for p, g in zip(ps, grads):
        p.grad = g.clone()
    for p, g in zip(ref, grads):
        p.grad = g.clone()
    opt.step()
    opt.step()
    # a fresh optimiser picks the state up (separate step tensors after loading: re-shared on the next step)
    opt2 = AdamW(ps, lr=2e-3, weight_decay=0.05)
    opt2.load_state_dict(opt.state_dict())
    # ... and steps inside a captured graph, replayed twice
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        opt2.step()
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        opt2.step()
    graph.replay()
    torch.cuda.synchronize()
    for _ in range(4):          # 2 eager + 1 warm-up + 1 replay (capturing does not execute)
        o_ref.step()
    assert float(opt2.state[ps[0]]["step"]) == 4.0
    for a, b in zip(ps, ref):
        assert torch.allclose(a, b, rtol=5e-6, atol=5e-7), (tuple(a.shape), (a - b).abs().max().item())


def test_adamw_refuses_cpu_parameters():
    from pose2room_b200.optim import AdamW
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError):
        AdamW([p]).step()
