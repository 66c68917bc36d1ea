"""Host-side behaviour of pose2room_b200.optim.AdamW that needs no GPU: hyper-parameter validation, torch's defaults, and
the loud failure on CPU parameters (there is no CPU path; the CUDA update itself is held to torch.optim.AdamW in
tests/test_optim_gpu.py)."""
import pytest
import torch


def test_defaults_are_torch_adamw_defaults():
    from pose2room_b200.optim import AdamW
    p = torch.nn.Parameter(torch.zeros(3))
    ours, ref = AdamW([p]), torch.optim.AdamW([p])
    for k in ("lr", "betas", "eps", "weight_decay"):
        assert ours.defaults[k] == ref.defaults[k], k


@pytest.mark.parametrize("kw", [dict(lr=-1.0), dict(eps=-1e-8), dict(weight_decay=-0.1), dict(betas=(1.0, 0.999)),
                                dict(betas=(0.9, -0.1))])
def test_invalid_hyper_parameters_raise(kw):
    from pose2room_b200.optim import AdamW
    with pytest.raises(ValueError):
        AdamW([torch.nn.Parameter(torch.zeros(2))], **kw)


def test_cpu_parameters_fail_loudly_and_parameters_without_gradient_are_skipped():
    from pose2room_b200.optim import AdamW
    a, b = torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.ones(4))
    opt = AdamW([a, b])
    opt.step()                                   # no gradients anywhere: nothing to do, nothing raised
    assert torch.equal(b.detach(), torch.ones(4)) and not opt.state
    a.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match="no CPU path"):
        opt.step()
